"""TEST INFRASTRUCTURE ONLY (see oracle/toyfhe_oracle.py): numpy float64 restatement of the CKKS encoding of the
reference, src/ckksencoding.jl -- the checker for tfb_ckks_encode / tfb_ckks_decode (tolerance parity: float64 FFT).

The FFT itself lives in a third-party dependency of the reference (FFTW via FFTW.jl, src/ckksencoding.jl:1); numpy's
pocketfft computes the same DFT definition (forward: sum x[n] exp(-2 pi i k n / N); inverse: 1/N sum ... exp(+...)).
Parity status: no reference test pins an encoded integer; the semantic pins are the decrypt-level tolerances of
test/ckks_*.jl (1e-3 .. 1e-8), which the GPU replays in tests/test_gpu_scheme.py meet."""
from __future__ import annotations

from fractions import Fraction
from typing import List, Sequence

import numpy as np


def zmstar_positions(N: int) -> List[int]:
    """ZmstarPermutation(2N)[1, :] .>> 1 (ckksencoding.jl:64, 82): slot i (0-based) <-> exponent 3^(i+1) mod 2N, position
    (3^(i+1) mod 2N) >> 1 of the length-N spectrum"""
    g, out = 1, []
    for _ in range(N // 2):
        g = g * 3 % (2 * N)
        out.append(g >> 1)
    return out


def encode(data: Sequence[complex], scale, N: int) -> List[int]:
    """convert(RingElement, ::CKKSEncoding) (ckksencoding.jl:76-101): slots and their conjugates scattered into a length-N
    spectrum, ifft, the negacyclic twist exp(2 pi i k / 2N), real part, FixedRational rounding (ckks.jl:30-58: round to
    nearest of x * scale, exact here through Fraction)"""
    n = N // 2
    cm = np.zeros(N, dtype=np.complex128)
    g = 1
    for i in range(n):
        g = g * 3 % (2 * N)
        cm[g >> 1] = data[i]                         # idxs[1, i+1]
        cm[(2 * N - g) >> 1] = np.conj(data[i])      # idxs[2, i+1]
    nip = np.fft.ifft(cm) * np.exp(2j * np.pi * np.arange(N) / (2 * N))
    assert np.abs(nip.imag).max() < 1e-9 * max(1.0, np.abs(nip).max())      # @assert isapprox(imag(p), 0, atol=10^-10)
    return [int(round(Fraction(float(x)) * Fraction(scale))) for x in nip.real]


def decode(coeffs_centred: Sequence[int], scale, N: int) -> np.ndarray:
    """CKKSEncoding{ScaleT}(plain) (ckksencoding.jl:60-70): centred coefficients / scale, twist exp(-2 pi i k / 2N), fft,
    the non-conjugated slots in ZmstarPermutation order"""
    cen = np.array([float(Fraction(int(x)) / Fraction(scale)) for x in coeffs_centred])
    F = np.fft.fft(cen * np.exp(-2j * np.pi * np.arange(N) / (2 * N)))
    return F[zmstar_positions(N)]
