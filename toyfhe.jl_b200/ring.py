"""Host-side mirror of the reference's ring layer -- ``module NTT`` of
src/pow2_cyc_rings.jl and the RNS pieces of src/crt.jl -- with the same names and
semantics, every numeric operation being one call into the CUDA engine.

(The reference is Julia; its toolchain is absent here, so this mirror is Python.
julia/ToyFHEB200.jl shows the equivalent `ccall` methods.)

    NegacyclicRing        pow2_cyc_rings.jl:27-47, crt.jl:282-295
    RingElement           pow2_cyc_rings.jl:93-145  (lazy primal/dual cache)
    nntt / inntt          pow2_cyc_rings.jl:295-318, crt.jl:247-267
    * + - ^               pow2_cyc_rings.jl:147-224
    apply_galois_element  pow2_cyc_rings.jl:321-329
    crtselect/drop_last   crt.jl:185-213
    modswitch(_drop)      crt.jl:215-236
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np

from . import engine as E


def _is_primitive_root(psi: int, n: int, q: int) -> bool:
    """is_primitive_root (pow2_cyc_rings.jl:22)"""
    return pow(psi, n, q) == 1


class NegacyclicRing:
    """F_q[x]/(x^N+1) over an RNS basis, with identified primitive 2N-th roots psi_i.

    ``NegacyclicRing(N, logqs=[60, 60])`` follows crt.jl:282-295 (prime chain + minimal roots);
    ``NegacyclicRing(N, qs=[...], psis=[...])`` follows pow2_cyc_rings.jl:27-47."""

    def __init__(self, N: int, logqs: Optional[Sequence[int]] = None, qs: Optional[Sequence[int]] = None,
                 psis: Optional[Sequence[int]] = None, device: Optional[int] = None):
        if logqs is not None:
            qs, psis = E.prime_chain(N, logqs)
        if qs is None:
            raise ValueError("need logqs or qs")
        if psis is None:
            psis = [E.minimal_primitive_root(q, 2 * N) for q in qs]   # pow2_cyc_rings.jl:38-41
        for q, p in zip(qs, psis):
            assert _is_primitive_root(p, 2 * N, q)                   # pow2_cyc_rings.jl:31
        self.N, self.qs, self.psis = int(N), [int(q) for q in qs], [int(p) for p in psis]
        self.L = len(self.qs)
        self.device = device
        self.ctx = E.Context(self.N, self.qs, self.psis, device=device)
        self._sub = {}

    # -- reference accessors
    def degree(self) -> int:
        return self.N

    def modulus(self) -> int:
        """NTT.modulus(CRTEncoded) = prod q_i (crt.jl:80)"""
        return math.prod(self.qs)

    def __repr__(self):
        return f"NegacyclicRing(N={self.N}, q={self.qs})"

    def __eq__(self, other):
        return isinstance(other, NegacyclicRing) and (self.N, self.qs, self.psis) == (other.N, other.qs, other.psis)

    def __hash__(self):
        return hash((self.N, tuple(self.qs)))

    # -- constructors of elements
    def zero(self) -> "RingElement":
        return RingElement(self, primal=self.ctx.empty((self.L, self.N)).zero_())

    def from_residues(self, a: np.ndarray, dual: bool = False) -> "RingElement":
        a = np.ascontiguousarray(a, dtype=np.uint64).reshape(self.L, self.N)
        t = self.ctx.to_device(a)
        return RingElement(self, dual=t) if dual else RingElement(self, primal=t)

    def __call__(self, coeffs: Sequence[int]) -> "RingElement":
        """ring(coeffs): integer coefficients -> CRTEncoded residues (crt.jl:91-95)"""
        assert len(coeffs) == self.N
        a = np.empty((self.L, self.N), dtype=np.uint64)
        for i, q in enumerate(self.qs):
            a[i] = [int(c) % q for c in coeffs]
        return self.from_residues(a)

    # -- crtselect / drop_last (crt.jl:185-213)
    def crtselect(self, which: Sequence[int]) -> "NegacyclicRing":
        key = tuple(which)
        if key not in self._sub:
            self._sub[key] = NegacyclicRing(self.N, qs=[self.qs[i] for i in which], psis=[self.psis[i] for i in which],
                                            device=self.device)
        return self._sub[key]

    def drop_last(self) -> "NegacyclicRing":
        return self.crtselect(range(self.L - 1))


class RingElement:
    """Element of the ring with lazily cached primal / dual coefficient tensors
    (pow2_cyc_rings.jl:93-138).  Tensors are int64 [L][N] on the ring's GPU."""

    def __init__(self, ring: NegacyclicRing, primal=None, dual=None):
        assert primal is not None or dual is not None      # pow2_cyc_rings.jl:116
        self.ring, self.primal, self.dual = ring, primal, dual

    # -- lazy conversions (pow2_cyc_rings.jl:124-138)
    def coeffs_primal(self):
        if self.primal is None:
            self.primal = self.ring.ctx.ntt_inv(self.dual)
        return self.primal

    def coeffs_dual(self):
        if self.dual is None:
            self.dual = self.ring.ctx.ntt_fwd(self.primal)
        return self.dual

    def copy(self) -> "RingElement":
        return RingElement(self.ring, None if self.primal is None else self.primal.clone(),
                           None if self.dual is None else self.dual.clone())

    # -- host views
    def residues(self) -> np.ndarray:
        return E.Context.to_host(self.coeffs_primal())

    def to_ints(self) -> List[int]:
        """convert(Integer, ::CRTEncoded) per coefficient: X in [0,Q) (crt.jl:105-112)"""
        res = self.residues()
        qs = self.ring.qs
        X = [int(v) for v in res[0]]
        M = qs[0]
        for i in range(1, len(qs)):
            q = qs[i]
            inv = pow(M, -1, q)
            ri = res[i]
            X = [x + M * (((int(r) - x) * inv) % q) for x, r in zip(X, ri)]
            M *= q
        return X

    def to_signed_ints(self) -> List[int]:
        """SignedMod lift (signedmod.jl:12-19)"""
        Q = self.ring.modulus()
        return [x - Q if x > Q // 2 else x for x in self.to_ints()]

    # -- ring_multiply (pow2_cyc_rings.jl:147-173): dual .* dual, result dual-only
    def __mul__(self, other):
        if isinstance(other, RingElement):
            assert other.ring == self.ring
            return RingElement(self.ring, dual=self.ring.ctx.mul(self.coeffs_dual(), other.coeffs_dual()))
        if isinstance(other, (int, np.integer)):
            return self.scalar_mul(int(other))
        return NotImplemented

    __rmul__ = __mul__

    def scalar_mul(self, s: int) -> "RingElement":
        """scalar_mul (pow2_cyc_rings.jl:177-185): applied to whichever caches exist"""
        c = self.ring.ctx
        return RingElement(self.ring, None if self.primal is None else c.scalar_mul(self.primal, s),
                           None if self.dual is None else c.scalar_mul(self.dual, s))

    def __neg__(self):
        c = self.ring.ctx
        return RingElement(self.ring, None if self.primal is None else c.neg(self.primal),
                           None if self.dual is None else c.neg(self.dual))

    def _addsub(self, other, fn):
        # pow2_cyc_rings.jl:192-219: operate in every domain both operands have; if they share
        # none, convert so that both results exist
        assert isinstance(other, RingElement) and other.ring == self.ring
        new_primal = new_dual = None
        if self.primal is not None and other.primal is not None:
            new_primal = fn(self.primal, other.primal)
        if self.dual is not None and other.dual is not None:
            new_dual = fn(self.dual, other.dual)
        if new_primal is None and new_dual is None:
            new_primal = fn(self.coeffs_primal(), other.coeffs_primal())
            new_dual = fn(self.coeffs_dual(), other.coeffs_dual())
        return RingElement(self.ring, new_primal, new_dual)

    def __add__(self, other):
        return self._addsub(other, self.ring.ctx.add)

    def __sub__(self, other):
        return self._addsub(other, self.ring.ctx.sub)

    def __pow__(self, n: int):
        """Base.power_by_squaring (pow2_cyc_rings.jl:221-224)"""
        assert n >= 1
        result, base = None, self
        while n:
            if n & 1:
                result = base if result is None else result * base
            n >>= 1
            if n:
                base = base * base
        return result

    def apply_galois_element(self, g: int) -> "RingElement":
        """pow2_cyc_rings.jl:321-329 (primal domain, primal-only result)"""
        return RingElement(self.ring, primal=self.ring.ctx.galois(self.coeffs_primal(), g))

    # -- RNS level changes (crt.jl:185-236)
    def crtselect(self, which: Sequence[int]) -> "RingElement":
        import torch
        sub = self.ring.crtselect(which)
        idx = torch.as_tensor(list(which), device=(self.primal if self.primal is not None else self.dual).device)
        sel = lambda t: None if t is None else t.index_select(0, idx).contiguous()
        return RingElement(sub, sel(self.primal), sel(self.dual))

    def drop_last(self) -> "RingElement":
        return self.crtselect(range(self.ring.L - 1))

    def modswitch(self) -> "RingElement":
        """exact division by the last prime (crt.jl:215-220, 226-228); primal-only result"""
        return RingElement(self.ring.drop_last(), primal=self.ring.ctx.rescale(self.coeffs_primal()))

    def modswitch_drop(self) -> "RingElement":
        """crt.jl:222-224, 230-232: drop the last residue of the primal coefficients"""
        return RingElement(self.ring.drop_last(), primal=self.coeffs_primal()[:-1].contiguous())

    def crt_expand(self, P: int, target: NegacyclicRing) -> "RingElement":
        """c .* CRTExpand{P} (crt.jl:35-40; modulusraising.jl:35-41)"""
        assert target.qs[:-1] == self.ring.qs and target.qs[-1] == P
        return RingElement(target, primal=self.ring.ctx.crt_expand(self.coeffs_primal(), P))


def nntt(r: RingElement):
    """NTT.nntt on the primal coefficients (pow2_cyc_rings.jl:295-303; crt.jl:247-256)"""
    return r.ring.ctx.ntt_fwd(r.coeffs_primal())


def inntt(r: RingElement):
    """NTT.inntt on the dual coefficients (pow2_cyc_rings.jl:308-318; crt.jl:258-267)"""
    return r.ring.ctx.ntt_inv(r.coeffs_dual())
