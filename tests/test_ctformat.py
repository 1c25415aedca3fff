"""On-disk ciphertext container (toyfhe.jl_b200/ctformat.py): round trip, header checks, corruption detection."""
import numpy as np
import pytest

import toyfhe_b200 as T
from toyfhe_b200 import ctformat as F


def _batch(rng, qs, N, B=3, comps=2):
    a = np.empty((B, comps, len(qs), N), dtype=np.uint64)
    for i, q in enumerate(qs):
        a[:, :, i] = rng.integers(0, q, size=(B, comps, N), dtype=np.uint64)
    return a


def test_round_trip_and_checks(tmp_path):
    N = 64
    qs, psis = T.prime_chain(N, [60, 40, 40])
    rng = np.random.default_rng(0)
    a = _batch(rng, qs, N)
    p = tmp_path / "ct.tfb"
    F.save(str(p), a, qs, psis, dual=False, scale=2.0 ** 40)
    got = F.load(str(p))
    assert got.N == N and got.qs == qs and got.psis == psis and not got.dual and got.scale == 2.0 ** 40
    assert np.array_equal(got.residues, a)
    raw = bytearray(p.read_bytes())
    assert len(raw) == 48 + 16 * 3 + a.nbytes + 4
    raw[100] ^= 1
    with pytest.raises(F.FormatError, match="checksum"):
        F.loads(bytes(raw))
    with pytest.raises(F.FormatError, match="size"):
        F.loads(bytes(p.read_bytes()[:-8]))
    with pytest.raises(F.FormatError, match="not a"):
        F.loads(b"x" * 200)
    bad = a.copy(); bad[0, 0, 1, 0] = qs[1]
    with pytest.raises(F.FormatError, match="canonical"):
        F.dumps(bad, qs, psis)
    empty = F.loads(F.dumps(a[:0], qs, psis, dual=True))
    assert empty.residues.shape == (0, 2, 3, N) and empty.dual and empty.scale is None


@pytest.mark.gpu
def test_file_feeds_the_engine(tmp_path):
    N = 1024
    qs, psis = T.prime_chain(N, [60, 60])
    rng = np.random.default_rng(1)
    a, b = _batch(rng, qs, N, 2, 2), _batch(rng, qs, N, 2, 2)
    F.save(str(tmp_path / "a.tfb"), a, qs, psis)
    F.save(str(tmp_path / "b.tfb"), b, qs, psis)
    fa, fb = F.load(str(tmp_path / "a.tfb")), F.load(str(tmp_path / "b.tfb"))
    ctx = T.Context(fa.N, fa.qs, fa.psis)
    prod = ctx.to_host(ctx.ct_tensor(ctx.to_device(fa.residues), ctx.to_device(fb.residues)))
    from oracle import c_oracle as CO
    assert np.array_equal(prod, CO.Rns(N, qs, psis).ct_tensor(a, b))
    F.save(str(tmp_path / "prod.tfb"), prod, qs, psis)
    assert np.array_equal(F.load(str(tmp_path / "prod.tfb")).residues, prod)
