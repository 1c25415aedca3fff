"""Golden fixtures under tests/golden/:
  reference_kats.json  -- every literal known-answer value the reference holds for the hot path, extracted from the
                          reference's own files by tests/golden/extract_reference_kats.py (re-checked here whenever
                          /root/reference is present; the GPU box only has the committed JSON);
  oracle_vectors.npz   -- seeded oracle outputs (tests/golden/make_oracle_vectors.py), a regression pin.
CPU half: the oracle reproduces both.  GPU half (marked gpu): the CUDA path reproduces both through the C-ABI without
executing anything under oracle/."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(GOLD, "reference_kats.json"), encoding="utf-8") as f:
        return json.load(f)


@pytest.fixture(scope="module")
def vec():
    return dict(np.load(os.path.join(GOLD, "oracle_vectors.npz")))


# ------------------------------------------------------------------ CPU: provenance and the oracle
def test_committed_kats_are_what_the_reference_files_say(kats, tmp_path):
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference tree not present (GPU box)")
    out = subprocess.run([sys.executable, os.path.join(GOLD, "extract_reference_kats.py")], capture_output=True, text=True, check=True).stdout
    assert json.loads(out) == kats


def test_oracle_reproduces_reference_kats(kats):
    from oracle import toyfhe_oracle as O
    for q, N, psi in kats["palisade_q_N_psi"]["value"]:
        assert pow(psi, N, q) == q - 1
    assert O.minimal_primitive_root(97, 8) == kats["minimal_root_q97_N4"]["value"]
    k = kats["rlwe_products_q97"]
    p = k["operands"]
    assert O.ring_multiply(p["p3"], p["p4"], 97, 33) == k["p3*p4"]
    assert O.ring_multiply(p["p1"], p["p1"], 97, 33) == k["p1^2"]
    assert O.ring_multiply(p["p1"], p["p2"], 97, 33) == k["p1*p2"]
    k = kats["naive_product_q7_N2"]
    assert O.ring_multiply_naive(k["a"], k["b"], 7) == k["value"]
    k = kats["slot_product_q65537_N2048"]
    q, N = 65537, 2048
    psi = O.minimal_primitive_root(q, 2 * N)
    a = O.inntt(k["a_slots_0_9"] + [0] * (N - 10), q, psi)
    b = O.inntt([k["b_slots_all"]] * N, q, psi)
    assert O.nntt(O.ring_multiply(a, b, q, psi), q, psi)[:11] == k["value_slots_0_10"]
    k = kats["crt_expand"]
    assert O.crt_encode(3, (5, 7)) == k["x_residues"]
    assert O.crt_expand(k["x_residues"], (5, 7), 11) == k["product_residues"]
    k = kats["crt_residual"]
    assert O.crt_reconstruct(O.crt_residual(3, 0, tuple(k["basis_of_the_arithmetic"])), tuple(k["basis_of_the_arithmetic"])) == k["value"] == k["formula_value"]


def test_oracle_reproduces_its_golden_vectors(vec):
    from oracle import c_oracle as CO
    for tag in ("ntt64", "ntt4096"):
        N = vec[tag + "_in"].shape[-1]
        orc = CO.Rns(N, [int(x) for x in vec[tag + "_q"]], [int(x) for x in vec[tag + "_psi"]])
        assert np.array_equal(orc.nntt(vec[tag + "_in"]), vec[tag + "_fwd"])
        assert np.array_equal(orc.inntt(vec[tag + "_in"]), vec[tag + "_inv"])
    for tag in ("bfv_a", "bfv_b"):
        N, L, Lb, t = (int(x) for x in vec[tag + "_meta"])
        q, psi = [int(x) for x in vec[tag + "_q"]], [int(x) for x in vec[tag + "_psi"]]
        oq, ob = CO.Rns(N, q[:L], psi[:L]), CO.Rns(N, q[L:], psi[L:])
        assert np.array_equal(oq.ct_tensor(vec[tag + "_c1"], vec[tag + "_c2"]), vec[tag + "_tensor"])
        assert np.array_equal(CO.bfv_mul(oq, ob, t, vec[tag + "_c1"], vec[tag + "_c2"]), vec[tag + "_mul"])


# ------------------------------------------------------------------ GPU: the CUDA path against the same fixtures
def _ctx(T, q, psi, N):
    return T.Context(N, [int(x) for x in q], [int(x) for x in psi])


@pytest.mark.gpu
def test_engine_reproduces_reference_kats(kats):
    import toyfhe_b200 as T
    H = T.Context.to_host
    assert T.minimal_primitive_root(97, 8) == kats["minimal_root_q97_N4"]["value"]
    ctx = T.Context(4, [97], [33])
    k = kats["rlwe_products_q97"]
    d = {n: ctx.to_device(np.array([v], dtype=np.uint64)) for n, v in k["operands"].items()}
    assert H(ctx.ring_mul(d["p3"], d["p4"])).tolist() == [k["p3*p4"]]
    assert H(ctx.ring_mul(d["p1"], d["p1"])).tolist() == [k["p1^2"]]
    assert H(ctx.ring_mul(d["p1"], d["p2"])).tolist() == [k["p1*p2"]]
    for q, N, psi in kats["palisade_q_N_psi"]["value"]:       # psi^N = -1: x^N + 1 factors; the transform inverts itself
        c = T.Context(N, [q], [psi])
        a = np.random.default_rng(N).integers(0, q, size=(1, N), dtype=np.uint64)
        assert np.array_equal(H(c.ntt_inv(c.ntt_fwd(c.to_device(a)))), a)
        one_x = np.zeros((1, N), dtype=np.uint64); one_x[0, 1] = 1                  # x
        xN1 = np.zeros((1, N), dtype=np.uint64); xN1[0, N - 1] = 1                  # x^(N-1)
        prod = H(c.ring_mul(c.to_device(one_x), c.to_device(xN1)))                  # x * x^(N-1) = x^N = -1
        assert int(prod[0, 0]) == q - 1 and not prod[0, 1:].any()
    k = kats["slot_product_q65537_N2048"]
    q, N = 65537, 2048
    c = T.Context(N, [q], [T.minimal_primitive_root(q, 2 * N)])
    a = np.zeros((1, N), dtype=np.uint64); a[0, :10] = k["a_slots_0_9"]
    b = np.full((1, N), k["b_slots_all"], dtype=np.uint64)
    prod = c.ntt_fwd(c.ring_mul(c.ntt_inv(c.to_device(a)), c.ntt_inv(c.to_device(b))))
    assert H(prod)[0, :11].tolist() == k["value_slots_0_10"]
    # (the CRTEncoded{(5,7)} / CRTResidual docstring KATs of src/crt.jl are scalar examples over primes that admit no
    #  negacyclic ring -- 7 != 1 mod 4 -- so they pin the oracle's formulas only; tfb_crt_expand / tfb_rescale are
    #  checked against the oracle, which reproduces them, in tests/test_gpu_parity.py)


@pytest.mark.gpu
def test_engine_reproduces_oracle_golden_vectors(vec):
    import math
    import toyfhe_b200 as T
    H = T.Context.to_host
    for tag in ("ntt64", "ntt4096"):
        c = _ctx(T, vec[tag + "_q"], vec[tag + "_psi"], vec[tag + "_in"].shape[-1])
        d = c.to_device(vec[tag + "_in"])
        assert np.array_equal(H(c.ntt_fwd(d)), vec[tag + "_fwd"])
        assert np.array_equal(H(c.ntt_inv(d)), vec[tag + "_inv"])
    for tag in ("bfv_a", "bfv_b"):
        N, L, Lb, t = (int(x) for x in vec[tag + "_meta"])
        cq, cb = _ctx(T, vec[tag + "_q"][:L], vec[tag + "_psi"][:L], N), _ctx(T, vec[tag + "_q"][L:], vec[tag + "_psi"][L:], N)
        d1, d2 = cq.to_device(vec[tag + "_c1"]), cq.to_device(vec[tag + "_c2"])
        assert np.array_equal(H(cq.ct_tensor(d1, d2)), vec[tag + "_tensor"])
        assert np.array_equal(H(cq.bfv_mul(cb, t, d1, d2)), vec[tag + "_mul"])
    c = _ctx(T, vec["ks_q"], vec["ks_psi"], 64)
    key = c.ntt_fwd(c.to_device(vec["ks_key"]))
    assert np.array_equal(H(c.keyswitch(key, c.to_device(vec["ks_ct"]), int(vec["ks_w"][0]))), vec["ks_out"])
    assert np.array_equal(H(c.rescale(c.to_device(vec["rs_in"]))), vec["rs_out"])
    for g, want in zip(vec["gal_g"], vec["gal_out"]):
        assert np.array_equal(H(c.galois(c.to_device(vec["rs_in"][:1]), int(g))), want)
    Q, t = math.prod(int(x) for x in vec["ks_q"]), int(vec["pt_t"][0])
    assert np.array_equal(H(c.bfv_encode(t, Q // t, c.to_device(vec["pt_m"]))), vec["pt_enc"])
    assert np.array_equal(H(c.bfv_decode(t, Q // t, c.to_device(vec["pt_b"]))), vec["pt_dec"])
