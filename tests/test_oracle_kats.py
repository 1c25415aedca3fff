"""Pin the CPU oracle against every literal known-answer value the reference
holds for the hot path.  CPU only."""
import math
import random

from oracle import toyfhe_oracle as O


# (1) src/cryptparams.jl:22-25 -- PALISADE (q, N, psi) triples
PALISADE = [
    (1099511627873, 8, 108163207722),
    (525313, 512, 513496),
    (34359724033, 1024, 7225104974),
    (1152921504606830593, 2048, 811032584449645127),
]


def test_cryptparams_triples_are_2n_th_roots():
    for q, N, psi in PALISADE:
        assert O.is_prime(q)
        assert pow(psi, 2 * N, q) == 1
        assert pow(psi, N, q) == q - 1


def test_cryptparams_roundtrip_and_product():
    rnd = random.Random(1)
    for q, N, psi in PALISADE[:3]:
        a = [rnd.randrange(q) for _ in range(N)]
        b = [rnd.randrange(q) for _ in range(N)]
        assert O.inntt(O.nntt(a, q, psi), q, psi) == a
        if N <= 512:
            assert O.ring_multiply(a, b, q, psi) == O.ring_multiply_naive(a, b, q)


# (2) docs/src/man/background/rlwe.md:183-187
def test_minimal_root_doc_kat():
    assert O.minimal_primitive_root(97, 8) == 33


# (3) docs/src/man/background/rlwe.md:193-212
def test_rlwe_doc_products():
    q, psi = 97, 33
    p1, p2, p3, p4 = [1, 1, 0, 0], [0, 0, 0, 1], [4, 0, 0, 0], [5, 0, 0, 0]
    assert O.ring_multiply(p3, p4, q, psi) == [20, 0, 0, 0]
    assert O.ring_multiply(p1, p1, q, psi) == [1, 2, 1, 0]
    assert O.ring_multiply(p1, p2, q, psi) == [96, 0, 0, 1]


def test_nntt_matches_definition_and_selfcheck_vectors():
    q, psi = 97, 33
    # SURVEY Appendix A derived vectors (from the definition)
    assert O.nntt([1, 1, 0, 0], q, psi) == [34, 48, 65, 51]
    assert O.nntt([0, 0, 0, 1], q, psi) == [47, 33, 50, 64]
    assert O.nntt([4, 0, 0, 0], q, psi) == [4, 4, 4, 4]
    rnd = random.Random(5)
    for (qq, N, ps) in [(97, 4, 33), PALISADE[0], (65537, 64, O.minimal_primitive_root(65537, 128))]:
        a = [rnd.randrange(qq) for _ in range(N)]
        assert O.nntt(a, qq, ps) == O.nntt_def(a, qq, ps)


# (4) docs/src/man/encoding.md:14-24 : GF(7), N=2, naive path 3*4 = [5,0]
def test_encoding_doc_naive_product():
    assert O.ring_multiply_naive([3, 0], [4, 0], 7) == [5, 0]


# (5) docs/src/man/encoding.md:69-92 : N=2048, p=65537 slot product
def test_slot_product_doc_kat():
    q, N = 65537, 2048
    psi = O.minimal_primitive_root(q, 2 * N)
    # SlotEncoding: slots are the dual coefficients (encoding.jl:35-56)
    a_slots = list(range(1, 11)) + [0] * (N - 10)
    b_slots = [10] * N
    a = O.inntt(a_slots, q, psi)
    b = O.inntt(b_slots, q, psi)
    prod = O.ring_multiply(a, b, q, psi)
    assert O.nntt(prod, q, psi)[:11] == [10, 20, 30, 40, 50, 60, 70, 80, 90, 100, 0]


# (6) src/crt.jl:23-33
def test_crt_expand_doc_kat():
    x = O.crt_encode(3, (5, 7))
    assert x == [3, 3]
    y = O.crt_expand(x, (5, 7), 11)
    assert y == [3, 5, 0]
    # the (non-doctest) docstring prints 333, which contradicts its own residues
    # (333 mod 7 == 4, not 5); the value of 3 * 11 with residues (3,5,0) is 33.
    assert O.crt_reconstruct(y, (5, 7, 11)) == 33
    assert [333 % 5, 333 % 7, 333 % 11] != y


# (7) src/crt.jl:50-58 (docstring writes F3(3) but the arithmetic shown is modulus 5)
def test_crt_residual_doc_kat():
    r = O.crt_residual(3, 1, (3, 5, 7))
    assert O.crt_reconstruct(r, (3, 5, 7)) == 3 * pow(21, -1, 5) % 5 * 21
    # the numeric example in the docstring: mod(3*invmod(77,5),5)*77 == 308 for basis (5,7,11)
    assert O.crt_reconstruct(O.crt_residual(3, 0, (5, 7, 11)), (5, 7, 11)) == 308


def test_appendix_a_vectors():
    # Galois g=3 on [1,2,3,4] over q=97
    assert O.apply_galois_element([1, 2, 3, 4], 3, 97) == [1, 4, 94, 2]
    # rescale of X=123 over (5,7,11) dropping 11
    poly = [[123 % 5], [123 % 7], [123 % 11]]
    assert O.modswitch(poly, (5, 7, 11)) == [[11 % 5], [11 % 7]]
    assert O.rha(5, 2) == 3 and O.rha(-5, 2) == -3 and O.rha(7, 2) == 4


def test_prime_chain_matches_survey(q8, psi8):
    qs, psis = O.prime_chain(2 ** 14, (60,) * 8)
    assert qs == q8
    assert psis == psi8
    q12, psi12 = O.prime_chain(2 ** 12, (60,))
    assert q12 == [1152921504606904321] and psi12 == [190237715829865]


def test_bfv_crt_test_primes():
    # test/bfv_crt.jl:8-10 : 2 primes ~2^50 at n=2048
    n = 2048
    p1 = O.nextprime(2 ** 50 + 1, 2 * n)
    p2 = O.nextprime(p1 + 2 * n, 2 * n)
    assert (p1, p2) == (1125899906949121, 1125899906977793)


def test_modswitch_is_exact_floor_division():
    rnd = random.Random(7)
    qs = [1099511627873, 97, 65537]
    Q = math.prod(qs)
    for _ in range(50):
        X = rnd.randrange(Q)
        out = O.modswitch([[X % q] for q in qs], qs)
        want = (X - X % qs[-1]) // qs[-1]
        assert [r[0] for r in out] == [want % q for q in qs[:-1]]
