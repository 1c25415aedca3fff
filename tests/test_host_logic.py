

def test_small_divisor_reciprocal_is_exact():
    """SmallDiv (csrc/rns_kernels.cu): q = (r * ceil(2^32 / d)) >> 32 must equal r // d for every row index r < 2^23 and every
    divisor the kernels use (numbers of primes up to TFB_MAX_L + 1, digit counts up to 512) -- the elementwise kernels
    split their row index with it instead of a 64-bit division."""
    import numpy as np
    r = np.arange(1 << 23, dtype=np.uint64)
    for d in list(range(2, 67)) + [97, 121, 127, 128, 241, 255, 256, 257, 481, 511, 512]:
        M = np.uint64(((1 << 32) + d - 1) // d)
        assert M < (1 << 32)
        q = (r * M) >> np.uint64(32)
        assert np.array_equal(q, r // np.uint64(d)), d


def test_digit_from_limbs_formula_matches_base_2w_digits():
    """pass1_pow2 (csrc/ntt_core3.cuh) cuts base-2^w digit k out of the 64-bit limbs of X: bits k w .. k w + w - 1, taking the
    upper part from the next limb when the digit straddles two limbs and nothing when there is no next limb.  Same digits as
    rlwe_she.jl:331-337 computes by repeated division."""
    import random
    rnd = random.Random(5)
    for nl in (1, 2, 8):
        for w in (1, 2, 3, 5, 7, 31, 63):
            X = rnd.getrandbits(64 * nl - 3)
            limbs = [(X >> (64 * i)) & (2**64 - 1) for i in range(nl)]
            mask = (1 << w) - 1
            D = (64 * nl + w - 1) // w
            for k in range(D):
                bit = k * w
                limb, off = bit >> 6, bit & 63
                v = limbs[limb] >> off
                if off + w > 64 and limb + 1 < nl:
                    v |= (limbs[limb + 1] << (64 - off)) & (2**64 - 1)
                assert v & mask == (X >> bit) & mask, (nl, w, k)
