

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
