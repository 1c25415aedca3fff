"""CPU-only: the sampler's CPU restatement (oracle/sampler_oracle.py) against the published Philox4x32-10
known-answer vectors (Random123 kat_vectors: zero, all-ones and pi-digit counters/keys), plus the distribution
properties the schemes rely on (ranges, centred Gaussian of the requested width)."""
import numpy as np

from oracle import sampler_oracle as SO


def test_philox4x32_10_known_answers():
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = SO.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(v[0]) for v in got) == want


def test_sampler_oracle_distributions():
    qs = [1152921504606904321, 1099511922689, 97]
    u = SO.sample_uniform(5, 1, 4, qs, 512)
    assert u.shape == (4, 3, 512)
    for i, q in enumerate(qs):
        assert int(u[:, i, :].max()) < q
    assert np.array_equal(u, SO.sample_uniform(5, 1, 4, qs, 512))
    assert not np.array_equal(u, SO.sample_uniform(5, 2, 4, qs, 512))
    x = SO.gaussian_ints(3.2, 9, 3, 64, 1024)
    assert abs(x.mean()) < 0.05 and abs(x.std() - np.sqrt(3.2 ** 2 + 1 / 12)) < 0.05
    g = SO.sample_gaussian(3.2, 9, 3, 2, qs, 1024)
    for i, q in enumerate(qs):
        assert np.array_equal(g[:, i, :].astype(np.int64), np.mod(x[:2], q))
