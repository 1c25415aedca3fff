import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


# Primes / roots the reference constructor NegacyclicRing(2^14, ntuple(_->60, 8))
# produces (crt.jl:282-295); re-derived by tests/test_oracle_kats.py.
Q8 = [1152921504607338497, 1152921504608747521, 1152921504609239041, 1152921504612646913,
      1152921504614023169, 1152921504614055937, 1152921504615628801, 1152921504615694337]
PSI8 = [109957280778515, 54778786028160, 14777115887834, 26847342347732,
        72317606385239, 58189445532409, 19679363742966, 150132853260056]


@pytest.fixture(scope="session")
def q8():
    return list(Q8)


@pytest.fixture(scope="session")
def psi8():
    return list(PSI8)
