"""toyfhe_b200 -- B200 (sm_100a) negacyclic-NTT / RNS engine behind ToyFHE.jl's
power-of-two cyclotomic ring path (src/pow2_cyc_rings.jl, src/crt.jl and the
ciphertext `*` / keyswitch / modswitch bodies of src/rlwe_she.jl).

Layout: ``csrc/`` holds the CUDA kernels and the C-ABI (include/toyfhe_b200.h);
``engine.py`` is the ctypes binding (the Python stand-in for the Julia ccall
shim); ``ring.py`` / ``scheme.py`` mirror the reference's host-side interface
(NegacyclicRing, RingElement, CipherText, keygen/encrypt/decrypt, ...)."""
from .engine import (force_generic, ntt_version, ntt_force_harvey, ntt_max_mode, ntt_cross, ABI_SYMBOLS, LIB_PATH, Context, PeerExchange, EngineError, kernel_launches, load_library,
                     minimal_primitive_root, ndigits, prime_chain, profile_enable, profile_read)
from ._build import build_library

__all__ = ["force_generic", "ntt_version", "ntt_force_harvey", "ntt_max_mode", "ntt_cross", "ABI_SYMBOLS", "LIB_PATH", "Context", "PeerExchange", "EngineError", "kernel_launches", "load_library",
           "minimal_primitive_root", "ndigits", "prime_chain", "profile_enable", "profile_read", "build_library"]

from .ring import NegacyclicRing, RingElement, nntt, inntt
from .scheme import (BFVParams, BGVParams, CKKSEncoding, CKKSParams, CKKSScale, CipherText, DropLastParams, EvalMultKey, GaloisKey,
                     KeyComponent, KeyPair, KeySwitchKey, ModulusRaised, PrivKey, PubKey, Sampler, SlotEncoding, UsageError,
                     apply_galois_element, ckks_mul_plain_vector, decrypt, enc_mul, encrypt, encrypt_zero,
                     galois_element_from_steps, keygen, keygen_evalmult, keygen_galois, keyswitch, make_eval_key,
                     modswitch, modswitch_drop, rotate)
from . import sharding
