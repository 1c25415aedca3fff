"""Host-side mirror of the reference's scheme layer for the RNS / power-of-two
path: src/rlwe_she.jl (keys, CipherText, keygen / encrypt / decrypt / + / * /
keyswitch / rotate), src/bfv.jl (pi, pi^-1, mul_expand / mul_contract),
src/ckks.jl + src/ckksencoding.jl (CKKS params, FixedRational scale, complex-FFT
encoding, ciphertext modswitch) and src/modulusraising.jl (special-prime
keyswitching).  Same names, argument meaning and error behaviour as the Julia
code; all polynomial arithmetic goes through the CUDA engine (ring.py)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import engine as E
from .ring import NegacyclicRing, RingElement


class UsageError(Exception):
    """rlwe_she.jl:218-225"""


# ----------------------------------------------------------------------------
# sampling (poly.jl:7-23, crt.jl:146-148, 277-279).  The reference draws from an
# unseeded global RNG; here the sampler is explicit and seedable.
# ----------------------------------------------------------------------------
class Sampler:
    """device=True draws on the GPU (tfb_sample_uniform / tfb_sample_gaussian, counter-based Philox keyed by the seed;
    every call takes the next stream id), so keygen and encrypt never touch host memory; device=False is the numpy
    PCG64 sampler the bit-exact scheme tests share with the oracle."""

    def __init__(self, seed: int, device: bool = False):
        self.rng = np.random.Generator(np.random.PCG64(seed))
        self.seed, self.device, self._stream = int(seed), bool(device), 0

    def _next_stream(self) -> int:
        self._stream += 1
        return self._stream

    def uniform(self, ring: NegacyclicRing) -> RingElement:
        """RingSampler(R, DiscreteUniform): every residue drawn independently (crt.jl:146-148)"""
        if self.device:
            return RingElement(ring, primal=ring.ctx.sample_uniform(self.seed, self._next_stream())[0])
        a = np.empty((ring.L, ring.N), dtype=np.uint64)
        for i, q in enumerate(ring.qs):
            a[i] = self.rng.integers(0, q, size=ring.N, dtype=np.uint64)
        return ring.from_residues(a)

    def gaussian_ints(self, N: int, sigma: float) -> List[int]:
        return [int(x) for x in np.rint(self.rng.normal(0.0, sigma, size=N))]

    def gaussian(self, ring: NegacyclicRing, sigma: float) -> RingElement:
        """RingSampler(R, DiscreteNormal(0, sigma)) (bfv.jl:31-32, ckks.jl:24-25)"""
        if self.device:
            return RingElement(ring, primal=ring.ctx.sample_gaussian(sigma, self.seed, self._next_stream())[0])
        return ring(self.gaussian_ints(ring.N, sigma))

    def zero(self, ring: NegacyclicRing) -> RingElement:
        return ring.zero()


# ----------------------------------------------------------------------------
# scheme parameters (rlwe_she.jl:9-65)
# ----------------------------------------------------------------------------
class SHEShemeParams:
    """Parameter objects compare STRUCTURALLY: the reference's checks are ``c1.params !== c2.params`` on immutable
    structs (rlwe_she.jl:223-225, 233, 248), i.e. egality by value -- two ``DropLastParams`` built by separate
    ``modswitch`` calls around the same parameters are the same parameters (examples/encrypted_mnist/infer.jl:137-160
    adds and multiplies such ciphertexts)."""
    relin_window: int = 0

    def _key(self):
        return (type(self),) + tuple(sorted((k, v) for k, v in self.__dict__.items()))

    def __eq__(self, other):
        return isinstance(other, SHEShemeParams) and type(self) is type(other) and self.__dict__ == other.__dict__

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash((type(self).__name__, self.relin_window))

    def R_cipher(self) -> NegacyclicRing: raise NotImplementedError
    def R_key(self) -> NegacyclicRing: return self.R_cipher()
    def R_plain(self): raise NotImplementedError
    def pi(self, b: RingElement): raise NotImplementedError
    def pi_inv(self, plaintext): raise NotImplementedError
    def noise(self, s: Sampler) -> RingElement: raise NotImplementedError     # N(params)
    def secret(self, s: Sampler) -> RingElement: raise NotImplementedError    # G(params)
    # optional hooks that change multiplication (rlwe_she.jl:38-40)
    def mul_expand(self, c: "CipherText"): return c.cs
    def mul_contract(self, cs): return cs


def default_relin_window(ring: NegacyclicRing) -> int:
    """crt.jl:297-298: CRT-encoded rings relinearise over the CRT basis by default"""
    return 0


class CKKSParams(SHEShemeParams):
    """src/ckks.jl:7-25"""

    def __init__(self, R: NegacyclicRing, relin_window: Optional[int] = None, sigma: float = 8 / math.sqrt(2 * math.pi)):
        self.R = R
        self.relin_window = default_relin_window(R) if relin_window is None else int(relin_window)
        self.sigma = float(sigma)

    def R_cipher(self): return self.R
    def R_plain(self): return self.R
    def pi_inv(self, plaintext): return plaintext if isinstance(plaintext, RingElement) else self.R(plaintext)
    def pi(self, b): return b
    def noise(self, s): return s.gaussian(self.R, self.sigma)
    def secret(self, s): return s.gaussian(self.R, self.sigma)


class BFVParams(SHEShemeParams):
    """src/bfv.jl:5-40 for the RNS route of test/bfv_crt.jl: ciphertext ring R, big ring
    R_big (disjoint basis), plaintext modulus t, Delta = floor(Q/t)."""

    def __init__(self, R: NegacyclicRing, Rbig: NegacyclicRing, t: int, relin_window: int = 1, sigma: float = 3.2,
                 Delta: Optional[int] = None):
        self.R, self.Rbig, self.t = R, Rbig, int(t)
        self.relin_window, self.sigma = int(relin_window), float(sigma)
        self.Delta = R.modulus() // self.t if Delta is None else int(Delta)

    def R_cipher(self): return self.R
    def R_plain(self): return self.t

    def _check_range(self):
        # engine limits of tfb_bfv_encode / tfb_bfv_decode: word-size t, Q / Delta < 2^40 (every reference test satisfies them)
        if not (0 < self.t < (1 << 63) and (self.Delta << 40) >= self.R.modulus()):
            raise UsageError("BFV plaintext maps: t must fit a machine word and Q / Delta must stay below 2^40")

    def pi_inv(self, plaintext: Sequence[int]) -> RingElement:
        """Delta * plaintext (bfv.jl:21-24) -- on the device (tfb_bfv_encode)"""
        self._check_range()
        ctx = self.R.ctx
        m = ctx.to_device(np.array([int(v) % self.t for v in plaintext], dtype=np.uint64).reshape(1, self.R.N))
        return RingElement(self.R, primal=ctx.bfv_encode(self.t, self.Delta, m)[0])

    def pi(self, b: RingElement) -> List[int]:
        """mod(divround(SignedMod(x), Delta), t) (bfv.jl:26-29; rounding div_hacks.jl:120-135) -- on the device
        (tfb_bfv_decode)"""
        self._check_range()
        ctx = self.R.ctx
        return [int(v) for v in ctx.to_host(ctx.bfv_decode(self.t, self.Delta, b.coeffs_primal()))]

    def noise(self, s): return s.gaussian(self.R, self.sigma)
    def secret(self, s): return s.gaussian(self.R, self.sigma)

    def mul_expand(self, c: "CipherText"):
        """map(c -> switch(R_big, c), c.cs) (bfv.jl:34, 202-226)"""
        return tuple(RingElement(self.Rbig, primal=self.R.ctx.bfv_switch(self.Rbig.ctx, x.coeffs_primal())) for x in c.cs)

    def mul_contract(self, cs):
        """switch(R, multround(e, t, Q)) (bfv.jl:35-40, 172-190)"""
        return tuple(RingElement(self.R, primal=self.R.ctx.bfv_contract(self.Rbig.ctx, self.t, x.coeffs_primal())) for x in cs)


class BGVParams(SHEShemeParams):
    """src/bgv.jl:5-34: plaintext embedded as is, noise scaled by the plaintext modulus t (ShiftedDiscreteNormal),
    pi = mod(SignedMod(x), t).  Uses the same ring product as the other schemes; both plaintext maps run on the device."""

    def __init__(self, R: NegacyclicRing, t: int, sigma: float):
        self.R, self.t, self.sigma = R, int(t), float(sigma)

    def R_cipher(self): return self.R
    def R_plain(self): return self.t

    def pi_inv(self, plaintext: Sequence[int]) -> RingElement:
        return self.R([int(m) % self.t for m in plaintext])

    def pi(self, b: RingElement) -> List[int]:
        ctx = self.R.ctx
        return [int(v) for v in ctx.to_host(ctx.centered_mod(self.t, b.coeffs_primal()))]

    def noise(self, s): return s.gaussian(self.R, self.sigma) * self.t      # t * DiscreteNormal (bgv.jl:27-33)
    def secret(self, s): return s.gaussian(self.R, self.sigma)


class ModulusRaised(SHEShemeParams):
    """src/modulusraising.jl: the last prime of the key ring is a special prime reserved
    for keys; ciphertexts live on the ring with it dropped."""

    def __init__(self, params: SHEShemeParams):
        self.params = params
        self.relin_window = params.relin_window

    def R_cipher(self): return self.params.R_cipher().drop_last()          # modulusraising.jl:18
    def R_key(self): return self.params.R_key()
    def R_plain(self): return self.params.R_plain().drop_last()            # :19
    def pi_inv(self, plaintext):
        p = self.params.pi_inv(plaintext)
        return p if p.ring == self.R_cipher() else p.modswitch_drop()
    def pi(self, b): return self.params.pi(b)
    def noise(self, s): return self.params.noise(s)
    def secret(self, s): return self.params.secret(s)
    def special_prime(self) -> int: return self.params.R_key().qs[-1]


def parent_params(p: SHEShemeParams) -> SHEShemeParams:
    return p.params if isinstance(p, ModulusRaised) else p


# ----------------------------------------------------------------------------
# keys and ciphertexts (rlwe_she.jl:67-149)
# ----------------------------------------------------------------------------
@dataclass
class KeyComponent:
    mask: RingElement
    masked: RingElement


@dataclass
class PrivKey:
    params: SHEShemeParams
    secret: RingElement


@dataclass
class PubKey:
    params: SHEShemeParams
    key: KeyComponent


@dataclass
class KeyPair:
    priv: PrivKey
    pub: PubKey


class KeySwitchKey:
    """Vector{KeyComponent} (rlwe_she.jl:91-94) plus a cache of the stacked NTT-domain key
    tensors per ciphertext level (the reference caches the duals inside each RingElement)."""

    def __init__(self, params: SHEShemeParams, key: List[KeyComponent]):
        self.params, self.key = params, key
        self._dual = {}

    def dual_for(self, which: Tuple[int, ...]):
        """[D][2][len(which)][N] NTT-domain key over the selected residues
        (downswitch_keyelement, crt.jl:238-244 / modulusraising.jl:43-49)"""
        import torch
        if which not in self._dual:
            idx = torch.as_tensor(list(which), device=self.key[0].mask.coeffs_dual().device)
            rows = [torch.stack([k.mask.coeffs_dual().index_select(0, idx), k.masked.coeffs_dual().index_select(0, idx)])
                    for k in self.key]
            self._dual[which] = torch.stack(rows).contiguous()
        return self._dual[which]


@dataclass
class EvalMultKey:
    key: KeySwitchKey


@dataclass
class GaloisKey:
    galois_element: int
    key: KeySwitchKey


class CipherText:
    """rlwe_she.jl:124-149; ``plain`` carries the encoding tag (e.g. a CKKS scale)"""

    def __init__(self, params: SHEShemeParams, cs: Sequence[RingElement], plain=None):
        self.params, self.cs, self.plain = params, tuple(cs), plain

    def __len__(self): return len(self.cs)
    def __getitem__(self, i): return self.cs[i]

    def ring(self) -> NegacyclicRing: return self.cs[0].ring

    def _addsub(self, other, sub: bool):
        if isinstance(other, CipherText):
            if other.params != self.params:
                raise UsageError("Attempting to add ciphertexts with differing parameters")
            n = max(len(self), len(other))
            out = []
            for i in range(n):
                if i >= len(self):
                    out.append(other[i])       # rlwe_she.jl:236: c2[i] is taken as is for + AND - (reference quirk)
                elif i >= len(other):
                    out.append(self[i])
                else:
                    out.append(self[i] - other[i] if sub else self[i] + other[i])
            return CipherText(self.params, out, self.plain)
        if isinstance(other, RingElement):                         # rlwe_she.jl:243-245
            first = self.cs[0] - other if sub else self.cs[0] + other
            return CipherText(self.params, (first,) + self.cs[1:], self.plain)
        return NotImplemented

    def __add__(self, other): return self._addsub(other, False)
    def __sub__(self, other): return self._addsub(other, True)

    def __mul__(self, other):
        if isinstance(other, CipherText):
            plain = None
            if isinstance(self.plain, CKKSScale) and isinstance(other.plain, CKKSScale):
                plain = CKKSScale(self.plain.scale * other.plain.scale)          # ckksencoding.jl:133-135
            return CipherText(self.params, enc_mul(self, other), plain)
        if isinstance(other, RingElement):                                       # plaintext polynomial multiply
            return CipherText(self.params, tuple(c * other for c in self.cs), self.plain)
        return NotImplemented


# ----------------------------------------------------------------------------
# key generation, encryption, decryption (rlwe_she.jl:151-216)
# ----------------------------------------------------------------------------
class SlotEncoding:
    """encoding.jl:28-56: a BFV plaintext given by its SLOTS, the values of the plaintext polynomial at
    psi_t^(2k+1), k = 0..N-1 -- the dual (NTT-domain) coefficients of the plaintext ring element over F_t
    (t prime, 2N | t-1, encoding.jl:35-43).  Slot-wise products of plaintexts are ring products, so a ciphertext
    multiply acts on all N slots at once (test/bfv_simd.jl).  Both directions run on the device (one inverse /
    forward NTT over the plaintext ring's own context)."""

    def __init__(self, plain_ring: NegacyclicRing, slots: Optional[Sequence[int]] = None):
        if plain_ring.L != 1:
            raise UsageError("SlotEncoding needs a single-prime plaintext ring (encoding.jl:38-41)")
        self.ring = plain_ring
        self.t = plain_ring.qs[0]
        self.slots = [0] * plain_ring.N if slots is None else [int(v) % self.t for v in slots]
        assert len(self.slots) == plain_ring.N

    def __getitem__(self, i): return self.slots[i]
    def __setitem__(self, i, v):
        if isinstance(i, slice):
            idx = range(*i.indices(len(self.slots)))
            vals = [v] * len(idx) if isinstance(v, (int, np.integer)) else list(v)
            for k, x in zip(idx, vals):
                self.slots[k] = int(x) % self.t
        else:
            self.slots[i] = int(v) % self.t
    def __len__(self): return len(self.slots)

    def coeffs(self) -> List[int]:
        """convert(RingElement, s): primal coefficients of the element whose dual is the slot vector (encoding.jl:49-56)"""
        return self.ring.from_residues(np.array([self.slots], dtype=np.uint64), dual=True).to_ints()

    @classmethod
    def from_coeffs(cls, plain_ring: NegacyclicRing, coeffs: Sequence[int]) -> "SlotEncoding":
        """SlotEncoding(r): the dual coefficients of plaintext r (encoding.jl:35-43)"""
        el = plain_ring([int(c) for c in coeffs])
        return cls(plain_ring, [int(v) for v in E.Context.to_host(el.coeffs_dual()).reshape(-1)])


def keygen(s: Sampler, params: SHEShemeParams) -> KeyPair:
    """rlwe_she.jl:155-167"""
    R = params.R_key()
    p0 = parent_params(params)
    mask = s.uniform(R)
    secret = p0.secret(s)
    error = p0.noise(s)
    masked = -(mask * secret + error)
    return KeyPair(PrivKey(params, secret), PubKey(params, KeyComponent(mask, masked)))


def encrypt_zero(s: Sampler, pk: PubKey) -> CipherText:
    """rlwe_she.jl:176-186; ModulusRaised drops the special prime (modulusraising.jl:21-24)"""
    p0 = parent_params(pk.params)
    u = p0.secret(s)
    e1, e2 = p0.noise(s), p0.noise(s)
    c1 = pk.key.masked * u + e1
    c2 = pk.key.mask * u + e2
    if isinstance(pk.params, ModulusRaised):
        c1, c2 = c1.modswitch_drop(), c2.modswitch_drop()
    return CipherText(pk.params, (c1, c2))


def encrypt(s: Sampler, key, plaintext) -> CipherText:
    """rlwe_she.jl:188-197"""
    pk = key.pub if isinstance(key, KeyPair) else key
    c = encrypt_zero(s, pk)
    tag = None
    if isinstance(plaintext, CKKSEncoding):
        tag = CKKSScale(plaintext.scale)
        plaintext = plaintext.to_ring_element(pk.params.R_cipher())
    if isinstance(plaintext, SlotEncoding):
        plaintext = plaintext.coeffs()
    m = pk.params.pi_inv(plaintext)
    return CipherText(pk.params, (c.cs[0] + m,) + c.cs[1:], tag)


def decrypt(key, c: CipherText):
    """rlwe_she.jl:199-216"""
    priv = key.priv if isinstance(key, KeyPair) else key
    secret = priv.secret
    while secret.ring != c[0].ring:
        secret = secret.modswitch_drop()
    b = c[0]
    spow = secret
    for i in range(1, len(c)):
        b = b + spow * c[i]
        spow = spow * secret
    dec = priv.params.pi(b)
    if isinstance(c.plain, CKKSScale):
        return CKKSEncoding.from_ring_element(dec, c.plain.scale)
    return dec


# ----------------------------------------------------------------------------
# homomorphic multiplication (rlwe_she.jl:247-266)
# ----------------------------------------------------------------------------
def enc_mul(c1: CipherText, c2: CipherText):
    if c1.params != c2.params:
        raise UsageError("Attempting to multiply ciphertexts with differing parameters")
    params = c1.params
    # the hooks are looked up on `params` itself: PassthroughParams such as ModulusRaised do
    # not forward mul_expand / mul_contract (rlwe_she.jl:38-40 vs :50-60)
    p0 = params
    R = c1.ring()
    if len(c1) == 2 and len(c2) == 2:
        # the fused engine paths: one call per ciphertext pair
        import torch
        a = torch.stack([x.coeffs_primal() for x in c1.cs]).contiguous()
        b = torch.stack([x.coeffs_primal() for x in c2.cs]).contiguous()
        if isinstance(p0, BFVParams):
            out = R.ctx.bfv_mul(p0.Rbig.ctx, p0.t, a, b)
        else:
            out = R.ctx.ct_tensor(a, b)
        return tuple(RingElement(R, primal=out[i].contiguous()) for i in range(3))
    # general component tensor c[i+j-1] += c1[i]*c2[j] through the hooks
    e1, e2 = p0.mul_expand(c1), p0.mul_expand(c2)
    n = len(e1) + len(e2) - 1
    acc: List[Optional[RingElement]] = [None] * n
    for i in range(len(e1)):
        for j in range(len(e2)):
            prod = e1[i] * e2[j]
            acc[i + j] = prod if acc[i + j] is None else acc[i + j] + prod
    return tuple(p0.mul_contract(acc))


# ----------------------------------------------------------------------------
# key switching (rlwe_she.jl:268-359, modulusraising.jl:26-49)
# ----------------------------------------------------------------------------
def make_eval_key(s: Sampler, old: RingElement, new: PrivKey) -> KeySwitchKey:
    """rlwe_she.jl:273-298; ModulusRaised multiplies `old` by the special prime first
    (modulusraising.jl:28-32)"""
    params = new.params
    p0 = parent_params(params)
    R = old.ring
    if isinstance(params, ModulusRaised):
        old = old * params.special_prime()
    w = params.relin_window
    if w != 0:
        nwindows = E.ndigits(R.qs, w)
        evala = [old * pow(2, i * w) for i in range(nwindows)]
    else:
        # CRT basis decomposition: residue i kept, zeros elsewhere (CRTResidual, crt.jl:60-77)
        res = old.coeffs_primal()
        evala = []
        for i in range(R.L):
            t = res.new_zeros(res.shape)
            t[i] = res[i]
            evala.append(RingElement(R, primal=t))
    key = []
    for a in evala:
        mask = s.uniform(R)
        e = p0.noise(s)
        masked = a - (mask * new.secret + e)
        key.append(KeyComponent(mask, masked))
    return KeySwitchKey(params, key)


def keygen_evalmult(s: Sampler, priv: PrivKey) -> EvalMultKey:
    """keygen(EvalMultKey, priv) (rlwe_she.jl:299)"""
    return EvalMultKey(make_eval_key(s, priv.secret ** 2, priv))


def galois_element_from_steps(steps: int, N: int) -> int:
    """rlwe_she.jl:304"""
    return pow(3, 2 * N - steps, 2 * N) if steps > 0 else pow(3, -steps, 2 * N)


def keygen_galois(s: Sampler, priv: PrivKey, galois_element: Optional[int] = None, steps: Optional[int] = None) -> GaloisKey:
    """keygen(GaloisKey, priv; galois_element | steps) (rlwe_she.jl:300-309)"""
    assert (galois_element is None) != (steps is None)
    if galois_element is None:
        galois_element = galois_element_from_steps(steps, priv.secret.ring.N)
    return GaloisKey(galois_element, make_eval_key(s, priv.secret.apply_galois_element(galois_element), priv))


def keyswitch(ek, c: CipherText) -> CipherText:
    """keyswitch(ek, c) (rlwe_she.jl:315-349) -- one fused engine call"""
    import torch
    ksk = ek.key if isinstance(ek, (EvalMultKey, GaloisKey)) else ek
    if len(c) not in (2, 3):
        raise UsageError("keyswitch expects a ciphertext of length 2 or 3")
    R = c.ring()
    ct = torch.stack([x.coeffs_primal() for x in c.cs]).contiguous()
    w = ksk.params.relin_window
    if isinstance(ksk.params, ModulusRaised):
        Rkey = ksk.key[0].mask.ring
        which = tuple(range(R.L)) + (Rkey.L - 1,)                      # modulusraising.jl:43-49
        ext = Rkey.crtselect(which)
        out = R.ctx.keyswitch(ksk.dual_for(which), ct, w, ext=ext.ctx)
    else:
        which = tuple(range(R.L))                                      # crt.jl:238-244
        out = R.ctx.keyswitch(ksk.dual_for(which), ct, w)
    return CipherText(c.params, (RingElement(R, primal=out[0].contiguous()), RingElement(R, primal=out[1].contiguous())), c.plain)


def apply_galois_element(c: CipherText, g: int) -> CipherText:
    """rlwe_she.jl:355-357"""
    return CipherText(c.params, tuple(x.apply_galois_element(g) for x in c.cs), c.plain)


def rotate(gk: GaloisKey, c: CipherText) -> CipherText:
    """rlwe_she.jl:359"""
    return keyswitch(gk, apply_galois_element(c, gk.galois_element))


# ----------------------------------------------------------------------------
# modulus switching of ciphertexts (crt.jl:161-183, 234-236; ckksencoding.jl:127-130)
# ----------------------------------------------------------------------------
class DropLastParams(SHEShemeParams):
    """crt.jl:161-183"""

    def __init__(self, params: SHEShemeParams):
        self.params = params
        self.relin_window = params.relin_window

    def R_cipher(self): return self.params.R_cipher().drop_last()
    def R_plain(self): return self.params.R_plain()
    def pi_inv(self, plaintext): return self.params.pi_inv(plaintext).modswitch_drop()
    def pi(self, b): return b
    def noise(self, s): return self.params.noise(s).modswitch_drop()
    def secret(self, s): return self.params.secret(s).modswitch_drop()


def modswitch_drop(c: CipherText) -> CipherText:
    """crt.jl:234-236"""
    return CipherText(DropLastParams(c.params), tuple(x.modswitch_drop() for x in c.cs), c.plain)


def modswitch(c: CipherText) -> CipherText:
    """CKKS rescale: divide every component by the last prime and the scale with it
    (ckksencoding.jl:127-130 -> crt.jl:215-228)"""
    if not isinstance(c.plain, CKKSScale):
        raise UsageError("modswitch(::CipherText) is only defined for CKKS ciphertexts")   # rlwe_she.jl:365-367
    qlast = c.ring().qs[-1]
    return CipherText(DropLastParams(c.params), tuple(x.modswitch() for x in c.cs), CKKSScale(c.plain.scale / qlast))


# ----------------------------------------------------------------------------
# CKKS encoding (ckks.jl:27-60 FixedRational; ckksencoding.jl:1-125)
# ----------------------------------------------------------------------------
@dataclass
class CKKSScale:
    scale: float


class CKKSEncoding:
    """N/2 complex slots at scale `scale` (ckksencoding.jl:3-9)"""

    def __init__(self, scale: float, data: np.ndarray):
        self.scale = scale
        self.data = np.asarray(data, dtype=np.complex128)

    @classmethod
    def zeros(cls, scale: float, N: int) -> "CKKSEncoding":
        return cls(scale, np.zeros(N // 2, dtype=np.complex128))

    @classmethod
    def from_ring_element(cls, plain: RingElement, scale: float) -> "CKKSEncoding":
        """decode (ckksencoding.jl:60-70) on the device (tfb_ckks_decode)"""
        ctx = plain.ring.ctx
        return cls(scale, ctx.ckks_decode(scale, plain.coeffs_primal()).cpu().numpy().reshape(-1))

    def to_ring_element(self, ring: NegacyclicRing) -> RingElement:
        """encode (ckksencoding.jl:76-101) on the device (tfb_ckks_encode; |scale * coefficient| < 2^126)"""
        n = len(self.data)
        N = 2 * n
        assert ring.N == N
        import torch
        d = torch.from_numpy(np.ascontiguousarray(self.data)).to(f"cuda:{ring.ctx.device}")
        return RingElement(ring, primal=ring.ctx.ckks_encode(self.scale, d.reshape(1, n))[0])


def ckks_mul_plain_vector(a: np.ndarray, c: CipherText) -> CipherText:
    """a .* c for a real vector a (ckksencoding.jl:106-111): encode at the ciphertext's scale,
    multiply every component, scale squares"""
    assert isinstance(c.plain, CKKSScale)
    enc = CKKSEncoding(c.plain.scale, np.asarray(a, dtype=np.complex128))
    re = enc.to_ring_element(c.ring())
    return CipherText(c.params, tuple(x * re for x in c.cs), CKKSScale(c.plain.scale ** 2))
