"""Batched, device-resident CKKS ciphertext operations for the workload pipelines (BASELINE.json configs 3 and 5).

A ``CtBatch`` is B independent ciphertexts at one level: one int64 tensor ``[B][comps][L][N]`` of primal residues
on the GPU plus the scale.  Every operation is one or a few engine calls over the whole batch (the reference's
``map`` over an array of ciphertexts, examples/encrypted_mnist/infer.jl:120-137); nothing returns to the host between
upload and the final download.  Keys, rings and single-ciphertext encrypt / decrypt come from the scheme mirror
(``toyfhe_b200.scheme``); this module only adds the batch dimension the C-ABI already has.

Operation -> reference:
  mul_scalar        ckksencoding.jl:100-103   c * b::AbstractFloat (scaled integer times every component)
  add / add_plain   rlwe_she.jl:228-245, ckksencoding.jl:113-125
  MatDiagonals      ckksencoding.jl:106-111   a .* c (encode at the ciphertext's scale, multiply components), operands cached
  rescale           ckksencoding.jl:127-130 -> crt.jl:215-228 (modswitch)
  square_relin      rlwe_she.jl:247-266 (c*c) + :315-349 keyswitch(ek, .)
  rotate            rlwe_she.jl:355-359  keyswitch(gk, apply_galois_element(c, g))
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

import toyfhe_b200 as T


class Level:
    """the ciphertext ring after dropping ``drops`` primes, with the key-ring view used by the keyswitch at this level"""

    def __init__(self, params: T.SHEShemeParams, drops: int):
        ring = params.R_cipher()
        for _ in range(drops):
            ring = ring.drop_last()
        self.ring = ring
        self.ctx = ring.ctx
        self.raised = isinstance(params, T.ModulusRaised)
        Rkey = params.R_key()
        if self.raised:
            self.which = tuple(range(ring.L)) + (Rkey.L - 1,)          # modulusraising.jl:43-49
            self.ext = Rkey.crtselect(self.which)
        else:
            self.which = tuple(range(ring.L))                          # crt.jl:238-244
            self.ext = None


class Pipeline:
    """rings per level and cached evaluation-key tensors for one parameter set"""

    def __init__(self, params: T.SHEShemeParams):
        self.params = params
        self.levels: Dict[int, Level] = {}
        self.N = params.R_cipher().N
        # plaintext operands derived from host values (scalars, bias vectors), encoded once per (level, scale, value): a pipeline
        # that runs again (or is replayed from a CUDA graph) touches no host memory
        self.consts: Dict[tuple, torch.Tensor] = {}

    def scalar_plain(self, drops: int, s: int) -> torch.Tensor:
        """[L][N] with every row filled with s mod q_i: the plaintext operand of `c * b` for tfb_mul_plain"""
        key = ("scalar", drops, s)
        if key not in self.consts:
            lvl = self.level(drops)
            self.consts[key] = lvl.ctx.to_device(np.array([[s % q] * lvl.ring.N for q in lvl.ring.qs], dtype=np.uint64))
        return self.consts[key]

    def encoded_plain(self, drops: int, scale: float, slots: np.ndarray) -> torch.Tensor:
        key = ("slots", drops, float(scale), slots.tobytes())
        if key not in self.consts:
            lvl = self.level(drops)
            d = torch.from_numpy(np.ascontiguousarray(slots)).to(f"cuda:{lvl.ctx.device}").reshape(1, -1)
            self.consts[key] = lvl.ctx.ckks_encode(scale, d)[0].contiguous()
        return self.consts[key]

    def level(self, drops: int) -> Level:
        if drops not in self.levels:
            self.levels[drops] = Level(self.params, drops)
        return self.levels[drops]


class CtBatch:
    def __init__(self, pipe: Pipeline, drops: int, ct: torch.Tensor, scale: float):
        self.pipe, self.drops, self.ct, self.scale = pipe, drops, ct, float(scale)

    @property
    def lvl(self) -> Level:
        return self.pipe.level(self.drops)

    @property
    def B(self) -> int:
        return self.ct.shape[0]

    # ---- construction / extraction through the scheme mirror (single ciphertexts)
    @classmethod
    def from_ciphertexts(cls, pipe: Pipeline, cts: Sequence[T.CipherText], drops: int = 0) -> "CtBatch":
        rows = [torch.stack([x.coeffs_primal() for x in c.cs]) for c in cts]
        return cls(pipe, drops, torch.stack(rows).contiguous(), cts[0].plain.scale)

    def ciphertext(self, i: int) -> T.CipherText:
        lvl = self.lvl
        params = self.pipe.params
        for _ in range(self.drops):
            params = T.DropLastParams(params)
        cs = tuple(T.RingElement(lvl.ring, primal=self.ct[i, k].contiguous()) for k in range(self.ct.shape[1]))
        return T.CipherText(params, cs, T.CKKSScale(self.scale))

    def replicate(self, B: int) -> "CtBatch":
        """B copies of the batch's ciphertexts (synthetic batches for throughput runs; the work per copy is identical)"""
        reps = (B + self.B - 1) // self.B
        return CtBatch(self.pipe, self.drops, self.ct.repeat(reps, 1, 1, 1)[:B].contiguous(), self.scale)

    # ---- arithmetic
    def mul_scalar(self, b: float, out: Optional["CtBatch"] = None, accumulate: bool = False) -> "CtBatch":
        """c * b: every component times round(b * scale) (ckksencoding.jl:100-103); result scale = scale^2.
        With ``out`` and ``accumulate`` the product is added to ``out`` (the convolution sums of infer.jl:118-121)."""
        lvl = self.lvl
        s = int(round(b * self.scale))
        if out is None or not accumulate:
            res = lvl.ctx.scalar_mul(self.ct, s, out=None if out is None else out.ct)
            return CtBatch(self.pipe, self.drops, res, self.scale * self.scale)
        lvl.ctx.mul_plain(self.ct, self.pipe.scalar_plain(self.drops, s), out=out.ct, accumulate=True)
        return out

    def add(self, other: "CtBatch") -> "CtBatch":
        assert self.drops == other.drops and abs(self.scale / other.scale - 1) < 1e-9
        return CtBatch(self.pipe, self.drops, self.lvl.ctx.add(self.ct, other.ct), self.scale)

    def add_plain(self, slots: np.ndarray) -> "CtBatch":
        """c .+ b for a scalar or a slot vector (ckksencoding.jl:113-125): encode at the ciphertext's scale, add to c[1]"""
        lvl = self.lvl
        n = lvl.ring.N // 2
        v = np.broadcast_to(np.asarray(slots, dtype=np.complex128), (n,)).copy()
        enc = self.pipe.encoded_plain(self.drops, self.scale, v)          # [L][N] primal, cached
        lvl.ctx.add_plain_first(self.ct, enc)                             # in place: this batch is an intermediate of the pipeline
        return self

    def rescale(self) -> "CtBatch":
        """modswitch: exact division by the last prime, scale divided with it (ckksencoding.jl:127-130, crt.jl:215-228)"""
        lvl = self.lvl
        qlast = lvl.ring.qs[-1]
        return CtBatch(self.pipe, self.drops + 1, lvl.ctx.rescale(self.ct), self.scale / qlast)

    def square_relin(self, ek: T.EvalMultKey) -> "CtBatch":
        """keyswitch(ek, c*c) (infer.jl:135-136, 158-159)"""
        lvl = self.lvl
        t3 = lvl.ctx.ct_tensor(self.ct, self.ct)                          # one operand buffer: transformed once
        key = ek.key.dual_for(lvl.which)
        out = lvl.ctx.keyswitch(key, t3, self.pipe.params.relin_window, ext=None if lvl.ext is None else lvl.ext.ctx)
        return CtBatch(self.pipe, self.drops, out, self.scale * self.scale)

    def rotate(self, gk: T.GaloisKey) -> "CtBatch":
        """ToyFHE.rotate(gk, c) = keyswitch(gk, apply_galois_element(c, g)) (rlwe_she.jl:355-359)"""
        lvl = self.lvl
        g = lvl.ctx.galois(self.ct, gk.galois_element)
        key = gk.key.dual_for(lvl.which)
        out = lvl.ctx.keyswitch(key, g, self.pipe.params.relin_window, ext=None if lvl.ext is None else lvl.ext.ctx)
        return CtBatch(self.pipe, self.drops, out, self.scale)


class MatDiagonals:
    """The plaintext operands of a diagonal-method matmul (test/ckks_matmul.jl:34-42, infer.jl:142-151): for k = 1..n the
    vector repeat(diag(circshift(W, (0, k-1))), inner|outer) encoded at the ciphertext's scale over its ring and kept
    in the dual domain -- encoded once per (weights, level), reused by every ciphertext of every batch."""

    def __init__(self, lvl: Level, scale: float, vectors: Sequence[np.ndarray]):
        n = lvl.ring.N // 2
        data = np.stack([np.asarray(v, dtype=np.complex128).reshape(n) for v in vectors])
        d = torch.from_numpy(data).to(f"cuda:{lvl.ctx.device}")
        enc = lvl.ctx.ckks_encode(scale, d)                               # [n_diag][L][N] primal
        self.dual = lvl.ctx.ntt_fwd(enc)
        self.scale = float(scale)

    def __len__(self):
        return self.dual.shape[0]


def diag_matmul(x: CtBatch, gk: T.GaloisKey, diags: MatDiagonals) -> CtBatch:
    """result = d_1 .* x; for k = 2..n: rotated = rotate(gk, rotated); result += d_k .* rotated
    (test/ckks_matmul.jl:34-42).  The products accumulate in the dual domain; one inverse transform at the end."""
    lvl = x.lvl
    assert abs(diags.scale / x.scale - 1) < 1e-9
    acc = torch.empty_like(x.ct)
    rotated = x
    for k in range(len(diags)):
        if k:
            rotated = rotated.rotate(gk)
        dual = lvl.ctx.ntt_fwd(rotated.ct)
        lvl.ctx.mul_plain(dual, diags.dual[k], out=acc, accumulate=k > 0)
    return CtBatch(x.pipe, x.drops, lvl.ctx.ntt_inv(acc, out=acc), x.scale * diags.scale)


class Graphed:
    """`fn()` -- a fixed sequence of engine calls over static device buffers -- captured ONCE into a CUDA graph and replayed.
    The engine's entry points are plain launches on the stream they are given, so stream capture records them all; the
    eager run that precedes the capture (on the capture stream) sizes the contexts' scratch and fills every operand cache.
    At these sizes a pipeline is thousands of short launches and the eager Python path is CPU-bound (MNIST, batch 64 on
    one B200: 172 -> 595 pipelines/s; results bit-identical, tests/test_gpu_workloads.py)."""

    def __init__(self, fn):
        self.fn, self.graph, self.out, self.eager = fn, None, None, False

    def __call__(self):
        if self.eager:
            return self.fn()
        if self.graph is None:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self.fn()
            side.synchronize()
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=side):
                    self.out = self.fn()
                self.graph = g
            except Exception as e:       # capture refused (driver / allocator state): same results from the eager path, only slower
                import warnings
                warnings.warn(f"CUDA graph capture failed ({e}); running the launch sequence eagerly")
                torch.cuda.synchronize()
                self.eager = True
                return self.fn()
        self.graph.replay()
        return self.out


def decrypt_slots(kp: T.KeyPair, batch: CtBatch, i: int = 0) -> np.ndarray:
    return T.decrypt(kp, batch.ciphertext(i)).data
