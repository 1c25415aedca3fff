"""BASELINE.json configs[4]: the encrypted-MNIST CKKS inference pipeline of examples/encrypted_mnist/infer.jl:96-177 as
a checked, device-resident workload.

Shape (infer.jl): N = 2^13, ring (q0 60-bit, 5 x 40-bit, special 60-bit), ModulusRaised(CKKSParams(R, 0, 3.2)) -- CRT-digit
keyswitch with a special prime -- scale 2^40; one pipeline classifies 64 images:

    49 input ciphertexts C_Iij (7x7 window offsets; slot k + 64 l = image k, window l)                 infer.jl:107-115
    conv      4 channels: sum_ij C_Iij * w[i,j,ch] (+ bias), rescale                                    :117-121
    square    x*x, keyswitch(ek, .), rescale                                                            :126-128
    fq1       4 diagonal-method 64x64 matmuls (63 rotations + 64 plaintext-vector multiplies each),
              summed, + bias, rescale                                                                   :132-156
    square    x*x, keyswitch(ek, .), rescale                                                            :158-160
    fq2       one 64x64 (zero-padded 10x64) diagonal matmul + bias                                      :162-170
    decrypt

per pipeline: 196 ct*scalar, 5 ct*ct + relinearisations, 10 ciphertext rescales, 315 rotations, 320 plaintext-vector
multiplies.  The trained model (mnist_conv.bson) and the MNIST images are not available offline: weights and images are
seeded synthetic data of the same shapes (``data: synthetic``); correctness = the decrypted result against the same
network evaluated in float64 on the same inputs.

    python -m workloads.mnist [--batch B] [--m 64] [--check]      # one JSON line
"""
from __future__ import annotations

import argparse
import json
import math
import os
import sys
import time
from typing import Dict, List

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

import toyfhe_b200 as T  # noqa: E402
from workloads.ckks_batch import CtBatch, Graphed, MatDiagonals, Pipeline, decrypt_slots, diag_matmul  # noqa: E402

CHANNELS = 4
KS = 7          # 7x7 convolution window
SCALE = float(2 ** 40)


def mnist_ring(N: int) -> T.NegacyclicRing:
    """infer.jl:96-110: q0 = nextprime(2^60+1), ps the next one, q1..q5 from nextprime(2^40+1), all = 1 mod 2N;
    ring order (q0, q1..q5, ps)"""
    qs, psis = T.prime_chain(N, [60, 40, 40, 40, 40, 40, 60])
    return T.NegacyclicRing(N, qs=qs, psis=psis)


def make_model(seed: int, m: int) -> Dict[str, np.ndarray]:
    """synthetic weights with the shapes of the trained network (Conv 7x7x1x4, Dense 4m -> m, Dense m -> 10)"""
    rng = np.random.default_rng(seed)
    return {
        "conv_w": rng.normal(0, 1 / KS, size=(KS, KS, CHANNELS)), "conv_b": rng.normal(0, 0.1, size=CHANNELS),
        "fq1_w": rng.normal(0, 1 / math.sqrt(4 * m), size=(m, CHANNELS * m)), "fq1_b": rng.normal(0, 0.1, size=m),
        "fq2_w": rng.normal(0, 1 / math.sqrt(m), size=(10, m)), "fq2_b": rng.normal(0, 0.1, size=10),
    }


def make_inputs(seed: int, m: int, n_img: int) -> np.ndarray:
    """I[i][j] = matrix [image k][window l] of pixel values in [0,1] (public_preprocess, infer.jl:57-64)"""
    rng = np.random.default_rng(seed)
    return rng.random(size=(KS, KS, n_img, m))


def plain_forward(model, I: np.ndarray) -> np.ndarray:
    """do_encrypted_inference on plaintext (infer.jl:66-90), float64: returns [10][n_img]"""
    KSa, _, n_img, m = I.shape
    conved = [sum(I[i, j] * model["conv_w"][i, j, ch] for i in range(KS) for j in range(KS)) + model["conv_b"][ch]
              for ch in range(CHANNELS)]
    sq1 = [(x ** 2).T for x in conved]                                  # [window l][image k]
    fq1 = sum(model["fq1_w"][:, ch * m:(ch + 1) * m] @ sq1[ch] for ch in range(CHANNELS)) + model["fq1_b"][:, None]
    sq2 = fq1 ** 2
    return model["fq2_w"] @ sq2 + model["fq2_b"][:, None]


def diag_vectors(W: np.ndarray, n_img: int) -> List[np.ndarray]:
    """repeat(diag(circshift(W, (0, k-1))), inner = n_img) for k = 1..m (infer.jl:142-151)"""
    m = W.shape[0]
    out = []
    for k in range(m):
        d = np.array([W[l, (l - k) % m] for l in range(m)])
        out.append(np.repeat(d, n_img))
    return out


class MnistPipeline:
    def __init__(self, N: int = 2 ** 13, m: int = 64, seed: int = 0, sampler_seed: int = 1):
        self.N, self.m = N, m
        self.n_img = (N // 2) // m
        assert self.n_img * m == N // 2
        self.R = mnist_ring(N)
        self.params = T.ModulusRaised(T.CKKSParams(self.R, 0, 3.2))
        self.s = T.Sampler(sampler_seed, device=True)
        self.kp = T.keygen(self.s, self.params)
        self.ek = T.keygen_evalmult(self.s, self.kp.priv)
        self.gk = T.keygen_galois(self.s, self.kp.priv, steps=self.n_img)
        self.pipe = Pipeline(self.params)
        self.model = make_model(seed, m)
        self._diags = None
        self._conv_w = None

    def encrypt_inputs(self, I: np.ndarray) -> List[CtBatch]:
        """C_Iij = encrypt(kp, CKKSEncoding(vec(I_ij))) (infer.jl:112-116): 49 batches of one ciphertext"""
        out = []
        for i in range(KS):
            for j in range(KS):
                slots = I[i, j].T.reshape(-1)                           # slot k + n_img * l (column-major vec of [k][l])
                c = T.encrypt(self.s, self.kp, T.CKKSEncoding(SCALE, slots.astype(np.complex128)))
                out.append(CtBatch.from_ciphertexts(self.pipe, [c]))
        return out

    def stack_inputs(self, C: List[CtBatch], batch: int) -> List[CtBatch]:
        """the 49 input batches replicated to `batch` pipelines as slices of ONE tensor [49][batch][2][L][N]"""
        stack = torch.stack([c.replicate(batch).ct for c in C]).contiguous()
        out = []
        for i, c in enumerate(C):
            b = CtBatch(self.pipe, c.drops, stack[i], c.scale)
            b.stack = stack
            out.append(b)
        return out

    def diagonals(self, scale1: float, scale2: float):
        """plaintext operands of the five matmuls, encoded once (levels 2 and 4 drops; the scales the ciphertexts have there)"""
        if self._diags is None:
            m, W1, W2 = self.m, self.model["fq1_w"], self.model["fq2_w"]
            W2p = np.vstack([W2, np.zeros((m - W2.shape[0], m))])        # naive_rectangular_matmul, infer.jl:162-166
            d1 = [MatDiagonals(self.pipe.level(2), scale1, diag_vectors(W1[:, ch * m:(ch + 1) * m], self.n_img)) for ch in range(CHANNELS)]
            d2 = MatDiagonals(self.pipe.level(4), scale2, diag_vectors(W2p, self.n_img))
            self._diags = (d1, d2)
        return self._diags

    def forward(self, C: List[CtBatch]) -> CtBatch:
        mdl, n_img = self.model, self.n_img
        # convolution: 4 channels x 49 ct*scalar summed (+ bias, rescale).  When the 49 input batches are slices of one
        # stacked tensor (encrypt_inputs_stacked) all 196 products are ONE launch that reads every input once (tfb_lincomb)
        conved = []
        base = getattr(C[0], "stack", None)
        if base is not None and all(getattr(c, "stack", None) is base for c in C):
            lvl = C[0].lvl
            if self._conv_w is None:
                s_int = [[int(round(float(mdl["conv_w"][i, j, ch]) * C[0].scale)) for i in range(KS) for j in range(KS)] for ch in range(CHANNELS)]
                self._conv_w = lvl.ctx.to_device(np.array([[[s % q for q in lvl.ring.qs] for s in row] for row in s_int], dtype=np.uint64))
            outs = lvl.ctx.lincomb(base, self._conv_w)                                   # [4][B][2][L][N]
            for ch in range(CHANNELS):
                acc = CtBatch(self.pipe, C[0].drops, outs[ch], C[0].scale * C[0].scale)
                conved.append(acc.add_plain(mdl["conv_b"][ch]).rescale())
        else:
            for ch in range(CHANNELS):
                acc = None
                for i in range(KS):
                    for j in range(KS):
                        w = float(mdl["conv_w"][i, j, ch])
                        if acc is None:
                            acc = C[i * KS + j].mul_scalar(w)
                        else:
                            C[i * KS + j].mul_scalar(w, out=acc, accumulate=True)
                conved.append(acc.add_plain(mdl["conv_b"][ch]).rescale())
        sq1 = [x.square_relin(self.ek).rescale() for x in conved]
        # the matmul operands are encoded at the scales the ciphertexts have at their levels: known once these exist
        if self._diags is None:
            s1 = sq1[0].scale
            # level/scale of fq2's input: fq1 (s1^2) rescaled, squared, rescaled
            q = self.pipe.level(2).ring.qs
            s_fq1 = s1 * s1 / q[-1]
            s2 = s_fq1 * s_fq1 / q[-2]
            self.diagonals(s1, s2)
        d1, d2 = self._diags
        fq1 = None
        for ch in range(CHANNELS):
            y = diag_matmul(sq1[ch], self.gk, d1[ch])
            fq1 = y if fq1 is None else fq1.add(y)
        fq1 = fq1.add_plain(np.repeat(mdl["fq1_b"], n_img)).rescale()
        sq2 = fq1.square_relin(self.ek).rescale()
        res = diag_matmul(sq2, self.gk, d2)
        bias = np.repeat(np.concatenate([mdl["fq2_b"], np.zeros(self.m - 10)]), n_img)
        return res.add_plain(bias)

    def forward_graphed(self, C: List[CtBatch]) -> CtBatch:
        """The whole pipeline (~2900 kernel launches) captured ONCE into a CUDA graph over the static input batches `C` and
        replayed: the engine's calls are plain launches on the stream they are given, so the capture sees them all; scratch and
        plaintext operands are sized / encoded by the eager run that precedes the capture.  New inputs go into C[i].ct
        (copy_) before replay(); the result batch is static as well."""
        key = tuple(int(c.ct.data_ptr()) for c in C)
        if getattr(self, "_graph_key", None) != key:
            self._graph, self._graph_key = Graphed(lambda: self.forward(C)), key
        return self._graph()

    def decode(self, res: CtBatch, i: int = 0) -> np.ndarray:
        """decrypt_matrix(kp, x)[1:10, :] (infer.jl:153, 167): [10][n_img]"""
        slots = np.real(decrypt_slots(self.kp, res, i))
        return slots.reshape(self.m, self.n_img)[:10]


def run(batch: int, m: int, N: int, check: bool = True, reps: int = 1, seed: int = 0, graph: bool = True) -> dict:
    torch.cuda.synchronize()
    P = MnistPipeline(N=N, m=m, seed=seed)
    I = make_inputs(seed + 100, m, P.n_img)
    want = plain_forward(P.model, I)
    C1 = P.encrypt_inputs(I)
    out = {"N": N, "m": m, "images_per_pipeline": P.n_img, "batch": batch}
    if check:
        got = P.decode(P.forward(C1))
        err = float(np.max(np.abs(got - want)))
        out.update({"max_abs_err": err, "max_abs_value": float(np.max(np.abs(want))),
                    "labels_agree": bool(np.array_equal(np.argmax(got, axis=0), np.argmax(want, axis=0)))})
    if batch > 0:
        C = P.stack_inputs(C1, batch)
        launches0 = T.kernel_launches()
        P.forward(C)                                                  # warm-up (allocations, operand encodes)
        torch.cuda.synchronize()
        launches = T.kernel_launches() - launches0
        fwd = P.forward_graphed if graph else P.forward
        fwd(C)                                                        # graph: eager run on the capture stream + capture + first replay
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            res = fwd(C)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        last = P.decode(res, batch - 1)
        out.update({"ms_per_batch": ms, "pipelines_per_s": batch / (ms * 1e-3), "images_per_s": batch * P.n_img / (ms * 1e-3),
                    "cuda_graph": bool(graph), "kernel_launches_per_batch": launches,
                    "last_of_batch_max_abs_err": float(np.max(np.abs(last - want)))})
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--m", type=int, default=64)
    ap.add_argument("--logn", type=int, default=13)
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    a = ap.parse_args()
    t0 = time.time()
    r = run(a.batch, a.m, 1 << a.logn, check=not a.no_check, reps=a.reps, graph=not a.no_graph)
    r["wall_s"] = time.time() - t0
    print(json.dumps({"workload": "encrypted_mnist (examples/encrypted_mnist/infer.jl:96-177)", **r}))


if __name__ == "__main__":
    main()
