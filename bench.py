#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 engine on BASELINE.json's metric:
homomorphic ciphertext multiplies/s (BFV, N=2^14, L=8 RNS primes, t=65537, R_big =
17 further 60-bit primes: BASELINE configs[1]; the engine multiplies over its own 17-prime joint basis
Q u P', P' = the first 9 primes of R_big -- same integers, same result) plus the forward-NTT rate at the
same (N, L).

  python bench.py --gpus N --steps K --warmup W          # our engine (one rank per GPU)
  python bench.py --impl reference ...                   # CPU restatement of the reference path

One step = one pass of tfb_bfv_mul over a batch of B independent ciphertext pairs
already resident in HBM.  Prints ONE JSON line (rank 0)."""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_RING = 2 ** 14
L_Q = 8
L_BIG = 17
T_PLAIN = 65537
METRIC = "bfv_ciphertext_muls_per_s"
UNIT = "ciphertext-muls/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="ciphertext pairs per GPU per step")
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs[2] / configs[4] workload pipelines")
    ap.add_argument("--mnist-batch", type=int, default=64, help="encrypted-MNIST pipelines per GPU (BASELINE configs[4]; 512 per GPU = 4096 on 8 GPUs)")
    ap.add_argument("--matmul-batch", type=int, default=16, help="CKKS 128x128 matmul ciphertexts per GPU (BASELINE configs[2])")
    return ap.parse_args()


def workload_name(batch):
    return (f"BFV ct*ct (rlwe_she.jl enc_mul + bfv.jl expand/contract), N=2^14, L=8x60-bit RNS primes, t=65537, "
            f"R_big=17x60-bit primes, batch={batch} ciphertext pairs per GPU")


def rings(reference=False):
    """primes / roots of NegacyclicRing(2^14, ntuple(_->60, 25)) (crt.jl:282-295).  The reference arm takes them from the
    oracle's own constructor so that it never loads the engine's library."""
    if reference:
        from oracle import toyfhe_oracle as O
        allq, allpsi = O.prime_chain(N_RING, [60] * (L_Q + L_BIG))
    else:
        import toyfhe_b200 as T
        allq, allpsi = T.prime_chain(N_RING, [60] * (L_Q + L_BIG))
    return allq[:L_Q], allpsi[:L_Q], allq[L_Q:], allpsi[L_Q:]


def config_dict(batch):
    """the same `config` on both arms (the driver compares them)"""
    Nb = N_RING * 8
    return {"workload": workload_name(batch), "sharding": "independent ciphertext pairs per GPU, no data-path collective",
            "l2": f"inputs {2 * batch * 2 * L_Q * Nb / 2**20:.0f} MiB + R_big intermediates per step, far larger than the 126 MB L2 (no flush needed)"}


def rand_ct(rng, qs, shape):
    import numpy as np
    out = np.empty(shape + (len(qs), N_RING), dtype=np.uint64)
    for i, q in enumerate(qs):
        out[..., i, :] = rng.integers(0, q, size=shape + (N_RING,), dtype=np.uint64)
    return out


# ------------------------------------------------------------------ CPU arm
def cpu_bfv_mul_rate(pairs_per_step, steps, warmup):
    """times the oracle's C restatement of the reference path on the host cores"""
    import numpy as np
    from oracle import c_oracle as CO
    # all host threads, whatever OMP_NUM_THREADS the launcher exported (torchrun sets it to 1)
    CO.set_threads(len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))
    qs, psis, qb, psib = rings(reference=True)
    oq, ob = CO.Rns(N_RING, qs, psis), CO.Rns(N_RING, qb, psib)
    rng = np.random.default_rng(0)
    c1, c2 = rand_ct(rng, qs, (pairs_per_step, 2)), rand_ct(rng, qs, (pairs_per_step, 2))
    for _ in range(warmup):
        CO.bfv_mul(oq, ob, T_PLAIN, c1, c2)
    t0 = time.perf_counter()
    for _ in range(steps):
        CO.bfv_mul(oq, ob, T_PLAIN, c1, c2)
    dt = time.perf_counter() - t0
    return pairs_per_step * steps / dt, dt / steps, CO.num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pairs = 2   # bounded sample of the workload per step (the rate is per pair, so it compares with the GPU arm's)
    value, s_per_step, cores = cpu_bfv_mul_rate(pairs, max(1, args.steps), max(0, args.warmup))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": s_per_step * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(args.batch),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{pairs} ciphertext pairs per step x {args.steps} steps (+{args.warmup} warm-up) of the same workload; "
                                   "C restatement of the Julia path (oracle/oracle.c), OpenMP on all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------ clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(self.gpu)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.f.read().splitlines():
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = sorted(sm)[len(sm) // 2:]  # upper half = samples under load
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import toyfhe_b200 as T

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1 and hasattr(os, "sched_setaffinity"):
        # each rank keeps to its own slice of the host cores, set BEFORE its pinned buffers are allocated and first touched
        # (tools/host_link_probe.py: on this pool's single-socket VM hosts it changes nothing, on a multi-socket host it
        # keeps a rank's staging memory next to the cores that drive its copies)
        cores = sorted(os.sched_getaffinity(0))
        per = max(1, len(cores) // world)
        os.sched_setaffinity(0, cores[local * per:(local + 1) * per] or cores)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL prints its banner ("NCCL version ...") on fd 1 when the first
        # communicator is created, so fd 1 points at stderr until that has happened
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
            probe = torch.zeros(1, device=f"cuda:{local}")
            dist.all_reduce(probe)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B, K, W = args.batch, args.steps, max(args.warmup, 3)
    qs, psis, qb, psib = rings()
    cq, cb = T.Context(N_RING, qs, psis, device=local), T.Context(N_RING, qb, psib, device=local)
    rng = np.random.default_rng(1234 + rank)
    # synthetic ciphertexts: uniform residues (what fresh ciphertext components look like, rlwe_she.jl:156-158);
    # a 16-pair host block replicated on the device keeps host RAM small
    blk = min(B, 16)
    h1, h2 = rand_ct(rng, qs, (blk, 2)), rand_ct(rng, qs, (blk, 2))
    reps = (B + blk - 1) // blk
    c1 = cq.to_device(h1).repeat(reps, 1, 1, 1)[:B].contiguous()
    c2 = cq.to_device(h2).repeat(reps, 1, 1, 1)[:B].contiguous()
    # decorrelate the replicas so every pair is distinct work (values stay canonical)
    c1 = torch.roll(c1, shifts=1, dims=3) if reps > 1 else c1
    out = cq.empty((B, 3, L_Q, N_RING))
    stream = torch.cuda.current_stream()

    def step():
        cq.bfv_mul(cb, T_PLAIN, c1, c2, out=out, stream=stream)

    for _ in range(W):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    T.profile_read(reset=True)
    T.profile_enable(True)
    launches0 = T.kernel_launches()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    for _ in range(K):
        step()
    ev1.record(stream)
    barrier()
    ms_total = ev0.elapsed_time(ev1)
    launches = T.kernel_launches() - launches0
    T.profile_enable(False)
    prof = T.profile_read(reset=True)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([ms_total], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    ms_step = ms_total / K
    value = world * B * K / (ms_total * 1e-3)

    # ---- standalone forward-NTT rate at (N=2^14, L=8): the other half of the metric
    ntt_in, ntt_out = c1, torch.empty_like(c1)
    for _ in range(3):
        cq.ntt_fwd(ntt_in, out=ntt_out, stream=stream)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    KN = max(K, 10)
    for _ in range(KN):
        cq.ntt_fwd(ntt_in, out=ntt_out, stream=stream)
    e1.record(stream)
    barrier()
    ntt_ms = e0.elapsed_time(e1) / KN
    tn = torch.tensor([ntt_ms], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(tn, op=dist.ReduceOp.MAX)
    ntt_ms = float(tn.item())
    ntt_polys = 2 * B
    ntt_bytes = ntt_polys * L_Q * N_RING * 8 * 2

    def rate(fn, reps):
        for _ in range(3):
            fn()
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record(stream)
        for _ in range(reps):
            fn()
        a1.record(stream)
        barrier()
        tms = torch.tensor([a0.elapsed_time(a1) / reps], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        return float(tms.item())

    # inverse NTT at the same shape, and both directions at N = 2^15 on the CKKS chain of BASELINE configs[2]
    # (60 + 9 x 40 + special 60 bits: 11 primes, 96 polynomials = 264 MiB per direction, out of place)
    inv_ms = rate(lambda: cq.ntt_inv(ntt_in, out=ntt_out, stream=stream), KN)
    q15, p15 = T.prime_chain(2 ** 15, [60] + [40] * 9 + [60])
    c15 = T.Context(2 ** 15, q15, p15, device=local)
    x15 = c15.sample_uniform(5, 1, (96,))
    y15 = torch.empty_like(x15)
    f15_ms = rate(lambda: c15.ntt_fwd(x15, out=y15, stream=stream), KN)
    i15_ms = rate(lambda: c15.ntt_inv(x15, out=y15, stream=stream), KN)
    bytes15 = 96 * len(q15) * (2 ** 15) * 8 * 2
    del x15, y15

    # ---- end to end through the host-buffer C-ABI call (pinned host memory, copies inside the timed region)
    Be = B   # the same batch as the device-resident step
    p1 = torch.from_numpy(rand_ct(rng, qs, (Be, 2)).view(np.int64)).pin_memory()
    p2 = torch.from_numpy(rand_ct(rng, qs, (Be, 2)).view(np.int64)).pin_memory()
    po = torch.empty((Be, 3, L_Q, N_RING), dtype=torch.int64).pin_memory()
    for _ in range(3):   # warm-up (allocates staging, faults in the pinned pages)
        cq.bfv_mul_host(cb, T_PLAIN, p1, p2, po, stream=stream)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        cq.bfv_mul_host(cb, T_PLAIN, p1, p2, po, stream=stream)   # synchronises before returning
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / args.e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=f"cuda:{local}")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e_value = world * Be / e2e_s

    # ---- BASELINE configs[2] and configs[4] as checked device-resident pipelines (workloads/): every rank runs its own
    # batch of independent pipelines (ciphertext-parallel, no data-path collective); rate = all ranks' pipelines / max time
    configs = None
    if not args.no_configs:
        from workloads import ckks_matmul as W3, mnist as W5
        configs = {}
        r3 = W3.run(args.matmul_batch, 128, 2 ** 15, 9, reps=2)
        r5 = W5.run(args.mnist_batch, 64, 2 ** 13, reps=1)
        tt = torch.tensor([r3["ms_per_batch"], r5["ms_per_batch"], r3["max_abs_err"], r5["max_abs_err"],
                           r3["last_of_batch_max_abs_err"], r5["last_of_batch_max_abs_err"]], dtype=torch.float64, device=f"cuda:{local}")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        m3, m5, e3, e5, l3, l5 = [float(v) for v in tt.tolist()]
        configs["c3_ckks_matmul"] = {
            "workload": "CKKS N=2^15, chain 60+9x40+special 60 (11 primes), ModulusRaised CRT-digit keyswitch, 128x128 diagonal matmul: "
                        "127 rotations + 128 plaintext-vector multiplies + 1 rescale per ciphertext (test/ckks_matmul.jl:30-44 scaled up)",
            "batch_per_gpu": args.matmul_batch, "ms_per_batch": m3, "matmuls_per_s": world * args.matmul_batch / (m3 * 1e-3),
            "rotations_per_s": world * args.matmul_batch * 127 / (m3 * 1e-3), "max_abs_err_vs_float64": max(e3, l3), "atol": 1e-5,
            "correct": max(e3, l3) < 1e-5, "kernel_launches_per_batch": r3["kernel_launches_per_batch"],
            "cuda_graph": "the batch's launch sequence is captured once and replayed (workloads/ckks_batch.py: Graphed)"}
        configs["c5_encrypted_mnist"] = {
            "workload": "encrypted-MNIST CKKS inference (examples/encrypted_mnist/infer.jl:96-177): N=2^13, primes 60+5x40+special 60, per pipeline "
                        "(64 images) 196 ct*scalar, 5 ct*ct + relinearisations, 10 rescales, 315 rotations, 320 plaintext-vector multiplies; "
                        "synthetic weights and images of the model's shapes",
            "batch_per_gpu": args.mnist_batch, "ms_per_batch": m5, "pipelines_per_s": world * args.mnist_batch / (m5 * 1e-3),
            "images_per_s": world * args.mnist_batch * 64 / (m5 * 1e-3), "max_abs_err_vs_float64": max(e5, l5),
            "labels_agree": bool(r5["labels_agree"]), "correct": max(e5, l5) < 1e-3 and bool(r5["labels_agree"]),
            "kernel_launches_per_batch": r5["kernel_launches_per_batch"],
            "cuda_graph": "the pipeline's launch sequence is captured once and replayed (workloads/ckks_batch.py: Graphed)"}

        # BASELINE configs[3]: BFV relinearisation keyswitch, N=2^14, 8 primes, base-4 digits (D = 241 digit polynomials).
        # One GPU: tfb_keyswitch.  N GPUs (N | 8): ONE ciphertext batch key-switched by all ranks together with the RNS
        # primes sharded over the ranks (tfb_keyswitch_shard) and the result rows assembled by one NCCL all-gather -- the
        # path's only data-path collective (toyfhe.jl_b200/sharding.py).
        if world in (1, 2, 4, 8):
            from toyfhe_b200 import sharding as S
            w4 = 2
            D4 = T.ndigits(qs, w4)
            lo4, hi4 = S.shard_range(L_Q, rank, world)
            shard = T.Context(N_RING, qs[lo4:hi4], psis[lo4:hi4], device=local) if world > 1 else cq
            full4 = cq.sample_uniform(99, 1, (D4, 2))                    # a (synthetic) evaluation key in the NTT domain, the same on every rank
            krows = S.key_rows_for_shard(full4, lo4, hi4) if world > 1 else full4   # the rows this rank keeps and streams
            c4, ok4 = {}, True
            xchg, xchg_note = None, None
            if world > 1:
                try:    # result slots of every rank mapped into every other rank (CUDA IPC, NVLink peer memory)
                    xchg = S.open_peer_exchange(cq, L_Q, 8)
                except Exception as e:   # no IPC between the ranks (container policy): the NCCL all-gather path below still runs
                    xchg_note = f"unavailable: {e}"
            for B4, mode4 in ((1, ""), (8, ""), (1, "_push"), (8, "_push")):
                if mode4 and xchg is None:
                    continue
                ct4 = cq.sample_uniform(7, 100 + B4, (B4, 3))           # the same ciphertexts on every rank (same seed and stream id)
                if world > 1 and mode4:   # the epilogue kernel stores this rank's rows into every rank's slot and waits on the peers' flags
                    run4 = lambda: S.keyswitch_residue_sharded_push(cq, shard, lo4, krows, ct4, w4, xchg)
                elif world > 1:           # shard kernels, then one NCCL all-gather of the result rows
                    run4 = lambda: S.keyswitch_residue_sharded(lambda a, b: cq.keyswitch_shard(shard, a, krows, ct4, w4), L_Q)
                else:
                    run4 = lambda: cq.keyswitch(krows, ct4, w4)
                for _ in range(3):
                    run4()
                barrier()
                k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                k0.record()
                for _ in range(10):
                    r4 = run4()
                k1.record()
                barrier()
                tk = torch.tensor([k0.elapsed_time(k1) / 10], dtype=torch.float64, device=f"cuda:{local}")
                if world > 1:
                    dist.all_reduce(tk, op=dist.ReduceOp.MAX)
                ms4 = float(tk.item())
                c4[f"batch{B4}{mode4}"] = {"ms_per_call": ms4, "keyswitches_per_s": B4 / (ms4 * 1e-3)}
                if world > 1:   # every rank checks the gathered result bit for bit against the whole-ring call with the whole key
                    ok4 = ok4 and bool(torch.equal(r4, cq.keyswitch(full4, ct4, w4)))
            del full4
            if xchg is not None:
                ok4 = ok4 and not xchg.timed_out()
                dist.barrier()
                xchg.close()
            if world > 1:
                okt = torch.tensor([1 if ok4 else 0], dtype=torch.int32, device=f"cuda:{local}")
                dist.all_reduce(okt, op=dist.ReduceOp.MIN)
                ok4 = bool(int(okt.item()))
            # (bit-exactness of the sharded path against the whole-ring call: tests/test_gpu_multi.py, tests/test_sharding_gloo.py)
            configs["c4_keyswitch_base4"] = {
                "workload": f"BFV relinearisation keyswitch, N=2^14, L=8x60-bit, relin_window=2 (D={D4} digit polynomials of 8 prime rows), "
                            + ("residues sharded over the ranks; batchB: tfb_keyswitch_shard + one NCCL all-gather of the result rows per call; "
                               "batchB_push: tfb_keyswitch_shard_push, the epilogue kernel stores its rows into every rank's buffer over NVLink "
                               "peer memory and waits on the peers' flags (no collective call)"
                               if world > 1 else "one GPU: tfb_keyswitch"),
                "sharding": "residue-parallel (strong scaling of ONE ciphertext batch)" if world > 1 else "none",
                "data_path_collective": "all_gather_into_tensor of [B][2][L/N][N] u64 per call" if world > 1 else None,
                "peer_push": (xchg_note or "CUDA IPC-mapped result slots, fused into ks_finish_push_kernel") if world > 1 else None,
                "correct": ok4, "checked": "sharded result == tfb_keyswitch over the whole ring with the whole key, on every rank" if world > 1 else "single GPU: the path the tests pin to the oracle",
                **c4}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    Nb = N_RING * 8
    # algorithmic bytes per step for each kernel class of this workload
    alg = {
        "ntt_fwd_row": 4 * B * L_BIG * 2 * Nb,
        "ntt_inv_row": 3 * B * L_BIG * 2 * Nb,
        "tensor_dual": 7 * B * L_BIG * Nb,
        "base_switch": 4 * B * (L_Q + L_BIG) * Nb,
        "bfv_contract": 3 * B * (L_Q + L_BIG) * Nb,
    }
    kernels = {}
    tot_prof_ms = sum(v[1] for v in prof.values()) or 1.0
    for name, (cnt, ms) in prof.items():
        if cnt == 0:
            continue
        kernels[name] = {"launches": cnt, "ms_total": round(ms, 4), "share": round(ms / tot_prof_ms, 4)}
        if name in alg and ms > 0:
            kernels[name]["achieved_gbs"] = round(alg[name] * K / (ms * 1e-3) / 1e9, 1)
            kernels[name]["avg_launch_ms"] = round(ms / cnt, 4)
    dom = max(kernels, key=lambda k: kernels[k]["ms_total"]) if kernels else None
    roofline = None
    if dom and "achieved_gbs" in kernels[dom]:
        a = kernels[dom]["achieved_gbs"]
        traffic, traffic_src = None, None
        try:   # DRAM bytes per launch of this kernel from the committed ncu capture (same shape), scaled to this batch
            tr = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
            if dom in tr:
                traffic = float(tr[dom]) * B / float(tr["batch"])
                traffic_src = "profiles/r02_traffic.json (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)"
        except Exception:
            pass
        roofline = {"kernel": dom, "bound": "hbm", "achieved": a, "peak": peak, "unit": "GB/s", "frac": round(a / peak, 4),
                    "traffic": traffic, "traffic_source": traffic_src,
                    "algorithmic_bytes_per_launch": alg[dom] // max(1, kernels[dom]["launches"] // K),
                    "peak_source": peak_src, "share_of_step": kernels[dom]["share"],
                    # the bound that actually binds: 30.9 FMA-heavy pipe cycles per warp-butterfly in the SASS of this kernel
                    # (IMAD.WIDE/IMAD.HI 4, IMAD 2), 896 warp-butterflies per SM sub-partition per 2^14 row, 148 SMs at 1.965 GHz
                    "pipe_ceiling_gbs": round(262144 * 148 / (896 * 30.9 / 1.965e9) / 1e9, 1),
                    "frac_of_pipe_ceiling": round(a / (262144 * 148 / (896 * 30.9 / 1.965e9) / 1e9), 4),
                    "note": "integer work on 61-bit residues: the FMA-heavy (IMAD) pipe is 66% busy in the transforms and 87% in the "
                            "base conversions at these rates and bounds them before HBM does (DESIGN.md section 5, "
                            "profiles/r02_ncu_bfv_step.txt, profiles/r02_ntt_ablation.txt)"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        v, s_step, cores = cpu_bfv_mul_rate(2, 3, 1)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "2 ciphertext pairs x 3 steps of the same workload; C restatement of the reference's Julia path "
                         "(oracle/oracle.c, OpenMP over primes/coefficients) -- the Julia reference cannot run offline"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": config_dict(B),
        "gpu_launches": int(launches),
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(2 * Be * 2 * L_Q * Nb),
                "d2h_bytes_per_step": int(Be * 3 * L_Q * Nb), "batch": Be, "ms_per_step": e2e_s * 1e3,
                "api": "tfb_bfv_mul_host (pinned host buffers; H2D, kernels and D2H pipelined over 8-pair chunks on three streams)",
                "bound": "host link: 48 GB/s per direction with both directions busy on one GPU, 39.6 GB/s per rank on two (aggregate H2D+D2H "
                         "stops growing with the number of ranks on this pool's single-socket VM hosts: tools/host_link_probe.py, "
                         "profiles/r02_host_link.txt); 4 MiB in + 3 MiB out per pair => ~11.5k pairs/s for one GPU's link"},
        "roofline": roofline,
        "kernels": kernels,
        "ntt_fwd": {"value": world * ntt_polys / (ntt_ms * 1e-3), "unit": "RNS-NTT/s (N=2^14, L=8)",
                    "prime_rows_per_s": world * ntt_polys * L_Q / (ntt_ms * 1e-3), "ms_per_step": ntt_ms,
                    "achieved_gbs": ntt_bytes / (ntt_ms * 1e-3) / 1e9, "frac_of_peak": ntt_bytes / (ntt_ms * 1e-3) / 1e9 / peak,
                    "algorithmic_bytes_per_launch": ntt_bytes},
        "ntt_inv": {"value": world * ntt_polys / (inv_ms * 1e-3), "unit": "RNS-NTT/s (N=2^14, L=8)", "ms_per_step": inv_ms,
                    "achieved_gbs": ntt_bytes / (inv_ms * 1e-3) / 1e9, "frac_of_peak": ntt_bytes / (inv_ms * 1e-3) / 1e9 / peak},
        "ntt_2p15": {"shape": "N=2^15, 11 primes (60 + 9x40 + 60 bits), 96 polynomials, out of place",
                     "fwd_gbs": bytes15 / (f15_ms * 1e-3) / 1e9, "inv_gbs": bytes15 / (i15_ms * 1e-3) / 1e9,
                     "fwd_frac_of_peak": bytes15 / (f15_ms * 1e-3) / 1e9 / peak, "inv_frac_of_peak": bytes15 / (i15_ms * 1e-3) / 1e9 / peak,
                     "note": "per rank; forward applies the row's last global level while loading (one HBM pass less)"},
        "cpu_baseline": cpu,
        "configs": configs,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
