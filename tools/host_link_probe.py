"""Host <-> device ceiling under N concurrent ranks (round-1 review item: the end-to-end curve stops scaling at 2 GPUs --
is it the engine or the host?).  Every rank copies 512 MiB of pinned memory H2D, D2H and both at once, all ranks together
(barrier before each phase); rank 0 prints per-rank and aggregate GB/s, the NUMA / affinity facts of the box, and the
ct-mul/s the link would allow (4 MiB in + 3 MiB out per BFV ciphertext pair at N = 2^14, L = 8).

    python tools/host_link_probe.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/host_link_probe.py
Option --pin: each rank restricts itself to a disjoint slice of the host cores (os.sched_setaffinity) BEFORE it allocates and
first-touches its pinned buffers, so that pages land next to the cores that drive the copies."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def pin_rank_to_cores(local: int, world: int):
    cores = sorted(os.sched_getaffinity(0))
    per = max(1, len(cores) // max(1, world))
    mine = cores[local * per:(local + 1) * per] or cores
    os.sched_setaffinity(0, mine)
    return mine


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    pin = "--pin" in sys.argv
    mine = pin_rank_to_cores(local, world) if pin else sorted(os.sched_getaffinity(0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = 512 << 20
    h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
    h_in.fill_(1); h_out.fill_(0)                                    # first touch by this rank's cores
    d_in = torch.empty(n, dtype=torch.uint8, device="cuda")
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    ck = 32 << 20

    def h2d():
        with torch.cuda.stream(s1):
            for o in range(0, n, ck):
                d_in[o:o + ck].copy_(h_in[o:o + ck], non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            for o in range(0, n, ck):
                h_out[o:o + ck].copy_(d_out[o:o + ck], non_blocking=True)

    res = {}
    for name, fns in (("h2d", (h2d,)), ("d2h", (d2h,)), ("both", (h2d, d2h))):
        for _ in range(2):
            for f in fns:
                f()
            torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            for f in fns:
                f()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        t = torch.tensor([dt], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[name] = n / float(t.item()) / 1e9                         # GB/s per direction per rank (slowest rank)
    if rank == 0:
        numa = {}
        try:
            for d in sorted(os.listdir("/sys/devices/system/node")):
                if d.startswith("node"):
                    numa[d] = open(f"/sys/devices/system/node/{d}/cpulist").read().strip()
        except OSError:
            pass
        link_pairs = min(res["both"] * 1e9 / (4 << 20), res["both"] * 1e9 / (3 << 20))
        out = {"n_gpus": world, "pinned_affinity": pin, "cores_rank0": len(mine), "host_cores": os.cpu_count(), "numa_nodes": numa,
               "per_rank_GBps": {k: round(v, 2) for k, v in res.items()},
               "aggregate_GBps": {k: round(v * world, 1) for k, v in res.items()},
               "bfv_pairs_per_s_the_link_allows": round(world * link_pairs),
               "note": "per-rank figure = slowest rank; 'both' = H2D and D2H concurrently, GB/s per direction"}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
