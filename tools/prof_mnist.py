"""Per-kernel-class device time of the encrypted-MNIST pipeline (eager run, CUDA events per launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import toyfhe_b200 as T
from workloads import mnist
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
P = mnist.MnistPipeline()
C = [c.replicate(B) for c in P.encrypt_inputs(mnist.make_inputs(1, 64, P.n_img))]
P.forward(C); torch.cuda.synchronize()
T.profile_read(reset=True); T.profile_enable(True)
P.forward(C); torch.cuda.synchronize()
T.profile_enable(False)
prof = T.profile_read(reset=True)
tot = sum(ms for _, ms in prof.values())
print(f"mnist pipeline batch {B}: {tot:.1f} ms of kernels; " + ", ".join(f"{k} {ms:.1f} ms ({100 * ms / tot:.0f}%, {cnt} launches)" for k, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1]) if cnt))
