"""How much of the 8-rank fused gather is the SM side?  Same kernel, 8 destination buffers, all LOCAL (one GPU)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpconstellation_b200 as M
from bench import make_constellation
N, K = 4096, 200
Y, const = make_constellation(N)
dev = torch.device("cuda:0")
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
tfd = torch.full((N,), 2.0, dtype=torch.float64, device=dev)
x, u, _ = M.propagate_batch_device(torch.from_numpy(Y).to(dev), tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
outs = [torch.empty((105, N * (K - 1)), dtype=torch.float64, device=dev) for _ in range(8)]
def best(fn):
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts[1:])
for nd in (1, 2, 4, 8):
    t = best(lambda: M.discretize_batch_device(x, u, tfd, const, out=outs[0], extra_dst=outs[1:nd] if nd > 1 else None))
    print(f"{nd} local destination(s): {t:.3f} ms  ({nd * 685 / t:.0f} GB/s of stores)")
