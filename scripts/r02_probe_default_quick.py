"""Quick probe: the default-mode kernel (K1b) and the fixed-step kernel on BASELINE configs[2] (CUDA events, L2 flushed)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from bench import make_constellation

dev = torch.device("cuda:0")
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return round(float(np.mean(ts)), 4), round(float(np.min(ts)), 4)


N, K, tf = 4096, 200, 2.0
Y, const = make_constellation(N)
y0 = torch.from_numpy(Y).to(dev)
tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
x, u, st = M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
ad = dict(rtol=1e-3, atol=1e-6, max_step=1e-2)
out = torch.empty((105, N * (K - 1)), dtype=torch.float64, device=dev)
res = {"default_ms": timed(lambda: M.discretize_batch_device(x, u, tfd, const, adaptive=ad, out=out)),
       "default_j2_ms": timed(lambda: M.discretize_batch_device(x, u, tfd, const, include_J2=True, adaptive=ad, out=out)),
       "uniform_ms": timed(lambda: M.discretize_batch_device(x, u, tfd, const, n_sub=100, out=out))}
o, s_ = M.discretize_batch_device(x, u, tfd, const, adaptive=ad, out=out)
torch.cuda.synchronize()
res["checksum_default"] = float(o.double().abs().sum().item())
print(json.dumps(res))
