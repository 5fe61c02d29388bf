"""Runs the default-mode (adaptive RK45 replica) discretization on BASELINE configs[2] a few times (ncu target)."""
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
out = torch.empty((105, N * (K - 1)), dtype=torch.float64, device=dev)
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    M.discretize_batch_device(x, u, tfd, const, out=out, adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2))
    e1.record(); torch.cuda.synchronize()
    print("adaptive ms", e0.elapsed_time(e1))
