"""Launches, once each after a warm-up, the small-batch kernels whose ncu summaries are kept under profiles/:
propagate_rk45_kernel on BASELINE config 2 (64 satellites, K=100) and on config 3 (4096), discretize_group_kernel on
configs 1 and 2.   ncu --set full -k regex:<kernel> --launch-skip <n> -c 1 python scripts/r02_ncu_targets.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

dev = torch.device("cuda:0")
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
for rep in range(2):           # launch order per repetition: prop(64), group(6336 forced), prop(1), group(49), prop(4096)
    for N, K, tf in ((64, 100, 1.0), (1, 50, 0.5), (4096, 200, 2.0)):
        Y, const = make_constellation(N)
        y0 = torch.from_numpy(Y).to(dev)
        tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
        x, u, _ = M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
        if N <= 64:
            _lib.lib().mpc_set_tuning(25)
            M.discretize_batch_device(x, u, tfd, const, n_sub=100)
            _lib.lib().mpc_set_tuning(24)
torch.cuda.synchronize()
