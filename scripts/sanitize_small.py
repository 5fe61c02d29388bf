"""Small end-to-end pass over every kernel configuration, meant to be run under compute-sanitizer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import mpconstellation_b200 as M
from bench import make_constellation
N, K = 5, 9
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
for drag, j2 in ((False, False), (True, True)):
    y, u, t, st = M.propagate_batch(Y, 0.3, ctrl, const, include_drag=drag, include_J2=j2, T=K, n_sub=4)
tab = 0.1 * np.ones((N, 3, 6))
M.propagate_batch(Y, 0.3, M.ControllerSpec(M._lib.CTRL_SEQUENCE, (0, 0, 0), tab, np.linspace(0.5, 2, N)), const, T=K, n_sub=3)
for j2 in (False, True):
    a = M.discretize_batch(y, u, 0.3, const, include_J2=j2, n_sub=10)
    b = M.discretize_batch(y, u, 0.3, const, include_J2=j2, adaptive=dict())
    c = M.discretize_batch(y, np.repeat(u, 3, axis=2)[:, :, :20], 0.3, const, include_J2=j2, n_sub=10)
    d = M.discretize_batch(y, np.repeat(u, 3, axis=2)[:, :, :20], 0.3, const, include_J2=j2, adaptive=dict())
res, x2, u2 = M.propagate_discretize(Y, 0.3, ctrl, const, T=K, n_sub_disc=8)
import torch
dev = torch.device("cuda:0")
xs, us = torch.from_numpy(np.ascontiguousarray(y)).to(dev), torch.from_numpy(np.ascontiguousarray(u)).to(dev)
tfv = torch.full((N,), 0.3, dtype=torch.float64, device=dev)
outs = [torch.zeros((105, N * (K - 1)), dtype=torch.float64, device=dev) for _ in range(8)]
for nd in (2, 4, 8):
    M.discretize_batch_device(xs, us, tfv, const, n_sub=6, out=outs[0], extra_dst=outs[1:nd])
torch.cuda.synchronize()
print("sanitize pass done; launches", M.launch_count(), "finite", bool(np.isfinite(res.soa).all() and np.isfinite(a.soa).all()))
