"""Two full-batch launches of discretize_pair_kernel on BASELINE configs[2] (4096 satellites x K=200, integrator_steps 101)
for ncu:   ncu --set full -k regex:discretize_pair_kernel --launch-skip 1 -c 1 python scripts/r02_ncu_pair.py [--all-nodes]
--all-nodes: every one of the 101 nodes evaluated (mpc_set_tuning(37)) instead of the 21-node form."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

dev = torch.device("cuda:0")
N, K, tf = 4096, 200, 2.0
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
x, u, _ = M.propagate_batch_device(torch.from_numpy(Y).to(dev), tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
if "--all-nodes" in sys.argv:
    _lib.check(_lib.lib().mpc_set_tuning(37))
out = torch.empty((105, N * (K - 1)), dtype=torch.float64, device=dev)
st = torch.empty(N * (K - 1), dtype=torch.int32, device=dev)
for _ in range(2):
    M.discretize_batch_device(x, u, tfd, const, n_sub=100, out=out, status=st)
torch.cuda.synchronize()
print("status max", int(st.max()))
