"""Times every tuning variant of the discretization kernel on BASELINE configs[2] (device-resident)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation
N, K = int(os.environ.get("N", 4096)), int(os.environ.get("K", 200))
Y, const = make_constellation(N)
dev = torch.device("cuda:0")
y0 = torch.from_numpy(Y).to(dev); tfd = torch.full((N,), 2.0, dtype=torch.float64, device=dev)
c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
y, u, st = M.propagate_batch_device(y0, tfd, c, const, include_drag=False, include_J2=False, T=K)
ref = None
variants = [int(v) for v in os.environ.get("VARIANTS", "0,1,2,3,4,5,6").split(",")]
for v in variants:
    _lib.check(_lib.lib().mpc_set_tuning(v))
    ts = []
    for it in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out, st2 = M.discretize_batch_device(y, u, tfd, const); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    o = out.cpu().numpy()
    if ref is None: ref = o
    print(f"variant {v}: best {min(ts[1:]):.3f} ms  median {np.median(ts[1:]):.3f} ms -> {N*(K-1)/min(ts[1:])*1e3:.4e} intervals/s; "
          f"max|diff vs v0| {np.max(np.abs(o-ref)):.2e} status {int(st2.max())}")
_lib.lib().mpc_set_tuning(0)
