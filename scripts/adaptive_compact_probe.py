"""A/B of the adaptive (default-mode) kernel on BASELINE configs[2]: inlined build vs the COMPACT build
(mpc_set_tuning(9), DESIGN.md section 9 plan item 1).  Prints times (CUDA events, L2 flushed) and the largest
difference between the two results.  NOT YET RUN ON A GPU at the end of round 1: first thing to run in round 2."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

N, K, tf = 4096, 200, 2.0
dev = torch.device("cuda:0")
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
x, u, _ = M.propagate_batch_device(torch.from_numpy(Y).to(dev), tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
ad = dict(rtol=1e-3, atol=1e-6, max_step=1e-2)
res = {}
for name, variant in (("inlined", 10), ("compact", 9)):
    _lib.check(_lib.lib().mpc_set_tuning(variant))
    out = torch.empty((105, N * (K - 1)), dtype=torch.float64, device=dev)
    st = torch.empty(N * (K - 1), dtype=torch.int32, device=dev)
    nn = torch.empty(N * (K - 1), dtype=torch.int32, device=dev)
    ts = []
    for i in range(7):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        M.discretize_batch_device(x, u, tfd, const, out=out, status=st, adaptive=ad, n_nodes=nn)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    res[name] = (out, st, nn)
    print(f"{name:8s}: mean {np.mean(ts):.3f} ms  min {np.min(ts):.3f} ms  status max {int(st.max())}  nodes {int(nn.min())}-{int(nn.max())}")
_lib.check(_lib.lib().mpc_set_tuning(10))
a, b = res["inlined"], res["compact"]
den = float(a[0].abs().max())
print("largest difference / largest entry:", float((a[0] - b[0]).abs().max()) / den, " node counts equal:", bool(torch.equal(a[2], b[2])))
