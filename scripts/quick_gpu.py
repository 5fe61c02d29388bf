import sys, time
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))); sys.path.insert(0, sys.path[0] + "/tests")
import numpy as np, torch
import mpconstellation_b200 as M
from conftest import synth_batch
from oracle.mpc_oracle import OracleConstants
g = np.load(sys.path[1] + "/tests/golden/discretize.npz"); const = OracleConstants(*g["const"])
print("fp64 peak", M.fp64_peak_tflops())
N, K = 4096, 200
y0, _, _ = synth_batch(N, 2, 2.0, const)
dev = torch.device("cuda:0")
y0d = torch.from_numpy(y0).to(dev); tfd = torch.full((N,), 2.0, dtype=torch.float64, device=dev)
c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    e[0].record()
    y, u, st = M.propagate_batch_device(y0d, tfd, c, const, include_drag=False, include_J2=False, T=K)
    e[1].record()
    out, st2 = M.discretize_batch_device(y, u, tfd, const)
    e[2].record(); torch.cuda.synchronize()
    tp, td = e[0].elapsed_time(e[1]), e[1].elapsed_time(e[2])
    print(f"propagate {tp:.3f} ms  discretize {td:.3f} ms  -> {N*(K-1)/td*1e3:.3e} intervals/s (disc only)")
