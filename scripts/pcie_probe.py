import torch, time, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda:0")
n = 815104 * 105
d = torch.empty(n, dtype=torch.float64, device=dev); d.fill_(1.0)
h = torch.empty(n, dtype=torch.float64, pin_memory=True)
for label, fn in (("1D D2H 685 MB", lambda: h.copy_(d, non_blocking=True)),):
    for _ in range(2): fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{label}: {dt*1e3:.2f} ms  {n*8/dt/1e9:.1f} GB/s")
d2 = d.view(105, -1); h2 = h.view(105, -1)
cols = 815104 // 8
def chunked():
    for c in range(8):
        h2[:, c*cols:(c+1)*cols].copy_(d2[:, c*cols:(c+1)*cols], non_blocking=True)
for _ in range(2): chunked(); torch.cuda.synchronize()
t0 = time.perf_counter(); chunked(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"8 strided 2D chunks: {dt*1e3:.2f} ms  {n*8/dt/1e9:.1f} GB/s")
import mpconstellation_b200 as M, numpy as np
from bench import make_constellation
Y, const = make_constellation(4096)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
y0 = M.pinned_empty((4096, 7)); y0[:] = Y
out = M.pinned_empty((105, 815104)); yh = M.pinned_empty((4096, 7, 200)); uh = M.pinned_empty((4096, 3, 200))
for i in range(4):
    t0 = time.perf_counter(); M.propagate_discretize(y0, 2.0, ctrl, const, T=200, out=out, y_out=yh, u_out=uh); dt = time.perf_counter() - t0
print(f"propagate_discretize host API: {dt*1e3:.2f} ms")
