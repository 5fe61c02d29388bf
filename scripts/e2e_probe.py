"""Times the public host API (propagate_discretize, pinned buffers) on BASELINE configs[2]; experiment helper."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpconstellation_b200 as M
from bench import make_constellation
N, K = int(os.environ.get("SATS", 4096)), 200
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
y0 = M.pinned_empty((N, 7)); y0[:] = Y
st = M.pinned_empty(N * (K - 1), np.int32)
out = M.pinned_empty((105, N * (K - 1))); yh = M.pinned_empty((N, 7, K)); uh = M.pinned_empty((N, 3, K))
ts = []
for i in range(8):
    t0 = time.perf_counter()
    res, _, _ = M.propagate_discretize(y0, 2.0, ctrl, const, T=K, out=out, y_out=yh, u_out=uh, status=st)
    ts.append((time.perf_counter() - t0) * 1e3)
print("propagate_discretize host API ms:", " ".join(f"{t:.2f}" for t in ts))
A = res.sat(N - 1)[0]
assert np.all(A[:, 6, :6] == 0) and np.all(A[:, 6, 6] == 1)
