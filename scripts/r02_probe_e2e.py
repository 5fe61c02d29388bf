"""Round-2 probe: the end-to-end host pass on BASELINE configs[2] -- satellite-major pipeline against the streamed
k-major pass with 16 / 32 / 48 / 64 k-windows (capped at two per kernel wave) (wall clock per call, pinned buffers)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

N, K, tf = 4096, 200, 2.0
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
n_int = N * (K - 1)
y0_h = M.pinned_empty((N, 7)); y0_h[:] = Y
out_h = M.pinned_empty((105, n_int)); y_h = M.pinned_empty((N, 7, K)); u_h = M.pinned_empty((N, 3, K)); st_h = M.pinned_empty(n_int, np.int32)
L = _lib.lib()


def run(layout, n=8):
    ts = []
    for i in range(n + 2):
        t0 = time.perf_counter()
        M.propagate_discretize(y0_h, tf, ctrl, const, T=K, out=out_h, y_out=y_h, u_out=u_h, status=st_h, layout=layout)
        if i >= 2:
            ts.append(time.perf_counter() - t0)
    return np.mean(ts) * 1e3, np.min(ts) * 1e3


print("satellite-major pipeline (chunks of one kernel wave): mean %.3f ms  min %.3f ms" % run("satmajor"))
for var, nw in ((26, 16), (27, 32), (28, 48), (29, 64)):
    L.mpc_set_tuning(var)
    print("k-major streamed pass, %2d windows: mean %.3f ms  min %.3f ms" % ((nw,) + run("kmajor")))
L.mpc_set_tuning(27)
print("bytes over PCIe per call: D2H %.1f MB (98 of 105 rows of the matrices, trajectory, inputs, status), H2D %.2f MB" %
      ((98 * n_int * 8 + y_h.nbytes + u_h.nbytes + n_int * 4) / 1e6, (y0_h.nbytes + N * 8) / 1e6))
