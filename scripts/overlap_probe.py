"""Times one SCP linearization pass on BASELINE configs[2] (4096 satellites x K=200): the two kernels back to back vs
mpc_propagate_discretize with the propagation overlapped, for several window counts.  CUDA events, L2 flushed."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from bench import make_constellation

N, K, tf, n_sub = 4096, 200, 2.0, 100
dev = torch.device("cuda:0")
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
y0 = torch.from_numpy(Y).to(dev)
tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
x = torch.empty((N, 7, K), dtype=torch.float64, device=dev)
u = torch.empty((N, 3, K), dtype=torch.float64, device=dev)
out = torch.empty((105, N * (K - 1)), dtype=torch.float64, device=dev)
sp = torch.empty(N, dtype=torch.int32, device=dev)
sd = torch.empty(N * (K - 1), dtype=torch.int32, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
n_prop = M.batch.default_n_sub(K) if "--rk4" in sys.argv else 0       # 0: the RK45 replay (the default propagator)
sys.argv = [a for a in sys.argv if a != "--rk4"]


def seq():
    M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, n_sub=n_prop, y=x, u_out=u, status=sp)
    M.discretize_batch_device(x, u, tfd, const, n_sub=n_sub, out=out, status=sd)


def ovl(nw):
    M.propagate_discretize_device(y0, tfd, ctrl, const, K, n_sub_prop=n_prop, n_sub_disc=n_sub, y=x, u_out=u, out=out,
                                  status_prop=sp, status_disc=sd, n_windows=nw)


def timeit(fn, reps=8):
    ts = []
    for i in range(reps + 2):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


seq(); torch.cuda.synchronize()
ref = out.clone()
print("back to back      : mean %.3f ms  min %.3f ms" % timeit(seq))
for nw in [int(a) for a in sys.argv[1:]] or [2, 4, 8, 12, 16, 24, 32]:
    out.zero_()
    m = timeit(lambda: ovl(nw))
    print("overlapped, %2d win: mean %.3f ms  min %.3f ms   identical=%s" % (nw, m[0], m[1], bool(torch.equal(out, ref))))
