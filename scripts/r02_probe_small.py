"""Round-2 probe: small-batch discretization (fixed-step mode), the thread-group kernel (8 lanes per interval) against the
one-thread-per-interval kernel, over batch sizes from BASELINE config 1 to beyond config 5; single-call latency of
Discretizer.discretize through the host API."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

dev = torch.device("cuda:0")
L = _lib.lib()
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)


def ev_ms(fn, n=8):
    ts = []
    for i in range(n + 3):
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts))


print("intervals | one thread per interval | thread-group (8 lanes) ")
for name, N, K, tf in (("config1", 1, 50, 0.5), ("10x100", 10, 100, 1.0), ("24x100", 24, 100, 1.0), ("config2", 64, 100, 1.0),
                       ("83x100", 83, 100, 1.0), ("config5", 256, 60, 2.0), ("190x100", 190, 100, 1.0), ("400x100", 400, 100, 1.0)):
    Y, const = make_constellation(N)
    y0 = torch.from_numpy(Y).to(dev)
    tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    x, u, _ = M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
    L.mpc_set_tuning(23)
    t_thread = ev_ms(lambda: M.discretize_batch_device(x, u, tfd, const, n_sub=100))
    a, _ = M.discretize_batch_device(x, u, tfd, const, n_sub=100)
    L.mpc_set_tuning(25)
    t_group = ev_ms(lambda: M.discretize_batch_device(x, u, tfd, const, n_sub=100))
    b, _ = M.discretize_batch_device(x, u, tfd, const, n_sub=100)
    L.mpc_set_tuning(24)
    torch.cuda.synchronize()
    d = float((a - b).abs().max() / a.abs().max())
    print(f"{name:9s} {N * (K - 1):6d} | {t_thread:.4f} ms | {t_group:.4f} ms | max difference / max entry {d:.1e}")

# single-call latency through the reference-shaped host API (one satellite, as optimizer.py:243-249 calls it)
hub = np.array([5371.4806e3, -4133.1393e3, 1399.9594e3, 4.6921e3, 4.9848e3, -3.2752e3, 12200.0])
for K, tf in ((50, 0.5), (100, 1.0)):
    sat = M.Satellite(hub[0:3].copy(), hub[3:6].copy(), float(hub[6]))
    scale = M.SatelliteScale(sat=sat)
    c = M.ConstantTangentialThrustController([sat], 0.5)
    sim = M.Simulator(sats=[sat], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=tf)
    x, u = sim.sim_data[sat.id], sim.sim_u[sat.id]
    d = M.Discretizer(scale.get_normalized_constants())
    f = M.Simulator.satellite_dynamics

    def wall(fn, n=40):
        for _ in range(5):
            fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3
    res = {}
    for mode, uni in (("default", False), ("uniform-101", True)):
        d.use_uniform_steps = uni
        for tag, var in (("thread", 23), ("group", 24)):
            L.mpc_set_tuning(var)
            res[f"{mode}/{tag}"] = wall(lambda: d.discretize(f, x, u, tf))
    L.mpc_set_tuning(24)
    t_run = wall(lambda: M.Simulator(sats=[M.Satellite(hub[0:3].copy(), hub[3:6].copy(), float(hub[6]))], controller=c, scale=scale,
                                     base_res=100, include_drag=False, include_J2=False).run(tf=tf), 10)
    print(f"K={K}: Discretizer.discretize per call: " + ", ".join(f"{k} {v:.3f} ms" for k, v in res.items()) + f" | Simulator.run {t_run:.3f} ms")
