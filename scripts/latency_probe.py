"""Per-call latency of the reference-shaped entry points for ONE satellite (the reference's calling pattern,
optimizer.py:243-249 / simulator.py:41-45): Discretizer.discretize and Simulator.run, wall clock per call."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpconstellation_b200 as M
from mpconstellation_b200 import batch, _lib

hub = np.array([5371.4806e3, -4133.1393e3, 1399.9594e3, 4.6921e3, 4.9848e3, -3.2752e3, 12200.0])
for K, tf in ((50, 0.5), (100, 1.0), (200, 2.0)):
    sat = M.Satellite(hub[0:3].copy(), hub[3:6].copy(), float(hub[6]))
    scale = M.SatelliteScale(sat=sat)
    ctrl = M.ConstantTangentialThrustController([sat], 0.5)
    sim = M.Simulator(sats=[sat], controller=ctrl, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=tf)
    x, t = sim.sim_data[sat.id], sim.sim_time[sat.id]
    u = sim.sim_u[sat.id]
    d = M.Discretizer(scale.get_normalized_constants())

    def timeit(fn, n=30):
        fn(); fn()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        return (time.perf_counter() - t0) / n * 1e3
    f = M.Simulator.satellite_dynamics
    d.use_uniform_steps = False
    t_def = timeit(lambda: d.discretize(f, x, u, tf))
    d.use_uniform_steps = True
    t_uni = timeit(lambda: d.discretize(f, x, u, tf))
    out_p = np.empty((105, K - 1))
    t_uni_pageable = timeit(lambda: d.discretize_batch(f, x[None], u[None], tf, out=out_p).sat(0))
    out_pin = _lib.pinned_empty((105, K - 1))
    t_uni_pinned = timeit(lambda: d.discretize_batch(f, x[None], u[None], tf, out=out_pin).sat(0))

    def run():
        s2 = M.Satellite(hub[0:3].copy(), hub[3:6].copy(), float(hub[6]))
        M.Simulator(sats=[s2], controller=M.ConstantTangentialThrustController([s2], 0.5), scale=scale, base_res=100,
                    include_drag=False, include_J2=False).run(tf=tf)
    t_run = timeit(run, 10)
    print(f"K={K}: discretize default-mode {t_def:.3f} ms | uniform-101 {t_uni:.3f} ms (caller's pageable out {t_uni_pageable:.3f}, "
          f"caller's pinned out {t_uni_pinned:.3f}) | Simulator.run {t_run:.3f} ms")
