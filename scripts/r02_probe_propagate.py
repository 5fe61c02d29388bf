"""Round-2 probe: the RK45 propagator (the reference's integrator, replayed) timed on BASELINE configs 2, 3, 5 and 1,
with / without the speculative first stage and for every satellites-per-warp mapping; the RK4 propagator beside it;
the overlapped pass (config 3) with both; parity of config 3 against the reference fixture (bench_workload.npz)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

dev = torch.device("cuda:0")
L = _lib.lib()
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(fn, n=6):
    ts = []
    for i in range(n + 2):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    return float(np.mean(ts)), float(np.min(ts))


res = {}
for name, N, K, tf in (("config1", 1, 50, 0.5), ("config2", 64, 100, 1.0), ("config5", 256, 60, 2.0), ("config3", 4096, 200, 2.0)):
    Y, const = make_constellation(N)
    y0 = torch.from_numpy(Y).to(dev)
    tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    r = {}
    n_rk4 = M.batch.default_n_sub(K)
    r["rk4"] = timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, n_sub=n_rk4))
    ref = None
    for spec in (12, 11):
        L.mpc_set_tuning(spec)
        for lv, lpw in ((13, "auto"), (14, 32), (15, 16), (16, 8), (17, 4), (18, 2), (19, 1)):
            if isinstance(lpw, int) and lpw * 148 * 4 * 8 < N:
                continue        # would need more than 8 warps per scheduler... skip the silly ones
            L.mpc_set_tuning(lv)
            key = f"rk45_spec{int(spec == 12)}_lpw{lpw}"
            r[key] = timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K))
            y, u, st = M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
            torch.cuda.synchronize()
            assert int(st.max()) == 0
            if ref is None:
                ref = y.clone()
            assert torch.equal(ref, y), key
    L.mpc_set_tuning(12)
    L.mpc_set_tuning(13)
    # drag + J2 (the Simulator's defaults)
    r["rk45_dragj2"] = timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=True, include_J2=True, T=K))
    r["rk4_dragj2"] = timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=True, include_J2=True, T=K, n_sub=n_rk4))
    # discretization, both modes, and the fused pass
    x, u, _ = M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
    r["disc_uniform"] = timed(lambda: M.discretize_batch_device(x, u, tfd, const, n_sub=100))
    r["disc_default"] = timed(lambda: M.discretize_batch_device(x, u, tfd, const, adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2)))
    r["pass_rk45"] = timed(lambda: M.propagate_discretize_device(y0, tfd, ctrl, const, K, n_sub_disc=100))
    r["pass_rk4"] = timed(lambda: M.propagate_discretize_device(y0, tfd, ctrl, const, K, n_sub_prop=n_rk4, n_sub_disc=100))
    res[name] = r
    print(name, json.dumps({k: round(v[0], 4) for k, v in r.items()}), flush=True)
    if name == "config3":
        gb = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "bench_workload.npz"))
        xs = x.cpu().numpy()
        e = max(float(np.max(np.abs(xs[i] - gb[f"s{j}_x"])) / np.max(np.abs(gb[f"s{j}_x"]))) for j, i in enumerate(gb["idx"]))
        print("config3 propagated states vs the unmodified reference:", e)
        res["config3_state_err_vs_reference"] = e
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_probe_propagate.json"), "w"), indent=1)
