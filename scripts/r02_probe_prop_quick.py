"""Quick probe: the RK45 propagator alone (BASELINE configs 1, 2, 5, 3) and the overlapped pass of config 3 (CUDA events,
L2 flushed), plus the distance of config 3's trajectories from the reference fixture (bench_workload.npz)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from bench import make_constellation

dev = torch.device("cuda:0")
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(fn, n=10):
    ts = []
    for i in range(n + 3):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        if i >= 3:
            ts.append(a.elapsed_time(b))
    return round(float(np.mean(ts)), 4), round(float(np.min(ts)), 4)


res = {}
for name, N, K, tf in (("config1", 1, 50, 0.5), ("config2", 64, 100, 1.0), ("config5", 256, 60, 2.0), ("config3", 4096, 200, 2.0)):
    Y, const = make_constellation(N)
    y0 = torch.from_numpy(Y).to(dev)
    tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    r = {"propagate_ms": timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K))}
    if N >= 64:
        r["propagate_drag_j2_ms"] = timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=True, include_J2=True, T=K))
    r["pass_ms"] = timed(lambda: M.propagate_discretize_device(y0, tfd, ctrl, const, K, n_sub_disc=100))
    res[name] = r
    print(name, json.dumps(r), flush=True)
g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "bench_workload.npz"))
Y, const = make_constellation(4096)
y0 = torch.from_numpy(Y).to(dev)
tfd = torch.full((4096,), 2.0, dtype=torch.float64, device=dev)
y, u, st = M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=200)
torch.cuda.synchronize()
worst = 0.0
for j, s_ in enumerate(g["idx"]):
    worst = max(worst, float(np.max(np.abs(y[s_].cpu().numpy() - g[f"s{j}_x"])) / np.max(np.abs(g[f"s{j}_x"]))))
print("config3 states vs the reference-flown satellites (norm-relative):", f"{worst:.2e}", "status max", int(st.max()))
