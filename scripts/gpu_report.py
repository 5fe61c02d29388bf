"""Parity + timing report on the GPU box: errors vs the reference fixtures (both quadrature modes) and kernel
times for BASELINE configs 1-3 in both modes.  Output is committed under profiles/."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mpconstellation_b200 as M
from bench import make_constellation
from oracle.mpc_oracle import OracleConstants, norm_rel_err
g = np.load(os.path.join(ROOT, "tests/golden/discretize.npz")); const = OracleConstants(*g["const"])
NAMES = ["A_k", "B_kp", "B_kn", "Sigma_k", "xi_k"]
print("== parity vs the unmodified reference (tests/golden), norm-relative max|d|/max|ref| per matrix")
for sc in ("d0", "d1", "d2", "d3", "d4"):
    for mode, uni in (("uni", True), ("def", False)):
        if f"{sc}_{mode}_A_k" not in g: continue
        d = M.Discretizer(const); d.use_uniform_steps = uni
        out = d.discretize(M.Simulator.satellite_dynamics, g[sc + "_x"], g[sc + "_u"], float(g[sc + "_tf"]))
        print(f"  {sc} K={g[sc+'_x'].shape[1]:3d} use_uniform_steps={uni!s:5}: " + "  ".join(f"{n} {norm_rel_err(o, g[f'{sc}_{mode}_{n}']):.1e}" for n, o in zip(NAMES, out)))
dev = torch.device("cuda:0")
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
print("== kernel times (device-resident, CUDA events, best of 5)")
for name, N, K, tf in (("config 1", 1, 50, 0.5), ("config 2", 64, 100, 1.0), ("config 3", 4096, 200, 2.0), ("1M sweep", 5025, 200, 2.0)):
    Y, c2 = make_constellation(N)
    y0 = torch.from_numpy(Y).to(dev); tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    def best(fn):
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return min(ts[1:])
    tp = best(lambda: M.propagate_batch_device(y0, tfd, ctrl, c2, include_drag=False, include_J2=False, T=K))
    tpd = best(lambda: M.propagate_batch_device(y0, tfd, ctrl, c2, include_drag=True, include_J2=True, T=K))
    y, u, _ = M.propagate_batch_device(y0, tfd, ctrl, c2, include_drag=False, include_J2=False, T=K)
    tu = best(lambda: M.discretize_batch_device(y, u, tfd, c2))
    tj = best(lambda: M.discretize_batch_device(y, u, tfd, c2, include_J2=True))
    nn = torch.zeros(N * (K - 1), dtype=torch.int32, device=dev)
    ta = best(lambda: M.discretize_batch_device(y, u, tfd, c2, adaptive=dict(), n_nodes=nn))
    n_int = N * (K - 1)
    print(f"  {name}: {N} sats x K={K} ({n_int} intervals): propagate {tp:.3f} ms (drag+J2 {tpd:.3f}) | discretize uniform-101 {tu:.3f} ms "
          f"({n_int/tu*1e3:.3e}/s), J2 {tj:.3f} ms | default/adaptive {ta:.3f} ms ({n_int/ta*1e3:.3e}/s, nodes {int(nn.min())}-{int(nn.max())})")
