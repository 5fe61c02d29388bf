"""Round-2 probe of the default-mode (adaptive) discretization on BASELINE configs[2]: the shipped build
(discretize_default_kernel) against the round-1 build (mpc_set_tuning(9)); times (CUDA events, L2 flushed), node counts,
difference between the two, and parity against the unmodified reference's default-mode fixture (bench_workload.npz).
`--once`: a single launch of the shipped build after warm-up (the launch ncu captures)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import mpconstellation_b200 as M
from mpconstellation_b200 import _lib
from bench import make_constellation

N, K, tf = 4096, 200, 2.0
dev = torch.device("cuda:0")
Y, const = make_constellation(N)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
x, u, _ = M.propagate_batch_device(torch.from_numpy(Y).to(dev), tfd, ctrl, const, include_drag=False, include_J2=False, T=K)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
ad = dict(rtol=1e-3, atol=1e-6, max_step=1e-2)
n_int = N * (K - 1)
if "--once" in sys.argv:
    out = torch.empty((105, n_int), dtype=torch.float64, device=dev)
    st = torch.empty(n_int, dtype=torch.int32, device=dev)
    for _ in range(2):
        M.discretize_batch_device(x, u, tfd, const, out=out, status=st, adaptive=ad)
    torch.cuda.synchronize()
    sys.exit(0)
res = {}
for name, variant in (("round-1 build", (9, 20)), ("shipped, 32-thread CTAs", (10, 20)),
                      ("shipped build", (10, 22))):
    for v in variant:
        _lib.check(_lib.lib().mpc_set_tuning(v))
    out = torch.full((105, n_int), float("nan"), dtype=torch.float64, device=dev)
    st = torch.empty(n_int, dtype=torch.int32, device=dev)
    nn = torch.empty(n_int, dtype=torch.int32, device=dev)
    ts = []
    for i in range(8):
        flush.fill_(1.0)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        M.discretize_batch_device(x, u, tfd, const, out=out, status=st, adaptive=ad, n_nodes=nn)
        b.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(a.elapsed_time(b))
    res[name] = (out, st, nn)
    print(f"{name}: mean {np.mean(ts):.3f} ms  min {np.min(ts):.3f} ms  status max {int(st.max())}  nodes {int(nn.min())}-{int(nn.max())}"
          f" (mean {float(nn.double().mean()):.3f})  {n_int / np.mean(ts) / 1e3:.3e} intervals/s")
_lib.check(_lib.lib().mpc_set_tuning(10))
_lib.check(_lib.lib().mpc_set_tuning(22))
assert torch.equal(res["shipped, 32-thread CTAs"][0], res["shipped build"][0]), "CTA size must not change a bit"
a, b = res["round-1 build"], res["shipped build"]
for r0, r1, nm in ((0, 49, "A_k"), (49, 70, "B_kp"), (70, 91, "B_kn"), (91, 98, "Sigma_k"), (98, 105, "xi_k")):
    den = float(a[0][r0:r1].abs().max())
    print(f"  {nm}: largest difference between the builds / largest entry = {float((a[0][r0:r1] - b[0][r0:r1]).abs().max()) / den:.2e}")
print("  node counts equal:", bool(torch.equal(a[2], b[2])))
gb = np.load(os.path.join(ROOT, "tests", "golden", "bench_workload.npz"))
soa = b[0].cpu().numpy().reshape(105, N, K - 1)
names = ["A_k", "B_kp", "B_kn", "Sigma_k", "xi_k"]
worst = 0.0
for j, i in enumerate(gb["idx"]):
    blk = soa[:, i][:, gb["ks"]]
    got = [blk[0:49].T.reshape(-1, 7, 7), blk[49:70].T.reshape(-1, 7, 3), blk[70:91].T.reshape(-1, 7, 3), blk[91:98], blk[98:105]]
    for n, g in zip(names, got):
        ref = gb[f"s{j}_def_{n}"]
        worst = max(worst, float(np.max(np.abs(g - ref)) / np.max(np.abs(ref))))
print(f"shipped build vs the unmodified reference's default mode (4 satellites x 6 intervals of this batch): {worst:.2e}")
