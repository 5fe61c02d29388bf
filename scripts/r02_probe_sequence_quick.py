"""Quick probe: the RK45 propagator with the table law (SequenceController, KIND 3) on BASELINE config 5's shape (256
satellites, per-satellite tables of 60 knots), without and with drag + J2; the tangential law beside it."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import mpconstellation_b200 as M
from bench import make_constellation

dev = torch.device("cuda:0")
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
for N, K, tf in ((256, 60, 2.0), (256, 200, 2.0), (4096, 60, 2.0), (4096, 200, 2.0)):
    Y, const = make_constellation(N)
    y0 = torch.from_numpy(Y).to(dev)
    tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    tang = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    x, u, st = M.propagate_batch_device(y0, tfd, tang, const, include_drag=False, include_J2=False, T=K)
    torch.cuda.synchronize()
    seq = M.SequenceController(u=u.cpu().numpy(), tf_u=tf, tf_sim=tf)          # per-satellite tables [N,3,K]
    r = {"tangential_ms": timed(lambda: M.propagate_batch_device(y0, tfd, tang, const, include_drag=False, include_J2=False, T=K)),
         "sequence_ms": timed(lambda: M.propagate_batch_device(y0, tfd, seq, const, include_drag=False, include_J2=False, T=K)),
         "sequence_drag_j2_ms": timed(lambda: M.propagate_batch_device(y0, tfd, seq, const, include_drag=True, include_J2=True, T=K))}
    xs, us, sts = M.propagate_batch_device(y0, tfd, seq, const, include_drag=False, include_J2=False, T=K)
    torch.cuda.synchronize()
    r["checksum_sequence"] = float(xs.abs().sum().item())
    r["seq_vs_tangential_states"] = float((xs - x).abs().max().item() / x.abs().max().item())
    res[f"{N}x{K}"] = r
    print(f"{N}x{K}", json.dumps(r), flush=True)
