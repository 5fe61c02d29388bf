"""torchrun script (round 2): the fused all-gather in both gathered layouts, verified against a single-rank computation
and timed -- BASELINE configs[2] weak-scaled (4096 satellites per rank) and configs[3] strong-scaled (5025 satellites in
total), two-kernel sequence against the overlapped pass.
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/r02_multi_gpu.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
import mpconstellation_b200 as M
from mpconstellation_b200 import distributed as D
from bench import make_constellation

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
os.environ.setdefault("NCCL_DEBUG", "WARN")
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
K, tf = 200, 2.0
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
res = {"world": world}


def log(*a):
    if rank == 0:
        print(*a, flush=True)


def timed(fn, reps=6):
    ts = []
    for i in range(reps + 2):
        flush.fill_(1.0)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        if i >= 2:
            ts.append(e0.elapsed_time(e1))
    t = torch.tensor([float(np.mean(ts))], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


# ---------------------------------------------------------------- 1. verification (small: 257 satellites in total, ragged)
Nv = 257
Yv, const = make_constellation(Nv)
yv = torch.from_numpy(Yv).to(dev)
tfv = torch.full((Nv,), tf, dtype=torch.float64, device=dev) * (1 + 0.01 * torch.arange(Nv, device=dev, dtype=torch.float64) / Nv)
x_all, u_all, _ = M.propagate_batch_device(yv, tfv, ctrl, const, include_drag=False, include_J2=True, T=K)
full, _ = M.discretize_batch_device(x_all, u_all, tfv, const, include_J2=True)
torch.cuda.synchronize()
n = K - 1
eq = lambda a, b: bool(torch.equal(a.view(torch.int64), b.view(torch.int64)))
ok_all = True
for layout in ("satmajor", "kmajor"):
    for mode in ("unicast", "multicast"):
        try:
            fg = D.FusedGather(Nv, K, device=dev, mode=mode, layout=layout)
        except Exception as exc:
            log(f"verify {layout}/{mode}: unavailable ({exc})")
            continue
        want = full if layout == "satmajor" else full.view(105, Nv, n).permute(0, 2, 1).reshape(105, Nv * n)
        s0, s1 = fg.s0, fg.s1
        for name, call in (("two kernels", lambda: fg.discretize(x_all[s0:s1].contiguous(), u_all[s0:s1].contiguous(), tfv[s0:s1].contiguous(), const, include_J2=True)),
                           ("fused pass", lambda: fg.propagate_discretize(yv[s0:s1].contiguous(), tfv[s0:s1].contiguous(), ctrl, const, include_J2=True, n_windows=4))):
            fg.buf[:42].fill_(float("nan"))
            fg.buf[49:].fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            call()
            torch.cuda.synchronize()
            dist.barrier()
            ok = eq(fg.buf, want)
            t = torch.tensor([int(ok)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok_all &= bool(int(t[0]))
            log(f"verify {layout}/{mode}/{name}: {'bit-identical to the single-rank result on every rank' if int(t[0]) else 'MISMATCH'}")
        A = fg.view().sat(Nv - 1)[0]
        assert np.array_equal(A, full[:49, -(K - 1):].T.reshape(K - 1, 7, 7).cpu().numpy())
        del fg
res["verified"] = ok_all

# ---------------------------------------------------------------- 2. timing
for tag, N_total in (("config3_weak", 4096 * world), ("config4_strong", 5025)):
    Y, const = make_constellation(N_total)
    s0, s1 = D.shard_range(N_total, rank, world)
    y0 = torch.from_numpy(np.ascontiguousarray(Y[s0:s1])).to(dev)
    Nl = s1 - s0
    tfd = torch.full((Nl,), tf, dtype=torch.float64, device=dev)
    x = torch.empty((Nl, 7, K), dtype=torch.float64, device=dev)
    u = torch.empty((Nl, 3, K), dtype=torch.float64, device=dev)
    stp = torch.empty(Nl, dtype=torch.int32, device=dev)
    r = {"satellites_total": N_total, "intervals_total": N_total * (K - 1)}
    # local work only (no exchange): the floor
    out_l = torch.empty((105, Nl * (K - 1)), dtype=torch.float64, device=dev)
    std = torch.empty(Nl * (K - 1), dtype=torch.int32, device=dev)
    r["local_pass_no_exchange_ms"] = timed(lambda: M.propagate_discretize_device(y0, tfd, ctrl, const, K, y=x, u_out=u, out=out_l, status_prop=stp, status_disc=std))
    r["local_propagate_ms"] = timed(lambda: M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, y=x, u_out=u, status=stp))
    r["local_discretize_ms"] = timed(lambda: M.discretize_batch_device(x, u, tfd, const, out=out_l, status=std))
    del out_l
    for layout in ("satmajor", "kmajor"):
        for stag in ((0, 4) if world >= 8 else (0,)):
            fg = D.FusedGather(N_total, K, device=dev, mode="unicast", layout=layout, stagger=stag)
            key = f"{layout}_stagger{stag}"
            r[key + "_two_kernels_ms"] = timed(lambda: (M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, y=x, u_out=u, status=stp),
                                                          fg.discretize(x, u, tfd, const)))
            if fg.overlap_ok:
                for nw in (0, 8, 32):
                    r[key + f"_fused_pass_w{nw}_ms"] = timed(lambda: fg.propagate_discretize(y0, tfd, ctrl, const, y=x, u_out=u, status_prop=stp, n_windows=nw))
            del fg
    try:
        fg = D.FusedGather(N_total, K, device=dev, mode="multicast", layout="kmajor")
        r["kmajor_multicast_fused_pass_ms"] = timed(lambda: fg.propagate_discretize(y0, tfd, ctrl, const, y=x, u_out=u, status_prop=stp))
        del fg
    except Exception as exc:
        r["kmajor_multicast_fused_pass_ms"] = f"unavailable: {exc}"
    res[tag] = r
    log(tag, json.dumps(r))
if rank == 0:
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", f"r02_multi_gpu_w{world}.json"), "w"), indent=1)
dist.barrier()
dist.destroy_process_group()
