"""torchrun script: the overlapped propagate -> discretize pass with the all-gather fused in (FusedGather.propagate_discretize)
against a single-rank computation of every satellite (bit-identical?) and against the back-to-back sequence (time).
    torchrun --nproc-per-node N scripts/multi_gpu_overlap_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
import mpconstellation_b200 as M
from mpconstellation_b200 import distributed as D
from bench import make_constellation

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("NCCL_DEBUG", "WARN")
dist.init_process_group("nccl", device_id=dev)
N, K, tf = int(os.environ.get("N", 4096)), int(os.environ.get("K", 200)), 2.0
Y, const = make_constellation(N * world)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
tfd_all = torch.full((N * world,), tf, dtype=torch.float64, device=dev)
y0_all = torch.from_numpy(Y).to(dev)
y_all, u_all, _ = M.propagate_batch_device(y0_all, tfd_all, ctrl, const, include_drag=False, include_J2=False, T=K)
full, _ = M.discretize_batch_device(y_all, u_all, tfd_all, const)
s0, s1 = D.shard_range(N * world, rank, world)
y0, tfd = y0_all[s0:s1].contiguous(), tfd_all[s0:s1].contiguous()
x = torch.empty((N, 7, K), dtype=torch.float64, device=dev)
u = torch.empty((N, 3, K), dtype=torch.float64, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)


def timed(fn, reps=6):
    ts = []
    for _ in range(reps):
        flush.fill_(1.0)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([sum(ts[1:]) / (reps - 1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])


lines, ok_all = [], 1
for mode in os.environ.get("MODES", "unicast,multicast").split(","):
    try:
        fg = D.FusedGather(N * world, K, device=dev, mode=mode, layout="satmajor")
    except Exception as exc:
        if rank == 0:
            print(f"{mode} unavailable: {exc}")
        continue

    fg.overlap_ok = True          # measure the windowed pass at any world size (the library enables it for world <= 2)

    def b2b():
        M.propagate_batch_device(y0, tfd, ctrl, const, include_drag=False, include_J2=False, T=K, y=x, u_out=u)
        fg.discretize(x, u, tfd, const)
    t_b2b = timed(b2b)
    for nw in (0, 8):
        fg.buf[:42].zero_(); fg.buf[49:].zero_(); x.zero_(); torch.cuda.synchronize(); dist.barrier()
        t_ov = timed(lambda: fg.propagate_discretize(y0, tfd, ctrl, const, y=x, u_out=u, n_windows=nw))
        ok = int(torch.equal(fg.buf, full)) & int(torch.equal(x, y_all[s0:s1])) & int(int(fg.status.max()) == 0)
        ok_all &= ok
        lines.append(f"   {mode} (stagger {fg.stagger}), windows {nw or 'default'}: back to back {t_b2b:.3f} ms | overlapped {t_ov:.3f} ms | identical to single-rank = {ok}")
    del fg
res = torch.tensor([ok_all], device=dev); dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"world {world}  N/rank {N}  K {K}: propagate + discretize + fused all-gather + barrier, max over ranks; verified on every rank = {int(res[0])}")
    print("\n".join(lines))
dist.destroy_process_group()
