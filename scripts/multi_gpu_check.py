"""torchrun script: validates the fused (peer-store) all-gather and the NCCL baseline against a single-rank
computation, and times both.   torchrun --nproc-per-node N scripts/multi_gpu_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
import mpconstellation_b200 as M
from mpconstellation_b200 import distributed as D
from bench import make_constellation

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
N, K, tf = int(os.environ.get("N", 1024)), int(os.environ.get("K", 200)), 2.0
Y, const = make_constellation(N * world)
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
tfd_all = torch.full((N * world,), tf, dtype=torch.float64, device=dev)
y_all, u_all, _ = M.propagate_batch_device(torch.from_numpy(Y).to(dev), tfd_all, ctrl, const, include_drag=False, include_J2=False, T=K)
s0, s1 = D.shard_range(N * world, rank, world)
x, u, tfd = y_all[s0:s1].contiguous(), u_all[s0:s1].contiguous(), tfd_all[s0:s1].contiguous()
n_int = N * (K - 1)

def timed(fn, reps=5):
    ts = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    t = torch.tensor([min(ts[1:])], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t[0])

# reference: every rank computes everything locally (small N) for verification
full, _ = M.discretize_batch_device(y_all, u_all, tfd_all, const)
torch.cuda.synchronize()
t_local = timed(lambda: M.discretize_batch_device(x, u, tfd, const))

fg = D.FusedGather(N * world, K, device=dev, skip_const=False, layout="satmajor")
fg.buf.zero_(); torch.cuda.synchronize(); dist.barrier()
t_fused = timed(lambda: fg.discretize(x, u, tfd, const))
variants = {}
ok_var = 1
for mode in ("unicast", "multicast"):
    for stag in (0, 4, 8, 16):
        try:
            fv = D.FusedGather(N * world, K, device=dev, mode=mode, skip_const=True, stagger=stag, layout="satmajor")
        except Exception:
            continue
        fv.buf[:42].zero_(); fv.buf[49:].zero_(); torch.cuda.synchronize(); dist.barrier()
        variants[f"{mode}+skip_const+stagger{stag}"] = timed(lambda: fv.discretize(x, u, tfd, const))
        ok_var &= int(torch.equal(fv.buf, full))
        del fv
ok_fused = torch.equal(fg.buf, full)
v = fg.view()
A = v.sat(N * world - 1)[0]
ok_view = np.array_equal(A, full[:49, -(K - 1):].T.reshape(K - 1, 7, 7).cpu().numpy())

t_mc, ok_mc = float("nan"), 1
try:
    fm = D.FusedGather(N * world, K, device=dev, mode="multicast", skip_const=False, layout="satmajor")
    fm.buf.zero_(); torch.cuda.synchronize(); dist.barrier()
    t_mc = timed(lambda: fm.discretize(x, u, tfd, const))
    ok_mc = int(torch.equal(fm.buf, full))
except Exception as exc:
    if rank == 0: print("multicast unavailable:", exc)
t_push, ok_push = {}, 1
for cw in (1, 2, "k1", "k2"):
    fp = D.FusedGather(N * world, K, device=dev, mode="pushk" if isinstance(cw, str) else "push",
                       chunk_waves=int(str(cw).lstrip("k")))
    fp.buf[:42].zero_(); fp.buf[49:].zero_(); torch.cuda.synchronize(); dist.barrier()
    t_push[cw] = timed(lambda: fp.discretize(x, u, tfd, const))
    ok_push &= int(torch.equal(fp.buf, full))
    del fp
loc = torch.empty((105, n_int), dtype=torch.float64, device=dev)
def nccl_step():
    def produce(c0, c1):
        a, b = c0 // (K - 1), c1 // (K - 1)
        M.discretize_batch_device(x[a:b], u[a:b], tfd[a:b], const, out=loc, out_offset=c0)
    nccl_step.res = D.nccl_gather_chunks(loc, 8, produce=produce, chunk_cols=((N + 7) // 8) * (K - 1))
t_nccl = timed(nccl_step)
chunks, bounds = nccl_step.res
rm = D.assemble_rank_major(chunks, bounds, world)
ok_nccl = all(torch.equal(rm[r], full[:, r * n_int:(r + 1) * n_int]) for r in range(world))
def nccl_plain():
    M.discretize_batch_device(x, u, tfd, const, out=loc)
    nccl_plain.res = torch.empty((world * 105, n_int), dtype=torch.float64, device=dev)
    dist.all_gather_into_tensor(nccl_plain.res, loc)
t_plain = timed(nccl_plain)
res = torch.tensor([int(ok_fused), int(ok_view), int(ok_nccl), int(ok_mc), int(ok_push), int(ok_var)], device=dev); dist.all_reduce(res, op=dist.ReduceOp.MIN)
if rank == 0:
    gb = (world - 1) * n_int * 840 / 1e9
    print(f"world {world}  N/rank {N}  K {K}: local-only {t_local:.3f} ms | fused peer-store gather {t_fused:.3f} ms | fused multicast-store gather {t_mc:.3f} ms | "
          f"copy-engine push (1 / 2 waves per chunk) {t_push[1]:.3f} / {t_push[2]:.3f} ms | copy-kernel push (1 / 2 waves) {t_push['k1']:.3f} / {t_push['k2']:.3f} ms | "
          f"NCCL chunked-overlap {t_nccl:.3f} ms | kernel then NCCL all-gather {t_plain:.3f} ms | "
          f"{gb:.2f} GB received per rank | verified fused/view/nccl/multicast/push/variants = {res.tolist()}")
    for k_, v_ in variants.items():
        print(f"   {k_}: {v_:.3f} ms")
dist.destroy_process_group()
