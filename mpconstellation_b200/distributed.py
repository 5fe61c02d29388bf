"""Multi-GPU: satellites sharded over ranks, one all-gather of the discretized matrices.

Every (satellite, interval) is independent given the reference trajectory (linearize_discretize.py:31-34),
so the only exchange step is the one north_star names: the SoA matrices of every shard are gathered to the
rank(s) driving the optimizer.  One process per GPU (torchrun); `torch.distributed` does the plumbing.

Layout after the gather: rank-major  G[r][row][col]  with  r in [0,world), row in [0,105),
col = local_sat * (K-1) + k  -- i.e. each rank's SoA block is kept contiguous (what NCCL all-gather
produces without an extra transpose).  `GatheredDiscretization.sat(s)` hands out the reference-shaped
views for a GLOBAL satellite index, so optimizer-side code indexes exactly as before (optimizer.py:327-339).
"""
import numpy as np

from . import _lib


def shard_sizes(n_sats, world):
    """Contiguous blocks of ceil(N/G) satellites; trailing ranks may be short or empty."""
    per = (n_sats + world - 1) // world if world > 0 else 0
    return [max(0, min(per, n_sats - r * per)) for r in range(world)]


def shard_range(n_sats, rank, world):
    per = (n_sats + world - 1) // world
    s0 = min(n_sats, rank * per)
    return s0, min(n_sats, s0 + per)


class GatheredDiscretization:
    """Rank-major gathered SoA blocks: array-like [world, 105, per*(K-1)] (numpy array or CPU/CUDA tensor)."""

    def __init__(self, gathered, n_sats, K, world):
        self.g, self.n_sats, self.K, self.world = gathered, n_sats, K, world
        self.per = (n_sats + world - 1) // world

    def locate(self, s):
        if not 0 <= s < self.n_sats:
            raise IndexError(s)
        return s // self.per, s % self.per

    def sat(self, s):
        """(A_k, B_kp, B_kn, Sigma_k, xi_k) of global satellite s, reference shapes/order
        (linearize_discretize.py:390); numpy views when the gathered buffer is a numpy array."""
        r, ls = self.locate(s)
        n = self.K - 1
        blk = self.g[r][:, ls * n:(ls + 1) * n]
        if not isinstance(blk, np.ndarray):
            blk = blk.cpu().numpy()
        A = blk[_lib.ROW_A:_lib.ROW_A + 49].T.reshape(n, 7, 7)
        Bp = blk[_lib.ROW_BP:_lib.ROW_BP + 21].T.reshape(n, 7, 3)
        Bn = blk[_lib.ROW_BN:_lib.ROW_BN + 21].T.reshape(n, 7, 3)
        return A, Bp, Bn, blk[_lib.ROW_SIGMA:_lib.ROW_SIGMA + 7], blk[_lib.ROW_XI:_lib.ROW_XI + 7]


def all_gather_soa(local, group=None, out=None):
    """All-gather equal-sized local SoA blocks [105, cols] -> [world, 105, cols] (NCCL for CUDA tensors,
    gloo for CPU tensors).  Ranks with fewer satellites pad their block to the common size first."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous(), group=group)
    return out


def discretize_sharded(x, u, tf, const, n_sats_total, include_J2=False, n_sub=100, chunks=4, group=None,
                       local_compute=None):
    """Discretize this rank's shard and all-gather the result.

    x [n_local,7,K], u [n_local,3,K], tf [n_local]: CUDA float64 tensors holding this rank's contiguous block of
    satellites (shard_range).  The shard is processed in `chunks` sub-blocks; the all-gather of chunk c runs on
    a side stream while chunk c+1 is being discretized.  Returns a list of GatheredDiscretization chunks'
    gathered tensors and a GatheredDiscretization-compatible accessor.

    `local_compute` lets a test substitute the per-chunk compute on a CPU-only box (gloo); the product path
    leaves it None and runs the CUDA kernel.
    """
    import torch
    import torch.distributed as dist
    from . import batch
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    per = (n_sats_total + world - 1) // world
    n_local = x.shape[0]
    K = x.shape[2]
    s0, s1 = shard_range(n_sats_total, rank, world)
    assert n_local == s1 - s0, "x must hold exactly this rank's shard"
    dev = x.device
    cuda = dev.type == "cuda"
    # pad the local block to `per` satellites so every rank contributes the same number of columns
    local = torch.zeros((_lib.MPC_OUT_ROWS, per * (K - 1)), dtype=torch.float64, device=dev)
    status = torch.zeros(per * (K - 1), dtype=torch.int32, device=dev)
    gathered = torch.empty((world, _lib.MPC_OUT_ROWS, per * (K - 1)), dtype=torch.float64, device=dev)
    csz = max(1, (per + chunks - 1) // chunks)
    side = torch.cuda.Stream(dev) if cuda else None
    views = []
    for c0 in range(0, per, csz):
        c1 = min(per, c0 + csz)
        l0, l1 = min(c0, n_local), min(c1, n_local)
        if l1 > l0:
            if local_compute is not None:
                local_compute(x[l0:l1], u[l0:l1], tf[l0:l1], local, l0 * (K - 1), status)
            else:
                batch.discretize_batch_device(x[l0:l1], u[l0:l1], tf[l0:l1], const, include_J2=include_J2, n_sub=n_sub,
                                              out=local, out_offset=l0 * (K - 1), status=status[l0 * (K - 1):l1 * (K - 1)])
        # gather the column range of this chunk from every rank into gathered[:, :, cols]
        cols = slice(c0 * (K - 1), c1 * (K - 1))
        piece = local[:, cols].contiguous()
        recv = torch.empty((world,) + tuple(piece.shape), dtype=torch.float64, device=dev)
        if cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(dev))
            side.wait_event(ev)
            with torch.cuda.stream(side):
                dist.all_gather_into_tensor(recv, piece, group=group)
                gathered[:, :, cols].copy_(recv)
            piece.record_stream(side)
            recv.record_stream(side)
        else:
            dist.all_gather_into_tensor(recv, piece, group=group)
            gathered[:, :, cols].copy_(recv)
        views.append(recv)
    if cuda:
        torch.cuda.current_stream(dev).wait_stream(side)
    return GatheredDiscretization(gathered, n_sats_total, K, world), status
