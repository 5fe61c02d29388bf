"""CPU: the N>1 host logic (sharding, chunked gather, gathered per-satellite views) under gloo, world_size 2.
The per-rank compute is stood in for by the plain-C oracle (tests may use oracle/); the exchange, layout and
view code is the product's (mpconstellation_b200/distributed.py)."""
import os
import socket

import numpy as np
import pytest

from conftest import NAMES, synth_batch
from mpconstellation_b200 import distributed as D


def test_shard_ranges_cover_and_pad():
    for n, w in [(4096, 8), (5025, 8), (7, 4), (3, 8), (0, 2), (9, 2)]:
        sizes = D.shard_sizes(n, w)
        assert sum(sizes) == n and len(sizes) == w
        edges = [D.shard_range(n, r, w) for r in range(w)]
        assert edges[0][0] == 0 and edges[-1][1] == n
        for r in range(w):
            assert edges[r][1] - edges[r][0] == sizes[r]
            if r:
                assert edges[r][0] == edges[r - 1][1]


def test_gathered_view_layouts_agree():
    n_sats, K, world = 5, 4, 2
    n = K - 1
    per = 3
    glob = np.arange(105 * n_sats * n, dtype=np.float64).reshape(105, n_sats * n)
    rankmajor = np.zeros((world, 105, per * n))
    for s in range(n_sats):
        r, ls = s // per, s % per
        rankmajor[r][:, ls * n:(ls + 1) * n] = glob[:, s * n:(s + 1) * n]
    vg = D.GatheredView(glob, n_sats, K, world, "global")
    vr = D.GatheredView(rankmajor, n_sats, K, world, "rank")
    for s in range(n_sats):
        for a, b in zip(vg.sat(s), vr.sat(s)):
            assert np.array_equal(a, b)
    A = vg.sat(3)[0]
    assert A.shape == (n, 7, 7) and A[1, 2, 5] == glob[2 * 7 + 5, 3 * n + 1]
    # k-major: column = k n_sats + s; per-satellite access is a strided view of the same buffer (no copy)
    kmaj = np.zeros_like(glob)
    for s in range(n_sats):
        for k in range(n):
            kmaj[:, k * n_sats + s] = glob[:, s * n + k]
    vk = D.GatheredView(kmaj, n_sats, K, world, "kmajor")
    for s in range(n_sats):
        for a, b in zip(vg.sat(s), vk.sat(s)):
            assert np.array_equal(a, b)
    assert np.shares_memory(vk.sat(2)[0], kmaj) and np.shares_memory(vk.sat(2)[3], kmaj)
    with pytest.raises(IndexError):
        vg.sat(5)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_sats, K, tf, tmp):
    import torch
    import torch.distributed as dist
    from oracle import c_oracle as C
    from oracle.mpc_oracle import OracleConstants
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "discretize.npz"))
    const = OracleConstants(*g["const"])
    _, x, u = synth_batch(n_sats, K, tf, const)
    s0, s1 = D.shard_range(n_sats, rank, world)
    per = (n_sats + world - 1) // world
    local = torch.zeros((105, per * (K - 1)), dtype=torch.float64)

    def produce(c0, c1):          # stand-in for the CUDA kernel launch on columns [c0, c1)
        lo, hi = c0 // (K - 1), -(-c1 // (K - 1))
        hi = min(hi, s1 - s0)
        if hi <= lo:
            return
        A, Bp, Bn, S, X, st = C.discretize_batch(x[s0 + lo:s0 + hi], u[s0 + lo:s0 + hi], tf, const, nthreads=2)
        assert st.max() == 0
        for i in range(hi - lo):
            blk = np.concatenate([A[i].reshape(K - 1, 49), Bp[i].reshape(K - 1, 21), Bn[i].reshape(K - 1, 21),
                                  S[i].T, X[i].T], axis=1).T
            local[:, (lo + i) * (K - 1):(lo + i + 1) * (K - 1)] = torch.from_numpy(np.ascontiguousarray(blk))

    chunks, bounds = D.nccl_gather_chunks(local, n_chunks=3, produce=produce)
    full = D.assemble_rank_major(chunks, bounds, world)
    if rank == 0:
        np.save(os.path.join(tmp, "gathered.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_chunked_gather_world2_matches_single_process(tmp_path, const):
    import torch.multiprocessing as mp
    from oracle import c_oracle as C
    n_sats, K, tf, world = 5, 7, 0.3, 2          # ragged: ranks hold 3 and 2 satellites
    mp.spawn(_worker, args=(world, _free_port(), n_sats, K, tf, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "gathered.npy")
    assert full.shape == (world, 105, 3 * (K - 1))
    _, x, u = synth_batch(n_sats, K, tf, const)
    ref = C.discretize_batch(x, u, tf, const)
    view = D.GatheredView(full, n_sats, K, world, "rank")
    for s in range(n_sats):
        for name, got, want in zip(NAMES, view.sat(s), (ref[0][s], ref[1][s], ref[2][s], ref[3][s], ref[4][s])):
            assert np.array_equal(got, want), (s, name)
    # the padded columns of the short rank stay zero
    assert not np.any(full[1][:, 2 * (K - 1):])


def _worker_kmajor(rank, world, port, n_sats, K, tf, tmp):
    """every rank discretizes its shard with the HOST BUILD of the kernel sources straight into its place of a k-major
    gathered buffer (DstTab.km_ntot / km_soff, as FusedGather._opts sets them), the buffers are summed over gloo (disjoint
    column sets: the stand-in for the peer stores of the fused gather)"""
    import torch
    import torch.distributed as dist
    import hostk
    from oracle.mpc_oracle import OracleConstants
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "discretize.npz"))
    const = OracleConstants(*g["const"])
    _, x, u = synth_batch(n_sats, K, tf, const)
    s0, s1 = D.shard_range(n_sats, rank, world)
    buf = np.zeros((105, n_sats * (K - 1)))
    hostk.discretize(x[s0:s1], u[s0:s1], tf, const, out=buf, pitch=n_sats * (K - 1), km_ntot=n_sats, km_soff=s0)
    t = torch.from_numpy(buf)
    dist.all_reduce(t)
    if rank == 0:
        np.save(os.path.join(tmp, "kmajor.npy"), t.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_k_major_gathered_layout_world2(tmp_path, const):
    import torch.multiprocessing as mp
    import hostk
    n_sats, K, tf, world = 5, 7, 0.3, 2          # ragged: ranks hold 3 and 2 satellites
    mp.spawn(_worker_kmajor, args=(world, _free_port(), n_sats, K, tf, str(tmp_path)), nprocs=world, join=True)
    full = np.load(tmp_path / "kmajor.npy")
    _, x, u = synth_batch(n_sats, K, tf, const)
    ref, _ = hostk.discretize(x, u, tf, const)                      # single process, satellite-major
    vk = D.GatheredView(full, n_sats, K, world, "kmajor")
    vs = D.GatheredView(ref, n_sats, K, world, "global")
    for s in range(n_sats):
        for name, got, want in zip(NAMES, vk.sat(s), vs.sat(s)):
            assert np.array_equal(got, want), (s, name)
