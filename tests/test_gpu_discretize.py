"""GPU parity tests of the discretization kernel, called through the C-ABI (ctypes)."""
import numpy as np
import pytest

from conftest import NAMES, rel_err, synth_batch

pytestmark = pytest.mark.gpu

# CUDA vs the reference's use_uniform_steps=True mode (golden fixtures): SURVEY 8c states <= 1e-8
TOL_REF = 1e-8
# CUDA vs the plain-C oracle (same RK4/trapezoid algorithm, different formulation: dense LU inverse,
# literal J2 Jacobian, 56-vector classical RK4): rounding-level agreement
TOL_ORACLE = 1e-10


@pytest.fixture(scope="module")
def M():
    import mpconstellation_b200 as m
    m._lib.require_gpu()
    return m


def _sel(o, ks):
    return o[ks] if o.ndim == 3 else o[:, ks]


@pytest.mark.parametrize("sc", ["d0", "d1", "d2", "d3", "d4"])
def test_discretizer_matches_reference_fixtures(M, gold_disc, const, sc):
    """the reference's own test scenarios (test_discretizer.py:30-150) through the reference signature"""
    g = gold_disc
    d = M.Discretizer(const, use_scipy_ZOH=False, include_drag=False, include_J2=False)
    d.use_uniform_steps = True
    out = d.discretize(M.Simulator.satellite_dynamics, g[sc + "_x"], g[sc + "_u"], float(g[sc + "_tf"]))
    K = g[sc + "_x"].shape[1]
    assert [o.shape for o in out] == [(K - 1, 7, 7), (K - 1, 7, 3), (K - 1, 7, 3), (7, K - 1), (7, K - 1)]
    for n, o in zip(NAMES, out):
        assert rel_err(o, g[f"{sc}_uni_{n}"]) < TOL_REF, (sc, n)


def test_j2_and_node_count_match_reference_fixtures(M, gold_disc, const):
    g = gold_disc
    ks = g["d3_j2_ks"]
    d = M.Discretizer(const, include_J2=True)
    d.use_uniform_steps = True
    out = d.discretize(M.Simulator.satellite_dynamics, g["d3_x"], g["d3_u"], 2.0)
    for n, o in zip(NAMES, out):
        assert rel_err(_sel(o, ks), g[f"d3_j2_uni_{n}"]) < TOL_REF, n
    d = M.Discretizer(const)
    d.use_uniform_steps = True
    d.integrator_steps = 21
    out = d.discretize(M.Simulator.satellite_dynamics, g["d3_x"], g["d3_u"], 2.0)
    for n, o in zip(NAMES, out):
        assert rel_err(_sel(o, ks), g[f"d3_n21_uni_{n}"]) < TOL_REF, n


@pytest.mark.parametrize("sc", ["d0", "d1", "d3", "d4"])
def test_default_mode_matches_reference_default_fixtures(M, gold_disc, const, sc):
    """Discretizer as shipped (use_uniform_steps=False): the device replays scipy's RK45 step controller, so the
    matrices match the unmodified reference's default output (<= 1e-8; observed ~1e-10, the symplectic-inverse
    defect of the RK45 solution)"""
    g = gold_disc
    d = M.Discretizer(const)
    assert d.use_uniform_steps is False
    out = d.discretize(M.Simulator.satellite_dynamics, g[sc + "_x"], g[sc + "_u"], float(g[sc + "_tf"]))
    for n, o in zip(NAMES, out):
        assert rel_err(o, g[f"{sc}_def_{n}"]) < TOL_REF, (sc, n)
    if sc == "d3":
        dj = M.Discretizer(const, include_J2=True)
        out = dj.discretize(M.Simulator.satellite_dynamics, g["d3_x"], g["d3_u"], 2.0)
        for n, o in zip(NAMES, out):
            assert rel_err(_sel(o, g["d3_j2_ks"]), g[f"d3_j2_def_{n}"]) < TOL_REF, n


def test_default_mode_batch_matches_c_oracle_and_its_node_counts(M, const):
    from oracle import c_oracle as C
    _, x, u = synth_batch(48, 60, 1.5, const)
    tfv = 1.5 * (1 + 0.1 * np.arange(48) / 48)
    ref = C.discretize_batch_adaptive(x, u, tfv, const)
    res = M.discretize_batch(x, u, tfv, const, adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2))
    assert res.status.max() == 0 and ref[5].max() == 0
    assert np.array_equal(res.n_nodes, ref[6])                  # the same steps were accepted
    for n, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < 1e-8, n
    # tighter tolerances -> more nodes, and convergence towards the uniform-node answer
    fine = M.discretize_batch(x, u, tfv, const, adaptive=dict(rtol=1e-9, atol=1e-12, max_step=2e-4))
    uni = M.discretize_batch(x, u, tfv, const)
    assert fine.n_nodes.min() > res.n_nodes.max()
    for n, o, r in zip(NAMES, fine.stacked(), uni.stacked()):
        assert rel_err(o, r) < 1e-6, n


def test_u_on_its_own_grid_like_reference_test_linearize_many(M, gold_disc, const):
    """test_discretizer.py:88-117 passes u = np.tile(T_init, (3, K)), a (3, 3K) array: the reference's hold takes its
    grid from u itself.  Same call here, checked against the reference's output (fixture d2q) and, in default mode,
    against the numpy/scipy restatement."""
    from oracle import mpc_oracle as O
    g = gold_disc
    ks = [int(k) for k in g["d2q_ks"]]
    d = M.Discretizer(const)
    d.use_uniform_steps = True
    out = d.discretize(M.Simulator.satellite_dynamics, g["d2_x"], g["d2q_u"], 1.0)
    for n, o in zip(NAMES, out):
        # this u jumps between 0.44 / 0.7 / 1.0 three times inside every interval; the reference's RK45 (rtol 1e-3,
        # dense output at the 101 nodes) integrates across those kinks with ~1e-4 error of its own, the device RK4
        # with 100 steps resolves them: agreement is limited by the reference's error here (observed 2e-4)
        assert rel_err(_sel(o, ks), g[f"d2q_uni_{n}"]) < 1e-3, n
    d = M.Discretizer(const)
    out = d.discretize(M.Simulator.satellite_dynamics, g["d2_x"], g["d2q_u"], 1.0)
    ref = [O.interval_matrices(k, g["d2_x"], g["d2q_u"], 1.0, const) for k in ks]
    for i, n in enumerate(NAMES):
        got = _sel(out[i], ks)
        want = np.stack([r[i] for r in ref]) if i < 3 else np.column_stack([r[i] for r in ref])
        assert rel_err(got, want) < 1e-8, n


def test_scipy_zoh_flag_is_the_same_hold(M, gold_disc, const):
    """test_discretizer.py:152-157 (test_custom_ZOH)"""
    g = gold_disc
    outs = []
    for flag in (False, True):
        d = M.Discretizer(const, use_scipy_ZOH=flag)
        d.use_uniform_steps = True
        outs.append(d.discretize(M.Simulator.satellite_dynamics, g["d1_x"], g["d1_u"], 0.1))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("n_sats,K,tf,j2,n_sub", [(64, 100, 1.0, False, 100), (7, 33, 0.7, True, 100),
                                                  (5, 2, 0.05, False, 100), (3, 50, 2.0, True, 16),
                                                  (130, 3, 0.02, False, 7)])
def test_batch_matches_c_oracle(M, const, n_sats, K, tf, j2, n_sub):
    """config 2 (64 sats x K=100) and ragged shapes against the plain-C oracle on the same seeded inputs"""
    from oracle import c_oracle as C
    _, x, u = synth_batch(n_sats, K, tf, const)
    tfv = tf * (1 + 0.05 * np.arange(n_sats) / n_sats)          # per-satellite tf
    ref = C.discretize_batch(x, u, tfv, const, include_J2=j2, n_sub=n_sub)
    assert ref[5].max() == 0
    res = M.discretize_batch(x, u, tfv, const, include_J2=j2, n_sub=n_sub)
    assert res.status.shape == (n_sats, K - 1) and res.status.max() == 0
    for n, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert o.shape == r.shape
        assert rel_err(o, r) < TOL_ORACLE, n
        # per-interval too, so a single bad unit cannot hide behind the batch maximum
        axes = tuple(range(2, o.ndim)) if n in ("A_k", "B_kp", "B_kn") else (1,)
        num = np.max(np.abs(o - r), axis=axes)
        den = np.max(np.abs(r), axis=axes)
        assert np.all(num <= 1e-9 * np.maximum(den, 1e-300)), n


def test_structure_of_outputs(M, const):
    """size-independent properties: last row of A_k is e7, B_kp + B_kn = A_k int Phi^-1 B (lambda weights sum to 1),
    the symplectic identity on the 6x6 block, and thrust-free intervals give zero mass rows"""
    _, x, u = synth_batch(16, 40, 1.0, const)
    res = M.discretize_batch(x, u, 1.0, const)
    A, Bp, Bn, S, X = res.stacked()
    assert np.array_equal(A[..., 6, :6], np.zeros_like(A[..., 6, :6])) and np.all(A[..., 6, 6] == 1.0)
    J = np.block([[np.zeros((3, 3)), np.eye(3)], [-np.eye(3), np.zeros((3, 3))]])
    P6 = A[..., :6, :6]
    assert np.max(np.abs(np.swapaxes(P6, -1, -2) @ J @ P6 - J)) < 1e-9
    u0 = np.zeros_like(u)
    res0 = M.discretize_batch(x, u0, 1.0, const)
    A0, Bp0, Bn0, S0, X0 = res0.stacked()
    assert np.all(Bp0[..., 6, :] == 0) and np.all(Bn0[..., 6, :] == 0) and np.all(S0[:, 6, :] == 0)   # eps guard (:208)
    assert np.all(np.isfinite(res0.soa))


def test_rollout_self_consistency_full_size_sample(M, const):
    """known-answer property at the bench's K: the discrete model on its own reference trajectory reproduces
    the nonlinear propagation (test_discretizer.py:110-113)"""
    _, x, u = synth_batch(8, 200, 2.0, const)
    res = M.discretize_batch(x, u, 2.0, const)
    for s in range(8):
        A, Bp, Bn, S, X = res.sat(s)
        for k in range(0, 199, 9):
            pred = A[k] @ x[s, :, k] + Bn[k] @ u[s, :, k] + Bp[k] @ u[s, :, k + 1] + S[:, k] * 2.0 + X[:, k]
            assert np.max(np.abs(pred - x[s, :, k + 1])) < 1e-5


def test_mass_failure_is_flagged_and_raises_like_reference(M, const):
    _, x, u = synth_batch(2, 10, 1.0, const)
    x = x.copy()
    x[1, 6, 4] = -0.5
    with pytest.raises(Exception, match="INVALID SATELLITE MASS"):
        M.discretize_batch(x, u, 1.0, const)
    res = M.discretize_batch(x, u, 1.0, const, check=False)
    assert res.status[1, 4] == M._lib.ST_MASS and res.status.sum() == M._lib.ST_MASS
    x[0, 0, 2] = np.nan
    res = M.discretize_batch(x, u, 1.0, const, check=False)
    assert res.status[0, 2] == M._lib.ST_NONFINITE


def test_drag_is_rejected_by_the_library_like_the_reference(M, const):
    _, x, u = synth_batch(1, 5, 1.0, const)
    with pytest.raises(M._lib.MpcError) as e:
        M.discretize_batch(x, u, 1.0, const, include_drag=True)
    assert e.value.code == M._lib.E_UNSUPPORTED


def test_device_api_pitch_offset_and_multi_destination(M, const):
    """device-pointer entry points on torch tensors: two 'ranks' write disjoint column ranges of one
    gathered buffer, and the multi-destination store replicates bit-identically"""
    import torch
    _, x, u = synth_batch(6, 12, 1.0, const)
    K = 12
    ref = M.discretize_batch(x, u, 1.0, const).soa.copy()
    dev = torch.device("cuda:0")
    n = 6 * (K - 1)
    gathered = torch.full((105, n + 5), float("nan"), dtype=torch.float64, device=dev)
    tfv = torch.ones(3, dtype=torch.float64, device=dev)
    for r in range(2):
        xs = torch.from_numpy(x[3 * r:3 * r + 3]).to(dev).contiguous()
        us = torch.from_numpy(u[3 * r:3 * r + 3]).to(dev).contiguous()
        M.discretize_batch_device(xs, us, tfv, const, out=gathered, out_offset=r * 3 * (K - 1))
    torch.cuda.synchronize()
    got = gathered.cpu().numpy()
    assert np.array_equal(got[:, :n], ref) and np.all(np.isnan(got[:, n:]))
    xs = torch.from_numpy(x).to(dev).contiguous()
    us = torch.from_numpy(u).to(dev).contiguous()
    tfv = torch.ones(6, dtype=torch.float64, device=dev)
    outs = [torch.zeros((105, n), dtype=torch.float64, device=dev) for _ in range(4)]
    M.discretize_batch_device(xs, us, tfv, const, out=outs[0], extra_dst=outs[1:])
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])                       # every destination gets the same bits
    # (the 4-destination launch runs the one-thread kernel, `ref` -- a 66-interval single-destination batch -- the
    # thread-group kernel: equal up to the rounding of the latter's cross-lane sums)
    assert rel_err(outs[0].cpu().numpy(), ref) < 1e-13


def test_full_size_config3_properties(M, const):
    """BASELINE config 3 (4096 satellites x K=200 = 815,104 intervals): full-size run checked through
    size-independent properties + a random sample of intervals against the C oracle"""
    from oracle import c_oracle as C
    N, K = 4096, 200
    y0, _, _ = synth_batch(N, 2, 2.0, const)            # initial states only (cheap)
    res, x, u = M.propagate_discretize(y0, 2.0, M.ConstantTangentialThrustController(tangential_thrust=0.5), const, T=K)
    assert res.status.max() == 0 and np.all(np.isfinite(res.soa))
    A, Bp, Bn, S, X = res.stacked()
    assert np.all(A[..., 6, 6] == 1.0) and not np.any(A[..., 6, :6])
    J = np.block([[np.zeros((3, 3)), np.eye(3)], [-np.eye(3), np.zeros((3, 3))]])
    P6 = A[::37, ::11, :6, :6]
    assert np.max(np.abs(np.swapaxes(P6, -1, -2) @ J @ P6 - J)) < 1e-9
    rng = np.random.default_rng(7)
    sats = np.sort(rng.choice(N, 12, replace=False))
    ref = C.discretize_batch(np.ascontiguousarray(x[sats]), np.ascontiguousarray(u[sats]), 2.0, const)
    for n, o, r in zip(NAMES, (A[sats], Bp[sats], Bn[sats], S[sats], X[sats]), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, n


def test_copy_engine_push_gather_on_one_gpu(M, const):
    """mpc_discretize_batch_push with local buffers standing in for the peers: several chunks (the batch spans
    three waves of the kernel), column offset inside a wider gathered buffer, constant rows pre-filled by
    mpc_fill_const_rows and never pushed; every destination ends up bit-identical to a plain launch"""
    import ctypes
    import torch
    from mpconstellation_b200 import _lib, batch
    dev = torch.device("cuda:0")
    sm = torch.cuda.get_device_properties(dev).multi_processor_count
    K = 40
    N = (3 * sm * 9 * 32) // (K - 1) - 5                 # just under three one-wave chunks
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    tfv = torch.full((N,), 1.0, dtype=torch.float64, device=dev)
    x, u, _ = M.propagate_batch_device(torch.from_numpy(y0).to(dev), tfv, M.ConstantTangentialThrustController(tangential_thrust=0.5),
                                       const, include_drag=False, include_J2=False, T=K)
    ref, st = M.discretize_batch_device(x, u, tfv, const, n_sub=20)
    n, pad = N * (K - 1), 77
    bufs = [torch.full((105, n + 2 * pad), float("nan"), dtype=torch.float64, device=dev) for _ in range(3)]
    L = _lib.lib()
    stream = torch.cuda.current_stream(dev).cuda_stream
    for b in bufs:
        _lib.check(L.mpc_fill_const_rows(b.data_ptr(), b.shape[1], stream))
    status = torch.full((n,), -1, dtype=torch.int32, device=dev)
    p = _lib.make_params(const, False, False)
    arr = (ctypes.c_void_p * 3)(*[b.data_ptr() for b in bufs])
    for waves, copy_kernel in ((1, 0), (2, 0), (1, 1)):
        _lib.check(L.mpc_discretize_batch_push(batch._ctx(0), x.data_ptr(), u.data_ptr(), tfv.data_ptr(), ctypes.byref(p),
                                               N, K, 20, arr, 3, n + 2 * pad, pad, status.data_ptr(), waves, copy_kernel,
                                               stream))
        torch.cuda.synchronize()
        assert int(status.max()) == 0 and int(status.min()) == 0
        for b in bufs:
            assert torch.equal(b[:, pad:pad + n], ref)
            for edge in (b[:42, :pad], b[49:, :pad], b[:42, pad + n:], b[49:, pad + n:]):
                assert bool(torch.isnan(edge).all())      # nothing written outside the batch's columns
            assert bool((b[42:48] == 0).all()) and bool((b[48] == 1).all())
        for b in bufs[1:]:
            b[:42].fill_(float("nan"))
            b[49:].fill_(float("nan"))


@pytest.mark.parametrize("adaptive", [False, True])
def test_zero_and_mixed_thrust_and_single_step(M, const, adaptive):
    """|u| <= eps guard (linearize_discretize.py:208): coasting satellites, satellites whose thrust switches off inside
    the horizon (thrust / no-thrust intervals side by side in one warp), and n_sub = 1; against the C oracle"""
    from oracle import c_oracle as C
    _, x, u = synth_batch(9, 21, 0.6, const)
    u[0] = 0.0                                   # coasting
    u[3, :, 10:] = 0.0                           # thrust ends at node 10: interval 9 has u_k != 0, u_k+1 == 0
    u[5, :, ::2] = 0.0                           # every other node
    u[7] = 1e-18                                 # below eps but not zero
    if adaptive:
        ref = C.discretize_batch_adaptive(x, u, 0.6, const)
        res = M.discretize_batch(x, u, 0.6, const, adaptive=dict())
        assert np.array_equal(res.n_nodes, ref[6])
    else:
        # n_sub = 200: short steps, even panel count -> the two-node-step path of the kernel (and of the oracle)
        r2 = C.discretize_batch(x, u, 0.6, const, n_sub=200)
        g2 = M.discretize_batch(x, u, 0.6, const, n_sub=200)
        assert g2.status.max() == 0 and r2[5].max() == 0
        for n, o, r in zip(NAMES, g2.stacked(), r2[:5]):
            assert rel_err(o, r) < TOL_ORACLE, ("n_sub=200", n)
        assert not np.any(g2.stacked()[1][0, :, 6, :]) and not np.any(g2.stacked()[1][7, :, 6, :])
        # n_sub = 25: odd -> one step per node
        ref = C.discretize_batch(x, u, 0.6, const, n_sub=25)
        res = M.discretize_batch(x, u, 0.6, const, n_sub=25)
        # n_sub = 1 (two quadrature nodes): on a short horizon, where one RK4 step is accurate and the symplectic
        # inverse of the numerical Phi equals its dense inverse to rounding ("very coarse steps")
        r1 = C.discretize_batch(x, u, 0.004, const, n_sub=1)
        g1 = M.discretize_batch(x, u, 0.004, const, n_sub=1)
        for n, o, r in zip(NAMES, g1.stacked(), r1[:5]):
            assert rel_err(o, r) < TOL_ORACLE, ("n_sub=1", n)
    assert res.status.max() == 0 and ref[5].max() == 0
    for n, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, n
    Bp, Bn = res.stacked()[1], res.stacked()[2]
    assert not np.any(Bp[0, :, 6, :]) and not np.any(Bn[0, :, 6, :])      # coasting: no mass-flow sensitivity (row 6 of B)
    assert not np.any(Bp[7, :, 6, :])                                     # |u| <= eps: the guard zeroes that row


def test_two_node_steps_and_one_step_per_node_agree(M, const):
    """mpc_set_tuning(7) switches the two-node steps (Hermite midpoint) off: on the reference's kind of grid the two
    integrators agree to ~1e-12, and a batch mixing short and long intervals (per-satellite tf: some threads take the
    two-node path, some fall back) matches the oracle, which makes the same per-interval choice"""
    from oracle import c_oracle as C
    from mpconstellation_b200 import _lib
    _, x, u = synth_batch(40, 60, 0.6, const)
    a = M.discretize_batch(x, u, 0.6, const, n_sub=100).soa.copy()
    try:
        _lib.check(_lib.lib().mpc_set_tuning(7))
        b = M.discretize_batch(x, u, 0.6, const, n_sub=100).soa.copy()
    finally:
        _lib.check(_lib.lib().mpc_set_tuning(8))
    assert not np.array_equal(a, b) and rel_err(a, b) < 1e-11
    tfv = np.where(np.arange(40) % 2 == 0, 0.3, 3.0)           # 0.005- and 0.05-orbit intervals side by side
    ref = C.discretize_batch(x, u, tfv, const, n_sub=100)
    res = M.discretize_batch(x, u, tfv, const, n_sub=100)
    assert res.status.max() == 0
    for n, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, n


def test_bench_step_at_full_size_matches_the_unmodified_reference(M):
    """The benchmark's own step -- 4096 satellites x K=200 through propagate_discretize_device, bench.py's arguments --
    against the unmodified reference flown on 4 satellites of that constellation (tests/golden/bench_workload.npz,
    made by tests/golden/make_golden.py bench): own propagation + own discretization vs the reference's RK45 propagation
    + its uniform-node discretization, end to end."""
    import os
    import torch
    import bench
    from conftest import GOLDEN
    gb = np.load(os.path.join(GOLDEN, "bench_workload.npz"))
    N, K, tf = int(gb["n_sats"]), 200, float(gb["tf"])
    Y, const_m = bench.make_constellation(N)
    dev = torch.device("cuda:0")
    y0, tfd = torch.from_numpy(Y).to(dev), torch.full((N,), tf, dtype=torch.float64, device=dev)
    ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    out, x, u, sp, sd = M.propagate_discretize_device(y0, tfd, ctrl, const_m, K, n_sub_prop=M.batch.default_n_sub(K),
                                                      n_sub_disc=100)
    nn = torch.empty(N * (K - 1), dtype=torch.int32, device=dev)
    out_def, sd2 = M.discretize_batch_device(x, u, tfd, const_m, adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2), n_nodes=nn)
    torch.cuda.synchronize()
    assert int(sp.max()) == 0 and int(sd.max()) == 0 and int(sd2.max()) == 0
    ks = gb["ks"]
    rows = {"A_k": (0, 49, (7, 7)), "B_kp": (49, 21, (7, 3)), "B_kn": (70, 21, (7, 3)), "Sigma_k": (91, 7, None), "xi_k": (98, 7, None)}
    worst = 0.0
    for j, s in enumerate(gb["idx"]):
        assert rel_err(x[s].cpu().numpy(), gb[f"s{j}_x"]) < 1e-6        # north_star tolerance on propagated states
        assert rel_err(u[s].cpu().numpy(), gb[f"s{j}_u"]) < 1e-6
        for tag, soa in (("uni", out), ("def", out_def)):
            blk = soa[:, s * (K - 1):(s + 1) * (K - 1)].cpu().numpy()
            for n, (r0, nr, shp) in rows.items():
                got = blk[r0:r0 + nr][:, ks]
                got = got.T.reshape((len(ks),) + shp) if shp else got
                e = rel_err(got, gb[f"s{j}_{tag}_{n}"])
                worst = max(worst, e)
                assert e < TOL_REF, (int(s), tag, n, e)
    print(f"bench step vs reference: worst norm-relative error {worst:.2e}")


@pytest.mark.parametrize("n_sats,K,tf,j2,n_sub", [(1, 50, 0.5, False, 100), (25, 100, 1.0, False, 100), (5, 17, 2.0, True, 10),
                                                   (3, 9, 3.0, True, 7), (24, 100, 1.0, True, 100)])
def test_small_batches_use_the_thread_group_kernel_and_agree_with_the_one_thread_kernels(M, const, n_sats, K, tf, j2, n_sub):
    """BASELINE configs 1-2 and every per-satellite Discretizer.discretize call run the 8-lanes-per-interval kernel
    (mpc_set_tuning(23) switches it off): A_k to the last bit or two (the same steps; bit-identical where both take two-node
    steps), everything else to the rounding of its cross-lane sums; also against the C oracle."""
    import torch
    from oracle import c_oracle as C
    dev = torch.device("cuda:0")
    _, x, u = synth_batch(n_sats, K, tf, const)
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    tfd = torch.full((n_sats,), tf, dtype=torch.float64, device=dev)
    L = M._lib.lib()
    before = M.launch_count()
    a, sa = M.discretize_batch_device(xd, ud, tfd, const, include_J2=j2, n_sub=n_sub)
    try:
        assert L.mpc_set_tuning(23) == 0
        b, sb = M.discretize_batch_device(xd, ud, tfd, const, include_J2=j2, n_sub=n_sub)
    finally:
        L.mpc_set_tuning(24)
    torch.cuda.synchronize()
    assert M.launch_count() - before == 2 and int(sa.max()) == 0 and torch.equal(sa, sb)
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert rel_err(a[0:49], b[0:49]) < 1e-14
    for r0, r1 in ((49, 70), (70, 91), (91, 98), (98, 105)):
        assert rel_err(a[r0:r1], b[r0:r1]) < 1e-12
    ref = C.discretize_batch(x, u, tf, const, include_J2=j2, n_sub=n_sub)
    got = M.DiscretizedBatch(a, sa.cpu().numpy().reshape(n_sats, K - 1), n_sats, K).stacked()
    # (only at the reference's 100 steps per interval: on coarse steps the kernels' symplectic Phi^-1 and the oracle's dense
    # inverse of the numerical Phi differ by the integrator's own defect -- DESIGN.md, known distances)
    if n_sub == 100:
        for n, g, r in zip(NAMES, got, ref[:5]):
            assert rel_err(g, r) < 1e-10, n


@pytest.mark.parametrize("tag", ["c11", "c24"])
def test_coast_to_thrust_switch_matches_the_unmodified_reference(M, tag):
    """u exactly 0 on the coasting nodes: the reference's global-grid lookup of an interval's end nodes decides the
    |u| <= eps guard of B_func there (linearize_discretize.py:208,308-315); fixture from the unmodified reference, through
    Discretizer.discretize in both quadrature modes (small batch: thread-group kernel; and the one-thread kernel)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "discretize_coast.npz"))
    from oracle.mpc_oracle import OracleConstants
    c = OracleConstants(*g["const"])
    x, u, tf = g[tag + "_x"], g[tag + "_u"], float(g[tag + "_tf"])
    d = M.Discretizer(c)
    L = M._lib.lib()
    for mode, uniform in (("uni", True), ("def", False)):
        ref = [g[f"{tag}_{mode}_{n}"] for n in NAMES]
        d.use_uniform_steps = uniform
        for variant in (24, 23):
            try:
                L.mpc_set_tuning(variant)
                got = d.discretize(M.Simulator.satellite_dynamics, x, u, tf)
            finally:
                L.mpc_set_tuning(24)
            for n, a, r in zip(NAMES, got, ref):
                assert rel_err(a, r) < (1e-7 if uniform else 1e-10), (mode, variant, n)
            assert rel_err(got[1][:, 6], ref[1][:, 6]) < 1e-11 and rel_err(got[2][:, 6], ref[2][:, 6]) < 1e-11


def test_reference_test_linearize_many_and_config2_chain_on_the_gpu(M):
    """the unmodified reference's own test_linearize_many call in its DEFAULT mode (test_discretizer.py:96-105; matching u
    and the test's malformed (3, 3K) u), through Discretizer.discretize; and BASELINE config 2 -- all 64 satellites x K=100
    through one batched pass, 4 of them against the reference's own propagation + both quadrature modes"""
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import make_constellation
    from oracle.mpc_oracle import OracleConstants
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "discretize_many.npz"))
    c = OracleConstants(*g["const"])
    sel = lambda o, ks, n: o[ks] if n in ("A_k", "B_kp", "B_kn") else o[:, ks]
    d = M.Discretizer(c)
    assert d.use_uniform_steps is False                       # the reference's default
    f = M.Simulator.satellite_dynamics
    for tag, u in (("m0", g["m0_u"]), ("m0q", g["m0_uq"])):
        got = d.discretize(f, g["m0_x"], u, 1.0)
        for n, a in zip(NAMES, got):
            assert rel_err(sel(a, g["m0_ks"], n), g[f"{tag}_def_{n}"]) < 1e-10, (tag, n)
    Y, const = make_constellation(64)
    idx, ks = g["m1_idx"], g["m1_ks"]
    assert np.array_equal(Y[idx], g["m1_y0"])
    ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    res, x, u = M.propagate_discretize(Y, 1.0, ctrl, const, T=100)
    resd = M.discretize_batch(x, u, 1.0, const, adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2))
    for j, i in enumerate(idx):
        assert rel_err(x[i], g[f"m1_s{j}_x"]) < 1e-9 and rel_err(u[i], g[f"m1_s{j}_u"]) < 1e-9
        for n, a, b in zip(NAMES, res.sat(int(i)), resd.sat(int(i))):
            assert rel_err(sel(a, ks, n), g[f"m1_s{j}_uni_{n}"]) < 1e-8, (j, n)
            assert rel_err(sel(b, ks, n), g[f"m1_s{j}_def_{n}"]) < 1e-10, (j, n)


@pytest.mark.parametrize("n_sats,K,tf,j2,tol_a", [(1, 50, 0.5, False, 2e-11), (64, 100, 1.0, False, 2e-11), (300, 200, 2.0, True, 2e-11),
                                                  (8, 60, 2.0, False, 1e-10), (256, 60, 2.0, True, 1e-10)])
def test_21_node_form_of_the_101_node_sums_on_the_device(M, const, n_sats, K, tf, j2, tol_a):
    """The fixed-step kernels as launched -- 20 steps and the 21-node Euler-Maclaurin rule wherever an interval allows it
    (kEmW, csrc/discretize_kernel.cuh) -- against the same launch with every one of the 101 nodes evaluated
    (mpc_set_tuning(37)) and against the plain-C oracle's literal sums: the quadrature to 1e-12, A_k (the integrator at
    the step 5 h) to 2e-11.  One-thread kernels and the thread-group kernel (n_sats = 1, 8).  K = 60, tf = 2 (BASELINE
    config 5, 0.034-orbit intervals): the 51-node rule on 50 steps, A_k to 5e-11."""
    import torch
    from oracle import c_oracle as C
    dev = torch.device("cuda:0")
    _, x, u = synth_batch(n_sats, K, tf, const)
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    tfd = torch.full((n_sats,), tf, dtype=torch.float64, device=dev)
    L = M._lib.lib()
    a, sa = M.discretize_batch_device(xd, ud, tfd, const, include_J2=j2)
    try:
        assert L.mpc_set_tuning(37) == 0
        b, sb = M.discretize_batch_device(xd, ud, tfd, const, include_J2=j2)
    finally:
        L.mpc_set_tuning(38)
    torch.cuda.synchronize()
    assert int(sa.max()) == 0 and torch.equal(sa, sb) and not torch.equal(a, b)
    a, b = a.cpu().numpy(), b.cpu().numpy()
    assert rel_err(a[0:49], b[0:49]) < tol_a
    for r0, r1 in ((49, 70), (70, 91), (91, 98), (98, 105)):
        assert rel_err(a[r0:r1], b[r0:r1]) < tol_a / 20      # (the integrands at the nodes carry the integrator's error too)
    ref = C.discretize_batch(x, u, tf, const, include_J2=j2)
    got = M.DiscretizedBatch(a, sa.cpu().numpy().reshape(n_sats, K - 1), n_sats, K).stacked()
    for n, o, r in zip(NAMES, got, ref[:5]):
        assert rel_err(o, r) < tol_a, n


def test_21_node_form_falls_back_to_the_101_nodes_per_interval_on_the_device(M, const):
    """a thrust jump, a sign flip through zero and coast-to-thrust switches take the 101 nodes (bit-identical to the launch
    without the rule), the intervals around them do not; the batch is what a re-planned SequenceController leaves behind"""
    import torch
    dev = torch.device("cuda:0")
    n_sats, K, tf = 4, 23, 0.23
    _, x, u = synth_batch(n_sats, K, tf, const)
    u = u.copy()
    u[0, :, 8:] *= 1.5
    u[1, :, 5:9] = 0.0
    u[2, :, 12:] *= -1.0
    xd, ud = torch.from_numpy(x).to(dev), torch.from_numpy(u).to(dev)
    tfd = torch.full((n_sats,), tf, dtype=torch.float64, device=dev)
    L = M._lib.lib()
    for grp in (23, 24):                        # one thread per interval / the thread-group kernel
        try:
            assert L.mpc_set_tuning(grp) == 0
            a, sa = M.discretize_batch_device(xd, ud, tfd, const)
            assert L.mpc_set_tuning(37) == 0
            b, sb = M.discretize_batch_device(xd, ud, tfd, const)
        finally:
            L.mpc_set_tuning(38)
            L.mpc_set_tuning(24)
        same = (a == b).all(dim=0).cpu().numpy().reshape(n_sats, K - 1)
        expect = np.zeros((n_sats, K - 1), dtype=bool)
        expect[0, 7] = expect[1, 4] = expect[1, 8] = expect[2, 11] = True
        assert int(sa.max()) == 0 and np.array_equal(same, expect), grp
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 2e-11
