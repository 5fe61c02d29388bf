"""The CUDA kernel SOURCES compiled for the host (tests/hostk, test infrastructure) against the reference fixtures and
the oracles: arithmetic, thread -> (satellite, interval) indexing, windowed launches, status words and progress words of
the code nvcc compiles for sm_100a, checked in the GPU-less container.  The GPU parity tests proper are tests/test_gpu_*.py;
this file makes sure a kernel edit cannot silently change the mathematics between two GPU runs."""
import os

import numpy as np
import pytest

import hostk
from conftest import GOLDEN, NAMES, rel_err, synth_batch
from oracle import c_oracle as C
from oracle import mpc_oracle as O

TOL_REF = 1e-8        # vs the unmodified reference (same tolerance as the GPU tests)
TOL_ORACLE = 1e-10    # vs the plain-C oracle on the same inputs


def _sel(o, ks):
    return o[ks] if o.ndim == 3 else o[:, ks]


@pytest.fixture(scope="module", autouse=True)
def _built():
    hostk.build()


@pytest.mark.parametrize("pair", [True, False])
@pytest.mark.parametrize("sc", ["d0", "d1", "d2", "d3", "d4"])
def test_fixed_step_kernels_match_reference_fixtures(gold_disc, const, sc, pair):
    """discretize_pair_kernel (two-node steps) and discretize_kernel (one step per node), test_discretizer.py:30-150"""
    g = gold_disc
    x, u, tf = g[sc + "_x"][None], g[sc + "_u"][None], float(g[sc + "_tf"])
    soa, st = hostk.discretize(x, u, tf, const, pair=pair)
    assert st.max() == 0 and st.min() == 0 and np.all(np.isfinite(soa))
    for n, o in zip(NAMES, hostk.stacked(soa, 1, x.shape[2])):
        assert rel_err(o[0], g[f"{sc}_uni_{n}"]) < TOL_REF, (sc, n)


def test_j2_and_node_count_match_reference_fixtures(gold_disc, const):
    g = gold_disc
    ks = g["d3_j2_ks"]
    soa, st = hostk.discretize(g["d3_x"][None], g["d3_u"][None], 2.0, const, include_J2=True)
    for n, o in zip(NAMES, hostk.stacked(soa, 1, 200)):
        assert rel_err(_sel(o[0], ks), g[f"d3_j2_uni_{n}"]) < TOL_REF, n
    soa, st = hostk.discretize(g["d3_x"][None], g["d3_u"][None], 2.0, const, n_sub=20)
    for n, o in zip(NAMES, hostk.stacked(soa, 1, 200)):
        assert rel_err(_sel(o[0], ks), g[f"d3_n21_uni_{n}"]) < TOL_REF, n


@pytest.mark.parametrize("sc", ["d0", "d1", "d3", "d4"])
def test_adaptive_kernel_matches_reference_default_mode(gold_disc, const, sc):
    """discretize_adaptive_kernel replays scipy's RK45 controller: the reference's DEFAULT mode, same node counts"""
    g = gold_disc
    x, u, tf = g[sc + "_x"][None], g[sc + "_u"][None], float(g[sc + "_tf"])
    soa, st, nodes = hostk.discretize_adaptive(x, u, tf, const)
    assert st.max() == 0
    for n, o in zip(NAMES, hostk.stacked(soa, 1, x.shape[2])):
        assert rel_err(o[0], g[f"{sc}_def_{n}"]) < TOL_ORACLE, (sc, n)
    ref = C.discretize_batch_adaptive(x, u, tf, const)
    assert np.array_equal(nodes, ref[6].reshape(-1))
    if sc == "d3":
        ks = g["d3_j2_ks"]
        soa, st, _ = hostk.discretize_adaptive(x, u, tf, const, include_J2=True)
        for n, o in zip(NAMES, hostk.stacked(soa, 1, 200)):
            assert rel_err(_sel(o[0], ks), g[f"d3_j2_def_{n}"]) < TOL_ORACLE, n


@pytest.mark.parametrize("j2", [False, True])
def test_both_builds_of_the_default_mode_kernel_agree(const, j2):
    """discretize_default_kernel (shipped: Phi ping-pongs through the output buffer, every node evaluated once with its
    full trapezoid weight) against discretize_adaptive_kernel (round 1: both ends of every panel): the same steps, the
    same node counts, results equal up to rounding (the panel sums are associated differently, and the shipped build steps
    the Phi columns in the units of the step, through the squared tableau for the position rows)"""
    _, x, u = synth_batch(5, 23, 1.1, const)
    a, sa, na = hostk.discretize_adaptive(x, u, 1.1, const, include_J2=j2)
    b, sb, nb = hostk.discretize_adaptive(x, u, 1.1, const, include_J2=j2, v1=True)
    assert sa.max() == 0 and np.array_equal(sa, sb) and np.array_equal(na, nb)
    for r0, r1 in ((0, 49), (49, 70), (70, 91), (91, 98), (98, 105)):
        assert rel_err(a[r0:r1], b[r0:r1]) < 1e-13


@pytest.mark.parametrize("n_sats,K,tf,j2,n_sub", [(16, 60, 1.0, False, 100), (7, 33, 0.7, True, 100), (3, 50, 2.0, True, 16),
                                                  (40, 3, 0.02, False, 7), (3, 2, 0.05, True, 100)])
def test_batch_matches_c_oracle(const, n_sats, K, tf, j2, n_sub):
    """ragged shapes, J2, per-satellite tf, odd panel counts (per-thread fall back to one step per node)"""
    _, x, u = synth_batch(n_sats, K, tf, const)
    tfv = tf * (1 + 0.01 * np.arange(n_sats))
    soa, st = hostk.discretize(x, u, tfv, const, include_J2=j2, n_sub=n_sub)
    ref = C.discretize_batch(x, u, tfv, const, include_J2=j2, n_sub=n_sub)
    assert st.max() == 0 and ref[5].max() == 0
    for n, o, r in zip(NAMES, hostk.stacked(soa, n_sats, K), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, n
    A = hostk.stacked(soa, n_sats, K)[0]
    assert np.all(A[..., 6, 6] == 1.0) and not np.any(A[..., 6, :6])          # structural constants of Phi's last row


def test_windowed_launches_assemble_the_full_launch_bit_for_bit(const):
    """the (k0, kc) launch window of the overlapped pass (mpc_propagate_discretize): ragged windows, pitch / offset, and
    nothing written outside the batch"""
    N, K = 9, 41
    _, x, u = synth_batch(N, K, 0.8, const)
    full, st_full = hostk.discretize(x, u, 0.8, const)
    n_int = N * (K - 1)
    out = np.full((105, n_int + 11), np.nan)
    st = np.full(n_int, -1, dtype=np.int32)
    for k0, kc in ((0, 7), (7, 13), (20, 1), (21, 19)):
        hostk.discretize(x, u, 0.8, const, k0=k0, kc=kc, out=out, pitch=n_int + 11, offset=4, status=st)
    assert np.array_equal(out[:, 4:4 + n_int], full) and np.array_equal(st, st_full)
    assert np.isnan(out[:, :4]).all() and np.isnan(out[:, 4 + n_int:]).all()


@pytest.mark.parametrize("N,K,tf,j2,n_sub", [(3, 6, 0.3, False, 100), (2, 5, 0.5, True, 16), (2, 4, 2.0, False, 7),
                                             (1, 3, 3.0, True, 10), (1, 2, 0.1, False, 100)])
def test_thread_group_kernel_matches_the_one_thread_kernels(const, N, K, tf, j2, n_sub):
    """discretize_group_kernel (8 lanes per interval, the small-batch mapping north_star names; the lanes of a group run
    as 8 host threads meeting at every shuffle): same scheme per interval as the one-thread kernels -- two-node steps,
    odd panel counts and over-long steps (one step per node), J2 -- so A_k is bit-identical and the rest agrees to the
    rounding of the epilogue's cross-lane sums; the idle groups of a last partial warp store nothing; a non-positive
    mass is flagged on the interval it occurs in."""
    _, x, u = synth_batch(N, K, tf, const)
    a, sa = hostk.discretize_group(x, u, tf, const, include_J2=j2, n_sub=n_sub, extra_groups=3)
    b, sb = hostk.discretize(x, u, tf, const, include_J2=j2, n_sub=n_sub)
    assert sa.max() == 0 and np.array_equal(sa, sb)
    assert np.array_equal(a[0:49], b[0:49])
    for r0, r1 in ((49, 70), (70, 91), (91, 98), (98, 105)):
        assert rel_err(a[r0:r1], b[r0:r1]) < 1e-13
    if K > 3:
        x = x.copy()
        x[0, 6, 1] = -1.0
        a, sa = hostk.discretize_group(x, u, tf, const, include_J2=j2, n_sub=n_sub)
        b, sb = hostk.discretize(x, u, tf, const, include_J2=j2, n_sub=n_sub)
        assert np.array_equal(sa != 0, sb != 0) and sa[1] == 1


@pytest.mark.parametrize("tag", ["c11", "c24"])
def test_kernels_follow_the_reference_lookup_at_a_coast_to_thrust_switch(tag):
    """every discretization kernel takes the inputs of an interval's two END nodes the way the reference looks them up
    (ref_node_input; see test_oracles_reproduce_the_reference_on_a_coast_to_thrust_switch): with u exactly 0 on the
    coasting nodes the mass rows of B_kp / B_kn -- 1-2 % of those rows where a straight hold is used -- match the
    unmodified reference to rounding."""
    g = np.load(os.path.join(GOLDEN, "discretize_coast.npz"))
    c = O.OracleConstants(*g["const"])
    x, u, tf = g[tag + "_x"], g[tag + "_u"], float(g[tag + "_tf"])
    K = x.shape[1]
    runs = {"pair": ("uni", lambda: hostk.discretize(x[None], u[None], tf, c)[0]),
            "one-step": ("uni", lambda: hostk.discretize(x[None], u[None], tf, c, pair=False)[0]),
            "group": ("uni", lambda: hostk.discretize_group(x[None], u[None], tf, c)[0]),
            "default": ("def", lambda: hostk.discretize_adaptive(x[None], u[None], tf, c)[0]),
            "default-v1": ("def", lambda: hostk.discretize_adaptive(x[None], u[None], tf, c, v1=True)[0])}
    for name, (mode, run) in runs.items():
        ref = [g[f"{tag}_{mode}_{n}"] for n in NAMES]
        got = hostk.stacked(run(), 1, K)
        for n, a, r in zip(NAMES, got, ref):
            assert rel_err(a[0], r) < (1e-7 if mode == "uni" else 1e-12), (name, n)
        assert rel_err(got[1][0][:, 6], ref[1][:, 6]) < 1e-12 and rel_err(got[2][0][:, 6], ref[2][:, 6]) < 1e-12, name


def _sel_many(o, ks, n):
    return o[ks] if n in ("A_k", "B_kp", "B_kn") else o[:, ks]


def test_reference_test_linearize_many_and_config2_chain():
    """fixtures generated by the unmodified reference (make_golden.py many): its own test_linearize_many call in the DEFAULT
    mode (test_discretizer.py:96-105), with the matching u and with the test's malformed (3, 3K) u on its own grid; and the
    BASELINE config-2 chain (4 of 64 satellites, K=100): replayed propagation -> both quadrature modes"""
    g = np.load(os.path.join(GOLDEN, "discretize_many.npz"))
    c = O.OracleConstants(*g["const"])
    x, u, uq, ks = g["m0_x"], g["m0_u"], g["m0_uq"], g["m0_ks"]
    K = x.shape[1]
    got = hostk.stacked(hostk.discretize_adaptive(x[None], u[None], 1.0, c)[0], 1, K)
    gotq = hostk.stacked(hostk.discretize_ugrid(x[None], uq[None], 1.0, c, adaptive=dict())[0], 1, K)
    for n, a, q in zip(NAMES, got, gotq):
        assert rel_err(_sel_many(a[0], ks, n), g[f"m0_def_{n}"]) < 1e-12, n
        assert rel_err(_sel_many(q[0], ks, n), g[f"m0q_def_{n}"]) < 1e-12, n
    ks = g["m1_ks"]
    y, uu, st, _, _ = hostk.propagate_rk45(g["m1_y0"], float(g["m1_tf"]), c, kind=2, thrust=(0.5, 0, 0), T=100)
    assert st.max() == 0
    uni = hostk.stacked(hostk.discretize(y, uu, 1.0, c)[0], 4, 100)
    dfl = hostk.stacked(hostk.discretize_adaptive(y, uu, 1.0, c)[0], 4, 100)
    for j in range(4):
        assert rel_err(y[j], g[f"m1_s{j}_x"]) < TOL_RK45 and rel_err(uu[j], g[f"m1_s{j}_u"]) < TOL_RK45
        for n, a, d in zip(NAMES, uni, dfl):
            assert rel_err(_sel_many(a[j], ks, n), g[f"m1_s{j}_uni_{n}"]) < TOL_REF, (j, n)
            assert rel_err(_sel_many(d[j], ks, n), g[f"m1_s{j}_def_{n}"]) < 1e-11, (j, n)


def test_k_major_layout_is_a_permutation_of_the_satellite_major_one(const):
    """DstTab.km_ntot / km_soff: column = k n_tot + s_off + s.  Same arithmetic per interval, so the k-major result is the
    satellite-major one permuted, bit for bit -- full launch, ragged k-windows, both kernels, a rank's block inside a
    larger gathered buffer (margins untouched)."""
    N, K, ntot, soff = 6, 9, 11, 3
    _, x, u = synth_batch(N, K, 0.8, const)
    n = K - 1
    full, st_full = hostk.discretize(x, u, 0.8, const)
    want = np.full((105, ntot * n), np.nan)
    for s_ in range(N):
        for k in range(n):
            want[:, k * ntot + soff + s_] = full[:, s_ * n + k]
    for pair in (True, False):
        out = np.full((105, ntot * n), np.nan)
        st = np.full(N * n, -1, dtype=np.int32)
        windows = ((0, 3), (3, 4), (7, 1)) if pair else ((0, -1),)
        for k0, kc in windows:
            hostk.discretize(x, u, 0.8, const, pair=pair, k0=k0, kc=kc, out=out, pitch=ntot * n, status=st,
                             km_ntot=ntot, km_soff=soff)
        ok = ~np.isnan(want)
        if pair:
            assert np.array_equal(out[ok], want[ok]) and np.isnan(out[~ok]).all() and np.array_equal(st, st_full)
        else:       # the one-step-per-node kernel integrates differently (~1e-13): same layout, own arithmetic
            ref1, st1 = hostk.discretize(x, u, 0.8, const, pair=False)
            want1 = np.full((105, ntot * n), np.nan)
            for s_ in range(N):
                for k in range(n):
                    want1[:, k * ntot + soff + s_] = ref1[:, s_ * n + k]
            assert np.array_equal(out[ok], want1[ok]) and np.isnan(out[~ok]).all() and np.array_equal(st, st1)


@pytest.mark.parametrize("case", ["p0", "p1", "p2", "p4"])
def test_propagate_kernel_vs_reference_and_c_oracle(gold_prop, const, case):
    gp = gold_prop
    y0 = O.normalize_state(gp["x0_dim"], O.scale_factors(gp["x0_dim"]))
    kw = {"p0": dict(kind=0, T=500, include_drag=True, include_J2=True),
          "p1": dict(kind=2, thrust=(0.5, 0, 0), T=200),
          "p2": dict(kind=1, thrust=tuple(gp["p2_thrust"]), T=300, include_drag=True, include_J2=True),
          "p4": dict(kind=2, thrust=(0.1, 0, 0), T=200, include_drag=True, include_J2=True)}[case]
    T, tf = kw["T"], float(gp[case + "_tf"])
    n_sub = int(np.ceil(1000 / (T - 1)))
    y, u, st, _ = hostk.propagate(y0[None], tf, const, n_sub=n_sub, **kw)
    assert st[0] == 0
    assert rel_err(y[0], gp[case + "_y"]) < 1e-6          # north_star tolerance on propagated states
    ckind = {0: C.CTRL_ZERO, 1: C.CTRL_CONSTANT, 2: C.CTRL_TANGENTIAL}[kw["kind"]]
    yr, ur, sr = C.propagate_batch(y0[None], tf, const, ckind, kw.get("thrust", (0, 0, 0)), T=T, n_sub=n_sub,
                                   include_drag=kw.get("include_drag", False), include_J2=kw.get("include_J2", False))
    assert rel_err(y, yr) < 1e-11 and (kw["kind"] == 0 or rel_err(u, ur) < 1e-10)


def test_propagate_progress_words_sequence_controller_and_mass_failure(const):
    """what the overlapped pass relies on: every warp adds 1 to every window's word, also when some of its satellites
    have run out of mass (NaN-filled trajectories, status 1) and when the batch does not fill its last warp"""
    N, T, seg = 70, 50, 6
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    y0 = y0.copy()
    y0[[2, 40], 6] = 1e-4
    rng = np.random.default_rng(5)
    tab = 0.3 * rng.standard_normal((3, 9))
    y, u, st, prog = hostk.propagate(y0, 1.0, const, kind=3, table=tab, end_tau=1.0, T=T, n_sub=4, seg_len=seg)
    n_win = (T - 1 + seg - 1) // seg
    assert list(prog[:n_win]) == [3] * n_win and not prog[n_win:].any()          # ceil(70 / 32) warps
    assert sorted(np.nonzero(st)[0]) == [2, 40] and np.isnan(y[2]).any() and np.isfinite(np.delete(y, [2, 40], 0)).all()
    yr, ur, sr = C.propagate_batch(y0, 1.0, const, C.CTRL_SEQUENCE, table=tab, end_tau=1.0, include_drag=False,
                                   include_J2=False, T=T, n_sub=4)
    ok = np.delete(np.arange(N), [2, 40])
    assert np.array_equal(sr, st) and rel_err(y[ok], yr[ok]) < 1e-11 and rel_err(u[ok], ur[ok]) < 1e-10


# ---- propagate_rk45_kernel: the reference's solve_ivp call replayed (simulator.py:185-187) -------------------------------
_RK45 = {"p0": dict(kind=0, T=500, include_drag=True, include_J2=True), "p1": dict(kind=2, thrust=(0.5, 0, 0), T=200),
         "p2": dict(kind=1, T=300, include_drag=True, include_J2=True),
         "p4": dict(kind=2, thrust=(0.1, 0, 0), T=200, include_drag=True, include_J2=True),
         "p5": dict(kind=3, end_tau=0.75, T=120)}
TOL_RK45 = 2e-12     # the same algorithm step for step: rounding only (observed <= 4e-13 over 5 orbits)


@pytest.mark.parametrize("spec", [True, False])
@pytest.mark.parametrize("case", sorted(_RK45))
def test_rk45_propagate_kernel_reproduces_reference_trajectories(gold_prop, const, case, spec):
    """every propagation fixture of the unmodified reference, the thrust cut-off inside the run (p5) included, to
    rounding; 1000 steps like scipy (nfev 6002); the speculative-first-stage build is the same arithmetic"""
    gp = gold_prop
    y0 = O.normalize_state(gp["x0_dim"], O.scale_factors(gp["x0_dim"]))
    kw = dict(_RK45[case])
    if case == "p2":
        kw["thrust"] = tuple(gp["p2_thrust"])
    if case == "p5":
        kw["table"] = gp["p5_u_tab"]
    y, u, st, steps, _ = hostk.propagate_rk45(y0[None], float(gp[case + "_tf"]), const, spec=spec, **kw)
    assert st[0] == 0 and steps[0] == 1000
    assert rel_err(y[0], gp[case + "_y"]) < TOL_RK45
    if case in ("p1", "p5"):
        assert rel_err(u[0], gp[case + "_u"]) < TOL_RK45


def test_rk45_propagate_kernel_three_sats_lanes_per_warp_and_progress(gold_prop):
    gp = gold_prop
    c3 = O.OracleConstants(*gp["p3_const"])
    sf = O.scale_factors(gp["p3_y0_dim"][0])
    y0 = np.stack([O.normalize_state(y, sf) for y in gp["p3_y0_dim"]])
    ref = None
    for lpw in (32, 2, 1):
        y, u, st, steps, prog = hostk.propagate_rk45(y0, 5.0, c3, include_drag=True, include_J2=True, T=500, lpw=lpw, seg_len=37)
        assert st.max() == 0 and rel_err(y, gp["p3_y"]) < TOL_RK45
        n_win = (499 + 36) // 37
        assert list(prog[:n_win]) == [3] * n_win and not prog[n_win:].any()     # one count per satellite and window
        ref = y if ref is None else ref
        assert np.array_equal(y, ref)                                            # the mapping does not touch the arithmetic


def test_rk45_propagate_kernel_step_control_and_failures(const):
    """step-size control engaged (max_step large: rejections at the thrust cut-off) against the C restatement of
    scipy: same step counts, same trajectory; per-satellite tables and end_tau; a satellite that runs out of mass is
    flagged, NaN-filled from the sample it failed at, and still releases every window; a NaN state ends with the
    step-size underflow scipy would report."""
    N, T = 9, 64
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    rng = np.random.default_rng(11)
    tab = 0.5 * rng.standard_normal((N, 3, 5))
    et = 0.2 + 0.8 * rng.random(N)
    for ms in (1.0, 0.05, 1e-3):
        yr, ur, sr, steps, rej = C.propagate_batch_rk45(y0, 1.7, const, C.CTRL_SEQUENCE, table=tab, end_tau=et,
                                                         include_drag=False, include_J2=True, T=T, max_step=ms)
        for spec in (True, False):
            y, u, st, n, _ = hostk.propagate_rk45(y0, 1.7, const, kind=3, table=tab, end_tau=et, include_J2=True, T=T,
                                                  max_step=ms, spec=spec)
            assert st.max() == 0 and np.array_equal(n, steps + rej)
            # (free step-size control amplifies rounding: err is a cancelling sum, h follows err^-1/5, and the O(h) error at
            # the cut-off follows h -- scipy itself and its C restatement differ by 1e-10 there)
            tol = 1e-9 if ms == 1.0 else TOL_RK45
            assert rel_err(y, yr) < tol and rel_err(u, ur) < max(tol, 1e-11)
        assert ms < 1.0 or rej.sum() > 0
    yb = y0.copy()
    yb[4, 6] = 2e-3                       # runs out of mass part way
    yb[7, 1] = np.nan                     # poisoned state
    big = np.tile(np.array([[3.0], [0.0], [0.0]]), (1, 4))
    y, u, st, n, prog = hostk.propagate_rk45(yb, 1.0, const, kind=3, table=big, end_tau=1.0, T=T, seg_len=9)
    assert st[4] == 1 and st[7] == 3 and np.delete(st, [4, 7]).max() == 0
    assert np.isfinite(y[4, :, 0]).all() and np.isnan(y[4, :, -1]).all() and np.isnan(y[7]).all()
    assert np.isfinite(np.delete(y, [4, 7], 0)).all()
    n_win = (T - 1 + 8) // 9
    assert list(prog[:n_win]) == [N] * n_win
    _, _, sr, _, _ = C.propagate_batch_rk45(yb, 1.0, const, C.CTRL_SEQUENCE, table=big, end_tau=1.0, include_drag=False,
                                            include_J2=False, T=T)
    assert np.array_equal(sr != 0, st != 0)


def test_bench_workload_chain_vs_reference():
    """propagate_kernel -> discretize_pair_kernel / discretize_adaptive_kernel on 4 satellites of the benchmark's
    constellation against the unmodified reference's own propagation + discretization (bench_workload.npz)"""
    gb = np.load(os.path.join(GOLDEN, "bench_workload.npz"))
    cb = O.OracleConstants(*gb["const"])
    ks, tf = gb["ks"], float(gb["tf"])
    y, u, st, _, _ = hostk.propagate_rk45(gb["y0"], tf, cb, kind=2, thrust=(0.5, 0, 0), T=200)
    assert st.max() == 0
    for j in range(len(gb["idx"])):       # the replayed integrator: the reference's own trajectory to rounding
        assert rel_err(y[j], gb[f"s{j}_x"]) < TOL_RK45 and rel_err(u[j], gb[f"s{j}_u"]) < TOL_RK45
    soa, sd = hostk.discretize(y, u, tf, cb)
    soa_d, sd2, _ = hostk.discretize_adaptive(y, u, tf, cb)
    assert sd.max() == 0 and sd2.max() == 0
    for j in range(len(gb["idx"])):
        assert rel_err(y[j], gb[f"s{j}_x"]) < 1e-6 and rel_err(u[j], gb[f"s{j}_u"]) < 1e-6
        for tag, s in (("uni", soa), ("def", soa_d)):
            for n, o in zip(NAMES, hostk.stacked(s, 4, 200)):
                assert rel_err(_sel(o[j], ks), gb[f"s{j}_{tag}_{n}"]) < TOL_REF, (j, tag, n)


def test_zero_thrust_and_mass_failure_status(const):
    """coasting intervals (the |u| <= eps guard, linearize_discretize.py:208) and a non-positive mass"""
    N, K = 4, 9
    _, x, u = synth_batch(N, K, 0.4, const)
    u = u.copy()
    u[1] = 0.0
    u[2, :, 4:] = 0.0
    x = x.copy()
    x[3, 6, 5] = -1.0
    soa, st = hostk.discretize(x, u, 0.4, const)
    ref = C.discretize_batch(x, u, 0.4, const)
    st2 = st.reshape(N, K - 1)
    assert st2[:3].max() == 0 and st2[3, 5] == 1 and np.array_equal(st2 != 0, ref[5] != 0)
    for n, o, r in zip(NAMES, hostk.stacked(soa, N, K), ref[:5]):
        assert rel_err(o[:3], r[:3]) < TOL_ORACLE, n


# ------------------------------------------------------------------------------------------- drag branch (K1c)
def _oracle_const(vec):
    return O.OracleConstants(*vec)


@pytest.mark.parametrize("tag", ["g0", "g1"])
@pytest.mark.parametrize("uniform", [True, False])
def test_drag_kernels_match_reference_fixtures(tag, uniform):
    """discretize_drag_kernel / the DRAG variant of the adaptive kernel vs the reference's drag branch
    (linearize_discretize.py:160-169) run with const.CD, rho_func, drho_func supplied (tests/golden/make_golden.py drag)"""
    g = np.load(os.path.join(GOLDEN, "discretize_drag.npz"))
    cst = _oracle_const(g[tag + "_const"])
    drag = (float(g[tag + "_cd"]), float(g[tag + "_rho_n"]))
    x, u, tf, ks = g[tag + "_x"][None], g[tag + "_u"][None], float(g[tag + "_tf"]), g[tag + "_ks"]
    soa, st, _ = hostk.discretize_drag(x, u, tf, cst, drag, include_J2=bool(g[tag + "_j2"]),
                                       adaptive=None if uniform else {})
    assert st.max() == 0
    mode = "uni" if uniform else "def"
    for n, o in zip(NAMES, hostk.stacked(soa, 1, x.shape[2])):
        assert rel_err(_sel(o[0], ks), g[f"{tag}_{mode}_{n}"]) < TOL_REF, (tag, mode, n)


def _power_law(law):
    """rho_func / drho_func of the radial fixture (tests/golden/make_golden.py: power_law_density, simulator.py:110)"""
    a, b, r0_m, r_e_m, rho_scale = [float(v) for v in law]

    def rho(r):
        return a * (np.linalg.norm(r) * r0_m - r_e_m) ** b / rho_scale

    def drho(r):
        return a * b * (np.linalg.norm(r) * r0_m - r_e_m) ** (b - 1.0) * r0_m / rho_scale
    return rho, drho


@pytest.mark.parametrize("uniform", [True, False])
def test_drag_kernels_with_altitude_dependent_density_match_the_reference(uniform):
    """rho_func = the power law of simulator.py:110, drho_func = its derivative: the reference's Dr_aD
    (linearize_discretize.py:166) is then not zero and G + Dr_aD is not symmetric.  The host fits both callables with
    Chebyshev series over the radii of the batch (mpconstellation_b200.discretizer.fit_density) and the kernels evaluate
    them at every stage; fixture from the unmodified reference (make_golden.py drag_radial; radii 1.0 ... 1.2, density
    falling by four orders of magnitude along the trajectory)."""
    from mpconstellation_b200.discretizer import fit_density
    g = np.load(os.path.join(GOLDEN, "discretize_drag_radial.npz"))
    cst = _oracle_const(g["const"])
    rho, drho = _power_law(g["law"])
    assert rho(g["x"][0:3, 0]) == float(g["rho_at_x0"]) and drho(g["x"][0:3, 0]) == float(g["drho_at_x0"])
    x, u, tf, ks = g["x"][None], g["u"][None], float(g["tf"]), g["ks"]
    model = fit_density(rho, drho, np.moveaxis(x[:, 0:3, :], 1, 2).reshape(-1, 3))
    assert isinstance(model, dict) and len(model["rho_c"]) == 32 and len(model["drho_c"]) == 32
    soa, st, _ = hostk.discretize_drag(x, u, tf, cst, (float(g["cd"]), model), include_J2=True,
                                       adaptive=None if uniform else {})
    assert st.max() == 0
    mode = "uni" if uniform else "def"
    for n, o in zip(NAMES, hostk.stacked(soa, 1, x.shape[2])):
        assert rel_err(_sel(o[0], ks), g[f"{mode}_{n}"]) < TOL_REF, (mode, n)
    if not uniform:
        # without the gradient (drho_func = 0) the reference's own answer is 7e-3 away on A_k and xi_k: the term is what is
        # being tested; and a kernel given that model reproduces THAT answer
        assert rel_err(g["def_A_k"], g["def_nograd_A_k"]) > 1e-3
        model0 = fit_density(rho, lambda r: 0.0, np.moveaxis(x[:, 0:3, :], 1, 2).reshape(-1, 3))
        assert len(model0["drho_c"]) == 0
        soa0, st0, _ = hostk.discretize_drag(x, u, tf, cst, (float(g["cd"]), model0), include_J2=True, adaptive={})
        for n, o in zip(NAMES, hostk.stacked(soa0, 1, x.shape[2])):
            assert rel_err(_sel(o[0], ks), g[f"def_nograd_{n}"]) < TOL_REF, n


@pytest.mark.parametrize("j2", [False, True])
def test_drag_batch_matches_c_oracle(const, j2):
    import copy
    cst = copy.copy(const)
    cst.S = const.S * 3e3
    _, x, u = synth_batch(11, 13, 0.4, cst)
    tf = np.linspace(0.3, 0.5, 11)
    drag = (2.2, 4.0e4)
    ref = C.discretize_batch(x, u, tf, cst, include_J2=j2, n_sub=40, drag=drag)
    soa, st, _ = hostk.discretize_drag(x, u, tf, cst, drag, include_J2=j2, n_sub=40)
    assert st.max() == 0 and ref[5].max() == 0
    for n, o, r in zip(NAMES, hostk.stacked(soa, 11, 13), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, n
    ref = C.discretize_batch_adaptive(x, u, tf, cst, include_J2=j2, drag=drag)
    soa, st, nodes = hostk.discretize_drag(x, u, tf, cst, drag, include_J2=j2, adaptive={})
    assert st.max() == 0 and np.array_equal(nodes, ref[6].reshape(-1))
    for n, o, r in zip(NAMES, hostk.stacked(soa, 11, 13), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, n


# ------------------------------------------------------------------- constraint terms and the sparse dynamics Jacobian
CT_KEYS = {"rf_hat": (0, 3), "Vc": (3, 0), "DrVc": (4, 3), "DrVc_rbar": (7, 0), "Vt": (8, 0), "DrVt_DvVt": (9, 6),
           "DrVt_DvVt_bar": (15, 0), "Vr": (16, 0), "DrVr_DvVr": (17, 6), "DrVr_DvVr_bar": (23, 0), "Vn": (24, 0),
           "DrVn_DvVn": (25, 6), "DrVn_DvVn_bar": (31, 0)}


@pytest.mark.parametrize("tag", ["c0", "c1", "c2"])
def test_constraint_terms_kernel_vs_reference_fixtures(tag):
    """constraint_terms_kernel vs Optimizer.get_constraint_terms of the unmodified reference (optimizer.py:80-170):
    unit vectors bit for bit (NaN positions of the inverted ubar_hat mask included), terminal terms to 5e-13"""
    g = np.load(os.path.join(GOLDEN, "constraint_terms.npz"))
    rbar, ubar, fin = hostk.constraint_terms(g[tag + "_x"][None], g[tag + "_u"][None], float(g["MU"]))
    np.testing.assert_array_equal(rbar[0], g[f"{tag}_rbar_hat"])
    np.testing.assert_array_equal(ubar[0], g[f"{tag}_ubar_hat"])
    for key, (off, ln) in CT_KEYS.items():
        got = fin[0, off:off + ln] if ln else fin[0, off]
        ref = np.asarray(g[f"{tag}_{key}"])
        assert np.max(np.abs(got - ref)) <= 5e-13 * max(np.max(np.abs(ref)), 1.0), key


def test_dynamics_jacobian_kernel_reproduces_the_pyomo_rule(const):
    """dynamics_jacobian_kernel: J z - rhs equals the residual of dynamics_const_rule (optimizer.py:327-339) written
    out with the rule's own indexing on the matrices the discretization kernel produced"""
    N, K = 3, 7
    _, x, u = synth_batch(N, K, 0.7, const)
    soa, st = hostk.discretize(x, u, 0.7, const, n_sub=10)
    A, Bp, Bn, S, X = hostk.stacked(soa, N, K)
    values, indices, rhs = hostk.dynamics_jacobian(soa, N, K)
    assert np.all(np.diff(indices, axis=1) > 0) and indices.min() >= 0 and indices.max() == 17 * N * K
    rng = np.random.default_rng(3)
    xz, uz, nuz, tfz = rng.standard_normal((N, 7, K)), rng.standard_normal((N, 3, K)), rng.standard_normal((N, 7, K)), 0.9
    z = np.concatenate([xz.ravel(), uz.ravel(), nuz.ravel(), [tfz]])
    res = ((values * z[indices]).sum(axis=1) - rhs).reshape(N, 7, K - 1)
    for s in range(N):
        for k in range(K - 1):
            for i in range(7):
                want = xz[s, i, k + 1] - (A[s, k, i] @ xz[s, :, k] + Bn[s, k, i] @ uz[s, :, k] + Bp[s, k, i] @ uz[s, :, k + 1]
                                          + S[s, i, k] * tfz + X[s, i, k] + nuz[s, i, k])
                assert abs(res[s, i, k] - want) < 1e-12 * max(1.0, abs(want)), (s, i, k)


# ----------------------------------------------------------------------------- memory safety of the kernels' indexing
def _run_sanitized(code_or_path, is_path):
    import subprocess
    import sys
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    if not os.path.isabs(asan) or not os.path.exists(asan):
        pytest.skip("no AddressSanitizer runtime next to this gcc")
    hostk.build(sanitize=True)
    env = dict(os.environ, HOSTK_SANITIZE="1", LD_PRELOAD=asan, ASAN_OPTIONS="detect_leaks=0:halt_on_error=1",
               PYTHONPATH=os.pathsep.join([os.path.dirname(os.path.abspath(__file__)), os.path.dirname(GOLDEN.rstrip("/")) + "/.."]))
    cmd = [sys.executable, code_or_path] if is_path else [sys.executable, "-c", code_or_path]
    return subprocess.run(cmd, capture_output=True, text=True, env=env, timeout=600)


def test_kernel_sources_under_address_sanitizer():
    """compute-sanitizer is closed on the GPU pool; the kernels' indexing is memchecked here instead: every kernel source
    on ragged shapes, launch windows, progress words, exactly sized heap buffers, under ASan + UBSan
    (tests/hostk/sanitize_driver.py).  A positive control makes sure the detector is live."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = _run_sanitized(os.path.join(root, "tests", "hostk", "sanitize_driver.py"), True)
    assert res.returncode == 0 and "launches clean" in res.stdout, res.stdout[-2000:] + res.stderr[-4000:]
    assert "AddressSanitizer" not in res.stderr and "runtime error" not in res.stderr
    control = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, {os.path.join(root, 'tests')!r}); sys.path.insert(0, {root!r})\n"
        "import hostk\nfrom conftest import synth_batch\nfrom oracle.mpc_oracle import OracleConstants\n"
        f"const = OracleConstants(*np.load({os.path.join(GOLDEN, 'discretize.npz')!r})['const'])\n"
        "y0, x, u = synth_batch(3, 5, 0.3, const)\n"
        "hostk.discretize(x, u, 0.3, const, out=np.full((105, 11), np.nan), pitch=12)   # one column short\n")
    res = _run_sanitized(control, False)
    assert res.returncode != 0 and "heap-buffer-overflow" in res.stderr


# ------------------------------------------------------------------- what the matrices MEAN (independent of the reference's code)
@pytest.mark.parametrize("j2", [False, True])
def test_matrices_are_the_derivatives_of_the_nonlinear_flow(const, j2):
    """x_{k+1} = F(x_k, u_k, u_{k+1}, tf) is the nonlinear flow over one interval under the first-order
    hold; the discretization must be its linearization: A_k = dF/dx_k, B_kn = dF/du_k, B_kp = dF/du_{k+1},
    Sigma_k = dF/dtf, and xi_k closes the affine model at the reference point.  F and its central differences come from
    the plain-C propagation oracle (RK4, FOH table of two columns), the matrices from the kernel source.  Phi is integrated
    (agreement 1e-9); B-, B+, Sigma, xi are the reference's 101-node TRAPEZOID rule (linearize_discretize.py:77-80), whose
    O(h^2) error on the quadratic-in-tau position rows of B (lambda * (tau - tau_k): 1.7e-5 of the row) is the
    reference's own and is mirrored, so those agree with the true derivatives to ~4e-6 of the matrix norm."""
    N, K, tf = 2, 12, 0.6
    _, x, u = synth_batch(N, K, tf, const)
    u = u + 0.05 * np.random.default_rng(2).standard_normal(u.shape)      # a hold that really varies inside the interval
    soa, st = hostk.discretize(x, u, tf, const, include_J2=j2)
    A, Bp, Bn, S, X = hostk.stacked(soa, N, K)
    dtau = 1.0 / (K - 1)

    def flow(xk, uk, uk1, tf_):
        y, _, s = C.propagate_batch(xk[None], tf_ * dtau, const, C.CTRL_SEQUENCE, table=np.column_stack([uk, uk1]),
                                    end_tau=1.0, include_drag=False, include_J2=j2, T=2, n_sub=200)
        assert s[0] == 0
        return y[0, :, 1]

    def jac(f, z, h):
        cols = []
        for j in range(len(z)):
            e = np.zeros(len(z))
            e[j] = h
            cols.append((f(z + e) - f(z - e)) / (2 * h))
        return np.column_stack(cols)

    for s_, k in ((0, 0), (1, 5), (1, K - 2)):
        xk, uk, uk1 = x[s_, :, k], u[s_, :, k], u[s_, :, k + 1]
        x1 = flow(xk, uk, uk1, tf)
        fdA = jac(lambda z: flow(z, uk, uk1, tf), xk, 1e-6)
        fdBn = jac(lambda z: flow(xk, z, uk1, tf), uk, 1e-6)
        fdBp = jac(lambda z: flow(xk, uk, z, tf), uk1, 1e-6)
        fdS = (flow(xk, uk, uk1, tf + 1e-6) - flow(xk, uk, uk1, tf - 1e-6)) / 2e-6
        assert rel_err(A[s_, k], fdA) < 2e-8 and rel_err(Bn[s_, k], fdBn) < 2e-5 and rel_err(Bp[s_, k], fdBp) < 2e-5
        assert rel_err(S[s_, :, k], fdS) < 2e-5
        affine = A[s_, k] @ xk + Bn[s_, k] @ uk + Bp[s_, k] @ uk1 + S[s_, :, k] * tf + X[s_, :, k]
        assert rel_err(affine, x1) < 5e-6          # closes up to the same trapezoid error (1e-6 with this rough hold)


# ------------------------------------------------------------------------------------------- seeded differential sweep
@pytest.mark.parametrize("seed", range(16))
def test_seeded_sweep_against_c_oracle(const, seed):
    """the cases of tests/test_gpu_fuzz.py (random shapes, horizons, step counts, drag / J2 flags, controller laws) on the
    host build of the kernel sources: propagation, fixed-step and adaptive discretization against the plain-C oracle"""
    from test_gpu_fuzz import _case
    c = _case(seed)
    rng, n, K, tf = c["rng"], c["n"], c["K"], c["tf"]
    y0, _, _ = synth_batch(n, 2, 0.1, const, seed=seed)
    tfv = tf * (1 + 0.1 * rng.random(n))
    tab, et, cp = None, 1.0, (0.0, 0.0, 0.0)
    if c["kind"] == 1:
        cp = tuple(rng.uniform(-0.5, 0.5, 3))
    elif c["kind"] == 2:
        cp = (float(rng.uniform(0.1, 1.0)), 0.0, 0.0)
    elif c["kind"] == 3:
        tab = rng.uniform(-0.4, 0.4, (n, 3, int(rng.integers(2, 30))))
        et = float(rng.uniform(1.0, 2.5))
    ck = [C.CTRL_ZERO, C.CTRL_CONSTANT, C.CTRL_TANGENTIAL, C.CTRL_SEQUENCE][c["kind"]]
    T = max(K, 2)
    n_prop = int(rng.integers(1, 12))
    y, u, st, _ = hostk.propagate(y0, tfv, const, kind=c["kind"], thrust=cp, table=tab, end_tau=et, include_drag=c["drag"],
                                  include_J2=c["j2"], T=T, n_sub=n_prop)
    yr, ur, sr = C.propagate_batch(y0, tfv, const, ck, cp, table=tab, end_tau=et, include_drag=c["drag"],
                                   include_J2=c["j2"], T=T, n_sub=n_prop)
    assert st.max() == 0 and sr.max() == 0
    assert rel_err(y, yr) < 1e-11 and np.max(np.abs(u - ur)) < 1e-11
    ref = C.discretize_batch(yr, ur, tfv, const, include_J2=c["j2"], n_sub=c["n_sub"])
    soa, sd = hostk.discretize(yr, ur, tfv, const, include_J2=c["j2"], n_sub=c["n_sub"])
    assert sd.max() == 0 and ref[5].max() == 0
    for nm, o, r in zip(NAMES, hostk.stacked(soa, n, T), ref[:5]):
        assert rel_err(o, r) < 1e-10, (nm, rel_err(o, r))
    ref = C.discretize_batch_adaptive(yr, ur, tfv, const, include_J2=c["j2"])
    soa, sd, nodes = hostk.discretize_adaptive(yr, ur, tfv, const, include_J2=c["j2"])
    assert sd.max() == 0 and np.array_equal(nodes, ref[6].reshape(-1))
    for nm, o, r in zip(NAMES, hostk.stacked(soa, n, T), ref[:5]):
        assert rel_err(o, r) < 1e-10, (nm, rel_err(o, r))


# ------------------------------------------------------------------------------------- poisoned inputs must terminate
@pytest.mark.timeout(300)
def test_poisoned_inputs_terminate_and_are_flagged(const):
    """A kernel that never returns hangs a GPU.  Every kernel source on NaN / Inf / zero-radius / non-positive-mass /
    huge inputs and absurd tf: the step loops end (the adaptive controller caps its node count, a NaN error accepts the
    step), and the affected units carry a non-zero status (mass 1, non-finite 2, step-size failure 3)."""
    _, x, u = synth_batch(4, 6, 0.3, const)
    idx = np.arange(x.size).reshape(x.shape)
    poisons = {"nan": (np.where(idx % 17 == 3, np.nan, x), u), "inf": (np.where(idx % 19 == 5, np.inf, x), u),
               "r0": (x * np.array([0, 0, 0, 1, 1, 1, 1.0])[None, :, None], u),
               "m-": (x * np.array([1, 1, 1, 1, 1, 1, -1.0])[None, :, None], u),
               "unan": (x, u * np.nan), "uhuge": (x, u * 1e200)}
    for tag, (xm, um) in poisons.items():
        for tf in (0.3, np.nan, np.inf, 1e6):
            st = [hostk.discretize(xm, um, tf, const)[1], hostk.discretize_adaptive(xm, um, tf, const)[1],
                  hostk.discretize_drag(xm, um, tf, const, (2.2, 4e4), n_sub=6)[1],
                  hostk.discretize_drag(xm, um, tf, const, (2.2, 4e4), adaptive={})[1]]
            nodes = hostk.discretize_adaptive(xm, um, tf, const)[2]
            assert nodes.max() <= 4097
            for s in st:
                assert s.min() >= 0 and s.max() <= 3 and s.max() > 0, (tag, tf)
                if tag in ("r0", "m-", "unan", "uhuge") or not np.isfinite(tf):
                    assert s.min() > 0, (tag, tf)          # every unit of the batch is affected
    y0 = x[:, :, 0]
    for ym in (y0 * np.nan, y0 * np.array([1, 1, 1, 1, 1, 1, -1.0])):
        for tf in (0.3, np.nan, 1e9):
            _, _, st, prog = hostk.propagate(ym, tf, const, kind=2, thrust=(0.5, 0, 0), include_drag=True, include_J2=True,
                                             T=20, n_sub=3, seg_len=4)
            assert st.min() == 1 and list(prog[:5]) == [1] * 5          # flagged, and the progress words still complete


def test_u_on_its_own_grid_kernels(gold_disc, const):
    """the GENU builds (mpc_discretize_batch_ugrid): the reference's own test_linearize_many passes a (3, 3K) u
    (test_discretizer.py:103); same checks as tests/test_gpu_discretize.py::test_u_on_its_own_grid_like_reference...,
    plus: with u given on x's own grid the general hold is the plain one"""
    g = gold_disc
    ks = [int(k) for k in g["d2q_ks"]]
    x, uq = g["d2_x"][None], g["d2q_u"][None]
    soa, st, _ = hostk.discretize_ugrid(x, uq, 1.0, const)
    assert st.max() == 0
    for n, o in zip(NAMES, hostk.stacked(soa, 1, x.shape[2])):
        assert rel_err(_sel(o[0], ks), g[f"d2q_uni_{n}"]) < 1e-3, n          # limited by the reference's own RK45 error
    soa, st, _ = hostk.discretize_ugrid(x, uq, 1.0, const, adaptive={})
    ref = [O.interval_matrices(k, g["d2_x"], g["d2q_u"], 1.0, const) for k in ks]
    for i, (n, o) in enumerate(zip(NAMES, hostk.stacked(soa, 1, x.shape[2]))):
        want = np.stack([r[i] for r in ref]) if i < 3 else np.column_stack([r[i] for r in ref])
        assert rel_err(_sel(o[0], ks), want) < TOL_REF, n
    xs, us = g["d4_x"][None], g["d4_u"][None]
    a, _ = hostk.discretize(xs, us, 0.5, const, pair=False)
    b, _, _ = hostk.discretize_ugrid(xs, us, 0.5, const)
    assert rel_err(b, a) < 1e-12


def test_end_node_input_lookup_follows_the_reference_on_every_node_of_every_grid():
    """ref_node_input decides int(tau // dtau) (Python's float floor division, linearize_discretize.py:308-311) with one
    FMA instead of fmod and a division.  Against the reference's u_FOH restated literally, for every node of every grid of
    2..600 nodes (and a few large ones): the same interval (the value is compared with both candidates' interpolation),
    the same exact zeros (what the |u| <= eps guard of B_func, :208, sees) and the value to rounding."""
    rng = np.random.default_rng(5)

    def u_foh(tau, u):                       # linearize_discretize.py:304-315, verbatim arithmetic
        if tau == 1:
            return u[:, -1]
        K = u.shape[1]
        dtau = 1 / (K - 1)
        k = int(tau // dtau)
        tau_k = k / (K - 1)
        tau_kp1 = (k + 1) / (K - 1)
        lambda_kn = (tau_kp1 - tau) / (tau_kp1 - tau_k)
        lambda_kp = (tau - tau_k) / (tau_kp1 - tau_k)
        return lambda_kn * u[:, k] + lambda_kp * u[:, k + 1]

    lower = 0
    for Ku in list(range(2, 600)) + [1000, 2048, 4097]:
        us = rng.normal(size=(3, Ku))
        us[:, rng.random(Ku) < 0.3] = 0.0                    # coast arcs: exact zeros next to thrusting nodes
        got = hostk.ref_node_input(us)
        tau = np.linspace(0, 1, Ku)                           # :356
        ref = np.array([u_foh(t, us) for t in tau])
        assert np.array_equal(got == 0.0, ref == 0.0), Ku
        assert np.max(np.abs(got - ref)) <= 4e-16 * np.max(np.abs(us)), Ku
        lower += sum(int(t // (1 / (Ku - 1))) < i for i, t in enumerate(tau[:-1]))
    assert lower > 50000          # the lookup lands in the interval LEFT of the node in a third of the cases: exercised


# ---- the 101-node trapezoid sums from 21 nodes (Euler-Maclaurin form, kEmW in csrc/discretize_kernel.cuh) -----------------

def _em_rule():
    """T(h) = sum_n w_n T(n h) for n = 5, 10, 20, 25 with sum w n^(2r) = 1 (r = 0..3), collected node by node: the weight of
    node j = 5 i in units of 5 h, as exact fractions"""
    from fractions import Fraction as F
    ns = [5, 10, 20, 25]
    A = [[F(n) ** (2 * r) for n in ns] for r in range(4)]
    b = [F(1)] * 4
    for c in range(4):
        for r in range(4):
            if r != c:
                f = A[r][c] / A[c][c]
                A[r] = [a - f * q for a, q in zip(A[r], A[c])]
                b[r] -= f * b[c]
    w = [b[i] / A[i][i] for i in range(4)]
    W = [F(0)] * 21
    for wn, n in zip(w, ns):
        p = 100 // n
        for j in range(p + 1):
            W[j * n // 5] += wn * n * (F(1, 2) if j in (0, p) else 1) / 5
    return w, W


def test_the_21_node_rule_is_the_euler_maclaurin_combination_of_four_trapezoid_sums():
    """the literals of kEmW are the exact rational weights; they are positive, symmetric and add up to the 20 coarse panels"""
    import re
    from fractions import Fraction as F
    w, W = _em_rule()
    assert [str(q) for q in w] == ["114114/78125", "-7904/15625", "4576/78125", "-209/15625"]
    src = open(os.path.join(os.path.dirname(hostk.CSRC), "csrc", "discretize_kernel.cuh")).read() if hasattr(hostk, "CSRC") else \
        open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "mpconstellation_b200", "csrc",
                          "discretize_kernel.cuh")).read()
    body = re.search(r"kEmW\[21\] = \{(.*?)\};", src, re.S).group(1)
    lits = [F(int(float(a)), int(b)) for a, b in re.findall(r"([0-9.]+) / ([0-9]+)", body)]
    assert lits == W and sum(W) == 20 and all(q > 0 for q in W) and W == W[::-1]
    # the rule returns the 101-node trapezoid sum -- not the integral -- exactly for polynomials through degree 7 and to
    # rounding for functions that vary across the interval the way the integrands do (0.06 ... 0.2 rad); the remainder
    # grows like the eighth power of the variation (1e-11 at 2 rad)
    t = np.linspace(0.0, 1.0, 101)

    def both(g):
        return 0.01 * (g[0] / 2 + g[-1] / 2 + g[1:-1].sum()), 0.05 * sum(float(W[i]) * g[5 * i] for i in range(21))

    for g in (t ** 7 - 0.3 * t ** 5 + t, np.exp(0.05 * t) * np.cos(0.2 * t + 0.4), 1.0 / (1.0 + 0.02 * t) ** 2):
        full, em = both(g)
        assert abs(em - full) <= 4e-16 * abs(full)
        assert abs(full - (0.01 * g[:-1].sum() + 0.005 * (g[-1] - g[0]))) < 1e-15      # (it IS the trapezoid sum)
    full, em = both(np.cos(2.0 * t + 0.4))
    assert 1e-13 < abs(em - full) < 1e-10


@pytest.mark.parametrize("j2", [False, True])
def test_21_node_form_of_the_fixed_step_kernels_reproduces_the_101_node_sums(const, j2):
    """discretize_pair_kernel / discretize_group_kernel as launched (em: 20 steps + the 21-node rule where the interval
    allows it) against the same kernels with every node evaluated and against the plain-C oracle's literal 101-node sums:
    the quadrature to 1e-12, A_k (the integrator at step 5 h) to 2e-11"""
    for n_sats, K, tf in ((5, 23, 0.23), (3, 50, 0.5)):
        _, x, u = synth_batch(n_sats, K, tf, const)
        a, sa = hostk.discretize(x, u, tf, const, include_J2=j2)
        b, sb = hostk.discretize(x, u, tf, const, include_J2=j2, em=False)
        g, sg = hostk.discretize_group(x, u, tf, const, include_J2=j2)
        assert sa.max() == 0 and sb.max() == 0 and sg.max() == 0 and not np.array_equal(a, b)
        assert rel_err(a[0:49], b[0:49]) < 2e-11
        for r0, r1 in ((49, 70), (70, 91), (91, 98), (98, 105)):
            assert rel_err(a[r0:r1], b[r0:r1]) < 1e-12, (r0, K)
        assert rel_err(a, g) < 1e-13 and rel_err(a[0:49], g[0:49]) < 1e-14      # the thread-group kernel takes the same form
        ref = C.discretize_batch(x, u, tf, const, include_J2=j2)
        for n, o in zip(NAMES, hostk.stacked(a, n_sats, K)):
            assert rel_err(o, ref[NAMES.index(n)]) < 2e-11, n


def test_21_node_form_is_taken_only_where_the_integrands_are_smooth(const):
    """per interval: a held input that changes by more than a quarter of its size between the two nodes (a thrust jump, a
    sign flip through zero, coast-to-thrust) and intervals too long for either rule take the 101 nodes -- bit-identical to
    the launch without the rules; coast arcs (u = 0 at both nodes) and slowly varying thrust take them"""
    n_sats, K, tf = 4, 23, 0.23
    _, x, u = synth_batch(n_sats, K, tf, const)
    u = u.copy()
    u[0, :, 8:] *= 1.5                    # jump by 50 % between nodes 7 and 8 of satellite 0
    u[1, :, 5:9] = 0.0                    # coast arc: intervals 5..7 are all-zero, 4 and 8 switch
    u[2, :, 12:] *= -1.0                  # the hold passes through zero inside interval 11
    u[3, :, 10] *= 1.2                    # 20 % up and down again: smooth enough on both sides
    a, sa = hostk.discretize(x, u, tf, const)
    b, sb = hostk.discretize(x, u, tf, const, em=False)
    assert sa.max() == 0 and sb.max() == 0
    same = np.all(a == b, axis=0).reshape(n_sats, K - 1)
    expect = np.zeros((n_sats, K - 1), dtype=bool)          # True: the 101 nodes were taken
    expect[0, 7] = expect[1, 4] = expect[1, 8] = expect[2, 11] = True
    assert np.array_equal(same, expect)
    assert rel_err(a, b) < 2e-11
    ref = C.discretize_batch(x, u, tf, const)
    for n, o in zip(NAMES, hostk.stacked(a, n_sats, K)):
        assert rel_err(o, ref[NAMES.index(n)]) < 2e-11, n
    # longer intervals (0.033 orbit, BASELINE config 5): the step 5 h would cost 1e-9; they take 50 steps and the 51-node
    # rule (kEmW2), the thread-group kernel too
    _, x2, u2 = synth_batch(2, 16, 0.5, const)
    a2, _ = hostk.discretize(x2, u2, 0.5, const)
    b2, _ = hostk.discretize(x2, u2, 0.5, const, em=False)
    g2, _ = hostk.discretize_group(x2, u2, 0.5, const)
    ref2 = C.discretize_batch(x2, u2, 0.5, const)
    assert not np.any(np.all(a2 == b2, axis=0)) and rel_err(a2[0:49], b2[0:49]) < 5e-11 and rel_err(a2[49:], b2[49:]) < 2e-12
    assert rel_err(a2, g2) < 1e-13
    for n, o in zip(NAMES, hostk.stacked(a2, 2, 16)):
        assert rel_err(o, ref2[NAMES.index(n)]) < 5e-11, n
    # still longer ones (0.06 orbit): all 101 nodes, one step per node
    _, x4, u4 = synth_batch(2, 9, 0.5, const)
    a4, _ = hostk.discretize(x4, u4, 0.5, const)
    b4, _ = hostk.discretize(x4, u4, 0.5, const, em=False)
    assert np.array_equal(a4, b4)
    # other node counts than the reference's 101: untouched
    a3, _ = hostk.discretize(x, u, tf, const, n_sub=50)
    b3, _ = hostk.discretize(x, u, tf, const, n_sub=50, em=False)
    assert np.array_equal(a3, b3)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_21_node_form_on_random_thrust_tables(const, seed):
    """seeded sweep: inputs whose direction and size change from node to node by random amounts (0 ... 40 % of their size:
    part of the intervals take the rule right at its smoothness bound, part fall back), J2 on and off, several grids -- the
    launch with the rule never differs from the launch that evaluates every node by more than the integrator's 2e-11
    (quadrature: 2e-12)"""
    rng = np.random.default_rng(seed)
    n_sats, K, tf = 6, int(rng.integers(12, 40)), None
    tf = 0.0101 * (K - 1) * float(rng.uniform(0.5, 1.05))          # intervals of 0.005 ... 0.0106 orbit
    _, x, u = synth_batch(n_sats, K, tf, const)
    u = u.copy()
    for s in range(n_sats):
        for k in range(1, K):
            d = rng.normal(size=3)
            d *= rng.uniform(0.0, 0.4) * np.linalg.norm(u[s, :, k - 1]) / np.linalg.norm(d)
            u[s, :, k] = u[s, :, k - 1] + d
    j2 = bool(seed % 2)
    a, sa = hostk.discretize(x, u, tf, const, include_J2=j2)
    b, sb = hostk.discretize(x, u, tf, const, include_J2=j2, em=False)
    assert sa.max() == 0 and sb.max() == 0
    took = ~np.all(a == b, axis=0)
    assert 0.2 < took.mean() < 0.98                                    # both branches exercised
    assert rel_err(a[0:49], b[0:49]) < 2e-11
    for r0, r1 in ((49, 70), (70, 91), (91, 98), (98, 105)):
        assert rel_err(a[r0:r1], b[r0:r1]) < 2e-12, (r0, seed)


def test_21_and_51_node_forms_in_the_drag_kernel(const):
    """discretize_drag_kernel (classical RK4 on the first-order system, dense 6x6 solve at the nodes) takes the same two
    rules under the same per-interval conditions: against the launch that evaluates every node and against the oracle"""
    drag = (2.5, 1.0e4 * 9.983e-13 / 3.6806e-17 * 3.6806e-17)          # const.CD, a density large enough to matter
    for n_sats, K, tf, tol in ((3, 23, 0.23, 2e-11), (2, 16, 0.5, 1e-10)):
        _, x, u = synth_batch(n_sats, K, tf, const)
        a, sa, _ = hostk.discretize_drag(x, u, tf, const, drag, include_J2=True)
        b, sb, _ = hostk.discretize_drag(x, u, tf, const, drag, include_J2=True, em=False)
        assert sa.max() == 0 and sb.max() == 0 and not np.array_equal(a, b)
        assert rel_err(a[0:49], b[0:49]) < tol and rel_err(a[49:], b[49:]) < tol / 10
        ref = C.discretize_batch(x, u, tf, const, include_J2=True, drag=drag)
        for n, o in zip(NAMES, hostk.stacked(a, n_sats, K)):
            assert rel_err(o, ref[NAMES.index(n)]) < tol, n
