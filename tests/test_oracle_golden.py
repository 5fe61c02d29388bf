"""CPU: pins oracle/ against fixtures produced by the unmodified reference (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, NAMES, rel_err
from oracle import c_oracle as C
from oracle import mpc_oracle as O

# tolerance of the RK4-on-uniform-nodes restatement against the reference's use_uniform_steps=True mode
# (<= 1e-8 norm-relative; the reference's own RK45 noise floor is ~1e-10)
TOL_UNIFORM = 1e-8


def _stack(res):
    return (np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[2] for r in res]),
            np.column_stack([r[3] for r in res]), np.column_stack([r[4] for r in res]))


def _sel(o, ks):
    return o[ks] if o.ndim == 3 else o[:, ks]


@pytest.mark.parametrize("sc", ["d0", "d1", "d4"])
@pytest.mark.parametrize("uniform", [False, True])
def test_py_oracle_small_scenarios(gold_disc, const, sc, uniform):
    """numpy/scipy restatement == reference, both quadrature modes (same library calls: ~1e-15)."""
    g = gold_disc
    out = O.discretize(g[sc + "_x"], g[sc + "_u"], float(g[sc + "_tf"]), const, use_uniform_steps=uniform)
    tag = sc + ("_uni" if uniform else "_def")
    for n, o in zip(NAMES, out):
        assert rel_err(o, g[f"{tag}_{n}"]) < 1e-13, (tag, n)


@pytest.mark.parametrize("tag,kw", [("d3_def", {}), ("d3_uni", dict(use_uniform_steps=True)),
                                    ("d3_j2_uni", dict(use_uniform_steps=True, include_J2=True)),
                                    ("d3_j2_def", dict(include_J2=True)),
                                    ("d3_n21_uni", dict(use_uniform_steps=True, integrator_steps=21))])
def test_py_oracle_tangential_sampled(gold_disc, const, tag, kw):
    g = gold_disc
    ks = [int(k) for k in g["d3_j2_ks"]][::4]
    out = _stack([O.interval_matrices(k, g["d3_x"], g["d3_u"], 2.0, const, **kw) for k in ks])
    full = tag in ("d3_def", "d3_uni")
    for n, o in zip(NAMES, out):
        ref = g[f"{tag}_{n}"]
        ref = _sel(ref, ks) if full else _sel(ref, list(range(0, len(g["d3_j2_ks"]), 4)))
        assert rel_err(o, ref) < 1e-13, (tag, n)


def test_py_oracle_u_on_other_grid(gold_disc, const):
    """the reference test's own malformed u (3,3K): FOH runs on u's grid (test_discretizer.py:103)"""
    g = gold_disc
    out = _stack([O.interval_matrices(int(k), g["d2_x"], g["d2q_u"], 1.0, const, use_uniform_steps=True)
                  for k in g["d2q_ks"]])
    for n, o in zip(NAMES, out):
        assert rel_err(o, g[f"d2q_uni_{n}"]) < 1e-13


@pytest.mark.parametrize("sc", ["d0", "d1", "d2", "d3", "d4"])
def test_c_oracle_vs_reference_uniform(gold_disc, const, sc):
    g = gold_disc
    A, Bp, Bn, S, X, st = C.discretize_batch(g[sc + "_x"][None], g[sc + "_u"][None], float(g[sc + "_tf"]), const)
    assert st.max() == 0
    for n, o in zip(NAMES, (A[0], Bp[0], Bn[0], S[0], X[0])):
        assert rel_err(o, g[f"{sc}_uni_{n}"]) < TOL_UNIFORM, (sc, n)


def test_c_oracle_j2_and_other_node_count(gold_disc, const):
    g = gold_disc
    ks = g["d3_j2_ks"]
    out = C.discretize_batch(g["d3_x"][None], g["d3_u"][None], 2.0, const, include_J2=True)
    for n, o in zip(NAMES, out[:5]):
        assert rel_err(_sel(o[0], ks), g[f"d3_j2_uni_{n}"]) < TOL_UNIFORM
    out = C.discretize_batch(g["d3_x"][None], g["d3_u"][None], 2.0, const, n_sub=20)
    for n, o in zip(NAMES, out[:5]):
        assert rel_err(_sel(o[0], ks), g[f"d3_n21_uni_{n}"]) < TOL_UNIFORM


@pytest.mark.parametrize("sc", ["d0", "d1", "d3", "d4"])
def test_c_oracle_rk45_replica_matches_reference_default_mode(gold_disc, const, sc):
    """the restated scipy RK45 controller reproduces the reference's DEFAULT mode (adaptive quadrature nodes)
    to rounding: same accepted steps, same matrices"""
    g = gold_disc
    out = C.discretize_batch_adaptive(g[sc + "_x"][None], g[sc + "_u"][None], float(g[sc + "_tf"]), const)
    assert out[5].max() == 0
    for n, o in zip(NAMES, out[:5]):
        assert rel_err(o[0], g[f"{sc}_def_{n}"]) < 1e-12, (sc, n)
    if sc == "d3":
        assert out[6].min() >= 4 and out[6].max() <= 8          # SURVEY: 4-8 nodes per interval
        ks = g["d3_j2_ks"]
        oj = C.discretize_batch_adaptive(g["d3_x"][None], g["d3_u"][None], 2.0, const, include_J2=True)
        for n, o in zip(NAMES, oj[:5]):
            assert rel_err(_sel(o[0], ks), g[f"d3_j2_def_{n}"]) < 1e-12, n


@pytest.mark.parametrize("tag", ["c11", "c24"])
def test_oracles_reproduce_the_reference_on_a_coast_to_thrust_switch(tag):
    """u exactly 0 on the coasting nodes (a bang-off-bang plan): the reference looks the end nodes of an interval up on the
    GLOBAL grid (u_FOH, linearize_discretize.py:308-315; tau_6 = 0.6000000000000001 of an 11-node grid lands in [6, 7]) and
    gets |u| ~ 1e-15 > eps, so the guard of B_func (:208) does not fire and the mass row of B at that node carries the
    neighbour's thrust direction.  Both restatements follow that lookup (ref_node_input in the C one); fixture from the
    unmodified reference (make_golden.py coast)."""
    g = np.load(os.path.join(GOLDEN, "discretize_coast.npz"))
    c = O.OracleConstants(*g["const"])
    x, u, tf = g[tag + "_x"], g[tag + "_u"], float(g[tag + "_tf"])
    for mode, uniform in (("uni", True), ("def", False)):
        ref = [g[f"{tag}_{mode}_{n}"] for n in NAMES]
        assert np.max(np.abs(ref[1][:, 6])) > 0 and np.max(np.abs(ref[2][:, 6])) > 0      # the mass rows are in play
        po = O.discretize(x, u, tf, c, use_uniform_steps=uniform)
        co = C.discretize_batch(x[None], u[None], tf, c) if uniform else C.discretize_batch_adaptive(x[None], u[None], tf, c)
        for n, p_, c_, r in zip(NAMES, po, co[:5], ref):
            assert rel_err(p_, r) < 1e-13, (mode, n)
            assert rel_err(c_[0], r) < (1e-7 if uniform else 1e-13), (mode, n)           # uniform: the reference's RK45 noise
        # the rows the lookup decides: the mass rows of B_kp / B_kn
        assert rel_err(co[1][0][:, 6], ref[1][:, 6]) < 1e-12 and rel_err(co[2][0][:, 6], ref[2][:, 6]) < 1e-12


def test_reference_default_mode_distance_is_its_own_quadrature_error(gold_disc, const):
    """Documented, not gated: default-mode B+- sit ~1e-3 from the uniform-node answer."""
    g = gold_disc
    e = [rel_err(g[f"d3_def_{n}"], g[f"d3_uni_{n}"]) for n in NAMES]
    assert e[0] < 1e-8 and 1e-4 < e[1] < 1e-2 and 1e-4 < e[2] < 1e-2 and e[3] < 1e-5 and e[4] < 1e-3


def test_rollout_reproduces_reference_trajectory(gold_disc):
    """the implicit check of test_discretizer.py:110-113 turned into assertions: the discrete model,
    evaluated on its own reference trajectory, reproduces the nonlinear propagation.  Constant thrust is
    exactly representable by the first-order hold (defect = integration error); the tangential law is not
    (defect = FOH error of a state-dependent input, ~1e-6 per step at K=200)."""
    g = gold_disc
    for sc, step_tol, roll_tol in (("d2", 1e-6, 1e-5), ("d3", 1e-5, 2e-3)):
        out = [g[f"{sc}_uni_{n}"] for n in NAMES]
        x, u, tf = g[sc + "_x"], g[sc + "_u"], float(g[sc + "_tf"])
        defect = max(np.max(np.abs(out[0][k] @ x[:, k] + out[2][k] @ u[:, k] + out[1][k] @ u[:, k + 1]
                                   + out[3][:, k] * tf + out[4][:, k] - x[:, k + 1])) for k in range(x.shape[1] - 1))
        assert defect < step_tol, (sc, defect)
        xr = O.rollout(out[0], out[1], out[2], out[3], out[4], x[:, 0], u, tf)
        assert rel_err(xr, x) < roll_tol, sc


def test_constants_and_scaling(gold_disc):
    g = gold_disc
    sf = O.scale_factors(g["x0_dim"])
    c = O.normalized_constants(sf)
    got = np.array([c.MU, c.R_E, c.J2, c.G0, c.ISP, c.S, c.R0, c.RHO])
    assert np.allclose(got, g["const"], rtol=1e-15, atol=0)
    assert np.allclose([sf[k] for k in ("r0", "s0", "v0", "a0", "m0", "T0", "mu0")], g["scale"], rtol=1e-15)
    y = O.normalize_state(g["x0_dim"], sf)
    assert np.allclose(O.redim_state(y, sf), g["x0_dim"], rtol=1e-14)


# ---------------------------------------------------------------------------------- propagation

def _y0(gp):
    sf = O.scale_factors(gp["x0_dim"])
    return O.normalize_state(gp["x0_dim"], sf), sf


def test_py_oracle_propagation(gold_prop, const):
    gp = gold_prop
    y0, _ = _y0(gp)
    y, t = O.propagate(y0, 2.0, O.ctrl_tangential(0.5), const, False, False, 200)
    assert rel_err(y, gp["p1_y"]) < 1e-13 and np.array_equal(t, gp["p1_t"])
    assert rel_err(O.extract_uk(y, t, O.ctrl_tangential(0.5)), gp["p1_u"]) < 1e-13
    y, t = O.propagate(y0, 2.0, O.ctrl_sequence(gp["p5_u_tab"], 1.5, 2.0), const, False, False, 120)
    assert rel_err(y, gp["p5_y"]) < 1e-13


# fixed-step RK4 (h <= 1e-3 in tau) against the reference's RK45: <= 1e-6 required on propagated states
@pytest.mark.parametrize("case", ["p0", "p1", "p2", "p4"])
def test_c_oracle_propagation_vs_reference(gold_prop, const, case):
    gp = gold_prop
    y0, _ = _y0(gp)
    kw = {"p0": dict(kind=C.CTRL_ZERO, T=500), "p1": dict(kind=C.CTRL_TANGENTIAL, cparams=(0.5, 0, 0), T=200, include_drag=False, include_J2=False),
          "p2": dict(kind=C.CTRL_CONSTANT, cparams=gp["p2_thrust"], T=300), "p4": dict(kind=C.CTRL_TANGENTIAL, cparams=(0.1, 0, 0), T=200)}[case]
    T = kw["T"]
    y, u, st = C.propagate_batch(y0[None], float(gp[case + "_tf"]), const, n_sub=int(np.ceil(1000 / (T - 1))), **kw)
    assert st[0] == 0
    assert rel_err(y[0], gp[case + "_y"]) < 1e-6   # north_star tolerance on propagated states
    if case == "p1":
        assert rel_err(u[0], gp["p1_u"]) < 1e-7


def test_c_oracle_three_perturbed_sats(gold_prop):
    gp = gold_prop
    c3 = O.OracleConstants(*gp["p3_const"])
    sf = O.scale_factors(gp["p3_y0_dim"][0])
    y0 = np.stack([O.normalize_state(y, sf) for y in gp["p3_y0_dim"]])
    y, _, st = C.propagate_batch(y0, 5.0, c3, C.CTRL_ZERO, T=500, n_sub=3)
    assert st.max() == 0 and rel_err(y, gp["p3_y"]) < 1e-6


def test_c_oracle_sequence_controller(gold_prop, const):
    """FOH kinks only (end_tau >= 1, as OptimalController uses it): RK4 agrees to <= 1e-6.  A table that ends
    inside the run (end_tau < 1) is a jump discontinuity the reference itself integrates across with O(h)
    error; there the two agree only to that error (documented in DESIGN.md)."""
    gp = gold_prop
    y0, _ = _y0(gp)
    tab = gp["p5_u_tab"]
    yr, _ = O.propagate(y0, 2.0, O.ctrl_sequence(tab, 2.0, 2.0), const, False, False, 120)
    y, u, st = C.propagate_batch(y0[None], 2.0, const, C.CTRL_SEQUENCE, table=tab, end_tau=1.0, include_drag=False,
                                 include_J2=False, T=120, n_sub=10)
    assert st[0] == 0 and rel_err(y[0], yr) < 1e-6
    y, u, st = C.propagate_batch(y0[None], 2.0, const, C.CTRL_SEQUENCE, table=tab, end_tau=0.75, include_drag=False,
                                 include_J2=False, T=120, n_sub=10)
    assert rel_err(y[0], gp["p5_y"]) < 5e-3 and rel_err(u[0], gp["p5_u"]) < 1e-12


# ---- the reference's own integrator restated (orc_propagate_rk45: scipy RK45 + controller + dense output) ----------
_RK45_CASES = {"p0": dict(kind=C.CTRL_ZERO, T=500),
               "p1": dict(kind=C.CTRL_TANGENTIAL, cparams=(0.5, 0, 0), T=200, include_drag=False, include_J2=False),
               "p2": dict(kind=C.CTRL_CONSTANT, T=300), "p4": dict(kind=C.CTRL_TANGENTIAL, cparams=(0.1, 0, 0), T=200),
               "p5": dict(kind=C.CTRL_SEQUENCE, T=120, end_tau=0.75, include_drag=False, include_J2=False)}


@pytest.mark.parametrize("case", sorted(_RK45_CASES))
def test_c_oracle_rk45_propagation_reproduces_the_reference(gold_prop, const, case):
    """simulator.py:185-187 restated in C against trajectories of the unmodified reference: rounding-level agreement,
    INCLUDING the thrust cut-off inside the run (p5: SequenceController with end_tau = 0.75, control.py:127-141), and
    scipy's step count (nfev 6002 = 2 + 6 * 1000 steps, none rejected)."""
    gp = gold_prop
    y0, _ = _y0(gp)
    kw = dict(_RK45_CASES[case])
    if case == "p2":
        kw["cparams"] = gp["p2_thrust"]
    if case == "p5":
        kw["table"] = gp["p5_u_tab"]
    y, u, st, steps, rej = C.propagate_batch_rk45(y0[None], float(gp[case + "_tf"]), const, **kw)
    assert st[0] == 0 and steps[0] == 1000 and rej[0] == 0
    assert rel_err(y[0], gp[case + "_y"]) < 1e-12
    if case in ("p1", "p5"):
        assert rel_err(u[0], gp[case + "_u"]) < 1e-12


def test_c_oracle_rk45_three_sats_and_segments(gold_prop):
    gp = gold_prop
    c3 = O.OracleConstants(*gp["p3_const"])
    sf = O.scale_factors(gp["p3_y0_dim"][0])
    y0 = np.stack([O.normalize_state(y, sf) for y in gp["p3_y0_dim"]])
    y, _, st, steps, _ = C.propagate_batch_rk45(y0, 5.0, c3, C.CTRL_ZERO, T=500)
    assert st.max() == 0 and rel_err(y, gp["p3_y"]) < 1e-12


def test_c_oracle_rk45_step_control_engages(const):
    """With the reference's max_step = 0.001 the controller never binds (1000 steps, no rejection).  With a larger
    max_step it does: steps are rejected at the thrust cut-off and retried exactly as scipy does it (checked against
    scipy itself through the numpy restatement of the reference, which takes max_step as an argument)."""
    y0 = np.array([1.0, 0, 0, 0, 6.28, 0.3, 1.0])
    tab = np.array([[0.5] * 4, [0.15] * 4, [0.0, 3.0, 0.0, 0.0]])
    n_rej = 0
    for tf, ms in ((0.5, 1.0), (2.0, 1.0), (2.0, 0.05)):
        yr, _ = O.propagate(y0, tf, O.ctrl_sequence(tab, 0.37 * tf, tf), const, False, False, 64, max_step=ms)
        y, u, st, steps, rej = C.propagate_batch_rk45(y0[None], tf, const, C.CTRL_SEQUENCE, table=tab, end_tau=0.37,
                                                       include_drag=False, include_J2=False, T=64, max_step=ms)
        assert st[0] == 0 and rel_err(y[0], yr) < 1e-12
        n_rej += rej[0]
    assert n_rej >= 4


def test_c_oracle_mass_failure_flag(const):
    y0 = np.array([[1.0, 0, 0, 0, 6.28, 0, 1e-3]])
    y, _, st = C.propagate_batch(y0, 5.0, const, C.CTRL_CONSTANT, (5.0, 0, 0), T=50, n_sub=20)
    assert st[0] == 1


CT_KEYS = ['rbar_hat', 'ubar_hat', 'rf_hat', 'Vc', 'DrVc', 'DrVc_rbar', 'Vt', 'DrVt_DvVt', 'DrVt_DvVt_bar',
           'Vr', 'DrVr_DvVr', 'DrVr_DvVr_bar', 'Vn', 'DrVn_DvVn', 'DrVn_DvVn_bar']


@pytest.mark.parametrize("tag", ["c0", "c1", "c2"])
def test_py_oracle_constraint_terms_vs_reference(tag):
    """oracle.constraint_terms == Optimizer.get_constraint_terms of the unmodified reference (optimizer.py:80-170),
    fixtures from tests/golden/make_golden.py; c2 has zero thrust (the reference's NaN ubar_hat)."""
    g = np.load(os.path.join(GOLDEN, "constraint_terms.npz"))
    out = O.constraint_terms(g[tag + "_x"], g[tag + "_u"], float(g["MU"]))
    assert sorted(out) == sorted(CT_KEYS)
    for k in CT_KEYS:
        np.testing.assert_array_equal(np.asarray(out[k]), g[f"{tag}_{k}"], err_msg=k)
    if tag == "c2":
        assert np.isnan(out["ubar_hat"]).all()
    else:
        assert (out["ubar_hat"] == 0).all()


@pytest.mark.parametrize("tag", ["g0", "g1"])
@pytest.mark.parametrize("uniform", [True, False])
def test_py_oracle_drag_branch_vs_reference(tag, uniform):
    """Discretizer(include_drag=True) of the unmodified reference with const.CD / rho_func / drho_func supplied
    (linearize_discretize.py:160-169; fixtures: make_golden.py drag).  g1 has S x 1e4 + J2: drag is 3e-4 of the answer."""
    g = np.load(os.path.join(GOLDEN, "discretize_drag.npz"))
    c = O.OracleConstants(*g[tag + "_const"])
    mode = "uni" if uniform else "def"
    for n, k in enumerate(g[tag + "_ks"][:4]):
        out = O.interval_matrices(int(k), g[tag + "_x"], g[tag + "_u"], 1.0, c, include_J2=bool(g[tag + "_j2"]),
                                  use_uniform_steps=uniform, drag=(float(g[tag + "_cd"]), float(g[tag + "_rho_n"])))
        ref = [g[f"{tag}_{mode}_{nm}"] for nm in NAMES]
        got_ref = (ref[0][n], ref[1][n], ref[2][n], ref[3][:, n], ref[4][:, n])
        for nm, a, b in zip(NAMES, out, got_ref):
            assert rel_err(a, b) < 1e-13, (nm, k)


@pytest.mark.parametrize("tag", ["g0", "g1"])
def test_c_oracle_drag_branch_vs_reference(tag):
    """plain-C restatement with the drag branch: RK4/101-node mode vs the reference's uniform mode (1e-8), RK45 replica
    vs the reference's default mode (1e-10)"""
    g = np.load(os.path.join(GOLDEN, "discretize_drag.npz"))
    c = O.OracleConstants(*g[tag + "_const"])
    drag = (float(g[tag + "_cd"]), float(g[tag + "_rho_n"]))
    ks = g[tag + "_ks"]
    j2 = bool(g[tag + "_j2"])
    out = C.discretize_batch(g[tag + "_x"][None], g[tag + "_u"][None], 1.0, c, include_J2=j2, drag=drag)
    assert out[5].max() == 0
    for nm, o, r in zip(NAMES, [_sel(a[0], ks) for a in out[:5]], [g[f"{tag}_uni_{n}"] for n in NAMES]):
        assert rel_err(o, r) < TOL_UNIFORM, nm
    out = C.discretize_batch_adaptive(g[tag + "_x"][None], g[tag + "_u"][None], 1.0, c, include_J2=j2, drag=drag)
    for nm, o, r in zip(NAMES, [_sel(a[0], ks) for a in out[:5]], [g[f"{tag}_def_{n}"] for n in NAMES]):
        assert rel_err(o, r) < 1e-10, nm


def test_c_oracle_integration_schemes_agree(gold_disc, const):
    """The three fixed-step schemes of the oracle on the reference's tangential scenario (K=200, tf=2): classical RK4
    on the 56-vector, Nystrom's 3-stage method per node, and the default (two-node steps + Hermite midpoint) differ by
    rounding / 1e-13, far below the reference's own 1e-10; all are 4th order (halving the step divides the error by ~16)."""
    x, u = gold_disc["d3_x"][None], gold_disc["d3_u"][None]
    res = {}
    try:
        for sch in ("rk4", "rkn4", "rkn4x2"):
            C.set_scheme(sch)
            res[sch] = C.discretize_batch(x, u, 2.0, const)
        for a, b in (("rkn4", "rk4"), ("rkn4x2", "rkn4")):
            for nm, p, q in zip(NAMES, res[a][:5], res[b][:5]):
                assert rel_err(p, q) < 1e-12, (a, b, nm)
        xs, us = x[:, :, :20], u[:, :, :20]             # 19 intervals of 0.42 orbit: coarse on purpose
        C.set_scheme("rkn4")
        fine = C.discretize_batch(xs, us, 8.0, const, n_sub=400)[0]
        e8, e16 = (rel_err(C.discretize_batch(xs, us, 8.0, const, n_sub=n)[0], fine) for n in (8, 16))
        assert 10.0 < e8 / e16 < 24.0
    finally:
        C.set_scheme("rkn4x2")


# ---------------------------------------------------------------- the benchmark's own workload, through the reference
@pytest.fixture(scope="module")
def gold_bench():
    import os
    from conftest import GOLDEN
    return np.load(os.path.join(GOLDEN, "bench_workload.npz"))


def test_bench_constellation_is_the_one_the_reference_flew(gold_bench):
    """bench.make_constellation (product-side SatelliteScale) builds the initial states the fixture's satellites were
    flown from by the unmodified reference (tests/golden/make_golden.py bench)"""
    import bench
    gb = gold_bench
    Y, cm = bench.make_constellation(int(gb["n_sats"]))
    assert rel_err(Y[gb["idx"]], gb["y0"]) < 1e-15
    assert rel_err(np.array([cm.MU, cm.R_E, cm.J2, cm.G0, cm.ISP, cm.S, cm.R0, cm.RHO]), gb["const"]) < 1e-15


def test_c_oracle_on_the_bench_workload_vs_reference(gold_bench):
    """4 satellites of the 4096-satellite benchmark constellation: propagation (<= 1e-6, north_star), uniform-node
    matrices (<= 1e-8) and default-mode matrices (RK45 replica, <= 1e-10) against the unmodified reference"""
    gb = gold_bench
    cb = O.OracleConstants(*gb["const"])
    ks, tf = gb["ks"], float(gb["tf"])
    y, u, st = C.propagate_batch(gb["y0"], tf, cb, C.CTRL_TANGENTIAL, (0.5, 0, 0), include_drag=False, include_J2=False,
                                 T=200, n_sub=6)
    assert st.max() == 0
    for j in range(len(gb["idx"])):
        assert rel_err(y[j], gb[f"s{j}_x"]) < 1e-6 and rel_err(u[j], gb[f"s{j}_u"]) < 1e-6
        x_ref, u_ref = gb[f"s{j}_x"][None], gb[f"s{j}_u"][None]
        out = C.discretize_batch(x_ref, u_ref, tf, cb)
        for n, o in zip(NAMES, out[:5]):
            assert rel_err(_sel(o[0], ks), gb[f"s{j}_uni_{n}"]) < TOL_UNIFORM, (j, n)
        out = C.discretize_batch_adaptive(x_ref, u_ref, tf, cb)
        for n, o in zip(NAMES, out[:5]):
            assert rel_err(_sel(o[0], ks), gb[f"s{j}_def_{n}"]) < 1e-10, (j, n)
