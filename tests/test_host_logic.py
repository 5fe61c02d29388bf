"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT

import mpconstellation_b200 as M
from mpconstellation_b200 import _lib, batch
from mpconstellation_b200.control import spec_from


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpc_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    if _lib.needs_build():
        _lib.build()
    L = ctypes.CDLL(_lib.SO_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 14
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/mpc_b200.h but not exported"
    assert sorted(_lib.EXPORTED_SYMBOLS) == declared
    assert _lib.lib().mpc_version() == 100


def test_struct_layout_matches_header():
    assert ctypes.sizeof(_lib.MpcParams) == 12 * 8 + 4 * 4 + 2 * 8 + 2 * 32 * 8
    assert (_lib.MpcParams.disc_n_rho.offset, _lib.MpcParams.disc_r_mid.offset, _lib.MpcParams.disc_rho_cheb.offset,
            _lib.MpcParams.disc_drho_cheb.offset) == (104, 112, 128, 384)
    assert ctypes.sizeof(_lib.MpcController) == 4 * 4 + 3 * 8 + 8 + 8 + 8
    assert _lib.MpcController.table.offset == 48
    assert ctypes.sizeof(_lib.MpcGatherOpts) == 4 * 4 + 2 * 8 and _lib.MpcGatherOpts.n_sats_total.offset == 16


def test_no_cpu_fallback_without_device():
    if _lib.lib().mpc_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        batch.discretize_batch(np.zeros((1, 7, 3)) + 1.0, np.zeros((1, 3, 3)), 1.0, M.SatelliteScale().get_normalized_constants())


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mpconstellation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("mpc_oracle-free", ""), f"{f} mentions the oracle"
                assert "hostk" not in text, f"{f} mentions the host build of the kernel sources (tests/hostk, test-only)"


def test_scale_and_constants_mirror_reference(gold_disc):
    sat = M.Satellite(gold_disc["x0_dim"][0:3], gold_disc["x0_dim"][3:6], gold_disc["x0_dim"][6])
    scale = M.SatelliteScale(sat=sat)
    c = scale.get_normalized_constants()
    got = np.array([c.MU, c.R_E, c.J2, c.G0, c.ISP, c.S, c.R0, c.RHO])
    assert np.array_equal(got, gold_disc["const"])
    x = sat.get_state_vector()
    xn = scale.normalize_state(x)
    assert np.allclose(scale.redim_state(xn), x, rtol=1e-15)
    X = np.column_stack([x, 2 * x])
    assert np.allclose(scale.redim_state(scale.normalize_state(X)), X, rtol=1e-15)
    assert np.isclose(np.linalg.norm(xn[0:3]), 1.0) and xn[6] == 1.0
    assert scale.normalize_thrust(scale.redim_thrust(0.5)) == pytest.approx(0.5)
    ids = {M.Satellite().id for _ in range(100)}
    assert len(ids) == 100


def test_controller_specs_and_host_laws(gold_prop):
    gp = gold_prop
    from oracle import mpc_oracle as O
    x = gp["p1_y"][:, 17]
    s = spec_from(M.Controller())
    assert s.kind == _lib.CTRL_ZERO and np.array_equal(M.Controller().get_u_func()(x, 0.3), np.zeros(3))
    c = M.ConstantThrustController(thrust=np.array([0.1, 0.2, 0.3]))
    assert spec_from(c).kind == _lib.CTRL_CONSTANT and spec_from(c).thrust == (0.1, 0.2, 0.3)
    c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    assert np.allclose(c.get_u_func()(x, 0.1), O.ctrl_tangential(0.5)(x, 0.1), rtol=1e-14, atol=1e-16)
    tab = gp["p5_u_tab"]
    c = M.SequenceController(u=tab, tf_u=1.5, tf_sim=2.0)
    sp = spec_from(c)
    assert sp.kind == _lib.CTRL_SEQUENCE and sp.end_tau == 0.75
    ref = O.ctrl_sequence(tab, 1.5, 2.0)
    for tau in (0.0, 0.1234, 0.5, 0.75, 0.7500001, 1.0):
        assert np.allclose(c.get_u_func()(x, tau), ref(x, tau), rtol=1e-13, atol=1e-15)
    # u_func closures carry their spec; arbitrary callables are refused (no Python on the device)
    assert spec_from(c.get_u_func()).kind == _lib.CTRL_SEQUENCE
    with pytest.raises(NotImplementedError):
        spec_from(lambda x, tau: np.zeros(3))


def test_reference_style_controller_objects_are_recognised():
    """duck-typed by class name, the way the reference's own controller classes look"""
    class ConstantThrustController:                      # noqa: shadows on purpose
        def __init__(self):
            self.thrust = np.array([0., 0., 0.1])

        def get_u_func(self):
            return lambda x, tau: self.thrust

    c = ConstantThrustController()
    assert spec_from(c).kind == _lib.CTRL_CONSTANT
    assert spec_from(c.get_u_func()).thrust == (0.0, 0.0, 0.1)

    class OptimalLike:
        sequence_controller = M.SequenceController(u=np.ones((3, 4)), tf_u=2, tf_sim=1)

    assert spec_from(OptimalLike()).end_tau == 2.0


def test_controllers_with_their_own_law_are_refused_not_flown_as_their_parent():
    """a subclass may override get_u_func: flying its parent's law would be silently wrong"""
    class PD(M.ConstantThrustController):
        def get_u_func(self, sat_id=None):
            return lambda x, tau: -0.1 * np.asarray(x[3:6])

    with pytest.raises(NotImplementedError):
        spec_from(PD())
    with pytest.raises(NotImplementedError):
        spec_from(PD().get_u_func())

    class Tuned(M.ConstantThrustController):              # same law, only the constructor differs: still fine
        def __init__(self):
            super().__init__(thrust=np.array([0.0, 0.2, 0.0]))

    assert spec_from(Tuned()).thrust == (0.0, 0.2, 0.0)

    class ConstantThrustController:                      # reference-style base ...
        thrust = np.array([0., 0., 0.1])

    class Feedback(ConstantThrustController):            # ... and a subclass of it: another name, another law
        pass

    with pytest.raises(NotImplementedError):
        spec_from(Feedback())

    class OptimalController:                             # the reference's raises AttributeError before update()
        pass

    with pytest.raises(AttributeError):
        spec_from(OptimalController())


def test_run_segment_serves_replanning_controllers_in_the_reference_order(monkeypatch):
    """simulator.py:58-64: update(), propagate, write back -- per satellite.  A controller that re-plans in update()
    must see that order (one launch per satellite); a stateless one is propagated in a single batched launch."""
    calls = []

    def fake_propagate(self, sats, tf, controller):
        calls.append((len(sats), spec_from(controller).thrust))
        T = int(self.eval_points)
        y = np.zeros((len(sats), 7, T))
        y[:, 6] = 1.0
        return y, np.zeros((len(sats), 3, T)), np.linspace(0, 1, T)

    monkeypatch.setattr(M.Simulator, "_propagate", fake_propagate)

    class Replanning(M.ConstantThrustController):
        n = 0

        def update(self):
            self.n += 1
            self.thrust = np.array([0.1 * self.n, 0.0, 0.0])

    sats = [M.Satellite() for _ in range(3)]
    sim = M.Simulator(sats=sats, controller=Replanning(), base_res=10)
    sim.run_segment(tf=1)
    assert [c[0] for c in calls] == [1, 1, 1] and np.allclose([c[1][0] for c in calls], [0.1, 0.2, 0.3])
    calls.clear()
    sim = M.Simulator(sats=sats, controller=M.ConstantThrustController(thrust=np.array([0.3, 0, 0])), base_res=10)
    sim.run_segments(tf=2, num_segments=2)
    assert [c[0] for c in calls] == [3, 3] and sim.sim_data[sats[0].id].shape == (7, 20)


def test_fit_density_of_the_drag_linearisation():
    """rho_func / drho_func (linearize_discretize.py:164-165) as the device takes them: a number when the density is
    constant along the batch, else Chebyshev series over the batch's radii, checked at the batch's own positions"""
    from mpconstellation_b200.discretizer import fit_density, _cheb_val
    rng = np.random.default_rng(5)
    d = rng.normal(size=(200, 3))
    pos = d / np.linalg.norm(d, axis=1)[:, None] * rng.uniform(1.0, 1.2, size=(200, 1))
    assert fit_density(lambda r: 4.0e4, lambda r: 0.0, pos) == 4.0e4
    a, b, r0, re_, sc = 8e26, -6.828, 6.9e6, 6.371e6, 3.7e-17          # the power law of simulator.py:110
    rho = lambda r: a * (np.linalg.norm(r) * r0 - re_) ** b / sc
    drho = lambda r: a * b * (np.linalg.norm(r) * r0 - re_) ** (b - 1) * r0 / sc
    m = fit_density(rho, drho, pos)
    rad = np.linalg.norm(pos, axis=1)
    t = (rad - m["r_mid"]) * m["r_ihalf"]
    assert np.all(np.abs(t) <= 1.0) and len(m["rho_c"]) == 32 and len(m["drho_c"]) == 32
    rr = np.array([rho(p) for p in pos])
    dd = np.array([drho(p) for p in pos])
    # the density falls by 3.5 orders of magnitude across these radii: 1e-10 of the largest value, 1e-6 of the local one
    assert np.max(np.abs(_cheb_val(m["rho_c"], t) - rr)) < 1e-10 * rr.max() and np.max(np.abs(_cheb_val(m["rho_c"], t) - rr) / rr) < 1e-6
    assert np.max(np.abs(_cheb_val(m["drho_c"], t) - dd)) < 1e-10 * np.abs(dd).max()
    m0 = fit_density(rho, lambda r: 0.0, pos)
    assert len(m0["drho_c"]) == 0 and np.array_equal(m0["rho_c"], m["rho_c"])
    with pytest.raises(NotImplementedError, match="smooth function of"):       # a kink inside the batch's radii
        fit_density(lambda r: abs(np.linalg.norm(r) - 1.1), lambda r: 0.0, pos)
    p = M._lib.make_params(M.SatelliteScale().get_normalized_constants(), include_drag=True, disc_drag=(2.5, m))
    assert (p.disc_n_rho, p.disc_n_drho) == (32, 32) and p.disc_rho_cheb[3] == m["rho_c"][3] and p.disc_cd == 2.5
    assert p.disc_r_ihalf == m["r_ihalf"] and p.disc_rho == m["rho_c"][0]


def test_discretizer_option_errors_mirror_reference():
    const = M.SatelliteScale().get_normalized_constants()
    f = M.Simulator.satellite_dynamics
    x = np.ones((7, 3))
    u = np.zeros((3, 3))
    d = M.Discretizer(const, include_drag=True)
    with pytest.raises(TypeError):                       # reference: 'NoneType' object is not callable
        d.discretize(f, x, u, 1.0)
    d = M.Discretizer(const, rho_func=lambda r: 1.0, drho_func=lambda r: 0.0, include_drag=True)
    with pytest.raises(AttributeError):                  # reference: 'Constants' object has no attribute 'CD'
        d.discretize(f, x, u, 1.0)
    const_cd = M.SatelliteScale().get_normalized_constants()
    const_cd.CD = 2.5
    # a density that depends on the DIRECTION of r has no device form (the kernels take rho(|r|), drho(|r|))
    d = M.Discretizer(const_cd, rho_func=lambda r: float(r[0]), drho_func=lambda r: 0.0, include_drag=True)
    xv = np.ones((7, 3))
    xv[0:3] = np.eye(3)
    with pytest.raises(NotImplementedError, match="smooth function of"):
        d.discretize(f, xv, u, 1.0)
    d = M.Discretizer(const_cd, rho_func=lambda r: 1.0, drho_func=lambda r: np.ones(3), include_drag=True)
    with pytest.raises(NotImplementedError, match="scalar"):                # drho_func is d rho / d|r|
        d.discretize(f, xv, u, 1.0)
    d = M.Discretizer(const)
    with pytest.raises(NotImplementedError):
        d.discretize(lambda *a, **k: None, x, u, 1.0)
    d.ivp_solver = "DOP853"
    with pytest.raises(NotImplementedError):
        d.discretize(f, x, u, 1.0)
    d = M.Discretizer(const)
    assert (d.ivp_max_step, d.ivp_solver, d.integrator_steps, d.use_uniform_steps) == (1e-2, 'RK45', 101, False)


def test_discretized_batch_views_have_reference_shapes_and_order():
    N, K = 3, 5
    n = N * (K - 1)
    soa = np.arange(105 * n, dtype=np.float64).reshape(105, n)
    b = batch.DiscretizedBatch(soa, np.zeros((N, K - 1), np.int32), N, K)
    A, Bp, Bn, S, X = b.sat(1)
    assert A.shape == (K - 1, 7, 7) and Bp.shape == (K - 1, 7, 3) and Bn.shape == (K - 1, 7, 3)
    assert S.shape == (7, K - 1) and X.shape == (7, K - 1)
    k, i, j = 2, 4, 5
    col = 1 * (K - 1) + k
    assert A[k, i, j] == soa[i * 7 + j, col]                 # model.A_k[s][k,i,j]   (optimizer.py:332)
    assert Bp[k, i, 2] == soa[49 + i * 3 + 2, col]           # model.B_kp[s][k,i,j]  (optimizer.py:334)
    assert Bn[k, i, 1] == soa[70 + i * 3 + 1, col]           # model.B_kn[s][k,i,j]  (optimizer.py:333)
    assert S[i, k] == soa[91 + i, col] and X[i, k] == soa[98 + i, col]   # [i,k]     (optimizer.py:335-336)
    assert A.base is not None                                 # views, not copies
    As, Bps, Bns, Ss, Xs = b.stacked()
    assert np.array_equal(As[1], A) and np.array_equal(Ss[1], S) and np.array_equal(Bns[1], Bn)
    # the k-major layout (column = k N + s, what propagate_discretize(layout="kmajor") returns): same views, still no copy
    km = np.ascontiguousarray(soa.reshape(105, N, K - 1).transpose(0, 2, 1)).reshape(105, n)
    bk = batch.DiscretizedBatch(km, np.zeros((N, K - 1), np.int32), N, K, layout="kmajor")
    for s_ in range(N):
        for a, c in zip(b.sat(s_), bk.sat(s_)):
            assert np.array_equal(a, c) and np.shares_memory(c, km)
    for a, c in zip(b.stacked(), bk.stacked()):
        assert np.array_equal(a, c) and np.shares_memory(c, km)


def test_default_n_sub_matches_reference_max_step():
    assert batch.default_n_sub(200) == 6 and batch.default_n_sub(1001) == 1 and batch.default_n_sub(2) == 1000
    assert batch.default_n_sub(1) == 1 and batch.default_n_sub(500) == 3


def test_host_dynamics_matches_oracle(gold_prop, const):
    from oracle import mpc_oracle as O
    y = gold_prop["p0_y"][:, 123]
    uf = O.ctrl_tangential(0.3)
    a = M.Simulator.satellite_dynamics(0.2, y, uf, 2.0, const, True, True)
    b = O.dynamics(y, uf(y, 0.2), 2.0, const, True, True)
    assert np.allclose(a, b, rtol=1e-14, atol=1e-18)
    with pytest.raises(Exception, match="INVALID SATELLITE MASS"):
        M.Simulator.satellite_dynamics(0.0, np.array([1, 0, 0, 0, 1, 0, -1.0]), uf, 1.0, const)


def test_save_to_csv_wire_format(tmp_path, gold_prop):
    """simulator.py:192-201: trajectory_<date>_<id><suffix>.csv, one row per sample, 7 columns, redimensionalized;
    the file np.loadtxt / MATLAB csvread (visualizer.m) read back is the trajectory to the last bit"""
    sat = M.Satellite(np.array([7e6, 0, 0]), np.array([0, 7.5e3, 0]), 1000.0)
    scale = M.SatelliteScale(sat=sat)
    sim = M.Simulator(sats=[sat], scale=scale)
    traj = np.asarray(gold_prop["p1_y"], dtype=float)          # a (7,T) trajectory from the unmodified reference
    sim.sim_data = {sat.id: traj}
    for redim in (True, False):
        (path,) = sim.save_to_csv(suffix="_t", redimensionalize=redim, directory=str(tmp_path))
        name = os.path.basename(path)
        assert re.fullmatch(r"trajectory_\d{4}-\d\d-\d\d-\d\d-\d\d-\d\d_%s_t\.csv" % sat.id, name)
        lines = open(path).read().splitlines()
        assert len(lines) == traj.shape[1] and all(len(ln.split(",")) == 7 for ln in lines)
        back = np.loadtxt(path, delimiter=",")
        want = scale.redim_state(traj).T if redim else traj.T
        assert np.array_equal(back, want)


def test_final_term_layout_matches_header_and_reference_keys():
    """_lib.FINAL_TERM_LAYOUT <-> MPC_FT_* in include/mpc_b200.h <-> the keys of Optimizer.get_constraint_terms"""
    from mpconstellation_b200 import constraints
    text = open(os.path.join(ROOT, "include", "mpc_b200.h")).read()
    defs = dict(re.findall(r"#define MPC_FT_([A-Z_]+) (\d+)", text))
    want = {"rf_hat": "RF_HAT", "Vc": "VC", "DrVc": "DRVC", "DrVc_rbar": "DRVC_RBAR", "Vt": "VT", "DrVt_DvVt": "DRVT_DVVT",
            "DrVt_DvVt_bar": "DRVT_DVVT_BAR", "Vr": "VR", "DrVr_DvVr": "DRVR_DVVR", "DrVr_DvVr_bar": "DRVR_DVVR_BAR",
            "Vn": "VN", "DrVn_DvVn": "DRVN_DVVN", "DrVn_DvVn_bar": "DRVN_DVVN_BAR"}
    assert set(want) == set(_lib.FINAL_TERM_LAYOUT) and set(want) | {"rbar_hat", "ubar_hat"} == set(constraints.KEYS)
    for key, (off, ln) in _lib.FINAL_TERM_LAYOUT.items():
        assert int(defs[want[key]]) == off
    spans = sorted((off, max(ln, 1)) for off, ln in _lib.FINAL_TERM_LAYOUT.values())
    assert spans[0][0] == 0 and all(a + n == b for (a, n), (b, _) in zip(spans, spans[1:]))
    assert spans[-1][0] + spans[-1][1] == _lib.FINAL_TERMS == int(re.search(r"#define MPC_FINAL_TERMS (\d+)", text).group(1))


def test_dynamics_jacobian_container_against_a_dense_restatement():
    """DynamicsJacobian (host container of mpc_dynamics_jacobian's CSR): pack / residual / to_scipy on synthetic
    values laid out exactly as the kernel documents them (include/mpc_b200.h), against the pyomo rule written densely"""
    from mpconstellation_b200.constraints import DynamicsJacobian
    N, K = 2, 4
    rng = np.random.default_rng(1)
    A, Bn, Bp = rng.standard_normal((N, K - 1, 7, 7)), rng.standard_normal((N, K - 1, 7, 3)), rng.standard_normal((N, K - 1, 7, 3))
    Sg, Xi = rng.standard_normal((N, 7, K - 1)), rng.standard_normal((N, 7, K - 1))
    rows = N * 7 * (K - 1)
    vals, idx, rhs = np.zeros((rows, 16)), np.zeros((rows, 16), dtype=np.int64), np.zeros(rows)
    NK = N * K
    for s in range(N):
        for i in range(7):
            for k in range(K - 1):
                r = (s * 7 + i) * (K - 1) + k
                ent = [((s * 7 + j) * K + k, -A[s, k, i, j]) for j in range(7)] + [((s * 7 + i) * K + k + 1, 1.0)]
                ent += [(7 * NK + (s * 3 + j) * K + k, -Bn[s, k, i, j]) for j in range(3)]
                ent += [(7 * NK + (s * 3 + j) * K + k + 1, -Bp[s, k, i, j]) for j in range(3)]
                ent += [(10 * NK + (s * 7 + i) * K + k, -1.0), (17 * NK, -Sg[s, i, k])]
                ent.sort()
                idx[r], vals[r], rhs[r] = [e[0] for e in ent], [e[1] for e in ent], Xi[s, i, k]
    jac = DynamicsJacobian(vals.ravel(), idx.ravel(), rhs, N, K)
    x, u, nu, tf = rng.standard_normal((N, 7, K)), rng.standard_normal((N, 3, K)), rng.standard_normal((N, 7, K)), 1.3
    res = jac.residual(jac.pack(x, u, nu, tf))
    for s in range(N):
        for k in range(K - 1):
            want = x[s, :, k + 1] - (A[s, k] @ x[s, :, k] + Bn[s, k] @ u[s, :, k] + Bp[s, k] @ u[s, :, k + 1]
                                     + Sg[s, :, k] * tf + Xi[s, :, k] + nu[s, :, k])
            assert np.allclose(res[s, :, k], want, rtol=0, atol=1e-12)
    J = jac.to_scipy()
    assert J.shape == (rows, 17 * NK + 1) and np.allclose(J @ jac.pack(x, u, nu, tf) - rhs, res.ravel(), atol=1e-12)
    assert np.array_equal(jac.indptr, np.arange(rows + 1) * 16)
