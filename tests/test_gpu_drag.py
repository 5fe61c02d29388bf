"""GPU: Discretizer(include_drag=True) -- the drag branch of the linearisation (linearize_discretize.py:160-169) that
the reference reaches once const.CD, rho_func and drho_func are supplied -- against fixtures produced by the unmodified
reference (tests/golden/make_golden.py drag) and against the plain-C oracle on a batch."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, NAMES, rel_err, synth_batch

pytestmark = pytest.mark.gpu

TOL_REF = 1e-8        # vs the reference (its RK45 noise floor is ~1e-10)
TOL_ORACLE = 1e-10    # vs the C oracle running the same algorithm


@pytest.fixture(scope="module")
def M():
    import mpconstellation_b200 as m
    m._lib.require_gpu()
    return m


def _const_with_cd(vec, cd):
    from oracle.mpc_oracle import OracleConstants
    c = OracleConstants(*vec)
    bag = type("Const", (), {k: getattr(c, k) for k in ("MU", "R_E", "J2", "G0", "ISP", "S", "R0", "RHO")})()
    bag.CD = cd
    return bag, c


@pytest.mark.parametrize("tag", ["g0", "g1"])
@pytest.mark.parametrize("uniform", [True, False])
def test_drag_discretizer_matches_reference_fixtures(M, tag, uniform):
    g = np.load(os.path.join(GOLDEN, "discretize_drag.npz"))
    const, _ = _const_with_cd(g[tag + "_const"], float(g[tag + "_cd"]))
    rho_n = float(g[tag + "_rho_n"])
    d = M.Discretizer(const, rho_func=lambda r: rho_n, drho_func=lambda r: 0.0, include_drag=True,
                      include_J2=bool(g[tag + "_j2"]))
    d.use_uniform_steps = uniform
    out = d.discretize(M.Simulator.satellite_dynamics, g[tag + "_x"], g[tag + "_u"], float(g[tag + "_tf"]))
    ks = g[tag + "_ks"]
    mode = "uni" if uniform else "def"
    for nm, o in zip(NAMES, out):
        ref = g[f"{tag}_{mode}_{nm}"]
        got = o[ks] if o.ndim == 3 else o[:, ks]
        assert rel_err(got, ref) < TOL_REF, (nm, rel_err(got, ref))
    # and the drag terms matter in g1: the no-drag discretization of the same inputs is 3e-4 away
    if tag == "g1" and uniform:
        d0 = M.Discretizer(const, include_J2=True)
        d0.use_uniform_steps = True
        A0 = d0.discretize(M.Simulator.satellite_dynamics, g[tag + "_x"], g[tag + "_u"], 1.0)[0]
        assert rel_err(A0[ks], g["g1_uni_A_k"]) > 1e-5


@pytest.mark.parametrize("uniform", [True, False])
def test_drag_discretizer_with_altitude_dependent_density_matches_the_reference(M, uniform):
    """Discretizer(rho_func = the power law of simulator.py:110, drho_func = its derivative): Dr_aD
    (linearize_discretize.py:166) is not zero, G + Dr_aD not symmetric.  Through the reference's own signature; the host
    hands the two callables over as Chebyshev series (discretizer.fit_density).  Fixture from the unmodified reference
    (make_golden.py drag_radial: radii 1.0 ... 1.2, the density falls by four orders of magnitude along the trajectory)."""
    g = np.load(os.path.join(GOLDEN, "discretize_drag_radial.npz"))
    const, _ = _const_with_cd(g["const"], float(g["cd"]))
    a, b, r0_m, r_e_m, rho_scale = [float(v) for v in g["law"]]
    rho = lambda r: a * (np.linalg.norm(r) * r0_m - r_e_m) ** b / rho_scale
    drho = lambda r: a * b * (np.linalg.norm(r) * r0_m - r_e_m) ** (b - 1.0) * r0_m / rho_scale
    ks = g["ks"]
    mode = "uni" if uniform else "def"
    for tag, dfun in ((mode, drho), ("def_nograd", lambda r: 0.0)):
        if tag == "def_nograd" and uniform:
            continue
        d = M.Discretizer(const, rho_func=rho, drho_func=dfun, include_drag=True, include_J2=True)
        d.use_uniform_steps = uniform
        out = d.discretize(M.Simulator.satellite_dynamics, g["x"], g["u"], float(g["tf"]))
        for nm, o in zip(NAMES, out):
            ref = g[f"{tag}_{nm}"]
            got = o[ks] if o.ndim == 3 else o[:, ks]
            assert rel_err(got, ref) < TOL_REF, (tag, nm, rel_err(got, ref))
    assert rel_err(g["def_A_k"], g["def_nograd_A_k"]) > 1e-3          # the gradient term is what is being tested
    # a batch: the same satellite three times with the model fitted over all of them == the single call
    if not uniform:
        x3, u3 = np.repeat(g["x"][None], 3, axis=0), np.repeat(g["u"][None], 3, axis=0)
        res = d.discretize_batch(M.Simulator.satellite_dynamics, x3, u3, float(g["tf"]))
        assert res.status.max() == 0
        s0, s2 = res.sat(0), res.sat(2)
        for o0, o2 in zip(s0, s2):
            assert np.array_equal(o0, o2)


@pytest.mark.parametrize("j2", [False, True])
def test_drag_batch_matches_c_oracle(M, const, j2):
    """37 satellites x K=13 (ragged against the 64-thread CTAs), exaggerated drag, per-satellite tf; both modes"""
    from oracle import c_oracle as C
    import copy
    cst = copy.copy(const)
    cst.S = const.S * 3e3
    y0, x, u = synth_batch(37, 13, 0.4, cst)
    tf = np.linspace(0.3, 0.5, 37)
    drag = (2.2, 4.0e4)
    ref = C.discretize_batch(x, u, tf, cst, include_J2=j2, n_sub=40, drag=drag)
    res = M.discretize_batch(x, u, tf, cst, include_J2=j2, include_drag=True, disc_drag=drag, n_sub=40)
    assert ref[5].max() == 0 and res.status.max() == 0
    for nm, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, (nm, rel_err(o, r))
    A = res.stacked()[0]
    assert np.all(A[..., 6, 6] == 1.0) and not np.any(A[..., 6, :6])
    ref = C.discretize_batch_adaptive(x, u, tf, cst, include_J2=j2, drag=drag)
    res = M.discretize_batch(x, u, tf, cst, include_J2=j2, include_drag=True, disc_drag=drag, adaptive=dict())
    assert res.status.max() == 0 and np.array_equal(res.n_nodes, ref[6])
    for nm, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < TOL_ORACLE, (nm, rel_err(o, r))


def test_drag_unsupported_combinations_fail_loudly(M, const):
    _, x, u = synth_batch(2, 5, 1.0, const)
    with pytest.raises(M._lib.MpcError) as e:           # u on its own grid + drag
        M.discretize_batch(x, np.repeat(u, 2, axis=2), 1.0, const, include_drag=True, disc_drag=(2.5, 1.0))
    assert e.value.code == M._lib.E_UNSUPPORTED
