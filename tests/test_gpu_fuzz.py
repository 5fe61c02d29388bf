"""GPU: seeded differential sweep -- random batch shapes, horizons, step counts, dynamics flags and controller laws,
every kernel against the plain-C oracle on the same inputs.  Sizes are small (the oracle finishes in seconds)."""
import numpy as np
import pytest

from conftest import NAMES, rel_err, synth_batch

pytestmark = pytest.mark.gpu

CASES = list(range(16))


@pytest.fixture(scope="module")
def M():
    import mpconstellation_b200 as m
    m._lib.require_gpu()
    return m


def _case(seed):
    rng = np.random.default_rng(1000 + seed)
    tf = float(rng.uniform(0.02, 1.5))
    K = max(int(rng.integers(2, 45)), int(np.ceil(1.1 * tf / 0.1)) + 1)      # intervals of at most 0.1 orbit
    # RK4 step of at most 5e-4 orbit: there the symplectic inverse of the numerical Phi (kernel) and its dense inverse
    # (oracle) agree to ~1e-11 (they differ by the integrator's own defect, DESIGN.md "very coarse steps")
    n_sub = max(int(rng.integers(4, 60)), int(np.ceil(1.1 * tf / (K - 1) / 5e-4)))
    return dict(n=int(rng.integers(1, 70)), K=K, tf=tf, n_sub=n_sub, j2=bool(rng.integers(0, 2)),
                drag=bool(rng.integers(0, 2)), kind=int(rng.integers(0, 4)), rng=rng)


@pytest.mark.parametrize("seed", CASES)
def test_propagate_then_discretize_against_c_oracle(M, const, seed):
    from oracle import c_oracle as C
    c = _case(seed)
    rng, n, K, tf = c["rng"], c["n"], c["K"], c["tf"]
    y0, _, _ = synth_batch(n, 2, 0.1, const, seed=seed)
    tfv = tf * (1 + 0.1 * rng.random(n))
    # --- controller law ---------------------------------------------------------------------------------------
    if c["kind"] == 0:
        ctrl, ck, cp, tab, et = M.Controller(), C.CTRL_ZERO, (0, 0, 0), None, 1.0
    elif c["kind"] == 1:
        cp = tuple(rng.uniform(-0.5, 0.5, 3))
        ctrl, ck, tab, et = M.ConstantThrustController(thrust=np.array(cp)), C.CTRL_CONSTANT, None, 1.0
    elif c["kind"] == 2:
        mag = float(rng.uniform(0.1, 1.0))
        ctrl, ck, cp, tab, et = M.ConstantTangentialThrustController(tangential_thrust=mag), C.CTRL_TANGENTIAL, (mag, 0, 0), None, 1.0
    else:
        Ku = int(rng.integers(2, 30))
        tab = rng.uniform(-0.4, 0.4, (n, 3, Ku))
        et = float(rng.uniform(1.0, 2.5))
        ctrl, ck, cp = M.ControllerSpec(M._lib.CTRL_SEQUENCE, (0, 0, 0), tab, et), C.CTRL_SEQUENCE, (0, 0, 0)
    T = max(K, 2)
    n_prop = int(rng.integers(1, 12))
    y, u, _, st = M.propagate_batch(y0, tfv, ctrl, const, include_drag=c["drag"], include_J2=c["j2"], T=T, n_sub=n_prop)
    yr, ur, sr = C.propagate_batch(y0, tfv, const, ck, cp, table=tab, end_tau=et, include_drag=c["drag"],
                                   include_J2=c["j2"], T=T, n_sub=n_prop)
    assert st.max() == 0 and sr.max() == 0
    assert rel_err(y, yr) < 1e-11 and np.max(np.abs(u - ur)) < 1e-11
    # --- discretization about that trajectory, both quadrature modes ---------------------------------------------
    ref = C.discretize_batch(yr, ur, tfv, const, include_J2=c["j2"], n_sub=c["n_sub"])
    res = M.discretize_batch(yr, ur, tfv, const, include_J2=c["j2"], n_sub=c["n_sub"])
    assert res.status.max() == 0 and ref[5].max() == 0
    for nm, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < 1e-10, (nm, rel_err(o, r))
    ref = C.discretize_batch_adaptive(yr, ur, tfv, const, include_J2=c["j2"])
    res = M.discretize_batch(yr, ur, tfv, const, include_J2=c["j2"], adaptive=dict())
    assert res.status.max() == 0 and np.array_equal(res.n_nodes, ref[6])
    for nm, o, r in zip(NAMES, res.stacked(), ref[:5]):
        assert rel_err(o, r) < 1e-10, (nm, rel_err(o, r))
