"""GPU: the batched SCP pass / closed-loop driver around the hot path (BASELINE config 5 shape), with the
pyomo/ipopt subproblem replaced by a stand-in (no solver is installed; see mpconstellation_b200/scp.py)."""
import numpy as np
import pytest

from conftest import rel_err, synth_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def M():
    import mpconstellation_b200 as m
    m._lib.require_gpu()
    return m


def test_hand_off_satisfies_the_dynamics_constraint_in_pyomo_indexing(M, const):
    """optimizer.py:327-339 evaluated on the reference point itself: residual = FOH/linearization defect only"""
    from mpconstellation_b200.scp import BatchedSCP
    y0, _, _ = synth_batch(6, 2, 1.0, const)
    for uniform in (False, True):
        scp = BatchedSCP(const, base_res=30, tf_horizon=2.0, use_uniform_steps=uniform)
        lin = scp.linearize(y0, 2.0, M.ConstantTangentialThrustController(tangential_thrust=0.5))
        assert lin.x_bar.shape == (6, 7, 60) and lin.u_bar.shape == (6, 3, 60)
        for s in (0, 5):
            A_k, B_kp, B_kn, Sigma_k, xi_k = lin.sat(s)
            assert A_k.shape == (59, 7, 7) and Sigma_k.shape == (7, 59)
            res = lin.dynamics_residual(s, lin.x_bar[s], lin.u_bar[s], 2.0)
            # K=60 over two orbits: coarse FOH of a state-dependent input (+ the default mode's own quadrature error)
            assert np.max(np.abs(res)) < 5e-3
            # virtual control nu absorbs exactly that defect (optimizer.py:337)
            assert np.max(np.abs(lin.dynamics_residual(s, lin.x_bar[s], lin.u_bar[s], 2.0, nu=res))) < 1e-15


def test_config5_closed_loop_256_satellites(M, const):
    """256 satellites, base_res=30, tf_horizon=2, two segments of one orbit (test_simulator.py:79-95 shape)"""
    from mpconstellation_b200.scp import BatchedSCP
    from oracle import c_oracle as C
    y0, _, _ = synth_batch(256, 2, 1.0, const)
    scp = BatchedSCP(const, base_res=30, tf_horizon=2.0, tf_interval=1.0, n_iterations=2)
    before = M.launch_count()
    traj, plans = scp.run_segments(y0, n_segments=2)
    assert M.launch_count() - before == 2 * (2 * 2 + 1)        # per segment: 2 x (propagate + discretize) + 1 flight
    assert traj.shape == (256, 7, 200) and np.all(np.isfinite(traj))
    assert np.all(np.diff(traj[:, 6, :], axis=1) <= 1e-15)      # thrusting: mass never increases
    assert len(plans) == 2 and len(plans[0]) == 2 and plans[0][0].matrices.status.max() == 0
    # first flight segment against the C restatement of the reference's integrator (the propagator's default):
    # per-satellite table, end_tau = tf_u / interval = 2
    u_tab = plans[0][1].u_bar            # the stand-in solver returns the reference input of the last iteration
    for s in (0, 100, 255):
        yr, _, st, _, _ = C.propagate_batch_rk45(y0[s:s + 1], 1.0, const, C.CTRL_SEQUENCE, table=u_tab[s], end_tau=2.0,
                                                 include_drag=True, include_J2=True, T=100)
        assert st[0] == 0 and rel_err(traj[s, :, :100], yr[0]) < 2e-12
    # the second segment starts where the first ended
    assert np.array_equal(traj[:, :, 100], traj[:, :, 99])


def test_per_satellite_end_tau(M, const):
    from oracle import c_oracle as C
    y0, _, _ = synth_batch(5, 2, 1.0, const)
    rng = np.random.default_rng(11)
    tabs = 0.2 * rng.standard_normal((5, 3, 12))
    end_tau = np.array([0.5, 1.0, 1.5, 2.0, 0.8])
    spec = M.ControllerSpec(M._lib.CTRL_SEQUENCE, (0.0, 0.0, 0.0), tabs, end_tau)
    y, u, t, st = M.propagate_batch(y0, 1.3, spec, const, include_drag=False, include_J2=True, T=40, n_sub=8)
    for s in range(5):
        yr, ur, sr = C.propagate_batch(y0[s:s + 1], 1.3, const, C.CTRL_SEQUENCE, table=tabs[s], end_tau=float(end_tau[s]),
                                       include_drag=False, include_J2=True, T=40, n_sub=8)
        assert rel_err(y[s], yr[0]) < 1e-11 and rel_err(u[s], ur[0]) < 1e-10
