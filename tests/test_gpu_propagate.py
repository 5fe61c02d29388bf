"""GPU parity tests of the propagation kernel and the Simulator mirror, through the C-ABI."""
import numpy as np
import pytest

from conftest import rel_err, synth_batch

pytestmark = pytest.mark.gpu

TOL_STATE = 1e-9      # the default integrator replays the reference's solve_ivp call step for step: rounding only
                      # (observed <= 4e-13; north_star asks for 1e-6)
TOL_RK4 = 1e-6        # Simulator(integrator="rk4"): a different method, smooth inputs only (north_star's bound)
TOL_ORACLE = 1e-11    # same method, same step grid, vs the plain-C oracle


@pytest.fixture(scope="module")
def M():
    import mpconstellation_b200 as m
    m._lib.require_gpu()
    return m


def _hubble(M, gp):
    x = gp["x0_dim"]
    sat = M.Satellite(x[0:3].copy(), x[3:6].copy(), float(x[6]))
    return sat, M.SatelliteScale(sat=sat)


def test_five_orbit_coast_drag_j2(M, gold_prop):
    """test_simulator.py:17-33"""
    sat, scale = _hubble(M, gold_prop)
    sim = M.Simulator(sats=[sat], scale=scale)
    data, time = sim.run(tf=5)
    assert data[sat.id].shape == (7, 500) and np.array_equal(time[sat.id], gold_prop["p0_t"])
    assert rel_err(data[sat.id], gold_prop["p0_y"]) < TOL_STATE


def test_tangential_reference_trajectory_and_extract_uk(M, gold_prop):
    """test_discretizer.py:120-138: the trajectory the discretizer is linearized about, and extract_uk"""
    gp = gold_prop
    sat, scale = _hubble(M, gp)
    c = M.ConstantTangentialThrustController([sat], 0.5)
    sim = M.Simulator(sats=[sat], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=2)
    assert rel_err(sim.sim_data[sat.id], gp["p1_y"]) < TOL_STATE
    assert rel_err(sim.sim_u[sat.id], gp["p1_u"]) < TOL_STATE
    u_host = M.Discretizer.extract_uk(sim.sim_data[sat.id], sim.sim_time[sat.id], c)
    assert rel_err(u_host, sim.sim_u[sat.id]) < 1e-12


def test_constant_and_tangential_with_perturbations(M, gold_prop):
    """test_simulator.py:175-203"""
    gp = gold_prop
    sat, scale = _hubble(M, gp)
    sim = M.Simulator(sats=[sat], controller=M.ConstantThrustController(thrust=gp["p2_thrust"]), scale=scale)
    sim.run(tf=3)
    assert rel_err(sim.sim_data[sat.id], gp["p2_y"]) < TOL_STATE
    sat, scale = _hubble(M, gp)
    sim = M.Simulator(sats=[sat], controller=M.ConstantTangentialThrustController(tangential_thrust=0.1), scale=scale)
    sim.run(tf=2)
    assert rel_err(sim.sim_data[sat.id], gp["p4_y"]) < TOL_STATE


def test_three_satellites_one_scale(M, gold_prop):
    """test_simulator.py:36-55"""
    gp = gold_prop
    sats = [M.Satellite(y[0:3].copy(), y[3:6].copy(), float(y[6])) for y in gp["p3_y0_dim"]]
    scale = M.SatelliteScale(sat=sats[0])
    sim = M.Simulator(sats=sats, scale=scale)
    data, _ = sim.run(tf=5)
    got = np.stack([data[s.id] for s in sats])
    assert rel_err(got, gp["p3_y"]) < TOL_STATE


def test_run_segments_bookkeeping(M, gold_prop):
    """test_simulator.py:149-173: concatenation, time offsets (+1e-7), satellite state write-back"""
    gp = gold_prop
    x = gp["x0_dim"]
    s1 = M.Satellite(x[0:3].copy(), x[3:6].copy(), float(x[6]))
    s2 = M.Satellite(x[0:3].copy(), x[3:6] * 1.1, float(x[6]))
    scale = M.SatelliteScale(sat=s1)
    c = M.ConstantTangentialThrustController(sats=[s1, s2], tangential_thrust=0.5)
    sim = M.Simulator(sats=[s1, s2], scale=scale, base_res=100, controller=c)
    sim.run_segments(tf=3, num_segments=4)
    got_y = np.stack([sim.sim_data[s1.id], sim.sim_data[s2.id]])
    got_t = np.stack([sim.sim_time[s1.id], sim.sim_time[s2.id]])
    assert got_y.shape == gp["p6_y"].shape
    assert np.allclose(got_t, gp["p6_t"], rtol=0, atol=1e-15)
    assert rel_err(got_y, gp["p6_y"]) < TOL_STATE
    final = np.stack([s1.get_state_vector(), s2.get_state_vector()])
    assert rel_err(final, gp["p6_final_dim"]) < TOL_STATE


def test_sequence_controller(M, gold_prop, const):
    """SequenceController as OptimalController drives it (control.py:217): FOH kinks (end_tau >= 1), and the table
    ending INSIDE the run (end_tau = 0.75: the thrust jumps to zero, control.py:127-141) -- the reference steps across
    the jump with an O(h) error of its own, which the replayed integrator reproduces like everything else."""
    from oracle import mpc_oracle as O
    gp = gold_prop
    sat, scale = _hubble(M, gp)
    tab = gp["p5_u_tab"]
    y0 = scale.normalize_state(sat.get_state_vector())
    yr, _ = O.propagate(y0, 2.0, O.ctrl_sequence(tab, 2.0, 2.0), const, False, False, 120)
    sim = M.Simulator(sats=[sat], controller=M.SequenceController(u=tab, tf_u=2.0, tf_sim=2.0), scale=scale,
                      base_res=60, include_drag=False, include_J2=False)
    sim.run(tf=2)
    assert rel_err(sim.sim_data[sat.id], yr) < TOL_STATE
    sat, scale = _hubble(M, gp)
    sim = M.Simulator(sats=[sat], controller=M.SequenceController(u=tab, tf_u=1.5, tf_sim=2.0), scale=scale,
                      base_res=60, include_drag=False, include_J2=False)
    sim.run(tf=2)
    assert rel_err(sim.sim_data[sat.id], gp["p5_y"]) < TOL_STATE
    assert rel_err(sim.sim_u[sat.id], gp["p5_u"]) < 1e-12


def test_fixed_step_rk4_option(M, gold_prop):
    """Simulator(integrator="rk4"): the round-1 propagator stays available; a different method than the reference's, so
    it agrees to north_star's 1e-6 on smooth inputs only (the thrust cut-off of p5 is out of its reach: ~5e-4)."""
    gp = gold_prop
    sat, scale = _hubble(M, gp)
    c = M.ConstantTangentialThrustController([sat], 0.5)
    sim = M.Simulator(sats=[sat], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.integrator = "rk4"
    sim.run(tf=2)
    e = rel_err(sim.sim_data[sat.id], gp["p1_y"])
    assert 1e-12 < e < TOL_RK4
    sim.integrator = "euler"
    with pytest.raises(ValueError):
        sim.run(tf=2)


@pytest.mark.parametrize("kind", ["zero", "tangential", "sequence_cutoff_per_sat"])
def test_rk45_batch_matches_c_restatement_of_scipy(M, const, kind):
    """the replayed integrator on a ragged batch (drag + J2, per-satellite tf / tables / end_tau) against the C
    restatement of scipy's algorithm: trajectories to rounding, step counts identical; every satellites-per-warp
    mapping and the build without the speculative first stage give bit-identical results"""
    from oracle import c_oracle as C
    N, T, tf = 37, 64, 1.3
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    rng = np.random.default_rng(3)
    tfv = tf * (1 + 0.1 * rng.random(N))
    tabs = 0.3 * rng.standard_normal((N, 3, 9))
    et = 0.3 + 0.9 * rng.random(N)
    ctrl, ckw = {
        "zero": (M.Controller(), dict(kind=C.CTRL_ZERO)),
        "tangential": (M.ConstantTangentialThrustController(tangential_thrust=0.4), dict(kind=C.CTRL_TANGENTIAL, cparams=(0.4, 0, 0))),
        "sequence_cutoff_per_sat": (M.ControllerSpec(M._lib.CTRL_SEQUENCE, (0.0, 0.0, 0.0), tabs, et),
                                    dict(kind=C.CTRL_SEQUENCE, table=tabs, end_tau=et)),
    }[kind]
    steps = np.zeros(N, dtype=np.int32)
    y, u, t, st = M.propagate_batch(y0, tfv, ctrl, const, include_drag=True, include_J2=True, T=T, n_steps=steps)
    yr, ur, sr, ns, nr = C.propagate_batch_rk45(y0, tfv, const, include_drag=True, include_J2=True, T=T, **ckw)
    assert st.max() == 0 and sr.max() == 0 and np.array_equal(steps, ns + nr)
    assert rel_err(y, yr) < 2e-12 and rel_err(u, ur) < 1e-11
    assert np.array_equal(t, np.linspace(0, 1, T))
    L = M._lib.lib()
    try:
        for variant in (11, 14, 16, 19):          # no speculation; 32, 8, 1 satellites per warp
            assert L.mpc_set_tuning(variant) == 0
            y2, u2, _, st2 = M.propagate_batch(y0, tfv, ctrl, const, include_drag=True, include_J2=True, T=T)
            assert np.array_equal(y2, y) and np.array_equal(u2, u) and st2.max() == 0, variant
    finally:
        L.mpc_set_tuning(12)
        L.mpc_set_tuning(13)
    # step-size control engaged (max_step well above what the error test allows)
    y, u, t, st = M.propagate_batch(y0, tfv, ctrl, const, include_drag=True, include_J2=True, T=T, n_steps=steps,
                                    rk45=dict(max_step=0.05))
    yr, ur, sr, ns, nr = C.propagate_batch_rk45(y0, tfv, const, include_drag=True, include_J2=True, T=T, max_step=0.05, **ckw)
    assert st.max() == 0 and np.array_equal(steps, ns + nr) and rel_err(y, yr) < 2e-12


def test_get_trajectory_ode_signature(M, gold_prop):
    gp = gold_prop
    sat, scale = _hubble(M, gp)
    c = M.ConstantTangentialThrustController([sat], 0.5)
    sim = M.Simulator(sats=[sat], controller=c, scale=scale, include_drag=False, include_J2=False)
    sim.eval_points = 200
    sol = sim.get_trajectory_ODE(sat, 2, c.get_u_func())
    assert sol.y.shape == (7, 200) and sol.t.shape == (200,)
    assert sol.nfev == 6002                                      # what scipy reports for the reference's call
    assert rel_err(sol.y, gp["p1_y"]) < TOL_STATE
    with pytest.raises(NotImplementedError):
        sim.get_trajectory_ODE(sat, 2, lambda x, tau: np.zeros(3))


@pytest.mark.parametrize("kind", ["zero", "constant", "tangential", "sequence", "sequence_per_sat"])
def test_batch_matches_c_oracle(M, const, kind):
    from oracle import c_oracle as C
    N, T, tf = 37, 64, 1.3
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    rng = np.random.default_rng(3)
    tfv = tf * (1 + 0.1 * rng.random(N))
    tab = 0.3 * rng.standard_normal((3, 17))
    tabs = 0.3 * rng.standard_normal((N, 3, 9))
    ctrl, ckw = {
        "zero": (M.Controller(), dict(kind=C.CTRL_ZERO)),
        "constant": (M.ConstantThrustController(thrust=np.array([0.1, -0.2, 0.05])), dict(kind=C.CTRL_CONSTANT, cparams=(0.1, -0.2, 0.05))),
        "tangential": (M.ConstantTangentialThrustController(tangential_thrust=0.4), dict(kind=C.CTRL_TANGENTIAL, cparams=(0.4, 0, 0))),
        "sequence": (M.SequenceController(u=tab, tf_u=2.0, tf_sim=1.0), dict(kind=C.CTRL_SEQUENCE, table=tab, end_tau=2.0)),
        "sequence_per_sat": (M.SequenceController(u=tabs, tf_u=1.0, tf_sim=1.0), dict(kind=C.CTRL_SEQUENCE, table=tabs, end_tau=1.0)),
    }[kind]
    y, u, t, st = M.propagate_batch(y0, tfv, ctrl, const, include_drag=True, include_J2=True, T=T, n_sub=9)
    yr, ur, sr = C.propagate_batch(y0, tfv, const, include_drag=True, include_J2=True, T=T, n_sub=9, **ckw)
    assert st.max() == 0 and sr.max() == 0
    assert rel_err(y, yr) < TOL_ORACLE and rel_err(u, ur) < 1e-10
    assert np.array_equal(t, np.linspace(0, 1, T))


def test_edge_shapes_and_mass_failure(M, const):
    y0, _, _ = synth_batch(3, 2, 1.0, const)
    y, u, t, st = M.propagate_batch(y0, 1.0, M.Controller(), const, T=1)
    assert y.shape == (3, 7, 1) and np.array_equal(y[:, :, 0], y0)
    y, u, t, st = M.propagate_batch(y0, 1.0, M.Controller(), const, T=1, n_sub=4)
    assert y.shape == (3, 7, 1) and np.array_equal(y[:, :, 0], y0)
    y, u, t, st = M.propagate_batch(y0[:0], 1.0, M.Controller(), const, T=10)
    assert y.shape == (0, 7, 10)
    y0 = y0.copy()
    y0[1, 6] = 1e-3
    with pytest.raises(Exception, match="INVALID SATELLITE MASS"):
        M.propagate_batch(y0, 5.0, M.ConstantThrustController(thrust=np.array([0.05, 0, 0])), const, T=50)
    y, u, t, st = M.propagate_batch(y0, 5.0, M.ConstantThrustController(thrust=np.array([0.05, 0, 0])), const, T=50, check=False)
    assert list(st) == [0, 1, 0] and np.all(np.isfinite(y[0])) and np.any(np.isnan(y[1]))
    y, u, t, st = M.propagate_batch(y0, 5.0, M.ConstantThrustController(thrust=np.array([0.05, 0, 0])), const, T=50, check=False, n_sub=21)
    assert list(st) == [0, 1, 0] and np.all(np.isfinite(y[0])) and np.any(np.isnan(y[1]))
    y0[2, 3] = np.nan       # a poisoned state ends like scipy would: step-size underflow, flagged
    with pytest.raises(RuntimeError, match="step size"):
        M.propagate_batch(y0[2:], 1.0, M.Controller(), const, T=20)


def test_device_tensor_api(M, const):
    import torch
    y0, _, _ = synth_batch(9, 2, 1.0, const)
    c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    yh, uh, _, _ = M.propagate_batch(y0, 2.0, c, const, include_drag=False, include_J2=False, T=50)
    dev = torch.device("cuda:0")
    y, u, st = M.propagate_batch_device(torch.from_numpy(y0).to(dev), torch.full((9,), 2.0, dtype=torch.float64, device=dev),
                                        c, const, include_drag=False, include_J2=False, T=50)
    torch.cuda.synchronize()
    assert np.array_equal(y.cpu().numpy(), yh) and np.array_equal(u.cpu().numpy(), uh) and int(st.max()) == 0


@pytest.mark.parametrize("n_prop", [0, 5])     # 0: the replayed RK45 (default), 5: fixed-step RK4
@pytest.mark.parametrize("case", ["windows_default", "windows_ragged_j2", "sequence_mass_failure", "small_batch", "odd_panels"])
def test_overlapped_propagate_discretize_is_bit_identical(M, const, case, n_prop):
    """mpc_propagate_discretize (propagation hidden behind the discretization, windows along k gated by stream memory
    operations) against the two kernels run back to back: same arithmetic, so every output must be bit-identical."""
    import torch
    dev = torch.device("cuda:0")
    N, T, tf, n_sub, nw, j2 = {"windows_default": (2048, 120, 1.5, 20, 0, False),
                               "windows_ragged_j2": (1500, 131, 2.0, 10, 7, True),
                               "sequence_mass_failure": (1280, 100, 1.0, 10, 4, False),
                               "small_batch": (5, 17, 0.5, 10, 0, False),
                               "odd_panels": (1300, 90, 1.0, 9, 3, False)}[case]
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    rng = np.random.default_rng(11)
    tfv = tf * (1 + 0.05 * rng.random(N))
    c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    if case == "sequence_mass_failure":
        c = M.SequenceController(u=0.3 * rng.standard_normal((N, 3, 9)), tf_u=1.0, tf_sim=1.0)
        y0 = y0.copy()
        y0[[3, 700, N - 1], 6] = 1e-4          # these satellites run out of mass: NaN trajectories, status set
    y0d, tfd = torch.from_numpy(y0).to(dev), torch.from_numpy(tfv).to(dev)
    y_a, u_a, sp_a = M.propagate_batch_device(y0d, tfd, c, const, include_drag=False, include_J2=j2, T=T, n_sub=n_prop)
    out_a, sd_a = M.discretize_batch_device(y_a, u_a, tfd, const, include_J2=j2, n_sub=n_sub)
    pitch = N * (T - 1) + 13
    out_b = torch.full((105, pitch), float("nan"), dtype=torch.float64, device=dev)
    for rep in range(2):                       # twice: the progress words must be re-armed by every call
        _, y_b, u_b, sp_b, sd_b = M.propagate_discretize_device(y0d, tfd, c, const, T, prop_J2=j2, disc_J2=j2,
                                                                n_sub_prop=n_prop, n_sub_disc=n_sub, out=out_b,
                                                                out_offset=6, n_windows=nw)
        torch.cuda.synchronize()
        eq = lambda a, b: bool(torch.equal(a.view(torch.int64), b.view(torch.int64)))   # NaN-aware bit comparison
        assert eq(y_a, y_b) and eq(u_a, u_b) and torch.equal(sp_a, sp_b) and torch.equal(sd_a, sd_b)
        assert eq(out_a, out_b[:, 6:6 + N * (T - 1)])
        assert bool(torch.isnan(out_b[:, :6]).all()) and bool(torch.isnan(out_b[:, 6 + N * (T - 1):]).all())
        out_b[:, 6:6 + N * (T - 1)] = float("nan")
    if case == "sequence_mass_failure":
        assert sorted(torch.nonzero(sp_b).flatten().tolist()) == [3, 700, N - 1] and int(sd_b.max()) > 0
    else:
        assert int(sp_b.max()) == 0 and int(sd_b.max()) == 0


@pytest.mark.parametrize("layout", ["kmajor", "satmajor"])
def test_fused_gather_entry_points_and_layouts_on_one_gpu(M, const, layout):
    """mpc_propagate_discretize_gather / mpc_discretize_batch_gather (options per call, k-major or satellite-major gathered
    layout) with two LOCAL buffers standing in for the ranks of an NVLink box: this rank's block lands in both, at its
    place inside a larger gathered buffer, bit-identical to the plain two-kernel sequence (k-major: permuted), margins
    untouched, for the overlapped windowed pass and for the plain launch."""
    import ctypes
    import torch
    from mpconstellation_b200 import _lib
    dev = torch.device("cuda:0")
    N, T, tf, n_sub, ntot, soff = 1500, 131, 2.0, 10, 2200, 300
    n = T - 1
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    rng = np.random.default_rng(5)
    tfv = tf * (1 + 0.05 * rng.random(N))
    c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    y0d, tfd = torch.from_numpy(y0).to(dev), torch.from_numpy(tfv).to(dev)
    y_a, u_a, sp_a = M.propagate_batch_device(y0d, tfd, c, const, include_drag=False, include_J2=True, T=T)
    out_a, sd_a = M.discretize_batch_device(y_a, u_a, tfd, const, include_J2=True, n_sub=n_sub)
    ref = out_a.view(105, N, n)
    want = torch.full((105, ntot * n), float("nan"), dtype=torch.float64, device=dev)
    if layout == "kmajor":
        want.view(105, n, ntot)[:, :, soff:soff + N] = ref.permute(0, 2, 1)
    else:
        want.view(105, ntot, n)[:, soff:soff + N, :] = ref
    eq = lambda a, b: bool(torch.equal(a.view(torch.int64), b.view(torch.int64)))   # NaN-aware bit comparison
    g = _lib.MpcGatherOpts(_lib.LAYOUT_K_MAJOR if layout == "kmajor" else _lib.LAYOUT_SAT_MAJOR, 0, 0, 0, ntot, soff)
    bufs = [torch.full((105, ntot * n), float("nan"), dtype=torch.float64, device=dev) for _ in range(2)]
    for rep in range(2):
        _, y_b, u_b, sp_b, sd_b = M.propagate_discretize_device(y0d, tfd, c, const, T, prop_J2=True, disc_J2=True,
                                                                n_sub_disc=n_sub, out=bufs[0], extra_dst=[bufs[1]],
                                                                n_windows=7, gather=g)
        torch.cuda.synchronize()
        assert eq(y_a, y_b) and eq(u_a, u_b) and torch.equal(sd_a, sd_b) and int(sp_b.max()) == 0
        for b in bufs:
            assert eq(b, want)
            b.fill_(float("nan"))
    # the plain (not overlapped) launch through mpc_discretize_batch_gather
    st = torch.empty(N * n, dtype=torch.int32, device=dev)
    p = _lib.make_params(const, True, False)
    arr = (ctypes.c_void_p * 2)(bufs[0].data_ptr(), bufs[1].data_ptr())
    _lib.check(_lib.lib().mpc_discretize_batch_gather(y_a.data_ptr(), u_a.data_ptr(), tfd.data_ptr(), ctypes.byref(p), N, T, n_sub,
                                                      arr, 2, ctypes.byref(g), st.data_ptr(),
                                                      torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize()
    assert eq(bufs[0], want) and eq(bufs[1], want) and torch.equal(st, sd_a)
    # bad options are refused
    bad = _lib.MpcGatherOpts(7, 0, 0, 0, ntot, soff)
    assert _lib.lib().mpc_discretize_batch_gather(y_a.data_ptr(), u_a.data_ptr(), tfd.data_ptr(), ctypes.byref(p), N, T, n_sub,
                                                  arr, 2, ctypes.byref(bad), st.data_ptr(), None) == _lib.E_INVALID
    small = _lib.MpcGatherOpts(0, 0, 0, 0, N - 1, 0)
    assert _lib.lib().mpc_discretize_batch_gather(y_a.data_ptr(), u_a.data_ptr(), tfd.data_ptr(), ctypes.byref(p), N, T, n_sub,
                                                  arr, 2, ctypes.byref(small), st.data_ptr(), None) == _lib.E_INVALID


@pytest.mark.parametrize("case", ["windows", "small", "mass_failure"])
def test_streamed_host_pass_in_the_k_major_layout(M, const, case):
    """propagate_discretize(layout="kmajor"): windows along k gated on the propagation's progress, each window read back
    while the next one runs; the per-satellite views, the trajectory, the inputs and the status words are bit-identical to
    the satellite-major host pass (same kernels per interval)"""
    N, T, tf, n_sub = {"windows": (1500, 131, 2.0, 20), "small": (7, 23, 0.7, 100), "mass_failure": (1300, 100, 1.0, 10)}[case]
    y0, _, _ = synth_batch(N, 2, 1.0, const)
    rng = np.random.default_rng(4)
    tfv = tf * (1 + 0.05 * rng.random(N))
    c = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    if case == "mass_failure":
        c = M.SequenceController(u=0.3 * rng.standard_normal((N, 3, 9)), tf_u=1.0, tf_sim=1.0)
        y0 = y0.copy()
        y0[[3, 700, N - 1], 6] = 1e-4
    a, ya, ua = M.propagate_discretize(y0, tfv, c, const, T=T, n_sub_disc=n_sub, disc_J2=True, prop_J2=True, check=False)
    for rep in range(2):
        b, yb, ub = M.propagate_discretize(y0, tfv, c, const, T=T, n_sub_disc=n_sub, disc_J2=True, prop_J2=True, check=False,
                                           layout="kmajor")
        def eq(p, q):      # bit comparison of everything that is a number; NaNs must sit in the same places (their sign /
            nan = np.isnan(p)                                                  # payload bits depend on the kernel that made them)
            return np.array_equal(nan, np.isnan(q)) and np.array_equal(p[~nan].view(np.int64), q[~nan].view(np.int64))
        assert b.layout == "kmajor" and eq(np.ascontiguousarray(ya), np.ascontiguousarray(yb)) and eq(np.ascontiguousarray(ua), np.ascontiguousarray(ub))
        assert np.array_equal(a.status, b.status)
        if case == "small":
            # (a batch this small runs the thread-group kernel in the satellite-major pass and the one-thread kernel in
            # the k-major one: equal up to the rounding of the former's cross-lane sums)
            for s_ in range(N):
                for p, q in zip(a.sat(s_), b.sat(s_)):
                    assert rel_err(q, p) < 1e-13
            continue
        for s_ in (0, N // 2, N - 1) + ((3, 700) if case == "mass_failure" else ()):
            for p, q in zip(a.sat(s_), b.sat(s_)):
                assert eq(np.ascontiguousarray(p), np.ascontiguousarray(q))
        # the whole result: the last (short) chunk of the satellite-major pipeline may run the thread-group kernel, so
        # rounding-level differences are allowed there; NaN patterns (failed satellites) must coincide
        ak = np.ascontiguousarray(a.soa.reshape(105, N, T - 1).transpose(0, 2, 1)).reshape(105, -1)
        bk = np.ascontiguousarray(b.soa)
        assert np.array_equal(np.isnan(ak), np.isnan(bk))
        fin = ~np.isnan(ak)
        assert np.max(np.abs(ak[fin] - bk[fin])) <= 1e-13 * np.max(np.abs(ak[fin]))
    if case == "mass_failure":
        assert a.status[3].max() == 1 and a.status[0].max() == 0
        with pytest.raises(Exception, match="INVALID SATELLITE MASS"):
            M.propagate_discretize(y0, tfv, c, const, T=T, n_sub_disc=n_sub, layout="kmajor")
    else:
        assert b.status.max() == 0
