"""Generates tests/golden/*.npz by running the UNMODIFIED reference (/root/reference).

Run in the build container only:  python tests/golden/make_golden.py
The reference pins no numeric result of its own on this path (its tests only plot),
so these fixtures -- outputs of the reference itself on its own test scenarios -- are
what pins oracle/ and, through it, the CUDA path.  Scenario sources are cited inline.
"""
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle.refshim import load_reference  # noqa: E402

R = load_reference()
F = R.Simulator.satellite_dynamics

# Hubble state, ref: test_discretizer.py:17-21
R_INIT = np.array([5371.4806, -4133.1393, 1399.9594]) * 1000
V_INIT = np.array([4.6921, 4.9848, -3.2752]) * 1000
M_INIT = 12200.0


def const_vec(c):
    return np.array([c.MU, c.R_E, c.J2, c.G0, c.ISP, c.S, c.R0, c.RHO])


def disc(const, x, u, tf, uniform, J2=False, steps=101, ks=None):
    d = R.Discretizer(const, include_J2=J2)
    d.use_uniform_steps = uniform
    d.integrator_steps = steps
    if ks is None:
        return d.discretize(F, x, u, tf)
    # in-process per-interval call (name-mangled privates, ref: linearize_discretize.py:115-116,357-358)
    K = x.shape[1]
    d._Discretizer__tau = np.linspace(0, 1, K)
    d._Discretizer__u = u
    opts = dict(use_uniform_steps=uniform, integrator_steps=steps, ivp_max_step=d.ivp_max_step, ivp_solver=d.ivp_solver)
    funcs = dict(dPhi_gen=d.dPhi_gen, f=F, u_func=d.u_func, B_func=d.B_func, Sigma_func=d.Sigma_func, xi_func=d.xi_func)
    res = [R.get_matrices(opts, funcs, tf, d._Discretizer__tau, x, k) for k in ks]
    return (np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[2] for r in res]),
            np.column_stack([r[3] for r in res]), np.column_stack([r[4] for r in res]))


def power_law_density(a, b, r0_m, r_e_m, rho_scale):
    """rho_func / drho_func for the drag linearisation from the power law the reference's authors tried for 400-600 km
    (simulator.py:110: `8E26 * altitude**-6.828`): density over rho_scale as a function of the NORMALIZED position, and
    its derivative with respect to the normalized radius (what linearize_discretize.py:166 multiplies r^T/|r| by)."""
    def rho(r):
        return a * (np.linalg.norm(r) * r0_m - r_e_m) ** b / rho_scale

    def drho(r):
        return a * b * (np.linalg.norm(r) * r0_m - r_e_m) ** (b - 1.0) * r0_m / rho_scale
    return rho, drho


def disc_drag(const, x, u, tf, uniform, rho_n, J2=False, ks=None, funcs_rho=None):
    """Discretizer with the drag branch enabled (linearize_discretize.py:160-169).  The shipped reference cannot
    reach it with its defaults (rho_func=None, Constants has no CD) but runs it once the caller supplies what the
    branch reads: const.CD, rho_func, drho_func.  Constant density, as Simulator.get_atmo_density (simulator.py:112)."""
    rho_f, drho_f = funcs_rho if funcs_rho is not None else ((lambda r: rho_n), (lambda r: 0.0))
    d = R.Discretizer(const, rho_func=rho_f, drho_func=drho_f, include_drag=True, include_J2=J2)
    d.use_uniform_steps = uniform
    K = x.shape[1]
    d._Discretizer__tau = np.linspace(0, 1, K)
    d._Discretizer__u = u
    opts = dict(use_uniform_steps=uniform, integrator_steps=d.integrator_steps, ivp_max_step=d.ivp_max_step,
                ivp_solver=d.ivp_solver)
    funcs = dict(dPhi_gen=d.dPhi_gen, f=F, u_func=d.u_func, B_func=d.B_func, Sigma_func=d.Sigma_func, xi_func=d.xi_func)
    res = [R.get_matrices(opts, funcs, tf, d._Discretizer__tau, x, k) for k in ks]
    return (np.stack([r[0] for r in res]), np.stack([r[1] for r in res]), np.stack([r[2] for r in res]),
            np.column_stack([r[3] for r in res]), np.column_stack([r[4] for r in res]))


def drag_fixtures():
    """tests/golden/discretize_drag.npz: tangential-thrust trajectory flown WITH drag, discretized with drag.
    g0: the physical parameters (S as scaled by SatelliteScale, C_D 2.5, 500 km density);
    g1: S x 1e4 (drag ~4e-4 of gravity, so that every drag term is well above the parity tolerance) + J2, and an
        A-side density 1.7x the one the dynamics use (the two are independent inputs of the reference)."""
    import constants as ref_constants
    import simulator as ref_simulator
    sat = R.Satellite(R_INIT, V_INIT, M_INIT)
    scale = R.SatelliteScale(sat=sat)
    c = R.ConstantTangentialThrustController([sat], 0.5)
    g = {}
    for tag, s_mult, rho_mult, J2 in (("g0", 1.0, 1.0, False), ("g1", 1e4, 1.7, True)):
        const = scale.get_normalized_constants()
        const.S = const.S * s_mult
        const.CD = ref_constants.C_D
        rho_n = ref_simulator.Simulator.get_atmo_density(np.array([1.0, 0.0, 0.0]), const.R0) / const.RHO * rho_mult
        sim = R.Simulator(sats=[R.Satellite(R_INIT, V_INIT, M_INIT)], controller=c, scale=scale, base_res=40,
                          include_drag=True, include_J2=J2)
        # the simulator asks its scale for the constants on every call: hand it the modified bag
        scale_mod = type("ScaleWithS", (), {"get_normalized_constants": lambda self: const,
                                            "normalize_state": scale.normalize_state,
                                            "redim_state": scale.redim_state})()
        sim.scale = scale_mod
        sim.run(tf=1)
        sid = sim.sats[0].id
        x, t = sim.sim_data[sid], sim.sim_time[sid]
        u = R.Discretizer.extract_uk(x, t, c)
        ks = list(range(0, x.shape[1] - 1, 3))
        g.update({f"{tag}_x": x, f"{tag}_u": u, f"{tag}_tf": 1.0, f"{tag}_ks": np.array(ks),
                  f"{tag}_const": const_vec(const), f"{tag}_cd": const.CD, f"{tag}_rho_n": rho_n, f"{tag}_j2": int(J2)})
        g.update(pack(f"{tag}_uni", disc_drag(const, x, u, 1.0, True, rho_n, J2=J2, ks=ks)))
        g.update(pack(f"{tag}_def", disc_drag(const, x, u, 1.0, False, rho_n, J2=J2, ks=ks)))
    np.savez(os.path.join(HERE, "discretize_drag.npz"), **g)
    print("discretize_drag.npz", os.path.getsize(os.path.join(HERE, "discretize_drag.npz")) // 1024, "KiB")


def drag_radial_fixtures():
    """tests/golden/discretize_drag_radial.npz: the g1 scenario of drag_fixtures (S x 1e4, J2, trajectory flown with drag)
    discretized with an ALTITUDE-DEPENDENT density in the drag branch of the linearisation: rho_func = the power law of
    simulator.py:110, drho_func = its derivative -- so that Dr_aD (linearize_discretize.py:166) is not zero.  The
    dynamics keep the simulator's constant density, as in the reference (f is Simulator.satellite_dynamics)."""
    import constants as ref_constants
    sat = R.Satellite(R_INIT, V_INIT, M_INIT)
    scale = R.SatelliteScale(sat=sat)
    c = R.ConstantTangentialThrustController([sat], 0.5)
    const = scale.get_normalized_constants()
    const.S = const.S * 1e4
    const.CD = ref_constants.C_D
    a, b = 8e26, -6.828
    r0_m, r_e_m = float(const.R0), float(ref_constants.R_EARTH)
    funcs_rho = power_law_density(a, b, r0_m, r_e_m, float(const.RHO))
    sim = R.Simulator(sats=[R.Satellite(R_INIT, V_INIT, M_INIT)], controller=c, scale=scale, base_res=40,
                      include_drag=True, include_J2=True)
    scale_mod = type("ScaleWithS", (), {"get_normalized_constants": lambda self: const,
                                        "normalize_state": scale.normalize_state, "redim_state": scale.redim_state})()
    sim.scale = scale_mod
    sim.run(tf=1)
    sid = sim.sats[0].id
    x, t = sim.sim_data[sid], sim.sim_time[sid]
    u = R.Discretizer.extract_uk(x, t, c)
    ks = list(range(0, x.shape[1] - 1, 3))
    g = {"x": x, "u": u, "tf": 1.0, "ks": np.array(ks), "const": const_vec(const), "cd": const.CD, "j2": 1,
         "law": np.array([a, b, r0_m, r_e_m, float(const.RHO)]),
         "rho_at_x0": funcs_rho[0](x[0:3, 0]), "drho_at_x0": funcs_rho[1](x[0:3, 0])}
    g.update(pack("uni", disc_drag(const, x, u, 1.0, True, None, J2=True, ks=ks, funcs_rho=funcs_rho)))
    g.update(pack("def", disc_drag(const, x, u, 1.0, False, None, J2=True, ks=ks, funcs_rho=funcs_rho)))
    # the same with drho_func = 0 (what a caller gets who passes the density but forgets its gradient): pins that the
    # gradient term is what differs
    g.update(pack("def_nograd", disc_drag(const, x, u, 1.0, False, None, J2=True, ks=ks,
                                          funcs_rho=(funcs_rho[0], lambda r: 0.0))))
    np.savez(os.path.join(HERE, "discretize_drag_radial.npz"), **g)
    print("discretize_drag_radial.npz", os.path.getsize(os.path.join(HERE, "discretize_drag_radial.npz")) // 1024, "KiB")


def bench_fixtures():
    """tests/golden/bench_workload.npz: satellites of the BENCHMARK's own synthetic constellation (bench.make_constellation,
    SURVEY 8d: Hubble state rotated about z by 2 pi i / N, speed scaled by 1 + 0.1 U[0,1), N = 4096) flown and
    discretized by the unmodified reference exactly as bench.py's step does it (tangential thrust 0.5, tf = 2,
    base_res = 100 -> K = 200, no drag / J2; use_uniform_steps=True, integrator_steps=101, and the default mode),
    so that the numbers the benchmark produces at full size are pinned to the reference itself, not only to the oracle."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
    HUBBLE = np.concatenate([R_INIT, V_INIT, [M_INIT]])
    N, tf = 4096, 2.0
    sat0 = R.Satellite(R_INIT, V_INIT, M_INIT)
    scale = R.SatelliteScale(sat=sat0)                    # one scale from satellite 0 (test_simulator.py:49)
    const = scale.get_normalized_constants()
    # bench.make_constellation, restated with the reference's own scale object
    y = scale.normalize_state(HUBBLE)
    rng = np.random.default_rng(20240531)
    ang = 2 * np.pi * np.arange(N) / N
    ca, sa = np.cos(ang), np.sin(ang)
    f = 1 + 0.1 * rng.random(N)
    Y = np.tile(y, (N, 1))
    Y[:, 0], Y[:, 1] = ca * y[0] - sa * y[1], sa * y[0] + ca * y[1]
    Y[:, 3], Y[:, 4] = (ca * y[3] - sa * y[4]) * f, (sa * y[3] + ca * y[4]) * f
    Y[:, 5] = y[5] * f
    idx = np.array([0, 1365, 2730, 4095])
    ks = np.array([0, 23, 57, 101, 150, 198])
    g = {"const": const_vec(const), "idx": idx, "ks": ks, "y0": Y[idx], "tf": tf, "n_sats": N}
    c = R.ConstantTangentialThrustController(tangential_thrust=0.5)
    for j, i in enumerate(idx):
        yd = scale.redim_state(Y[i])
        s = R.Satellite(yd[0:3], yd[3:6], float(yd[6]))
        sim = R.Simulator(sats=[s], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
        sim.run(tf=tf)
        x, t = sim.sim_data[s.id], sim.sim_time[s.id]
        u = R.Discretizer.extract_uk(x, t, c)
        g.update({f"s{j}_x": x, f"s{j}_u": u})
        g.update(pack(f"s{j}_uni", disc(const, x, u, tf, True, ks=list(ks))))
        g.update(pack(f"s{j}_def", disc(const, x, u, tf, False, ks=list(ks))))
        print("satellite", i, "done")
    np.savez(os.path.join(HERE, "bench_workload.npz"), **g)
    print("bench_workload.npz", os.path.getsize(os.path.join(HERE, "bench_workload.npz")) // 1024, "KiB")


def many_fixtures():
    """tests/golden/discretize_many.npz (verdict r1, missing item 5):
    m0: the reference's own test_linearize_many call (test_discretizer.py:88-117) in its DEFAULT mode -- constant-thrust
        trajectory, K=100, once with the matching u (3,K) and once with the test's own malformed u = np.tile(T_init,
        (3, K)) of shape (3, 3K) (u on its own grid), sampled intervals;
    m1: BASELINE configs[1] size -- 4 of the 64 satellites of bench.make_constellation(64) at K=100, tf=1, flown and
        discretized by the unmodified reference in both quadrature modes (sampled intervals)."""
    sat = R.Satellite(R_INIT, V_INIT, M_INIT)
    scale = R.SatelliteScale(sat=sat)
    const = scale.get_normalized_constants()
    T_init = np.array([0.44, 0.7, 1.0])
    g = {"const": const_vec(const)}
    sim = R.Simulator(sats=[sat], controller=R.ConstantThrustController([sat], T_init), scale=scale,
                      base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=1)
    x = sim.sim_data[sat.id]
    K = x.shape[1]
    u = np.tile(T_init.reshape(3, 1), (1, K))
    uq = np.tile(T_init, (3, K))
    ks = [0, 1, 17, 50, 77, 98]
    g.update(m0_x=x, m0_u=u, m0_uq=uq, m0_tf=1.0, m0_ks=np.array(ks))
    g.update(pack("m0_def", disc(const, x, u, 1, False, ks=ks)))
    g.update(pack("m0q_def", disc(const, x, uq, 1, False, ks=ks)))
    HUBBLE = np.concatenate([R_INIT, V_INIT, [M_INIT]])
    N, tf = 64, 1.0
    y = scale.normalize_state(HUBBLE)
    rng = np.random.default_rng(20240531)
    ang = 2 * np.pi * np.arange(N) / N
    ca, sa = np.cos(ang), np.sin(ang)
    f = 1 + 0.1 * rng.random(N)
    Y = np.tile(y, (N, 1))
    Y[:, 0], Y[:, 1] = ca * y[0] - sa * y[1], sa * y[0] + ca * y[1]
    Y[:, 3], Y[:, 4] = (ca * y[3] - sa * y[4]) * f, (sa * y[3] + ca * y[4]) * f
    Y[:, 5] = y[5] * f
    idx = np.array([0, 21, 42, 63])
    ks = np.array([0, 13, 49, 98])
    g.update(m1_idx=idx, m1_ks=ks, m1_y0=Y[idx], m1_tf=tf, m1_n_sats=N)
    c = R.ConstantTangentialThrustController(tangential_thrust=0.5)
    for j, i in enumerate(idx):
        yd = scale.redim_state(Y[i])
        s = R.Satellite(yd[0:3], yd[3:6], float(yd[6]))
        sim = R.Simulator(sats=[s], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
        sim.run(tf=tf)
        xx, tt = sim.sim_data[s.id], sim.sim_time[s.id]
        uu = R.Discretizer.extract_uk(xx, tt, c)
        g.update({f"m1_s{j}_x": xx, f"m1_s{j}_u": uu})
        g.update(pack(f"m1_s{j}_uni", disc(const, xx, uu, tf, True, ks=list(ks))))
        g.update(pack(f"m1_s{j}_def", disc(const, xx, uu, tf, False, ks=list(ks))))
    np.savez(os.path.join(HERE, "discretize_many.npz"), **g)
    print("discretize_many.npz", os.path.getsize(os.path.join(HERE, "discretize_many.npz")) // 1024, "KiB")


def coast_fixtures():
    """tests/golden/discretize_coast.npz: a coast-to-thrust switch with u EXACTLY zero on the coasting nodes (what a
    bang-off-bang plan looks like).  The reference looks the end nodes of every interval up on the global grid
    (u_FOH, linearize_discretize.py:308-315): tau_6 = 0.6000000000000001 lands in [6, 7], where it interpolates
    1e-15 (u_7 - u_6) -- above eps, so B_func's |u| <= eps guard (:208) does not fire and the mass row of B at that node
    carries the unit direction of u_7 although u_6 = 0.  Pins the kernels' end-node lookup (ref_node_input)."""
    sat = R.Satellite(R_INIT, V_INIT, M_INIT)
    scale = R.SatelliteScale(sat=sat)
    const = scale.get_normalized_constants()
    g = {"const": const_vec(const)}
    for tag, K, tf, n_coast in (("c11", 11, 0.6, 7), ("c24", 24, 1.3, 10)):
        c = R.ConstantTangentialThrustController([sat], 0.5)
        sim = R.Simulator(sats=[R.Satellite(R_INIT, V_INIT, M_INIT)], controller=c, scale=scale, base_res=int(round(K / tf)) + 1,
                          include_drag=False, include_J2=False)
        sim.eval_points = K
        s0 = sim.sats[0]
        sol = sim.get_trajectory_ODE(s0, tf, c.get_u_func())
        x, t = sol.y, sol.t
        u = R.Discretizer.extract_uk(x, t, c)
        u[:, :n_coast] = 0.0                       # coast on the first nodes, thrust afterwards
        u[:, -2:] = 0.0                            # ... and a thrust-to-coast switch at the end
        g.update({f"{tag}_x": x, f"{tag}_u": u, f"{tag}_tf": tf})
        g.update(pack(f"{tag}_uni", disc(const, x, u, tf, True)))
        g.update(pack(f"{tag}_def", disc(const, x, u, tf, False)))
    np.savez(os.path.join(HERE, "discretize_coast.npz"), **g)
    print("discretize_coast.npz", os.path.getsize(os.path.join(HERE, "discretize_coast.npz")) // 1024, "KiB")


def pack(prefix, out):
    names = ["A_k", "B_kp", "B_kn", "Sigma_k", "xi_k"]
    return {f"{prefix}_{n}": o for n, o in zip(names, out)}


def main():
    sat = R.Satellite(R_INIT, V_INIT, M_INIT)
    scale = R.SatelliteScale(sat=sat)
    const = scale.get_normalized_constants()
    g = {"const": const_vec(const), "x0_dim": sat.get_state_vector(),
         "scale": np.array([scale._r0, scale._s0, scale._v0, scale._a0, scale._m0, scale._T0, scale._mu0])}
    T_init = np.array([0.44, 0.7, 1.0])

    # --- D0: ref test_discretizer.py:30-54 (K=2, DIMENSIONAL x with normalized constants, tf=1)
    x = np.column_stack([sat.get_state_vector()] * 2)
    u = np.column_stack([T_init] * 2)
    g.update(d0_x=x, d0_u=u, d0_tf=1.0)
    g.update(pack("d0_def", disc(const, x, u, 1, False)))
    g.update(pack("d0_uni", disc(const, x, u, 1, True)))

    # --- D1: ref test_discretizer.py:57-85 (K=3 identical columns, tf=0.1)
    xn = scale.normalize_state(sat.get_state_vector())
    x = np.column_stack([xn] * 3)
    u = np.column_stack([T_init] * 3)
    g.update(d1_x=x, d1_u=u, d1_tf=0.1)
    g.update(pack("d1_def", disc(const, x, u, 0.1, False)))
    g.update(pack("d1_uni", disc(const, x, u, 0.1, True)))

    # --- D2: ref test_discretizer.py:88-117 (constant thrust, tf=1, base_res=100 -> K=100)
    sim = R.Simulator(sats=[sat], controller=R.ConstantThrustController([sat], T_init), scale=scale,
                      base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=1)
    x = sim.sim_data[sat.id]
    K = x.shape[1]
    u = np.tile(T_init.reshape(3, 1), (1, K))
    out = disc(const, x, u, 1, True)
    # The test's own u is np.tile(T_init,(3,K)) (:103): shape (3,3K) whose COLUMNS cycle through
    # [.44]*3,[.7]*3,[1.]*3 -- not the constant thrust the trajectory was flown with.  The reference
    # accepts it (u_FOH uses u's own column count for its grid, :308-315); keep it as a fixture of the
    # "u on a different grid than x" behaviour.
    uq = np.tile(T_init, (3, K))
    g.update(d2q_u=uq, d2q_ks=np.array([0, 50, 98]))
    g.update(pack("d2q_uni", disc(const, x, uq, 1, True, ks=[0, 50, 98])))
    g.update(d2_x=x, d2_u=u, d2_tf=1.0, d2_t=sim.sim_time[sat.id])
    g.update(pack("d2_uni", out))

    # --- D3: ref test_discretizer.py:120-150 (tangential 0.5, tf=2 -> K=200); also the MPC seed (control.py:178-180)
    c = R.ConstantTangentialThrustController([sat], 0.5)
    sim = R.Simulator(sats=[sat], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=2)
    x = sim.sim_data[sat.id]
    t = sim.sim_time[sat.id]
    u = R.Discretizer.extract_uk(x, t, c)
    g.update(d3_x=x, d3_u=u, d3_tf=2.0, d3_t=t)
    g.update(pack("d3_def", disc(const, x, u, 2, False)))
    g.update(pack("d3_uni", disc(const, x, u, 2, True)))
    ks = list(range(0, 199, 10))
    g["d3_j2_ks"] = np.array(ks)
    g.update(pack("d3_j2_uni", disc(const, x, u, 2, True, J2=True, ks=ks)))
    g.update(pack("d3_j2_def", disc(const, x, u, 2, False, J2=True, ks=ks)))
    g["d3_n21_ks"] = np.array(ks)
    g.update(pack("d3_n21_uni", disc(const, x, u, 2, True, steps=21, ks=ks)))

    # --- D4: BASELINE config 1 (single satellite, K=50): tangential 0.5, tf=0.5, base_res=100
    sim = R.Simulator(sats=[sat], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=0.5)
    x = sim.sim_data[sat.id]
    t = sim.sim_time[sat.id]
    u = R.Discretizer.extract_uk(x, t, c)
    g.update(d4_x=x, d4_u=u, d4_tf=0.5, d4_t=t)
    g.update(pack("d4_def", disc(const, x, u, 0.5, False)))
    g.update(pack("d4_uni", disc(const, x, u, 0.5, True)))
    np.savez(os.path.join(HERE, "discretize.npz"), **g)

    # ------------------------------------------------------------------ propagation
    p = {"const": const_vec(const), "x0_dim": sat.get_state_vector()}
    # P0: ref test_simulator.py:17-33 (5-orbit coast, default drag+J2)
    s = R.Satellite(R_INIT, V_INIT, M_INIT)
    sim = R.Simulator(sats=[s], scale=scale)
    sim.run(tf=5)
    p.update(p0_y=sim.sim_data[s.id], p0_t=sim.sim_time[s.id], p0_tf=5.0)
    # P1: tangential 0.5, tf=2, no drag / no J2 (test_discretizer.py:120-131)
    s = R.Satellite(R_INIT, V_INIT, M_INIT)
    sim = R.Simulator(sats=[s], controller=c, scale=scale, base_res=100, include_drag=False, include_J2=False)
    sim.run(tf=2)
    p.update(p1_y=sim.sim_data[s.id], p1_t=sim.sim_time[s.id], p1_tf=2.0, p1_mag=0.5,
             p1_u=R.Discretizer.extract_uk(sim.sim_data[s.id], sim.sim_time[s.id], c))
    # P2: constant thrust [0,0,0.1], drag+J2 (test_simulator.py:175-188, shortened to 3 orbits)
    s = R.Satellite(R_INIT, V_INIT, M_INIT)
    sim = R.Simulator(sats=[s], controller=R.ConstantThrustController(thrust=np.array([0., 0., 0.1])), scale=scale)
    sim.run(tf=3)
    p.update(p2_y=sim.sim_data[s.id], p2_t=sim.sim_time[s.id], p2_tf=3.0, p2_thrust=np.array([0., 0., 0.1]))
    # P3: three perturbed satellites v0*(1+0.1*rand) (test_simulator.py:36-55), one scale from sat 0
    rng = np.random.default_rng(20240531)
    fac = 1 + 0.1 * rng.random(3)
    sats = [R.Satellite(R_INIT, V_INIT * f, M_INIT) for f in fac]
    sc3 = R.SatelliteScale(sat=sats[0])
    sim = R.Simulator(sats=sats, scale=sc3)
    sim.run(tf=5)
    p.update(p3_fac=fac, p3_const=const_vec(sc3.get_normalized_constants()),
             p3_y0_dim=np.stack([q.get_state_vector() for q in sats]),
             p3_y=np.stack([sim.sim_data[q.id] for q in sats]), p3_t=sim.sim_time[sats[0].id], p3_tf=5.0)
    # P4: tangential 0.1 with drag+J2 (test_simulator.py:190-203), 2 orbits
    s = R.Satellite(R_INIT, V_INIT, M_INIT)
    c01 = R.ConstantTangentialThrustController(tangential_thrust=0.1)
    sim = R.Simulator(sats=[s], controller=c01, scale=scale)
    sim.run(tf=2)
    p.update(p4_y=sim.sim_data[s.id], p4_t=sim.sim_time[s.id], p4_tf=2.0, p4_mag=0.1)
    # P5: SequenceController (control.py:86-143) -- table shorter than the run (end_tau = 0.75)
    kk = 30
    tt = np.linspace(0, 1, kk)
    u_tab = np.vstack([0.3 * np.cos(2 * np.pi * tt), 0.4 * np.sin(3 * tt) + 0.1, 0.05 * (tt - 0.5)])
    s = R.Satellite(R_INIT, V_INIT, M_INIT)
    cs = R.SequenceController(u=u_tab, tf_u=1.5, tf_sim=2.0)
    sim = R.Simulator(sats=[s], controller=cs, scale=scale, base_res=60, include_drag=False, include_J2=False)
    sim.run(tf=2)
    p.update(p5_y=sim.sim_data[s.id], p5_t=sim.sim_time[s.id], p5_tf=2.0, p5_u_tab=u_tab, p5_tf_u=1.5,
             p5_u=R.Discretizer.extract_uk(sim.sim_data[s.id], sim.sim_time[s.id], cs))
    # P6: run_segments bookkeeping, two satellites, tangential 0.5, tf=3 in 4 segments (test_simulator.py:149-173)
    s1 = R.Satellite(R_INIT, V_INIT, M_INIT)
    s2 = R.Satellite(R_INIT, V_INIT * 1.1, M_INIT)
    cc = R.ConstantTangentialThrustController(sats=[s1, s2], tangential_thrust=0.5)
    sim = R.Simulator(sats=[s1, s2], scale=scale, base_res=100, controller=cc)
    sim.run_segments(tf=3, num_segments=4)
    p.update(p6_y=np.stack([sim.sim_data[s1.id], sim.sim_data[s2.id]]),
             p6_t=np.stack([sim.sim_time[s1.id], sim.sim_time[s2.id]]),
             p6_final_dim=np.stack([s1.get_state_vector(), s2.get_state_vector()]))
    np.savez(os.path.join(HERE, "propagate.npz"), **p)
    # ------------------------------------------------------------------ constraint terms (optimizer.py:80-170)
    import optimizer as ref_optimizer
    gd = np.load(os.path.join(HERE, "discretize.npz"))
    zero_u = np.zeros((3, gd["d4_x"].shape[1]))
    cases = {"c0": (gd["d3_x"], gd["d3_u"]), "c1": (gd["d2_x"], gd["d2_u"]), "c2": (gd["d4_x"], zero_u)}
    ct = {"MU": const.MU}
    for tag, (xb, ub) in cases.items():
        opt = ref_optimizer.Optimizer([xb], [ub], [np.zeros_like(xb)], 1.0, None, None, scale, verbose=False)
        out = opt.get_constraint_terms()
        ct[tag + "_x"], ct[tag + "_u"] = xb, ub
        for key, val in out.items():
            ct[f"{tag}_{key}"] = np.asarray(val[0])
    np.savez(os.path.join(HERE, "constraint_terms.npz"), **ct)
    for f in ("discretize.npz", "propagate.npz", "constraint_terms.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)) // 1024, "KiB")


if __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "coast":
    coast_fixtures()
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "many":
    many_fixtures()
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "drag":
    drag_fixtures()
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "drag_radial":
    drag_radial_fixtures()
elif __name__ == "__main__" and len(sys.argv) > 1 and sys.argv[1] == "bench":
    bench_fixtures()
elif __name__ == "__main__":
    main()
