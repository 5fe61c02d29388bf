"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SCvx discretize / propagate hot path.

This module is a plain numpy/scipy restatement of the reference algorithm
(rgovindjee/mpconstellation).  It exists to CHECK the CUDA path; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import it.  The product package (mpconstellation_b200/) never does.

Pinned: tests/test_oracle_golden.py compares every function here against
fixtures produced by the unmodified reference (tests/golden/make_golden.py,
run in the build container where /root/reference is mounted).

The third-party arithmetic the reference leans on -- scipy.integrate.solve_ivp
(RK45), numpy.linalg.inv, numpy.trapz -- is not vendored by the reference and
is not pinned by it (no requirements file); this oracle calls the same
libraries from this image (scipy 1.18.1, numpy 2.3.5) with the same arguments
the reference passes.

All `ref:` citations are file:line under /root/reference.
"""
from __future__ import annotations

import multiprocessing as mp
from dataclasses import dataclass
from functools import partial

import numpy as np
from scipy import integrate

# ref: constants.py:1-8 (dimensional values; the hot path only sees normalized ones)
MU_EARTH = 3.986004418e14
R_EARTH = 6.371e6
J2_EARTH = 1.08262668e-3
G0_EARTH = 9.80665
ISP_DEFAULT = 500.0
C_D = 2.5
S_AREA = 55.44
RHO_ATMO_500KM = 9.983e-13  # ref: simulator.py:112 (constant density model)


@dataclass
class OracleConstants:
    """Normalized constants bag.  ref: constants.py:11-20, satellite_scale.py:36-44."""
    MU: float
    R_E: float
    J2: float
    G0: float
    ISP: float
    S: float
    R0: float
    RHO: float


def scale_factors(x_dim):
    """Designer units from a dimensional state.  ref: satellite_scale.py:27-34."""
    r0 = float(np.linalg.norm(x_dim[0:3]))
    s0 = 2 * np.pi * np.sqrt(r0 ** 3 / MU_EARTH)
    return dict(r0=r0, s0=s0, v0=r0 / s0, a0=r0 / s0 ** 2, m0=float(x_dim[6]),
                T0=float(x_dim[6]) * r0 / s0 ** 2, mu0=r0 ** 3 / s0 ** 2)


def normalized_constants(sf) -> OracleConstants:
    """ref: satellite_scale.py:36-44."""
    return OracleConstants(MU=MU_EARTH / sf["mu0"], R_E=R_EARTH / sf["r0"], J2=J2_EARTH,
                           G0=G0_EARTH / sf["a0"], ISP=ISP_DEFAULT / sf["s0"], S=S_AREA / sf["r0"] ** 2,
                           R0=sf["r0"], RHO=sf["m0"] / sf["r0"] ** 3)


def normalize_state(x, sf):
    """ref: satellite_scale.py:62-78 (1-D or 7xN)."""
    x = np.asarray(x, dtype=float)
    out = np.array(x, dtype=float, copy=True)
    out[0:3] = x[0:3] / sf["r0"]
    out[3:6] = x[3:6] / sf["v0"]
    out[6] = x[6] / sf["m0"]
    return out


def redim_state(x, sf):
    """ref: satellite_scale.py:46-60."""
    x = np.asarray(x, dtype=float)
    out = np.array(x, dtype=float, copy=True)
    out[0:3] = x[0:3] * sf["r0"]
    out[3:6] = x[3:6] * sf["v0"]
    out[6] = x[6] * sf["m0"]
    return out


# --------------------------------------------------------------------------- dynamics

def dynamics(y, u, tf, const, include_drag=True, include_J2=True):
    """Right-hand side tf*f(y,u) with the thrust already evaluated.

    ref: simulator.py:115-161.  Raises like the reference on non-positive mass
    (simulator.py:135-136).
    """
    r = y[0:3]
    v = y[3:6]
    m = y[6]
    if m <= 0:
        raise Exception(f"ERROR: INVALID SATELLITE MASS: {m}")
    rn = np.linalg.norm(r)
    dy = np.zeros(7)
    dy[0:3] = v
    dy[3:6] = -const.MU / rn ** 3 * r + u / m
    if include_drag:
        # ref: simulator.py:150-153
        dy[3:6] += -1 / 2 * C_D * const.S * (1 / m) * (RHO_ATMO_500KM / const.RHO) * np.linalg.norm(v) * v
    if include_J2:
        # ref: simulator.py:154-158
        q = 5 * (r[2] / rn) ** 2
        shape = np.array([q - 1, q - 1, q - 3])
        dy[3:6] += 1.5 * const.J2 * const.MU * const.R_E ** 2 / rn ** 5 * (shape * r)
    dy[6] = -np.linalg.norm(u) / (const.G0 * const.ISP)
    return tf * dy


def jac_x(x, u, tf, const, include_J2=False, drag=None):
    """A = tf * d f / d x (7x7).  ref: linearize_discretize.py:119-183.  The drag branch (:160-169) is unreachable
    with the reference's defaults (Constants has no CD, rho_func is None); it runs once the caller supplies them:
    drag = (CD, rho, drho) with rho = rho_func(r), drho = drho_func(r) already evaluated."""
    r = x[0:3].reshape(3, 1)
    rx, ry, rz = x[0], x[1], x[2]
    rn = np.linalg.norm(r)
    m = x[6]
    T = np.asarray(u, dtype=float).reshape(3, 1)
    grav = -const.MU / rn ** 3 * np.eye(3) + 3 * const.MU / rn ** 5 * (r @ r.T)      # :146-147
    if include_J2:
        kJ2 = 1.5 * const.J2 * const.MU * const.R_E ** 2                              # :150
        zz = (rz / rn) ** 2
        GJ2 = np.diag([5 * zz - 1, 5 * zz - 1, 5 * zz - 3])                           # :152
        ddr = 5 * rz ** 2 * (-2 * (r.T / rn ** 4)) + (5 / rn ** 2) * np.array([[0, 0, 2 * rz]])  # :153-154
        j2 = ((kJ2 * GJ2 @ r) @ (-5 * r.T / rn ** 7)
              + kJ2 / rn ** 5 * np.vstack([rx * ddr, ry * ddr, rz * ddr])
              + kJ2 / rn ** 5 * GJ2)                                                   # :155-158
    else:
        j2 = np.zeros((3, 3))
    A = np.zeros((7, 7))
    A[0:3, 3:6] = np.eye(3)
    A[3:6, 0:3] = grav + j2
    A[3:6, 6:7] = -T / m ** 2                                                          # :175
    if drag is not None:
        CD, rho, drho = drag
        v = x[3:6].reshape(3, 1)
        vn = np.linalg.norm(v)
        A[3:6, 0:3] += ((-CD * const.S / (2 * m)) * vn * v) @ (drho * r.T / rn)        # :165
        A[3:6, 3:6] = ((-rho * CD * const.S) / (2 * m)) * (vn * np.eye(3) + (1 / vn) * (v @ v.T))   # :166-167
        A[3:6, 6:7] += ((rho * CD * const.S) / (2 * m ** 2)) * vn * v                  # :168
    return tf * A


def jac_u(x, u, tf, const):
    """B = tf * d f / d u (7x3).  ref: linearize_discretize.py:186-215."""
    m = x[6]
    T = np.asarray(u, dtype=float)
    B = np.zeros((7, 3))
    B[3:6, :] = np.eye(3) / m
    nT = np.linalg.norm(T)
    if nT > np.finfo(float).eps:                                                       # :208
        B[6, :] = -T / (const.G0 * const.ISP * nT)
    return tf * B


def foh(tau, u_nodes):
    """First-order hold over the global tau grid.  ref: linearize_discretize.py:294-315."""
    if tau == 1:
        return u_nodes[:, -1]
    K = u_nodes.shape[1]
    dtau = 1 / (K - 1)
    k = int(tau // dtau)
    lo = k / (K - 1)
    hi = (k + 1) / (K - 1)
    return (hi - tau) / (hi - lo) * u_nodes[:, k] + (tau - lo) / (hi - lo) * u_nodes[:, k + 1]


# --------------------------------------------------------------------------- discretization

def interval_matrices(k, x, u, tf, const, include_J2=False, use_uniform_steps=False,
                      integrator_steps=101, ivp_max_step=1e-2, ivp_solver="RK45", drag=None):
    """One interval: (A_k, B_kp, B_kn, Sigma_k, xi_k).  ref: linearize_discretize.py:8-82.
    drag = (CD, rho_n) enables include_drag with a constant density: the dynamics then carry the simulator's drag
    (global C_D, 500 km density, simulator.py:150-153) and A the terms of linearize_discretize.py:160-169."""
    dr = None if drag is None else (drag[0], drag[1], 0.0)
    has_drag = drag is not None
    K = x.shape[1]
    tau = np.linspace(0, 1, K)
    t0, t1 = tau[k], tau[k + 1]
    nodes = np.linspace(t0, t1, integrator_steps) if use_uniform_steps else None

    def rhs(t, y):
        # ref: linearize_discretize.py:262-290
        ut = foh(t, u)
        Phi = y[0:49].reshape(7, 7)
        xs = y[49:56]
        dPhi = jac_x(xs, ut, tf, const, include_J2, dr) @ Phi
        dx = dynamics(xs, ut, tf, const, include_drag=has_drag, include_J2=include_J2)
        return np.concatenate([dPhi.ravel(), dx])

    y0 = np.concatenate([np.eye(7).ravel(), x[:, k]])
    sol = integrate.solve_ivp(rhs, [t0, t1], y0, max_step=ivp_max_step, method=ivp_solver, t_eval=nodes)
    pts = nodes if use_uniform_steps else sol.t
    n = pts.size
    Phi_end = sol.y[0:49, -1].reshape(7, 7)
    Phi_all = sol.y[0:49, :].T.reshape(n, 7, 7)
    xs_all = sol.y[49:56, :]
    lam_n = (t1 - pts) / (t1 - t0)
    lam_p = (pts - t0) / (t1 - t0)
    Bs = np.zeros((n, 7, 3))
    Ss = np.zeros((7, n))
    Xs = np.zeros((7, n))
    for i, t in enumerate(pts):
        ut = foh(t, u)
        xi = xs_all[:, i]
        Bs[i] = jac_u(xi, ut, tf, const)                                               # :65
        Ss[:, i] = dynamics(xi, ut, 1, const, include_drag=has_drag, include_J2=include_J2)  # :66, :252-253
        Xs[:, i] = -(jac_x(xi, ut, tf, const, include_J2, dr) @ xi + Bs[i] @ ut)       # :67, :232-235
    Pinv = np.linalg.inv(Phi_all)                                                      # :69
    Bn_int = Pinv @ (Bs * lam_n[:, None, None])
    Bp_int = Pinv @ (Bs * lam_p[:, None, None])
    S_int = np.einsum("nij,jn->in", Pinv, Ss)
    X_int = np.einsum("nij,jn->in", Pinv, Xs)
    trap = np.trapezoid if hasattr(np, "trapezoid") else np.trapz
    B_kp = Phi_end @ trap(Bp_int, x=pts, axis=0)
    B_kn = Phi_end @ trap(Bn_int, x=pts, axis=0)
    Sig = Phi_end @ trap(S_int, x=pts, axis=1)
    Xi = Phi_end @ trap(X_int, x=pts, axis=1)
    return Phi_end, B_kp, B_kn, Sig, Xi


def discretize(x, u, tf, const, include_J2=False, use_uniform_steps=False, integrator_steps=101,
               ivp_max_step=1e-2, ivp_solver="RK45", processes=1):
    """All K-1 intervals of one satellite, reference shapes and return order
    (A_k, B_kp, B_kn, Sigma_k, xi_k).  ref: linearize_discretize.py:334-390.
    processes>1 fans the intervals over a process pool exactly as the reference does
    (a fresh mp.Pool per call, :377-380)."""
    K = x.shape[1]
    g = partial(interval_matrices, x=x, u=u, tf=tf, const=const, include_J2=include_J2,
                use_uniform_steps=use_uniform_steps, integrator_steps=integrator_steps,
                ivp_max_step=ivp_max_step, ivp_solver=ivp_solver)
    if processes and processes > 1:
        with mp.Pool(processes) as pool:
            res = pool.map(g, range(K - 1))
    else:
        res = [g(k) for k in range(K - 1)]
    A_k = np.zeros((K - 1, 7, 7))
    B_kp = np.zeros((K - 1, 7, 3))
    B_kn = np.zeros((K - 1, 7, 3))
    Sigma_k = np.zeros((7, K - 1))
    xi_k = np.zeros((7, K - 1))
    for i, r in enumerate(res):
        A_k[i], B_kp[i], B_kn[i], Sigma_k[:, i], xi_k[:, i] = r
    return A_k, B_kp, B_kn, Sigma_k, xi_k


# --------------------------------------------------------------------------- controllers / propagation

def ctrl_zero():
    """ref: control.py:20-29."""
    z = np.zeros(3)
    return lambda x, tau: z


def ctrl_constant(thrust):
    """ref: control.py:47-53."""
    th = np.asarray(thrust, dtype=float)
    return lambda x, tau: th


def ctrl_tangential(mag):
    """Thrust `mag` along t_hat = h_hat x r_hat.  ref: control.py:66-84."""
    def u(x, tau):
        r = x[0:3]
        v = x[3:6]
        r_hat = r / np.linalg.norm(r)
        h = np.cross(r, v)
        h_hat = h / np.linalg.norm(h)
        t_hat = np.cross(h_hat, r_hat)
        return np.column_stack([r_hat, t_hat, h_hat]) @ np.array([0, mag, 0])
    return u


def ctrl_sequence(u_tab, tf_u=1, tf_sim=1):
    """FOH of a (3,Ku) table on tau/end_tau while tau <= end_tau, zero after.
    ref: control.py:86-143."""
    end_tau = tf_u / tf_sim
    u_tab = np.asarray(u_tab, dtype=float)

    def u(x, tau):
        if tau <= end_tau:
            return foh(tau / end_tau, u_tab)
        return np.zeros(3)
    return u


def propagate(y0, tf, u_func, const, include_drag=True, include_J2=True, eval_points=100,
              max_step=0.001):
    """Normalized IVP over tau in [0,1].  ref: simulator.py:164-189.
    Returns (y[7,T], t[T])."""
    ts = np.linspace(0, 1, eval_points)

    def rhs(t, y):
        return dynamics(y, u_func(y, t), tf, const, include_drag, include_J2)

    sol = integrate.solve_ivp(rhs, [0, 1], np.asarray(y0, dtype=float), t_eval=ts, max_step=max_step)
    return sol.y, sol.t


def extract_uk(x, tau, u_func):
    """ref: linearize_discretize.py:393-411."""
    return np.column_stack([u_func(x[:, i], tau[i]) for i in range(x.shape[1])])



# ------------------------------------------------------------------ constraint terms (optimizer.py:80-170)

def _skew(x):                                              # optimizer.py:41-45
    return np.array([[0, -x[2], x[1]], [x[2], 0, -x[0]], [-x[1], x[0], 0]])


def constraint_terms(x_bar, u_bar, MU):
    """Optimizer.get_constraint_terms for ONE satellite (optimizer.py:106-169), restated statement by
    statement -- including the precedence of `Dv_h_hat` (:121) and the inverted `ubar_hat` mask (:137-138).
    x_bar (7,K), u_bar (3,Ku) -> dict with the reference's keys."""
    I = np.eye(3)
    r, v = x_bar[0:3, -1], x_bar[3:6, -1]
    nr = np.linalg.norm(r)
    rv = np.concatenate([r, v])
    h = np.cross(r, v)
    nh = np.linalg.norm(h)
    r_hat, h_hat = r / nr, h / nh
    t_hat = np.cross(h_hat, r_hat)
    Dr_h = ((nh**-1 * I) - (nh**-3 * np.outer(h, h))) @ (-_skew(v))
    Dv_h = (nh**-1 * I) - (nh**-3 * np.outer(h, h)) @ (_skew(r))
    Dr_r = (nr**-1 * I) - (nr**-3 * np.outer(r, r))
    Dr_t = (-_skew(r_hat) @ Dr_h) + (_skew(h_hat) @ Dr_r)
    Dv_t = -_skew(r_hat) @ Dv_h
    out = {}
    rb = x_bar[0:3, :-1]
    out["rbar_hat"] = rb / np.linalg.norm(rb, axis=0)
    un = np.linalg.norm(u_bar, axis=0)
    uh = np.zeros(u_bar.shape)
    idx = un <= np.finfo(float).eps
    with np.errstate(invalid="ignore", divide="ignore"):
        uh[:, idx] = u_bar[:, idx] / un[idx]
    out["ubar_hat"] = uh
    out["rf_hat"] = r_hat
    out["Vc"] = np.sqrt(MU / nr)
    DrVc = (-1 / 2) * (MU**0.5) * (nr**(-5 / 2)) * r
    out["DrVc"], out["DrVc_rbar"] = DrVc, np.dot(DrVc, r)
    out["Vt"] = np.dot(v, t_hat)
    g = np.concatenate([np.dot(v, Dr_t), t_hat + np.dot(v, Dv_t)])
    out["DrVt_DvVt"], out["DrVt_DvVt_bar"] = g, np.dot(g, rv)
    out["Vr"] = np.dot(v, r_hat)
    g = np.concatenate([np.dot(v, Dr_r), r_hat])
    out["DrVr_DvVr"], out["DrVr_DvVr_bar"] = g, np.dot(g, rv)
    out["Vn"] = np.dot(v, h_hat)
    g = np.concatenate([np.dot(v, Dr_h), h_hat + np.dot(v, Dv_h)])
    out["DrVn_DvVn"], out["DrVn_DvVn_bar"] = g, np.dot(g, rv)
    return out


def rollout(A_k, B_kp, B_kn, Sigma_k, xi_k, x0, u, tf):
    """Discrete model rolled forward -- the implicit check the reference's tests make.
    ref: test_discretizer.py:110-113, optimizer.py:327-339."""
    K = A_k.shape[0] + 1
    xs = [np.asarray(x0, dtype=float)]
    for k in range(K - 1):
        xs.append(A_k[k] @ xs[-1] + B_kn[k] @ u[:, k] + B_kp[k] @ u[:, k + 1] + Sigma_k[:, k] * tf + xi_k[:, k])
    return np.column_stack(xs)


def norm_rel_err(a, b):
    """Parity metric: max|a-b| / max|b| (hard part 2)."""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))
