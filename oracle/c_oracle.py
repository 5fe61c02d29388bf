"""TEST INFRASTRUCTURE ONLY -- ctypes front end of oracle/mpc_oracle.c (the plain-C CPU oracle).

Importers: tests/, __graft_entry__.smoke(), bench.py cpu_baseline / --impl reference.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmpc_oracle.so")

CTRL_ZERO, CTRL_CONSTANT, CTRL_TANGENTIAL, CTRL_SEQUENCE = 0, 1, 2, 3


class OrcParams(ctypes.Structure):
    _fields_ = [(n, ctypes.c_double) for n in ("MU", "R_E", "J2", "G0", "ISP", "S", "R0", "RHO", "C_D", "RHO_ATM", "CD_A", "RHO_A")] + \
               [("include_J2", ctypes.c_int), ("include_drag", ctypes.c_int)]


def build(force=False):
    src = os.path.join(_HERE, "mpc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmpc_oracle.so"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.orc_discretize_rk4.argtypes = [dp, dp, dp, ctypes.POINTER(OrcParams), ctypes.c_int, ctypes.c_int,
                                            ctypes.c_int, dp, ip, ctypes.c_int]
        _lib.orc_discretize_rk4.restype = ctypes.c_int
        _lib.orc_propagate_rk4.argtypes = [dp, dp, ctypes.POINTER(OrcParams), ctypes.c_int, dp, dp, ctypes.c_int,
                                           ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                           dp, dp, ip, ctypes.c_int]
        _lib.orc_propagate_rk4.restype = ctypes.c_int
        _lib.orc_discretize_rk45.argtypes = [dp, dp, dp, ctypes.POINTER(OrcParams), ctypes.c_int, ctypes.c_int,
                                             ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, ip, ip, ctypes.c_int]
        _lib.orc_discretize_rk45.restype = ctypes.c_int
        _lib.orc_propagate_rk45.argtypes = [dp, dp, ctypes.POINTER(OrcParams), ctypes.c_int, dp, dp, ctypes.c_int,
                                            ctypes.c_int, ctypes.c_double, dp, ctypes.c_int, ctypes.c_int,
                                            ctypes.c_double, ctypes.c_double, ctypes.c_double, dp, dp, ip, ip, ip,
                                            ctypes.c_int]
        _lib.orc_propagate_rk45.restype = ctypes.c_int
        _lib.orc_max_threads.restype = ctypes.c_int
    return _lib


def _dp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double)) if a is not None else None


def make_params(const, include_J2=False, include_drag=False, c_d=2.5, rho_atm=9.983e-13, drag=None):
    """drag = (CD, rho_n): the discretizer's drag branch (linearize_discretize.py:160-169) with constant density"""
    cd_a, rho_a = (0.0, 0.0) if drag is None else (float(drag[0]), float(drag[1]))
    return OrcParams(const.MU, const.R_E, const.J2, const.G0, const.ISP, const.S, const.R0, const.RHO,
                     c_d, rho_atm, cd_a, rho_a, int(include_J2), int(include_drag or drag is not None))


def set_scheme(name):
    """Integration scheme of the fixed-step mode: "rkn4x2" (default; what the CUDA kernels run: Nystrom's 3-stage
    4th-order method, steps spanning two quadrature nodes with a Hermite midpoint where the step is short enough),
    "rkn4" (one Nystrom step per node) or "rk4" (classical RK4 on the first-order 56-vector)."""
    lib().orc_set_scheme({"rk4": 0, "rkn4": 1, "rkn4x2": 2}[name])


def max_threads():
    return lib().orc_max_threads()


def discretize_batch(x, u, tf, const, include_J2=False, n_sub=100, nthreads=0, drag=None):
    """x [N,7,K], u [N,3,K], tf scalar or [N] -> (A[N,K-1,7,7], B_kp[N,K-1,7,3], B_kn, Sigma[N,7,K-1], xi[N,7,K-1], status)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    N, _, K = x.shape
    tf = np.ascontiguousarray(np.broadcast_to(np.asarray(tf, dtype=np.float64), (N,)))
    out = np.zeros((N, K - 1, 105))
    status = np.zeros(N * (K - 1), dtype=np.int32)
    p = make_params(const, include_J2, drag=drag)
    lib().orc_discretize_rk4(_dp(x), _dp(u), _dp(tf), ctypes.byref(p), N, K, n_sub, _dp(out),
                             status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), nthreads)
    A = out[:, :, 0:49].reshape(N, K - 1, 7, 7)
    Bp = out[:, :, 49:70].reshape(N, K - 1, 7, 3)
    Bn = out[:, :, 70:91].reshape(N, K - 1, 7, 3)
    S = out[:, :, 91:98].transpose(0, 2, 1)
    X = out[:, :, 98:105].transpose(0, 2, 1)
    return A, Bp, Bn, S, X, status.reshape(N, K - 1)


def discretize_batch_adaptive(x, u, tf, const, include_J2=False, rtol=1e-3, atol=1e-6, max_step=1e-2, nthreads=0,
                              drag=None):
    """The reference's default mode (quadrature on the accepted RK45 steps).  Returns the same tuple as
    discretize_batch plus the node count per interval."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    N, _, K = x.shape
    tf = np.ascontiguousarray(np.broadcast_to(np.asarray(tf, dtype=np.float64), (N,)))
    out = np.zeros((N, K - 1, 105))
    status = np.zeros(N * (K - 1), dtype=np.int32)
    nodes = np.zeros(N * (K - 1), dtype=np.int32)
    p = make_params(const, include_J2, drag=drag)
    ip = ctypes.POINTER(ctypes.c_int)
    lib().orc_discretize_rk45(_dp(x), _dp(u), _dp(tf), ctypes.byref(p), N, K, rtol, atol, max_step, _dp(out),
                              status.ctypes.data_as(ip), nodes.ctypes.data_as(ip), nthreads)
    A = out[:, :, 0:49].reshape(N, K - 1, 7, 7)
    Bp = out[:, :, 49:70].reshape(N, K - 1, 7, 3)
    Bn = out[:, :, 70:91].reshape(N, K - 1, 7, 3)
    S = out[:, :, 91:98].transpose(0, 2, 1)
    X = out[:, :, 98:105].transpose(0, 2, 1)
    return A, Bp, Bn, S, X, status.reshape(N, K - 1), nodes.reshape(N, K - 1)


def propagate_batch(y0, tf, const, kind=CTRL_ZERO, cparams=(0.0, 0.0, 0.0), table=None, end_tau=1.0,
                    include_drag=True, include_J2=True, T=100, n_sub=10, nthreads=0):
    """y0 [N,7] -> (y[N,7,T], u[N,3,T], status[N]); samples at linspace(0,1,T)."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    N = y0.shape[0]
    tf = np.ascontiguousarray(np.broadcast_to(np.asarray(tf, dtype=np.float64), (N,)))
    cp = np.ascontiguousarray(np.resize(np.asarray(cparams, dtype=np.float64), 3))
    Ku, per_sat = 0, 0
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
        Ku = table.shape[-1]
        per_sat = int(table.ndim == 3)
    y = np.zeros((N, 7, T))
    uo = np.zeros((N, 3, T))
    status = np.zeros(N, dtype=np.int32)
    p = make_params(const, include_J2, include_drag)
    lib().orc_propagate_rk4(_dp(y0), _dp(tf), ctypes.byref(p), kind, _dp(cp), _dp(table), Ku, per_sat,
                            float(end_tau), N, T, n_sub, _dp(y), _dp(uo),
                            status.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), nthreads)
    return y, uo, status


def propagate_batch_rk45(y0, tf, const, kind=CTRL_ZERO, cparams=(0.0, 0.0, 0.0), table=None, end_tau=1.0,
                         include_drag=True, include_J2=True, T=100, rtol=1e-3, atol=1e-6, max_step=1e-3, nthreads=0):
    """The reference's own integrator (simulator.py:185-187: solve_ivp RK45, max_step=0.001, t_eval=linspace(0,1,T)),
    restated in C: step-size controller + dense output.  y0 [N,7] -> (y[N,7,T], u[N,3,T], status[N], n_steps[N],
    n_rejected[N]).  end_tau: scalar or [N]."""
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    N = y0.shape[0]
    tf = np.ascontiguousarray(np.broadcast_to(np.asarray(tf, dtype=np.float64), (N,)))
    cp = np.ascontiguousarray(np.resize(np.asarray(cparams, dtype=np.float64), 3))
    Ku, per_sat = 0, 0
    if table is not None:
        table = np.ascontiguousarray(table, dtype=np.float64)
        Ku = table.shape[-1]
        per_sat = int(table.ndim == 3)
    et_arr = None
    if np.ndim(end_tau) != 0:
        et_arr = np.ascontiguousarray(end_tau, dtype=np.float64)
        end_tau = float(et_arr[0])
    y = np.zeros((N, 7, T))
    uo = np.zeros((N, 3, T))
    status = np.zeros(N, dtype=np.int32)
    steps = np.zeros(N, dtype=np.int32)
    rej = np.zeros(N, dtype=np.int32)
    p = make_params(const, include_J2, include_drag)
    ip = ctypes.POINTER(ctypes.c_int)
    lib().orc_propagate_rk45(_dp(y0), _dp(tf), ctypes.byref(p), kind, _dp(cp), _dp(table), Ku, per_sat,
                             float(end_tau), _dp(et_arr), N, T, rtol, atol, max_step, _dp(y), _dp(uo),
                             status.ctypes.data_as(ip), steps.ctypes.data_as(ip), rej.ctypes.data_as(ip), nthreads)
    return y, uo, status, steps, rej
