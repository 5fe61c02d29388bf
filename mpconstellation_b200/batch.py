"""Batched entry points over the C-ABI: every (satellite, interval) / satellite is one GPU work unit.

Host arrays (numpy) go through the library's host API (chunked copy/compute pipeline); torch CUDA
tensors go through the device API on torch's current stream.  Shapes follow the reference with a
leading batch axis: x [N,7,K], u [N,3,K], tf [N] -> SoA [105, N*(K-1)].
"""
import ctypes
import math
import threading

import numpy as np

from . import _lib
from .control import ControllerSpec, spec_from

_ctx_lock = threading.Lock()
_ctxs = {}
_call_locks = {}


def _lock(device=0):
    """One lock per device context: an mpc_ctx owns ONE workspace (device buffers, events, pinned staging), and ctypes
    releases the GIL during a call, so two Python threads must not be inside host-API calls of the same ctx at once."""
    with _ctx_lock:
        return _call_locks.setdefault(device, threading.RLock())


def _ctx(device=0):
    with _ctx_lock:
        if device not in _ctxs:
            _lib.require_gpu()
            h = ctypes.c_void_p()
            _lib.check(_lib.lib().mpc_ctx_create(int(device), ctypes.byref(h)))
            _ctxs[device] = h
        return _ctxs[device]


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != tuple(shape):
        raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
    return a


def _tf_vec(tf, n):
    return np.ascontiguousarray(np.broadcast_to(np.asarray(tf, dtype=np.float64), (n,)))


class DiscretizedBatch:
    """SoA result of a batched discretization plus reference-shaped views.

    soa: [105, N*(K-1)] float64 (rows: A 49 | B_kp 21 | B_kn 21 | Sigma 7 | xi 7), status: [N, K-1] int32.
    layout "satmajor": column = s (K-1) + k (per-satellite blocks); "kmajor": column = k N + s (what the streamed host
    pass propagate_discretize(layout="kmajor") writes).  sat(s) returns (A_k, B_kp, B_kn, Sigma_k, xi_k) with the shapes
    and order of Discretizer.discretize (linearize_discretize.py:390); they are views in either layout, no copy is made.
    """

    def __init__(self, soa, status, n_sats, K, layout="satmajor"):
        self.soa, self.status, self.n_sats, self.K, self.layout = soa, status, n_sats, K, layout

    def _rows(self, r0, nrows, s):
        n = self.K - 1
        if self.layout == "kmajor":
            return self.soa[r0:r0 + nrows, s::self.n_sats]
        return self.soa[r0:r0 + nrows, s * n:(s + 1) * n]

    def sat(self, s):
        n = self.K - 1
        A = self._rows(_lib.ROW_A, 49, s).T.reshape(n, 7, 7)
        Bp = self._rows(_lib.ROW_BP, 21, s).T.reshape(n, 7, 3)
        Bn = self._rows(_lib.ROW_BN, 21, s).T.reshape(n, 7, 3)
        return A, Bp, Bn, self._rows(_lib.ROW_SIGMA, 7, s), self._rows(_lib.ROW_XI, 7, s)

    def stacked(self):
        """(A[N,K-1,7,7], B_kp[N,K-1,7,3], B_kn[N,K-1,7,3], Sigma[N,7,K-1], xi[N,7,K-1]) as views."""
        N, n = self.n_sats, self.K - 1
        if self.layout == "kmajor":
            v = self.soa.reshape(_lib.MPC_OUT_ROWS, n, N).transpose(0, 2, 1)      # [105, N, n], strided
        else:
            v = self.soa.reshape(_lib.MPC_OUT_ROWS, N, n)
        A = v[0:49].transpose(1, 2, 0).reshape(N, n, 7, 7)
        Bp = v[49:70].transpose(1, 2, 0).reshape(N, n, 7, 3)
        Bn = v[70:91].transpose(1, 2, 0).reshape(N, n, 7, 3)
        return A, Bp, Bn, v[91:98].transpose(1, 0, 2), v[98:105].transpose(1, 0, 2)

    def raise_on_error(self):
        raise_on_status(self.status)


def raise_on_status(status):
    st = np.asarray(status)
    if st.size and st.max() != 0:
        idx = np.argwhere(st != 0)[0]
        code = int(st[tuple(idx)])
        if code == _lib.ST_MASS:
            # same exception type / text as simulator.py:135-136
            raise Exception(f"ERROR: INVALID SATELLITE MASS: non-positive mass at unit {tuple(int(i) for i in idx)}")
        if code == _lib.ST_STEP:
            raise RuntimeError(f"adaptive integration failed (step size underflow) at unit {tuple(int(i) for i in idx)}")
        raise FloatingPointError(f"non-finite result at unit {tuple(int(i) for i in idx)}")


def discretize_batch(x, u, tf, const, include_J2=False, include_drag=False, n_sub=100, out=None, status=None,
                     device=0, check=True, adaptive=None, disc_drag=None):
    """Discretize every interval of every satellite in one launch sequence (host arrays).

    x [N,7,K], u [N,3,K], tf scalar or [N]; `out` may be a preallocated (ideally pinned) [105, N*(K-1)]
    array.  Returns a DiscretizedBatch.  ref: linearize_discretize.py:334-390 / :8-82.

    adaptive=None: fixed-step fourth-order Runge-Kutta-Nystrom, trapezoid on n_sub+1 nodes (the reference's
    use_uniform_steps=True node set).  adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2): the reference's
    default mode, quadrature on the steps scipy's RK45 controller accepts; the result carries `.n_nodes`.

    include_drag=True needs disc_drag=(CD, rho): const.CD and the (constant) value of rho_func the reference's drag
    branch reads (linearize_discretize.py:162-168); without them the call is rejected like the reference's.
    """
    ctx = _ctx(device)
    x = _f64(x)
    if x.ndim != 3 or x.shape[1] != 7:
        raise ValueError(f"x must be [N,7,K], got {x.shape}")
    N, _, K = x.shape
    u = _f64(u)
    if u.ndim != 3 or u.shape[:2] != (N, 3) or u.shape[2] < 2:
        raise ValueError(f"u must be [N,3,Ku] with Ku >= 2, got {u.shape}")
    Ku = u.shape[2]                      # Ku != K: u lives on its own grid (linearize_discretize.py:308-315)
    tfv = _tf_vec(tf, N)
    n_int = N * (K - 1)
    if out is None:
        out = _lib.pinned_empty((_lib.MPC_OUT_ROWS, n_int))
    elif out.shape != (_lib.MPC_OUT_ROWS, n_int) or out.dtype != np.float64 or not out.flags["C_CONTIGUOUS"]:
        raise ValueError("out must be C-contiguous float64 [105, N*(K-1)]")
    if status is None:
        status = np.zeros(n_int, dtype=np.int32)
    p = _lib.make_params(const, include_J2, include_drag, disc_drag=disc_drag)
    n_nodes = None
    with _lock(device):
        if Ku != K:
            ad = adaptive or {}
            n_nodes = np.zeros(n_int, dtype=np.int32) if adaptive is not None else None
            _lib.check(_lib.lib().mpc_discretize_batch_ugrid_host(
                ctx, _lib.addr(x), _lib.addr(u), Ku, _lib.addr(tfv), ctypes.byref(p), N, K, int(adaptive is not None),
                int(n_sub), float(ad.get("rtol", 1e-3)), float(ad.get("atol", 1e-6)), float(ad.get("max_step", 1e-2)),
                _lib.addr(out), _lib.addr(status), _lib.addr(n_nodes)))
        elif adaptive is None:
            _lib.check(_lib.lib().mpc_discretize_batch_host(ctx, _lib.addr(x), _lib.addr(u), _lib.addr(tfv),
                                                            ctypes.byref(p), N, K, int(n_sub), _lib.addr(out),
                                                            _lib.addr(status)))
        else:
            n_nodes = np.zeros(n_int, dtype=np.int32)
            _lib.check(_lib.lib().mpc_discretize_batch_adaptive_host(
                ctx, _lib.addr(x), _lib.addr(u), _lib.addr(tfv), ctypes.byref(p), N, K, float(adaptive.get("rtol", 1e-3)),
                float(adaptive.get("atol", 1e-6)), float(adaptive.get("max_step", 1e-2)), _lib.addr(out), _lib.addr(status),
                _lib.addr(n_nodes)))
    res = DiscretizedBatch(out, status.reshape(N, K - 1), N, K)
    res.n_nodes = None if n_nodes is None else n_nodes.reshape(N, K - 1)
    if check:
        res.raise_on_error()
    return res


def _ctrl_struct(spec, n_sats, table_addr=None, end_tau_addr=None):
    c = _lib.MpcController()
    c.kind = spec.kind
    c.thrust[0], c.thrust[1], c.thrust[2] = [float(t) for t in spec.thrust]
    keep = None
    if np.ndim(spec.end_tau) == 0:
        c.end_tau = float(spec.end_tau)
    else:
        et = np.ascontiguousarray(spec.end_tau, dtype=np.float64)
        if et.shape != (n_sats,):
            raise ValueError("per-satellite end_tau must have shape [N]")
        c.end_tau = float(et[0])
        c.end_tau_per_sat = end_tau_addr if end_tau_addr is not None else et.ctypes.data
        keep = [et]
    if spec.kind == _lib.CTRL_SEQUENCE:
        tab = spec.table
        if tab.ndim == 3 and tab.shape[0] != n_sats:
            raise ValueError("per-satellite sequence table must be [N,3,Ku]")
        c.table_len = tab.shape[-1]
        c.table_per_sat = int(tab.ndim == 3)
        keep = [keep, tab]
        c.table = table_addr if table_addr is not None else tab.ctypes.data
    return c, keep


def default_n_sub(T, max_step=0.001):
    """RK4 steps between samples so that the step in tau is <= max_step (simulator.py:186)."""
    if T < 2:
        return 1
    return max(1, int(math.ceil(1.0 / (max_step * (T - 1)) - 1e-9)))


def propagate_batch(y0, tf, controller, const, include_drag=True, include_J2=True, T=100, n_sub=None,
                    want_u=True, device=0, check=True, y_out=None, u_out=None, rk45=None, n_steps=None):
    """Propagate N satellites over tau in [0,1] and sample at linspace(0,1,T) (host arrays).

    y0 [N,7] normalized states; returns (y [N,7,T], u [N,3,T] or None, t [T], status [N]).
    ref: simulator.py:164-189 (+ extract_uk, linearize_discretize.py:393-411).

    n_sub=None (default): the reference's own integrator, replayed on the device -- scipy's RK45 with its step-size
    controller exactly as simulator.py:185-187 calls it (max_step=0.001, rtol 1e-3, atol 1e-6, dense-output samples);
    `rk45=dict(rtol=, atol=, max_step=)` overrides those numbers, `n_steps` (int32 [N]) receives the steps attempted.
    n_sub >= 1: fixed-step RK4 with n_sub steps between samples (default_n_sub(T) mirrors max_step=0.001).
    """
    ctx = _ctx(device)
    y0 = _f64(y0)
    if y0.ndim != 2 or y0.shape[1] != 7:
        raise ValueError(f"y0 must be [N,7], got {y0.shape}")
    N = y0.shape[0]
    T = int(T)
    spec = spec_from(controller)
    tfv = _tf_vec(tf, N)
    if n_sub is None:
        n_sub = 0
    if n_sub != 0 and rk45 is not None:
        raise ValueError("rk45 options only apply to the reference integrator (n_sub=None)")
    y = y_out if y_out is not None else _lib.pinned_empty((N, 7, T))
    uo = (u_out if u_out is not None else _lib.pinned_empty((N, 3, T))) if want_u else None
    status = np.zeros(N, dtype=np.int32)
    t = np.linspace(0, 1, T)
    if N == 0 or T == 0:
        return y, uo, t, status
    p = _lib.make_params(const, include_J2, include_drag)
    c, _keep = _ctrl_struct(spec, N)
    if n_sub == 0:
        o = rk45 or {}
        if n_steps is not None and (n_steps.shape != (N,) or n_steps.dtype != np.int32):
            raise ValueError("n_steps must be int32 [N]")
        with _lock(device):
            _lib.check(_lib.lib().mpc_propagate_batch_rk45_host(
                ctx, _lib.addr(y0), _lib.addr(tfv), ctypes.byref(p), ctypes.byref(c), N, T, float(o.get("rtol", 1e-3)),
                float(o.get("atol", 1e-6)), float(o.get("max_step", 1e-3)), _lib.addr(y), _lib.addr(uo),
                _lib.addr(status), _lib.addr(n_steps)))
    else:
        with _lock(device):
            _lib.check(_lib.lib().mpc_propagate_batch_host(ctx, _lib.addr(y0), _lib.addr(tfv), ctypes.byref(p),
                                                           ctypes.byref(c), N, T, int(n_sub), _lib.addr(y),
                                                           _lib.addr(uo), _lib.addr(status)))
    if check:
        raise_on_status(status)
    return y, uo, t, status


def propagate_discretize(y0, tf, controller, const, T, prop_drag=False, prop_J2=False, disc_J2=False,
                         n_sub_prop=None, n_sub_disc=100, out=None, y_out=None, u_out=None, device=0, check=True,
                         status=None, layout="satmajor"):
    """One SCP linearization pass on the device: propagate -> extract_uk -> discretize, the reference
    trajectory staying in HBM between the kernels (control.py:180-188 pattern).  K = T.
    layout="kmajor": the matrices come back in the k-major layout (DiscretizedBatch views hide it) through the streamed
    pass -- windows along k gated on the propagation's progress, every window read back while the next one runs.
    Returns (DiscretizedBatch, y [N,7,T], u [N,3,T])."""
    ctx = _ctx(device)
    y0 = _f64(y0)
    N = y0.shape[0]
    T = int(T)
    spec = spec_from(controller)
    tfv = _tf_vec(tf, N)
    n_int = N * (T - 1)
    if n_sub_prop is None:
        n_sub_prop = 0                    # the reference's RK45, replayed (see propagate_batch)
    if out is None:
        out = _lib.pinned_empty((_lib.MPC_OUT_ROWS, n_int))
    y = y_out if y_out is not None else _lib.pinned_empty((N, 7, T))
    uo = u_out if u_out is not None else _lib.pinned_empty((N, 3, T))
    if status is None:
        status = np.zeros(n_int, dtype=np.int32)
    elif status.shape != (n_int,) or status.dtype != np.int32 or not status.flags["C_CONTIGUOUS"]:
        raise ValueError("status must be C-contiguous int32 [N*(T-1)] (pinned: pinned_empty(n, np.int32))")
    pp = _lib.make_params(const, prop_J2, prop_drag)
    pd = _lib.make_params(const, disc_J2, False)
    c, _keep = _ctrl_struct(spec, N)
    if layout not in ("satmajor", "kmajor"):
        raise ValueError(f"unknown layout {layout!r}")
    with _lock(device):
        _lib.check(_lib.lib().mpc_propagate_discretize_host_layout(
            ctx, _lib.addr(y0), _lib.addr(tfv), ctypes.byref(pp), ctypes.byref(pd), ctypes.byref(c), N, T, int(n_sub_prop),
            int(n_sub_disc), _lib.addr(y), _lib.addr(uo), _lib.addr(out), _lib.addr(status),
            _lib.LAYOUT_K_MAJOR if layout == "kmajor" else _lib.LAYOUT_SAT_MAJOR))
    res = DiscretizedBatch(out, status.reshape(N, T - 1), N, T, layout)
    if check:
        res.raise_on_error()
    return res, y, uo


# ------------------------------------------------------------------ device-tensor API (torch CUDA tensors)

def _torch():
    import torch
    return torch


def discretize_batch_device(x, u, tf, const, include_J2=False, n_sub=100, out=None, out_pitch=None, out_offset=0,
                            status=None, extra_dst=None, adaptive=None, n_nodes=None):
    """Device form: x [N,7,K], u [N,3,K], tf [N] are float64 CUDA tensors; work is enqueued on torch's
    current stream, nothing synchronises.  `out` is [105, pitch]; `extra_dst` is an optional list of
    further [105, pitch] buffers (e.g. peer-mapped) every result is also stored to (total 1,2,4 or 8)."""
    torch = _torch()
    N, _, K = x.shape
    n_int = N * (K - 1)
    assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous()
    assert u.shape == (N, 3, K) and u.is_contiguous() and tf.shape == (N,)
    if out is None:
        out = torch.empty((_lib.MPC_OUT_ROWS, n_int), dtype=torch.float64, device=x.device)
    pitch = out.shape[1] if out_pitch is None else int(out_pitch)
    if status is None:
        status = torch.empty(n_int, dtype=torch.int32, device=x.device)
    p = _lib.make_params(const, include_J2, False)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    L = _lib.lib()
    if adaptive is not None:
        _lib.check(L.mpc_discretize_batch_adaptive(
            x.data_ptr(), u.data_ptr(), tf.data_ptr(), ctypes.byref(p), N, K, float(adaptive.get("rtol", 1e-3)),
            float(adaptive.get("atol", 1e-6)), float(adaptive.get("max_step", 1e-2)), out.data_ptr(), pitch,
            int(out_offset), status.data_ptr(), n_nodes.data_ptr() if n_nodes is not None else None, stream))
    elif extra_dst:
        ptrs = [out.data_ptr()] + [int(d if isinstance(d, int) else d.data_ptr()) for d in extra_dst]
        arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
        _lib.check(L.mpc_discretize_batch_multi(x.data_ptr(), u.data_ptr(), tf.data_ptr(), ctypes.byref(p), N, K,
                                                int(n_sub), arr, len(ptrs), pitch, int(out_offset),
                                                status.data_ptr(), stream))
    else:
        _lib.check(L.mpc_discretize_batch(x.data_ptr(), u.data_ptr(), tf.data_ptr(), ctypes.byref(p), N, K,
                                          int(n_sub), out.data_ptr(), pitch, int(out_offset), status.data_ptr(),
                                          stream))
    return out, status


def propagate_batch_device(y0, tf, controller, const, include_drag=True, include_J2=True, T=100, n_sub=None,
                           y=None, u_out=None, status=None):
    """Device form of propagate_batch: y0 [N,7], tf [N] float64 CUDA tensors -> y [N,7,T], u [N,3,T], status [N]."""
    torch = _torch()
    N = y0.shape[0]
    spec = spec_from(controller)
    if n_sub is None:
        n_sub = 0                         # the reference's RK45, replayed (see propagate_batch)
    if y is None:
        y = torch.empty((N, 7, T), dtype=torch.float64, device=y0.device)
    if u_out is None:
        u_out = torch.empty((N, 3, T), dtype=torch.float64, device=y0.device)
    if status is None:
        status = torch.empty(N, dtype=torch.int32, device=y0.device)
    tab_dev = et_dev = None
    if spec.kind == _lib.CTRL_SEQUENCE:
        tab_dev = torch.as_tensor(spec.table, dtype=torch.float64).to(y0.device).contiguous()
        if np.ndim(spec.end_tau) != 0:
            et_dev = torch.as_tensor(np.asarray(spec.end_tau, dtype=np.float64)).to(y0.device).contiguous()
    c, _keep = _ctrl_struct(spec, N, table_addr=tab_dev.data_ptr() if tab_dev is not None else None,
                            end_tau_addr=et_dev.data_ptr() if et_dev is not None else None)
    p = _lib.make_params(const, include_J2, include_drag)
    stream = torch.cuda.current_stream(y0.device).cuda_stream
    _lib.check(_lib.lib().mpc_propagate_batch(y0.data_ptr(), tf.data_ptr(), ctypes.byref(p), ctypes.byref(c), N,
                                              int(T), int(n_sub), y.data_ptr(), u_out.data_ptr(),
                                              status.data_ptr(), stream))
    for t_ in (tab_dev, et_dev):
        if t_ is not None:
            t_.record_stream(torch.cuda.current_stream(y0.device))
    return y, u_out, status


def propagate_discretize_device(y0, tf, controller, const, T, prop_drag=False, prop_J2=False, disc_J2=False,
                                n_sub_prop=None, n_sub_disc=100, y=None, u_out=None, out=None, out_pitch=None,
                                out_offset=0, status_prop=None, status_disc=None, n_windows=0, extra_dst=None,
                                out_ptr=None, gather=None):
    """Device form of propagate_discretize: one SCP linearization pass (control.py:180-188) on CUDA tensors, enqueued on
    torch's current stream, with the propagation overlapped with the discretization (mpc_propagate_discretize: the
    intervals are discretized window by window along k as the propagation publishes its progress).  Bit-identical to
    propagate_batch_device followed by discretize_batch_device.  y0 [N,7], tf [N] float64 CUDA tensors; K = T.
    `extra_dst`: further [105, pitch] buffers (tensors or raw peer-mapped addresses) every result is also stored to
    (total 1, 2, 4 or 8: the fused all-gather); `out_ptr`: raw address used instead of out.data_ptr() (a multicast
    mapping of `out`); `gather`: an _lib.MpcGatherOpts -- layout (satellite- or k-major), n_sats_total, sat_offset and the
    gather options for this call (then out_pitch / out_offset are implied).
    Returns (out [105, pitch], y [N,7,T], u [N,3,T], status_prop [N], status_disc [N*(T-1)])."""
    torch = _torch()
    N = y0.shape[0]
    T = int(T)
    dev = y0.device
    assert y0.is_cuda and y0.dtype == torch.float64 and y0.is_contiguous() and y0.shape == (N, 7) and tf.shape == (N,)
    n_int = N * (T - 1)
    spec = spec_from(controller)
    if n_sub_prop is None:
        n_sub_prop = 0                    # the reference's RK45, replayed (see propagate_batch)
    if y is None:
        y = torch.empty((N, 7, T), dtype=torch.float64, device=dev)
    if u_out is None:
        u_out = torch.empty((N, 3, T), dtype=torch.float64, device=dev)
    if out is None:
        out = torch.empty((_lib.MPC_OUT_ROWS, n_int), dtype=torch.float64, device=dev)
    pitch = out.shape[1] if out_pitch is None else int(out_pitch)
    if status_prop is None:
        status_prop = torch.empty(N, dtype=torch.int32, device=dev)
    if status_disc is None:
        status_disc = torch.empty(n_int, dtype=torch.int32, device=dev)
    assert y.shape == (N, 7, T) and u_out.shape == (N, 3, T) and y.is_contiguous() and u_out.is_contiguous()
    tab_dev = et_dev = None
    if spec.kind == _lib.CTRL_SEQUENCE:
        tab_dev = torch.as_tensor(spec.table, dtype=torch.float64).to(dev).contiguous()
        if np.ndim(spec.end_tau) != 0:
            et_dev = torch.as_tensor(np.asarray(spec.end_tau, dtype=np.float64)).to(dev).contiguous()
    c, _keep = _ctrl_struct(spec, N, table_addr=tab_dev.data_ptr() if tab_dev is not None else None,
                            end_tau_addr=et_dev.data_ptr() if et_dev is not None else None)
    pp = _lib.make_params(const, prop_J2, prop_drag)
    pd = _lib.make_params(const, disc_J2, False)
    cur = torch.cuda.current_stream(dev)
    ptrs = [int(out_ptr) if out_ptr is not None else out.data_ptr()]
    ptrs += [int(d_ if isinstance(d_, int) else d_.data_ptr()) for d_ in (extra_dst or [])]
    arr = (ctypes.c_void_p * len(ptrs))(*ptrs)
    dev_i = dev.index if dev.index is not None else torch.cuda.current_device()
    with _lock(dev_i):   # the ctx owns the internal streams / progress words of the overlapped pass
        if gather is not None:      # fused all-gather with per-call options / layout (mpc_propagate_discretize_gather)
            _lib.check(_lib.lib().mpc_propagate_discretize_gather(
                _ctx(dev_i), y0.data_ptr(), tf.data_ptr(),
                ctypes.byref(pp), ctypes.byref(pd), ctypes.byref(c), N, T, int(n_sub_prop), int(n_sub_disc), y.data_ptr(),
                u_out.data_ptr(), arr, len(ptrs), ctypes.byref(gather), status_prop.data_ptr(), status_disc.data_ptr(),
                int(n_windows), cur.cuda_stream))
        else:
            _lib.check(_lib.lib().mpc_propagate_discretize_multi(
                _ctx(dev_i), y0.data_ptr(), tf.data_ptr(),
                ctypes.byref(pp), ctypes.byref(pd), ctypes.byref(c), N, T, int(n_sub_prop), int(n_sub_disc), y.data_ptr(),
                u_out.data_ptr(), arr, len(ptrs), pitch, int(out_offset), status_prop.data_ptr(), status_disc.data_ptr(),
                int(n_windows), cur.cuda_stream))
    for t_ in (tab_dev, et_dev):
        if t_ is not None:
            t_.record_stream(cur)
    return out, y, u_out, status_prop, status_disc


def fp64_peak_tflops(device=0, repeats=5):
    _lib.require_gpu()
    tf_, ms = ctypes.c_double(), ctypes.c_double()
    _lib.check(_lib.lib().mpc_fp64_peak_probe(int(device), int(repeats), ctypes.byref(tf_), ctypes.byref(ms)))
    return tf_.value, ms.value


def launch_count():
    return int(_lib.lib().mpc_launch_count())
