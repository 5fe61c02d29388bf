"""Optimizer.get_constraint_terms (optimizer.py:80-170) for a batch of satellites, on the GPU.

This is the step that follows the discretization inside Optimizer.solve_OPT (optimizer.py:243-262): unit vectors
of the reference positions / thrusts and the linearised circular-orbit terminal conditions.  The reference's
formulas are mirrored literally, including its inverted `ubar_hat` mask (optimizer.py:137-138) and the operator
precedence of `Dv_h_hat` (optimizer.py:121) -- they define what the optimizer is handed.
"""
import ctypes

import numpy as np

from . import _lib
from .batch import _ctx, _f64, _lock

KEYS = ['rbar_hat', 'ubar_hat', 'rf_hat', 'Vc', 'DrVc', 'DrVc_rbar', 'Vt', 'DrVt_DvVt', 'DrVt_DvVt_bar',
        'Vr', 'DrVr_DvVr', 'DrVr_DvVr_bar', 'Vn', 'DrVn_DvVn', 'DrVn_DvVn_bar']    # optimizer.py:101-104


def _split_final(fin):
    """[N,32] terminal block -> {key: [N] or [N,len]} views (layout: include/mpc_b200.h MPC_FT_*)."""
    out = {}
    for key, (off, ln) in _lib.FINAL_TERM_LAYOUT.items():
        out[key] = fin[:, off] if ln == 0 else fin[:, off:off + ln]
    return out


def constraint_terms_batch(x, u, const, device=0):
    """x [N,7,K], u [N,3,Ku] (host arrays) -> dict of stacked arrays: rbar_hat [N,3,K-1], ubar_hat [N,3,Ku],
    rf_hat [N,3], Vc [N], DrVc [N,3], ... with the reference's keys."""
    ctx = _ctx(device)
    x = _f64(x)
    u = _f64(u)
    if x.ndim != 3 or x.shape[1] != 7 or x.shape[2] < 2:
        raise ValueError(f"x must be [N,7,K] with K >= 2, got {x.shape}")
    N, _, K = x.shape
    if u.ndim != 3 or u.shape[:2] != (N, 3) or u.shape[2] < 1:
        raise ValueError(f"u must be [N,3,Ku], got {u.shape}")
    Ku = u.shape[2]
    rbar = np.empty((N, 3, K - 1))
    ubar = np.empty((N, 3, Ku))
    fin = np.empty((N, _lib.FINAL_TERMS))
    with _lock(device):
        _lib.check(_lib.lib().mpc_constraint_terms_host(ctx, _lib.addr(x), _lib.addr(u), N, K, Ku, float(const.MU),
                                                        _lib.addr(rbar), _lib.addr(ubar), _lib.addr(fin)))
    out = {"rbar_hat": rbar, "ubar_hat": ubar}
    out.update(_split_final(fin))
    return out


def constraint_terms_device(x, u, const, rbar_hat=None, ubar_hat=None, final_terms=None):
    """Device form: x [N,7,K], u [N,3,Ku] float64 CUDA tensors; enqueued on torch's current stream.
    Returns (rbar_hat [N,3,K-1], ubar_hat [N,3,Ku], final_terms [N,32])."""
    import torch
    N, _, K = x.shape
    Ku = u.shape[2]
    assert x.is_cuda and x.dtype == torch.float64 and x.is_contiguous() and u.is_contiguous() and u.shape[:2] == (N, 3)
    if rbar_hat is None:
        rbar_hat = torch.empty((N, 3, K - 1), dtype=torch.float64, device=x.device)
    if ubar_hat is None:
        ubar_hat = torch.empty((N, 3, Ku), dtype=torch.float64, device=x.device)
    if final_terms is None:
        final_terms = torch.empty((N, _lib.FINAL_TERMS), dtype=torch.float64, device=x.device)
    stream = torch.cuda.current_stream(x.device).cuda_stream
    _lib.check(_lib.lib().mpc_constraint_terms(x.data_ptr(), u.data_ptr(), N, K, Ku, float(const.MU),
                                               rbar_hat.data_ptr(), ubar_hat.data_ptr(), final_terms.data_ptr(),
                                               stream))
    return rbar_hat, ubar_hat, final_terms


def get_constraint_terms(x_bar, u_bar, const, device=0):
    """Drop-in for Optimizer.get_constraint_terms: x_bar / u_bar are the optimizer's N-element lists of (7,K) /
    (3,K) arrays; returns the same dict of N-element lists (optimizer.py:84-100)."""
    if len(x_bar) == 0:
        return {k: [] for k in KEYS}
    stacked = constraint_terms_batch(np.stack([np.asarray(a, dtype=np.float64) for a in x_bar]),
                                     np.stack([np.asarray(a, dtype=np.float64) for a in u_bar]), const, device=device)
    n = len(x_bar)
    return {k: [stacked[k][i] if stacked[k].ndim > 1 else float(stacked[k][i]) for i in range(n)] for k in KEYS}


# ------------------------------------------------------------------ sparse dynamics constraint (optimizer.py:327-339)

NNZ_PER_ROW = 16


class DynamicsJacobian:
    """CSR form  J z = rhs  of the reference's dynamics constraint (dynamics_const_rule, optimizer.py:327-339) for N
    satellites: rows in pyomo's generation order (s, i, k), row = (s*7+i)*(K-1)+k; 16 non-zeros per row; variables
    z = [x (N,7,K) | u (N,3,K) | nu (N,7,K) | tf] flattened C-order (include/mpc_b200.h, mpc_dynamics_jacobian)."""

    def __init__(self, values, indices, rhs, n_sats, K):
        self.values, self.indices, self.rhs, self.n_sats, self.K = values, indices, rhs, n_sats, K
        self.rows = n_sats * 7 * (K - 1)
        self.cols = 17 * n_sats * K + 1

    @property
    def indptr(self):
        return np.arange(self.rows + 1, dtype=np.int64) * NNZ_PER_ROW

    def pack(self, x, u, nu, tf):
        """decision vector z from x [N,7,K], u [N,3,K], nu [N,7,K] (last column unused by the constraint), tf scalar"""
        return np.concatenate([np.ravel(x), np.ravel(u), np.ravel(nu), [float(tf)]])

    def residual(self, z):
        """J z - rhs, reshaped [N,7,K-1]: zero where the reference's constraint holds"""
        v = np.asarray(self.values).reshape(self.rows, NNZ_PER_ROW)
        c = np.asarray(self.indices).reshape(self.rows, NNZ_PER_ROW)
        return ((v * np.asarray(z)[c]).sum(axis=1) - np.asarray(self.rhs)).reshape(self.n_sats, 7, self.K - 1)

    def to_scipy(self):
        from scipy.sparse import csr_matrix
        return csr_matrix((np.asarray(self.values), np.asarray(self.indices), self.indptr), shape=(self.rows, self.cols))


def dynamics_jacobian_device(soa, n_sats, K, pitch=None, offset=0, values=None, indices=None, rhs=None,
                             with_indices=True):
    """Device form: soa is the [105, pitch] float64 CUDA tensor a discretization wrote (columns offset.. hold the
    batch).  Returns (values [rows*16] f64, indices [rows*16] i64 or None, rhs [rows] f64) CUDA tensors; enqueued on
    torch's current stream.  Between SCP iterations only the values change: pass with_indices=False."""
    import torch
    rows = n_sats * 7 * (K - 1)
    pitch = soa.shape[1] if pitch is None else int(pitch)
    assert soa.is_cuda and soa.dtype == torch.float64 and soa.is_contiguous()
    if values is None:
        values = torch.empty(rows * NNZ_PER_ROW, dtype=torch.float64, device=soa.device)
    if indices is None and with_indices:
        indices = torch.empty(rows * NNZ_PER_ROW, dtype=torch.int64, device=soa.device)
    if rhs is None:
        rhs = torch.empty(rows, dtype=torch.float64, device=soa.device)
    stream = torch.cuda.current_stream(soa.device).cuda_stream
    _lib.check(_lib.lib().mpc_dynamics_jacobian(soa.data_ptr(), pitch, int(offset), int(n_sats), int(K),
                                                values.data_ptr(), indices.data_ptr() if indices is not None else None,
                                                rhs.data_ptr(), stream))
    return values, indices, rhs


def dynamics_jacobian(matrices):
    """Host form: `matrices` is a DiscretizedBatch (e.g. Linearization.matrices); the assembly runs on the GPU.
    Returns a DynamicsJacobian with numpy arrays."""
    import torch
    _lib.require_gpu()
    host = matrices.soa
    if getattr(matrices, "layout", "satmajor") == "kmajor":        # the assembly kernel reads per-satellite column blocks
        N, n = matrices.n_sats, matrices.K - 1
        host = host.reshape(_lib.MPC_OUT_ROWS, n, N).transpose(0, 2, 1).reshape(_lib.MPC_OUT_ROWS, N * n)
    soa = torch.from_numpy(np.ascontiguousarray(host)).cuda()
    v, c, r = dynamics_jacobian_device(soa, matrices.n_sats, matrices.K)
    return DynamicsJacobian(v.cpu().numpy(), c.cpu().numpy(), r.cpu().numpy(), matrices.n_sats, matrices.K)
