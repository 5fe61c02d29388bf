"""GPU: Optimizer.get_constraint_terms (optimizer.py:80-170) on the device, through the C-ABI, against the
fixtures produced by the unmodified reference and against the numpy restatement on a larger batch."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, synth_batch

pytestmark = pytest.mark.gpu

KEYS = ['rbar_hat', 'ubar_hat', 'rf_hat', 'Vc', 'DrVc', 'DrVc_rbar', 'Vt', 'DrVt_DvVt', 'DrVt_DvVt_bar',
        'Vr', 'DrVr_DvVr', 'DrVr_DvVr_bar', 'Vn', 'DrVn_DvVn', 'DrVn_DvVn_bar']
BIT_EXACT = ("rbar_hat", "ubar_hat")     # written without contraction: identical to numpy's roundings
TOL = 5e-13                               # terminal terms: a few 3x3 products, relative to the term's own scale


@pytest.fixture(scope="module")
def M():
    import mpconstellation_b200 as m
    m._lib.require_gpu()
    return m


def _check(got, ref, key):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, key
    if key in BIT_EXACT:
        np.testing.assert_array_equal(got, ref, err_msg=key)      # NaN == NaN positions included
    else:
        assert np.max(np.abs(got - ref)) <= TOL * max(np.max(np.abs(ref)), 1.0), key


@pytest.mark.parametrize("tag", ["c0", "c1", "c2"])
def test_constraint_terms_vs_reference_fixtures(M, tag):
    from mpconstellation_b200.constraints import get_constraint_terms
    g = np.load(os.path.join(GOLDEN, "constraint_terms.npz"))

    class Const:
        MU = float(g["MU"])
    out = get_constraint_terms([g[tag + "_x"]], [g[tag + "_u"]], Const)       # the optimizer's list-of-arrays form
    assert list(out) == KEYS and all(len(v) == 1 for v in out.values())
    for k in KEYS:
        _check(out[k][0], g[f"{tag}_{k}"], k)
    if tag == "c2":
        assert np.isnan(out["ubar_hat"][0]).all()       # zero thrust: the reference divides 0/0 (optimizer.py:137-138)


def test_constraint_terms_batch_vs_oracle(M, const):
    """300 satellites, thrust partly exactly zero / tiny (both sides of the eps mask), u on its own grid"""
    from oracle import mpc_oracle as O
    from mpconstellation_b200.constraints import constraint_terms_batch, constraint_terms_device
    import torch
    y0, x, u = synth_batch(300, 37, 1.3, const)
    rng = np.random.default_rng(7)
    u = np.repeat(u, 2, axis=2)[:, :, :53].copy()               # Ku = 53 != K = 37
    u[:, :, 5] = 0.0
    u[:, :, 9] = 1e-17 * rng.standard_normal((300, 3))
    u[:, :, 11] = 1e-15 * rng.standard_normal((300, 3))
    out = constraint_terms_batch(x, u, const)
    for s in range(0, 300, 13):
        ref = O.constraint_terms(x[s], u[s], const.MU)
        for k in KEYS:
            _check(out[k][s], ref[k], k)
    assert np.isnan(out["ubar_hat"][:, :, 5]).all() and np.isfinite(out["ubar_hat"][:, :, 9]).all()
    assert (out["ubar_hat"][:, :, 11] == 0).all() or np.isfinite(out["ubar_hat"][:, :, 11]).all()
    # device-tensor form gives the same bits as the host form
    xd, ud = torch.from_numpy(x).cuda(), torch.from_numpy(u).cuda()
    rb, ub, fin = constraint_terms_device(xd, ud, const)
    np.testing.assert_array_equal(rb.cpu().numpy(), out["rbar_hat"])
    np.testing.assert_array_equal(ub.cpu().numpy(), out["ubar_hat"])
    np.testing.assert_array_equal(fin.cpu().numpy()[:, 3], out["Vc"])


def test_constraint_terms_edges(M, const):
    from mpconstellation_b200.constraints import constraint_terms_batch, get_constraint_terms
    assert get_constraint_terms([], [], const) == {k: [] for k in KEYS}
    _, x, u = synth_batch(1, 2, 0.1, const)                     # K = 2: a single rbar_hat column
    out = constraint_terms_batch(x, u, const)
    assert out["rbar_hat"].shape == (1, 3, 1) and out["DrVn_DvVn"].shape == (1, 6)
    with pytest.raises(ValueError):
        constraint_terms_batch(x[:, :6], u, const)


def test_sparse_dynamics_jacobian_reproduces_the_pyomo_rule(M, const):
    """mpc_dynamics_jacobian: J z - rhs equals the residual of dynamics_const_rule (optimizer.py:327-339) evaluated
    with the rule's own indexing, for random decision vectors; structure: 16 nnz per row, ascending columns"""
    from mpconstellation_b200.constraints import dynamics_jacobian
    from mpconstellation_b200.scp import Linearization
    y0, x, u = synth_batch(5, 9, 0.7, const)
    mats = M.discretize_batch(x, u, 0.7, const, n_sub=10)
    lin = Linearization(x, u, np.full(5, 0.7), mats)
    jac = dynamics_jacobian(mats)
    N, K = 5, 9
    assert jac.rows == N * 7 * (K - 1) and jac.cols == 17 * N * K + 1
    idx = jac.indices.reshape(jac.rows, 16)
    assert np.all(np.diff(idx, axis=1) > 0) and idx.min() >= 0 and idx.max() == jac.cols - 1
    rng = np.random.default_rng(3)
    xz, uz, nuz, tfz = rng.standard_normal((N, 7, K)), rng.standard_normal((N, 3, K)), rng.standard_normal((N, 7, K)), 0.9
    res = jac.residual(jac.pack(xz, uz, nuz, tfz))
    for s in range(N):
        want = lin.dynamics_residual(s, xz[s], uz[s], tfz, nu=nuz[s])
        assert np.max(np.abs(res[s] - want)) < 1e-12 * max(1.0, np.max(np.abs(want)))
    # scipy view of the same CSR
    J = jac.to_scipy()
    assert J.shape == (jac.rows, jac.cols) and J.nnz == 16 * jac.rows
    assert np.allclose(J @ jac.pack(xz, uz, nuz, tfz) - jac.rhs, res.ravel(), rtol=0, atol=1e-12)
    # on the reference point with nu = the FOH defect the constraint holds
    nu0 = np.zeros((N, 7, K))
    for s in range(N):
        nu0[s, :, :K - 1] = lin.dynamics_residual(s, x[s], u[s], 0.7)
    assert np.max(np.abs(jac.residual(jac.pack(x, u, nu0, 0.7)))) < 1e-13
    # a k-major result (the streamed host pass) assembles to the same system
    ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
    ms, _, _ = M.propagate_discretize(y0, 0.7, ctrl, const, T=K, n_sub_disc=10)
    mk, _, _ = M.propagate_discretize(y0, 0.7, ctrl, const, T=K, n_sub_disc=10, layout="kmajor")
    js, jk = dynamics_jacobian(ms), dynamics_jacobian(mk)
    assert np.array_equal(js.indices, jk.indices) and np.max(np.abs(js.values - jk.values)) < 1e-12 and np.max(np.abs(js.rhs - jk.rhs)) < 1e-12
