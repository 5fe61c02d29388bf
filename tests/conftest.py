import os
import sys
import warnings

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

warnings.filterwarnings("ignore", category=DeprecationWarning)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


GOLDEN = os.path.join(ROOT, "tests", "golden")
NAMES = ["A_k", "B_kp", "B_kn", "Sigma_k", "xi_k"]


@pytest.fixture(scope="session")
def gold_disc():
    return np.load(os.path.join(GOLDEN, "discretize.npz"))


@pytest.fixture(scope="session")
def gold_prop():
    return np.load(os.path.join(GOLDEN, "propagate.npz"))


@pytest.fixture(scope="session")
def const(gold_disc):
    from oracle.mpc_oracle import OracleConstants
    return OracleConstants(*gold_disc["const"])


def rel_err(a, b):
    """norm-relative parity metric: max|a-b| / max|b|"""
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    den = np.max(np.abs(b)) if b.size else 0.0
    if a.size == 0:
        return 0.0
    return float(np.max(np.abs(a - b)) / den) if den > 0 else float(np.max(np.abs(a - b)))


def synth_batch(n_sats, K, tf, const, seed=20240531, thrust=0.5):
    """Synthetic constellation of SURVEY.md section 8(d): Hubble state rotated about z by 2*pi*i/N,
    speed scaled by 1 + 0.1*U[0,1), tangential thrust; reference trajectories from the C oracle."""
    from oracle import c_oracle as C
    g = np.load(os.path.join(GOLDEN, "discretize.npz"))
    from oracle import mpc_oracle as O
    sf = O.scale_factors(g["x0_dim"])
    y0 = O.normalize_state(g["x0_dim"], sf)
    rng = np.random.default_rng(seed)
    ang = 2 * np.pi * np.arange(n_sats) / max(n_sats, 1)
    ca, sa = np.cos(ang), np.sin(ang)
    Y = np.tile(y0, (n_sats, 1))
    Y[:, 0], Y[:, 1] = ca * y0[0] - sa * y0[1], sa * y0[0] + ca * y0[1]
    f = 1 + 0.1 * rng.random(n_sats)
    Y[:, 3], Y[:, 4] = (ca * y0[3] - sa * y0[4]) * f, (sa * y0[3] + ca * y0[4]) * f
    Y[:, 5] = y0[5] * f
    x, u, st = C.propagate_batch(Y, tf, const, C.CTRL_TANGENTIAL, (thrust, 0, 0), include_drag=False,
                                 include_J2=False, T=K, n_sub=max(1, int(np.ceil(1000 / max(K - 1, 1)))))
    assert st.max() == 0
    return Y, x, u
