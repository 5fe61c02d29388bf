"""TEST INFRASTRUCTURE ONLY -- runs every kernel source (host build, tests/hostk) on ragged shapes under AddressSanitizer
+ UBSan.  compute-sanitizer is closed on the GPU pool, so this is the memcheck of the kernels' INDEXING: every input,
output and status buffer is a heap array of exactly the documented size (numpy), the per-thread shared-memory slots are a
global array of the instrumented library, and any read or write outside them aborts the process.

Started by tests/test_kernel_source_on_host.py::test_kernel_sources_under_address_sanitizer as
    HOSTK_SANITIZE=1 LD_PRELOAD=<libasan> python tests/hostk/sanitize_driver.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import hostk                                   # noqa: E402
from conftest import synth_batch               # noqa: E402
from oracle.mpc_oracle import OracleConstants  # noqa: E402

assert os.environ.get("HOSTK_SANITIZE"), "run with HOSTK_SANITIZE=1"
const = OracleConstants(*np.load(os.path.join(os.path.dirname(HERE), "golden", "discretize.npz"))["const"])
n = 0
for N, K in ((1, 2), (3, 5), (33, 4), (7, 13)):         # K = 2, batches that do not fill a warp / that straddle two
    y0, x, u = synth_batch(N, K, 0.3, const)
    for j2 in (False, True):
        for n_sub in (1, 2, 7, 10):
            for pair in (True, False):
                soa, st = hostk.discretize(x, u, 0.3, const, include_J2=j2, n_sub=n_sub, pair=pair)
                assert st.min() >= 0
                n += 1
        hostk.discretize_adaptive(x, u, 0.3, const, include_J2=j2)
        hostk.discretize_drag(x, u, 0.3, const, (2.2, 4.0e4), include_J2=j2, n_sub=6)
        hostk.discretize_drag(x, u, 0.3, const, (2.2, 4.0e4), include_J2=j2, adaptive={})
        n += 3
    # launch windows of the overlapped pass: first / middle / last (ragged) window, exactly sized destination
    n_int = N * (K - 1)
    out = np.full((105, n_int), np.nan)
    st = np.full(n_int, -1, dtype=np.int32)
    seg = max(1, (K - 1) // 3)
    for k0 in range(0, K - 1, seg):
        hostk.discretize(x, u, 0.3, const, k0=k0, kc=min(seg, K - 1 - k0), out=out, status=st)
        n += 1
    assert st.min() == 0 and np.isfinite(out).all()
    # propagation: every controller law, drag / J2, progress words, T = 1
    tab = 0.2 * np.random.default_rng(1).standard_normal((3, 5))
    for kind in (0, 1, 2, 3):
        for dj in ((False, False), (True, True)):
            for T, seg in ((1, 0), (2, 1), (K, 2), (K + 3, 0)):
                hostk.propagate(y0, 0.3, const, kind=kind, thrust=(0.1, 0.2, -0.1), table=tab if kind == 3 else None,
                                end_tau=0.7, include_drag=dj[0], include_J2=dj[1], T=T, n_sub=2, seg_len=seg)
                n += 1
    # the replayed RK45 propagator: controller laws, drag / J2, progress words (one count per satellite), T = 1, a step
    # controller that rejects (max_step 1), fewer satellites per warp
    for kind in (0, 2, 3):
        for T, seg, ms, lpw in ((1, 0, 1e-3, 32), (2, 1, 0.05, 32), (K, 2, 1.0, 4), (K + 3, 0, 0.05, 1)):
            hostk.propagate_rk45(y0, 0.3, const, kind=kind, thrust=(0.1, 0.2, -0.1), table=tab if kind == 3 else None,
                                 end_tau=0.7, include_drag=(kind == 0), include_J2=(kind == 0), T=T, max_step=ms, lpw=lpw,
                                 seg_len=seg, spec=(kind != 2))
            n += 1
    # the thread-group kernel (8 host threads per interval), with the idle groups of a last partial warp
    if N <= 3:
        for n_sub in (2, 7):
            hostk.discretize_group(x, u, 0.3, const, include_J2=True, n_sub=n_sub, extra_groups=2)
            n += 1
    # the k-major gathered layout: this batch as the middle shard of a larger buffer, window by window
    ntot, soff = N + 5, 2
    outk = np.full((105, ntot * (K - 1)), np.nan)
    for k0 in range(0, K - 1, seg_k := max(1, (K - 1) // 2)):
        hostk.discretize(x, u, 0.3, const, k0=k0, kc=min(seg_k, K - 1 - k0), out=outk, pitch=ntot * (K - 1), status=st,
                         km_ntot=ntot, km_soff=soff)
        n += 1
    assert np.isfinite(outk.reshape(105, K - 1, ntot)[:, :, soff:soff + N]).all()
    # the round-1 build of the default-mode kernel (kept for A/B)
    hostk.discretize_adaptive(x, u, 0.3, const, v1=True)
    n += 1
    # constraint terms with u on its own (longer and shorter) grid, and the sparse Jacobian
    for Ku in (1, K, K + 4):
        hostk.constraint_terms(x, np.ascontiguousarray(np.resize(u, (N, 3, Ku))), const.MU)
        n += 1
    soa, _ = hostk.discretize(x, u, 0.3, const, n_sub=4)
    hostk.dynamics_jacobian(soa, N, K)
    n += 1
print(f"hostk sanitize: {n} launches clean")
