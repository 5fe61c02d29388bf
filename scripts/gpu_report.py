"""Parity + timing report on the GPU box: errors vs the reference fixtures (both quadrature modes) and kernel
times for BASELINE configs 1-3 in both modes.  Output is committed under profiles/."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import mpconstellation_b200 as M
from bench import make_constellation
from types import SimpleNamespace


def OracleConstants(MU, R_E, J2, G0, ISP, S, R0, RHO):      # the fixtures' constant vector as a Constants-like bag
    return SimpleNamespace(MU=MU, R_E=R_E, J2=J2, G0=G0, ISP=ISP, S=S, R0=R0, RHO=RHO)


def norm_rel_err(a, b):                                      # the parity metric of the tests: max|a-b| / max|b|
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / max(float(np.max(np.abs(b))), 1e-300))


g = np.load(os.path.join(ROOT, "tests/golden/discretize.npz")); const = OracleConstants(*g["const"])
NAMES = ["A_k", "B_kp", "B_kn", "Sigma_k", "xi_k"]
print("== parity vs the unmodified reference (tests/golden), norm-relative max|d|/max|ref| per matrix")
for sc in ("d0", "d1", "d2", "d3", "d4"):
    for mode, uni in (("uni", True), ("def", False)):
        if f"{sc}_{mode}_A_k" not in g: continue
        d = M.Discretizer(const); d.use_uniform_steps = uni
        out = d.discretize(M.Simulator.satellite_dynamics, g[sc + "_x"], g[sc + "_u"], float(g[sc + "_tf"]))
        print(f"  {sc} K={g[sc+'_x'].shape[1]:3d} use_uniform_steps={uni!s:5}: " + "  ".join(f"{n} {norm_rel_err(o, g[f'{sc}_{mode}_{n}']):.1e}" for n, o in zip(NAMES, out)))
gd = np.load(os.path.join(ROOT, "tests/golden/discretize_drag.npz"))
for tag in ("g0", "g1"):
    cv = OracleConstants(*gd[tag + "_const"])
    bag = type("Const", (), {k: getattr(cv, k) for k in ("MU", "R_E", "J2", "G0", "ISP", "S", "R0", "RHO")})()
    bag.CD = float(gd[tag + "_cd"]); rho_n = float(gd[tag + "_rho_n"]); ks = gd[tag + "_ks"]
    for mode, uni in (("uni", True), ("def", False)):
        d = M.Discretizer(bag, rho_func=lambda r: rho_n, drho_func=lambda r: 0.0, include_drag=True, include_J2=bool(gd[tag + "_j2"]))
        d.use_uniform_steps = uni
        out = d.discretize(M.Simulator.satellite_dynamics, gd[tag + "_x"], gd[tag + "_u"], 1.0)
        sel = [o[ks] if o.ndim == 3 else o[:, ks] for o in out]
        print(f"  {tag} include_drag=True (S x{1 if tag == 'g0' else 10000}) use_uniform_steps={uni!s:5}: " + "  ".join(f"{n} {norm_rel_err(o, gd[f'{tag}_{mode}_{n}']):.1e}" for n, o in zip(NAMES, sel)))
gc = np.load(os.path.join(ROOT, "tests/golden/constraint_terms.npz"))
for tag in ("c0", "c1"):
    cm = type("Const", (), {"MU": float(gc["MU"])})()
    ct = M.get_constraint_terms([gc[tag + "_x"]], [gc[tag + "_u"]], cm)
    worst = max(float(np.max(np.abs(np.asarray(ct[k][0]) - gc[f"{tag}_{k}"])) / max(1.0, float(np.max(np.abs(gc[f"{tag}_{k}"])))))
                for k in ct if k != "ubar_hat")
    print(f"  {tag} get_constraint_terms: max scaled error over all terms {worst:.1e}")
# ---- round 2: propagation with the replayed integrator, the coast-to-thrust fixture, the reference's test_linearize_many
gp = np.load(os.path.join(ROOT, "tests/golden/propagate.npz"))
x0 = gp["x0_dim"]
hub = lambda: M.Satellite(x0[0:3].copy(), x0[3:6].copy(), float(x0[6]))
print("== propagated states vs the unmodified reference (Simulator, default integrator = its solve_ivp call replayed)")
for tag, tf, kw, ctl in (("p0 5-orbit coast, drag + J2", 5, {}, None),
                         ("p1 tangential 0.5, tf 2", 2, dict(base_res=100, include_drag=False, include_J2=False), ("tan", 0.5)),
                         ("p2 constant thrust, drag + J2", 3, {}, ("const", gp["p2_thrust"])),
                         ("p4 tangential 0.1, drag + J2", 2, {}, ("tan", 0.1)),
                         ("p5 sequence table ENDING INSIDE the run (end_tau 0.75)", 2, dict(base_res=60, include_drag=False, include_J2=False), ("seq", None))):
    sat = hub()
    scale = M.SatelliteScale(sat=sat)
    c = (M.Controller() if ctl is None else M.ConstantTangentialThrustController([sat], ctl[1]) if ctl[0] == "tan"
         else M.ConstantThrustController(thrust=ctl[1]) if ctl[0] == "const" else M.SequenceController(u=gp["p5_u_tab"], tf_u=1.5, tf_sim=2.0))
    errs = []
    for integ in ("rk45", "rk4"):
        sim = M.Simulator(sats=[hub()], controller=c, scale=scale, **kw)
        sim.integrator = integ
        sim.run(tf=tf)
        errs.append(norm_rel_err(sim.sim_data[sim.sats[0].id], gp[tag[:2] + "_y"]))
    print(f"  {tag}: replayed RK45 {errs[0]:.1e} | fixed-step RK4 (round 1) {errs[1]:.1e}")
gco = np.load(os.path.join(ROOT, "tests/golden/discretize_coast.npz")); cco = OracleConstants(*gco["const"])
for tag in ("c11", "c24"):
    for mode, uni in (("uni", True), ("def", False)):
        d = M.Discretizer(cco); d.use_uniform_steps = uni
        out = d.discretize(M.Simulator.satellite_dynamics, gco[tag + "_x"], gco[tag + "_u"], float(gco[tag + "_tf"]))
        print(f"  coast-to-thrust {tag} use_uniform_steps={uni!s:5}: " + "  ".join(f"{n} {norm_rel_err(o, gco[f'{tag}_{mode}_{n}']):.1e}" for n, o in zip(NAMES, out))
              + f"  | mass rows of B_kp / B_kn {norm_rel_err(out[1][:, 6], gco[f'{tag}_{mode}_B_kp'][:, 6]):.1e} / {norm_rel_err(out[2][:, 6], gco[f'{tag}_{mode}_B_kn'][:, 6]):.1e}")
gm = np.load(os.path.join(ROOT, "tests/golden/discretize_many.npz")); cm_ = OracleConstants(*gm["const"])
selm = lambda o, ks: o[ks] if o.ndim == 3 else o[:, ks]
d = M.Discretizer(cm_)
for tag, uu in (("m0", gm["m0_u"]), ("m0q", gm["m0_uq"])):
    out = d.discretize(M.Simulator.satellite_dynamics, gm["m0_x"], uu, 1.0)
    print(f"  reference test_linearize_many, DEFAULT mode, u {uu.shape}: " + "  ".join(f"{n} {norm_rel_err(selm(o, gm['m0_ks']), gm[f'{tag}_def_{n}']):.1e}" for n, o in zip(NAMES, out)))
dev = torch.device("cuda:0")
ctrl = M.ConstantTangentialThrustController(tangential_thrust=0.5)
print("== kernel times (device-resident, CUDA events, best of 5)")
for name, N, K, tf in (("config 1", 1, 50, 0.5), ("config 2", 64, 100, 1.0), ("config 3", 4096, 200, 2.0), ("1M sweep", 5025, 200, 2.0)):
    Y, c2 = make_constellation(N)
    y0 = torch.from_numpy(Y).to(dev); tfd = torch.full((N,), tf, dtype=torch.float64, device=dev)
    def best(fn):
        ts = []
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return min(ts[1:])
    tp = best(lambda: M.propagate_batch_device(y0, tfd, ctrl, c2, include_drag=False, include_J2=False, T=K))
    tpd = best(lambda: M.propagate_batch_device(y0, tfd, ctrl, c2, include_drag=True, include_J2=True, T=K))
    y, u, _ = M.propagate_batch_device(y0, tfd, ctrl, c2, include_drag=False, include_J2=False, T=K)
    tu = best(lambda: M.discretize_batch_device(y, u, tfd, c2))
    tj = best(lambda: M.discretize_batch_device(y, u, tfd, c2, include_J2=True))
    nn = torch.zeros(N * (K - 1), dtype=torch.int32, device=dev)
    ta = best(lambda: M.discretize_batch_device(y, u, tfd, c2, adaptive=dict(), n_nodes=nn))
    n_int = N * (K - 1)
    if N == 4096:
        import ctypes
        from mpconstellation_b200 import _lib
        o2 = torch.empty((105, n_int), dtype=torch.float64, device=dev); st2 = torch.empty(n_int, dtype=torch.int32, device=dev)
        pd = _lib.make_params(c2, False, True, disc_drag=(2.5, 27123.37))
        sm_ = torch.cuda.current_stream(dev).cuda_stream
        tdg = best(lambda: _lib.check(_lib.lib().mpc_discretize_batch(y.data_ptr(), u.data_ptr(), tfd.data_ptr(), ctypes.byref(pd), N, K, 100, o2.data_ptr(), n_int, 0, st2.data_ptr(), sm_)))
        tda = best(lambda: _lib.check(_lib.lib().mpc_discretize_batch_adaptive(y.data_ptr(), u.data_ptr(), tfd.data_ptr(), ctypes.byref(pd), N, K, 1e-3, 1e-6, 1e-2, o2.data_ptr(), n_int, 0, st2.data_ptr(), None, sm_)))
        from mpconstellation_b200.constraints import constraint_terms_device, dynamics_jacobian_device
        tct = best(lambda: constraint_terms_device(y, u, c2))
        full, _ = M.discretize_batch_device(y, u, tfd, c2)
        tjac = best(lambda: dynamics_jacobian_device(full, N, K))
        print(f"  {name}: with the drag branch of the linearisation: uniform-101 {tdg:.3f} ms, default/adaptive {tda:.3f} ms | "
              f"constraint terms {tct:.3f} ms | sparse dynamics Jacobian ({N*7*(K-1)*16*8/1e6:.0f} MB of values) {tjac:.3f} ms")
    print(f"  {name}: {N} sats x K={K} ({n_int} intervals): propagate {tp:.3f} ms (drag+J2 {tpd:.3f}) | discretize uniform-101 {tu:.3f} ms "
          f"({n_int/tu*1e3:.3e}/s), J2 {tj:.3f} ms | default/adaptive {ta:.3f} ms ({n_int/ta*1e3:.3e}/s, nodes {int(nn.min())}-{int(nn.max())})")
