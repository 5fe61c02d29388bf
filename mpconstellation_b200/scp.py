"""Batched SCP linearization pass and closed-loop driver: the callers either side of the hot path.

The reference's OptimalController.update (control.py:170-235) does, for ONE satellite and serially:
    reference trajectory (run_nonlinear, :180)  ->  extract_uk (:188)  ->  Optimizer(...).solve_OPT (:198-199, which calls
    Discretizer.discretize, optimizer.py:243-249)  ->  new SequenceController  ->  run_nonlinear again (:227)
and Simulator.run_segments (simulator.py:79-94) advances the satellites segment by segment.

`BatchedSCP` keeps that control flow but evaluates both hot-path legs for ALL satellites at once on the GPU.  The
convex subproblem itself (pyomo + ipopt, optimizer.py:251-613) is out of scope and is injected as a callable:
    solver(hand_off) -> (u_new [N,3,K], tf_new [N])
where `hand_off.sat(s)` yields exactly the five arrays optimizer.py:327-339 indexes.  With pyomo/ipopt installed one
wraps the reference's Optimizer in such a callable; `hold_reference_solver` is a deterministic stand-in (it returns the
reference input unchanged) for boxes without a solver, clearly NOT an optimizer.
"""
from dataclasses import dataclass

import numpy as np

from . import batch
from .control import ConstantTangentialThrustController, ControllerSpec, spec_from
from . import _lib


@dataclass
class Linearization:
    """Result of one batched linearization pass (the hand-off to the subproblem)."""
    x_bar: np.ndarray        # [N,7,K] reference trajectories
    u_bar: np.ndarray        # [N,3,K] reference inputs (extract_uk)
    tf: np.ndarray           # [N]
    matrices: "batch.DiscretizedBatch"

    def sat(self, s):
        """(A_k, B_kp, B_kn, Sigma_k, xi_k) of satellite s -- what Optimizer.solve_OPT appends at optimizer.py:243-249."""
        return self.matrices.sat(s)

    def dynamics_residual(self, s, x, u, tf, nu=None):
        """Residual of the reference's dynamics constraint (optimizer.py:327-339) for a candidate (x [7,K], u [3,K], tf),
        indexed exactly as the pyomo rule does: one value per (i, k)."""
        A_k, B_kp, B_kn, Sigma_k, xi_k = self.sat(s)
        K = x.shape[1]
        res = np.zeros((7, K - 1))
        for k in range(K - 1):
            for i in range(7):
                rhs = (sum(A_k[k, i, j] * x[j, k] for j in range(7))
                       + sum(B_kn[k, i, j] * u[j, k] for j in range(3))
                       + sum(B_kp[k, i, j] * u[j, k + 1] for j in range(3))
                       + Sigma_k[i, k] * tf + xi_k[i, k] + (nu[i, k] if nu is not None else 0.0))
                res[i, k] = x[i, k + 1] - rhs
        return res


def hold_reference_solver(lin):
    """Stand-in for the pyomo/ipopt subproblem: keeps the reference input and final time (NOT an optimizer)."""
    return np.array(lin.u_bar), np.array(lin.tf)


class BatchedSCP:
    def __init__(self, const, base_res=30, tf_horizon=2.0, tf_interval=1.0, n_iterations=2, seed_thrust=0.5,
                 include_J2=False, use_uniform_steps=False, integrator_steps=101, device=0):
        self.const = const
        self.base_res = base_res                  # OptimalController.base_res        (control.py:161)
        self.horizon = float(tf_horizon)          # OptimalController.horizon         (control.py:159)
        self.interval = float(tf_interval)        # OptimalController.interval        (control.py:160)
        self.n_iterations = n_iterations          # SCPn_iterations                   (control.py:166)
        self.seed_thrust = seed_thrust            # T_tan_mag of the seed trajectory  (control.py:178)
        self.include_J2 = include_J2
        self.use_uniform_steps = use_uniform_steps
        self.integrator_steps = integrator_steps
        self.device = device

    def linearize(self, y0, tf, controller):
        """run_nonlinear + extract_uk + discretize for every satellite (control.py:180-188 without the solve).
        The reference trajectory is propagated without drag / J2, as OptimalController.run_nonlinear does (:238-239)."""
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        N = y0.shape[0]
        tfv = np.ascontiguousarray(np.broadcast_to(np.asarray(tf, dtype=np.float64), (N,)))
        K = int(self.base_res * float(np.max(tfv)))
        if self.use_uniform_steps:
            # (k-major layout: the streamed host pass; the per-satellite views of the result hide it)
            res, x, u = batch.propagate_discretize(y0, tfv, controller, self.const, T=K, disc_J2=self.include_J2,
                                                   n_sub_disc=self.integrator_steps - 1, device=self.device,
                                                   layout="kmajor")
        else:
            x, u, _, _ = batch.propagate_batch(y0, tfv, controller, self.const, include_drag=False, include_J2=False,
                                               T=K, device=self.device)
            res = batch.discretize_batch(x, u, tfv, self.const, include_J2=self.include_J2, device=self.device,
                                         adaptive=dict(rtol=1e-3, atol=1e-6, max_step=1e-2))
        return Linearization(np.asarray(x), np.asarray(u), tfv, res)

    def plan(self, y0, solver=hold_reference_solver):
        """OptimalController.update for N satellites: seed trajectory, then n_iterations of (linearize, solve,
        re-propagate with the solved inputs).  Returns the per-satellite input tables [N,3,K] and their tf_u."""
        N = y0.shape[0]
        ctrl = ConstantTangentialThrustController(tangential_thrust=self.seed_thrust)
        tf_u = np.full(N, self.horizon)
        history = []
        for _ in range(self.n_iterations):
            lin = self.linearize(y0, tf_u, ctrl)
            u_new, tf_u = solver(lin)
            u_new = np.ascontiguousarray(u_new, dtype=np.float64)
            history.append(lin)
            # SequenceController(u=u_opt, tf_u=tf_u, tf_sim=tf_u) -> end_tau = 1 (control.py:221)
            ctrl = ControllerSpec(_lib.CTRL_SEQUENCE, (0.0, 0.0, 0.0), u_new, 1.0)
        return u_new, tf_u, history

    def run_segments(self, y0, n_segments, solver=hold_reference_solver, include_drag=True, include_J2=True,
                     sim_base_res=100):
        """Simulator.run_segments with an OptimalController (test_simulator.py:79-95) for N satellites: before every
        segment the controller re-plans (plan), then the satellites fly `interval` with the planned inputs
        (SequenceController(u, tf_u, tf_sim=interval), control.py:217) under the full dynamics."""
        y = np.ascontiguousarray(y0, dtype=np.float64)
        N = y.shape[0]
        traj, plans = [], []
        for _ in range(n_segments):
            u_tab, tf_u, hist = self.plan(y, solver)
            plans.append(hist)
            end_tau = np.asarray(tf_u, dtype=np.float64) / self.interval     # one end_tau per satellite
            ctrl = ControllerSpec(_lib.CTRL_SEQUENCE, (0.0, 0.0, 0.0), u_tab, end_tau)
            T = int(sim_base_res * self.interval)
            ys, _, _, _ = batch.propagate_batch(y, self.interval, ctrl, self.const, include_drag=include_drag,
                                                include_J2=include_J2, T=T, device=self.device)
            traj.append(np.array(ys))
            y = np.ascontiguousarray(ys[:, :, -1])
            if self.horizon - self.interval > 0.1:          # control.py:234-235
                self.horizon -= self.interval
        return np.concatenate(traj, axis=2), plans
