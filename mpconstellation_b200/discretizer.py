"""Discretizer: the reference's class (linearize_discretize.py:85-411) with the per-interval work on the GPU.

Constructor, attributes, `discretize(f, x, u, tf)` signature, return order (A_k, B_kp, B_kn, Sigma_k, xi_k)
and shapes are the reference's, so `optimizer.py:243-249` can hold one of these as `self.d` unchanged.
"""
import numpy as np

from . import batch
from .control import spec_from


class Discretizer:
    def __init__(self, const, rho_func=None, drho_func=None, include_drag=False, include_J2=False,
                 use_scipy_ZOH=False):
        self.const = const
        self.include_drag = include_drag
        self.include_J2 = include_J2
        # the reference's two input holds (u_FOH / scipy interp1d 'linear') are the same function of tau
        # (linearize_discretize.py:327-331); the kernel evaluates that first-order hold directly
        self.use_scipy_ZOH = use_scipy_ZOH
        self.rho_func = rho_func
        self.drho_func = drho_func
        # ODE solver parameters (linearize_discretize.py:104-105)
        self.ivp_max_step = 1e-2
        self.ivp_solver = 'RK45'
        # numerical integration parameters (linearize_discretize.py:108-109)
        self.integrator_steps = 101
        self.use_uniform_steps = False
        # scipy.integrate.solve_ivp defaults (the reference passes neither, linearize_discretize.py:37-41)
        self.ivp_rtol = 1e-3
        self.ivp_atol = 1e-6
        self.device = 0

    # -- option handling ---------------------------------------------------------------------------
    def _check_options(self, f):
        name = getattr(f, "__name__", None)
        if name != "satellite_dynamics":
            raise NotImplementedError(
                "the GPU discretizer integrates the reference's Simulator.satellite_dynamics on the device; "
                f"an arbitrary Python f ({f!r}) cannot be linearized there and there is no CPU fallback")
        if self.include_drag:
            # the reference cannot run this branch either: rho_func defaults to None and Constants has no
            # CD attribute (linearize_discretize.py:162-169, constants.py:11-20)
            if self.rho_func is None:
                raise TypeError("'NoneType' object is not callable")
            if not hasattr(self.const, "CD"):
                raise AttributeError("'Constants' object has no attribute 'CD'")
            if self.drho_func is None:
                raise TypeError("'NoneType' object is not callable")
        if self.ivp_solver != 'RK45':
            raise NotImplementedError(f"ivp_solver={self.ivp_solver!r}: the device integrator is a fixed-step 4th-order Runge-Kutta method "
                                      "on the reference's node grid (stated against RK45)")
        if self.use_uniform_steps and int(self.integrator_steps) < 2:
            raise ValueError("integrator_steps must be >= 2")

    def _disc_drag(self, x):
        """(CD, rho) for the device, or None.  The drag branch (linearize_discretize.py:160-169) calls rho_func(r) and
        drho_func(r) at every state; the kernels implement a CONSTANT density (what Simulator.get_atmo_density
        returns, simulator.py:112), so both callables are sampled on the reference trajectory and anything else is
        rejected loudly -- there is no CPU fallback."""
        if not self.include_drag:
            return None
        pos = np.moveaxis(x[:, 0:3, :], 1, 2).reshape(-1, 3)
        if pos.shape[0] > 64:
            pos = pos[np.linspace(0, pos.shape[0] - 1, 64).astype(int)]
        rho = np.array([float(self.rho_func(r)) for r in pos])
        drho = np.array([float(np.max(np.abs(self.drho_func(r)))) for r in pos])
        if rho.size and (np.ptp(rho) != 0.0 or np.any(drho != 0.0)):
            raise NotImplementedError("the GPU discretizer linearizes drag for a constant density only "
                                      "(rho_func constant, drho_func zero along the trajectory)")
        return float(self.const.CD), float(rho[0]) if rho.size else 0.0

    # -- the reference entry point -------------------------------------------------------------------
    def discretize(self, f, x, u, tf):
        """x (7,K), u (3,K), tf scalar -> A_k (K-1,7,7), B_kp (K-1,7,3), B_kn (K-1,7,3), Sigma_k (7,K-1),
        xi_k (7,K-1).  ref: linearize_discretize.py:334-390."""
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        if x.ndim != 2 or x.shape[0] != 7:
            raise ValueError(f"x must be (7, K), got {x.shape}")
        res = self.discretize_batch(f, x[None], u[None] if u.ndim == 2 else u, tf)
        return res.sat(0)

    # -- batched form ------------------------------------------------------------------------------
    def discretize_batch(self, f, x, u, tf, out=None, check=True):
        """x [N,7,K], u [N,3,K], tf scalar or [N] -> batch.DiscretizedBatch (SoA + reference-shaped views)."""
        self._check_options(f)
        x = np.asarray(x, dtype=np.float64)
        u = np.asarray(u, dtype=np.float64)
        K = x.shape[2]
        if K < 2:
            raise ValueError("need at least K = 2 temporal nodes")
        # u may have its own column count: the reference's hold takes its grid from u (linearize_discretize.py:308-315)
        # use_uniform_steps=False (the reference default): quadrature on the steps scipy's RK45 controller accepts
        # (replayed on the device, scipy's default rtol/atol as the reference passes none);
        # use_uniform_steps=True: fixed-step fourth-order Runge-Kutta(-Nystrom) on the uniform integrator_steps grid.
        adaptive = None if self.use_uniform_steps else dict(rtol=self.ivp_rtol, atol=self.ivp_atol,
                                                            max_step=float(self.ivp_max_step))
        return batch.discretize_batch(x, u, tf, self.const, include_J2=self.include_J2,
                                      include_drag=bool(self.include_drag), disc_drag=self._disc_drag(x),
                                      n_sub=max(1, int(self.integrator_steps) - 1), out=out, device=self.device,
                                      check=check, adaptive=adaptive)

    @staticmethod
    def extract_uk(x_k, tau_k, controller):
        """3 x K controller outputs at the sampled states.  ref: linearize_discretize.py:393-411.
        (Simulator.run on the GPU already returns these next to the trajectory: sim.sim_u.)"""
        u_func = spec_from(controller).host_u_func() if not callable(controller) or hasattr(controller, "device_spec") \
            else controller
        return np.column_stack([u_func(x_k[:, i], tau_k[i]) for i in range(x_k.shape[1])])
