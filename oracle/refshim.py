"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

Only usable in the build container, where /root/reference is mounted.  Nothing
on the GPU box may import this module (the reference does not travel); it is
used solely by tests/golden/make_golden.py to generate the committed fixtures
that pin oracle/mpc_oracle.py and oracle/mpc_oracle.c.

The reference's control.py pulls in optimizer.py (pyomo) and sim_plotter.py
(matplotlib); neither is installed and neither touches the hot path, so they
are stubbed in sys.modules.  `simulator` must be imported before `control`
(circular import: simulator.py:5 / control.py:2).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MPC_REFERENCE_ROOT", "/root/reference")


class _Anything:
    """Absorbs any attribute access / call made on a stubbed plotting module."""

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *args, **kwargs):
        return _Anything()


def _stub(name, **attrs):
    mod = types.ModuleType(name)
    mod.__dict__.update(attrs)
    sys.modules[name] = mod
    return mod


def load_reference():
    """Returns a namespace with the reference's hot-path classes."""
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"reference tree not present at {REFERENCE_ROOT}")
    if "matplotlib" not in sys.modules:
        plt = _stub("matplotlib.pyplot", subplots=_Anything(), show=_Anything(), Circle=_Anything(),
                    axes=_Anything(), title=_Anything(), gca=_Anything(), legend=_Anything())
        _stub("matplotlib").pyplot = plt
        _stub("mpl_toolkits").mplot3d = _stub("mpl_toolkits.mplot3d")
    if "pyomo" not in sys.modules:
        _stub("pyomo")
        _stub("pyomo.core")
        _stub("pyomo.core.base")
        _stub("pyomo.core.base.expression", ScalarExpression=object)
        _stub("pyomo.environ")
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import simulator as ref_simulator  # noqa: F401  (must precede control)
    import control as ref_control
    import linearize_discretize as ref_ld
    import satellite as ref_satellite
    import satellite_scale as ref_scale
    import constants as ref_constants
    return types.SimpleNamespace(
        Simulator=ref_simulator.Simulator,
        Discretizer=ref_ld.Discretizer,
        get_matrices=ref_ld.get_matrices,
        Satellite=ref_satellite.Satellite,
        SatelliteScale=ref_scale.SatelliteScale,
        Constants=ref_constants.Constants,
        constants=ref_constants,
        Controller=ref_control.Controller,
        ConstantThrustController=ref_control.ConstantThrustController,
        ConstantTangentialThrustController=ref_control.ConstantTangentialThrustController,
        SequenceController=ref_control.SequenceController,
    )
