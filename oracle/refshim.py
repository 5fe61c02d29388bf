"""TEST INFRASTRUCTURE ONLY -- import shim for the *unmodified* reference.

Two places the reference can be loaded from:
  * /root/reference (the build container): used by tests/golden/make_golden.py to generate the committed fixtures
    that pin oracle/mpc_oracle.py and oracle/mpc_oracle.c;
  * oracle/_ref/ (git-ignored build output, travels to the GPU box like the built .so files): the reference's six
    hot-path modules COMPILED to Python bytecode by `compile_reference()` (called from __graft_entry__.build() where
    /root/reference is mounted) -- no reference source is copied into this repository.  bench.py --impl reference
    times exactly that code: the unmodified reference on the box's host cores.

The reference's control.py pulls in optimizer.py (pyomo) and sim_plotter.py
(matplotlib); neither is installed and neither touches the hot path, so they
are stubbed in sys.modules.  `simulator` must be imported before `control`
(circular import: simulator.py:5 / control.py:2).
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MPC_REFERENCE_ROOT", "/root/reference")
COMPILED_ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
HOT_PATH_MODULES = ("simulator", "control", "linearize_discretize", "satellite", "satellite_scale", "constants")


def compile_reference(src_root=None, out_root=None):
    """py_compile the reference's hot-path modules from where they lie into oracle/_ref/<module>.bc (sourceless
    bytecode, importable on the GPU box where /root/reference does not exist).  Returns the list of files written,
    [] when the reference tree is not present."""
    import py_compile
    src_root = src_root or REFERENCE_ROOT
    out_root = out_root or COMPILED_ROOT
    if not os.path.isdir(src_root):
        return []
    os.makedirs(out_root, exist_ok=True)
    done = []
    for m in HOT_PATH_MODULES:
        # (".bc": `*.pyc` files are dropped by the snapshot that ships the tree to the GPU box)
        done.append(py_compile.compile(os.path.join(src_root, m + ".py"), cfile=os.path.join(out_root, m + ".bc"),
                                       doraise=True))
    return done


class _BytecodeFinder:
    """Meta-path finder: imports the reference's hot-path modules from oracle/_ref/<module>.bc (sourceless bytecode)."""

    @staticmethod
    def find_spec(name, path=None, target=None):
        import importlib.machinery
        import importlib.util
        if name not in HOT_PATH_MODULES:
            return None
        f = os.path.join(COMPILED_ROOT, name + ".bc")
        if not os.path.exists(f):
            return None
        return importlib.util.spec_from_loader(name, importlib.machinery.SourcelessFileLoader(name, f), origin=f)


def reference_location():
    """(path, kind): where the reference can be imported from here, kind 'source' | 'bytecode' | None."""
    if os.path.isdir(REFERENCE_ROOT):
        return REFERENCE_ROOT, "source"
    if all(os.path.exists(os.path.join(COMPILED_ROOT, m + ".bc")) for m in HOT_PATH_MODULES):
        return COMPILED_ROOT, "bytecode"
    return None, None


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
    root, kind = reference_location()
    if root is None:
        raise RuntimeError(f"reference neither at {REFERENCE_ROOT} nor compiled under {COMPILED_ROOT}")
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
    if kind == "bytecode":
        # control.py imports optimizer (pyomo) and sim_plotter (matplotlib): not on the hot path, not compiled
        if "optimizer" not in sys.modules:
            _stub("optimizer", Optimizer=_Anything())
        if "sim_plotter" not in sys.modules:
            _stub("sim_plotter", plot_orbit_3D=_Anything(), __all__=["plot_orbit_3D"])
        if not any(f is _BytecodeFinder for f in sys.meta_path):
            sys.meta_path.insert(0, _BytecodeFinder)
    elif root not in sys.path:
        sys.path.insert(0, root)
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
