"""Import shim for the UNMODIFIED reference (mlfpm/deepof @ /root/reference).

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` (run in the
build container, where /root/reference is mounted) to generate the committed
golden vectors.  Nothing that runs on the GPU box imports this module: the
reference tree does not exist there.

``import deepof`` itself fails in this image (deepof/__init__.py pulls
matplotlib, shapely, ...), so we register a bare namespace package and stub the
five non-arithmetic imports.  All arithmetic modules
(models_new, losses, training, censNetConv_pt, model_utils_new) load from the
read-only tree unmodified.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("DEEPOF_REFERENCE", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "deepof", "clustering"))


def load():
    """Return (models_new, losses, training, model_utils_new) of the reference."""
    if not available():
        raise RuntimeError("reference tree not mounted at %s" % REF_ROOT)
    if "deepof" not in sys.modules or not hasattr(sys.modules["deepof"], "_b200_shim"):
        pkg = types.ModuleType("deepof")
        pkg.__path__ = [os.path.join(REF_ROOT, "deepof")]
        pkg._b200_shim = True
        sys.modules["deepof"] = pkg

        def stub(name, **attrs):
            m = types.ModuleType(name)
            m.__dict__.update(attrs)
            sys.modules[name] = m
            return m

        if "h5py" not in sys.modules:
            stub("h5py")
        if "optuna" not in sys.modules:
            stub("optuna", Trial=object, TrialPruned=type("TrialPruned", (Exception,), {}))
        if "IPython" not in sys.modules:
            ip = stub("IPython")
            ip.display = stub("IPython.display", clear_output=lambda *a, **k: None)
        stub("deepof.utils", validate_parameter=lambda *a, **k: None)
        stub("deepof.data_loading", get_dt=lambda *a, **k: None)
    import deepof.clustering.models_new as M
    import deepof.clustering.losses as L
    import deepof.clustering.training as T
    import deepof.clustering.model_utils_new as U
    return M, L, T, U
