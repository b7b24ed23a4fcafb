"""Load the UNMODIFIED reference modules under a private package name.  Pinning / benchmark infrastructure only (never imported by desed_task_b200).

The repository ships an import shim called `desed_task` (a regular package), which wins over the reference's namespace
package of the same name whenever the repo root is on sys.path.  Pinning scripts and the benchmark's reference legs must
see the reference's own code, so they import it as `_desed_ref.*`: a synthetic package whose `__path__` is the reference's
`desed_task/` directory (relative imports such as `from .CNN import CNN`, desed_task/nnet/CRNN.py:7, keep working).
`utils/__init__.py` pulls `dcase_util` (absent in this image), so utils modules are loaded by file path instead.
"""
import importlib
import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
CANDIDATES = ("/root/reference", os.path.join(_HERE, "_ref"))
PKG = "_desed_ref"


def reference_root():
    """Checkout (build container) or the installed copy under baseline/_ref (travels to the GPU box); None if neither."""
    for c in CANDIDATES:
        if os.path.isfile(os.path.join(c, "desed_task", "nnet", "CRNN.py")):
            return c
    return None


def _package(root):
    pkg = sys.modules.get(PKG)
    if pkg is None:
        pkg = types.ModuleType(PKG)
        pkg.__path__ = [os.path.join(root, "desed_task")]
        sys.modules[PKG] = pkg
    return pkg


def load(root=None):
    """Returns a namespace with the reference's hot-path classes / functions."""
    root = root or reference_root()
    if root is None:
        raise ImportError("no reference checkout (/root/reference) or install (baseline/_ref) found")
    _package(root)
    ns = types.SimpleNamespace(root=root)
    ns.CRNN = importlib.import_module(PKG + ".nnet.CRNN").CRNN
    ns.CNN = importlib.import_module(PKG + ".nnet.CNN")
    ns.RNN = importlib.import_module(PKG + ".nnet.RNN")
    ns.data_augm = importlib.import_module(PKG + ".data_augm")

    def by_path(rel, name):
        s = importlib.util.spec_from_file_location(PKG + "_" + name, os.path.join(root, "desed_task", rel))
        m = importlib.util.module_from_spec(s)
        s.loader.exec_module(m)
        return m
    ns.scaler = by_path("utils/scaler.py", "scaler")
    ns.schedulers = by_path("utils/schedulers.py", "schedulers")
    ns.postprocess = by_path("utils/postprocess.py", "postprocess")
    ns.TorchScaler = ns.scaler.TorchScaler
    ns.ExponentialWarmup = ns.schedulers.ExponentialWarmup
    ns.ClassWiseMedianFilter = ns.postprocess.ClassWiseMedianFilter
    return ns
