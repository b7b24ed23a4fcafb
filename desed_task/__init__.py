"""Import-path shim: `desed_task.*` names of the reference's hot path, served by desed_task_b200 (B200 kernels).

Lets the DCASE recipes keep their imports (`from desed_task.nnet.CRNN import CRNN`, `from desed_task.data_augm import mixup`,
`from desed_task.utils.scaler import TorchScaler`, ...) when this repository precedes the reference on sys.path.  Only the
hot-path modules live here.  Everything else the recipes import (`desed_task.dataio`, `desed_task.utils.encoder`,
`desed_task.utils.torch_utils`, `desed_task.evaluation`, ...; recipes/dcase2023_task4_baseline/train_sed.py:11-16) stays with
the reference: this package extends its `__path__` over every other `desed_task` directory it can find (later sys.path
entries - the reference is a namespace package -, an installed distribution, or $DESED_TASK_REFERENCE), so those submodules
resolve to the reference's own files while the hot-path names resolve here first."""
import os as _os
import pkgutil as _pkgutil
import sys as _sys

__path__ = _pkgutil.extend_path(__path__, __name__)


def _reference_dirs(sub=""):
    """`desed_task[/sub]` directories of the reference that are not this shim: $DESED_TASK_REFERENCE (a checkout root or the
    package directory itself) and sys.path entries (covers `pip install -e` .pth entries and plain PYTHONPATH)."""
    here = _os.path.dirname(_os.path.abspath(__file__))
    roots = []
    env = _os.environ.get("DESED_TASK_REFERENCE")
    if env:
        roots += [env, _os.path.dirname(env.rstrip("/"))]
    roots += [p for p in _sys.path if isinstance(p, str)]
    out = []
    for r in roots:
        d = _os.path.join(r or ".", "desed_task")
        if _os.path.isdir(d) and _os.path.abspath(d) != here:
            d = _os.path.join(d, sub) if sub else d
            if _os.path.isdir(d) and d not in out:
                out.append(d)
    return out


def _extend(path, sub=""):
    for d in _reference_dirs(sub):
        if d not in path:
            path.append(d)
    return path


_extend(__path__)
