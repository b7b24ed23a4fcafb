"""`desed_task.utils` of the reference (desed_task/utils/__init__.py:1-2 exports ManyHotEncoder and ExponentialWarmup).

scaler / schedulers / postprocess are served by desed_task_b200; `encoder`, `torch_utils`, `download` resolve to the
reference's own files through the extended `__path__`.  `ManyHotEncoder` (needs the third-party `dcase_util`) is imported
lazily so that the hot-path names stay importable where that dependency is absent."""
import pkgutil as _pkgutil

from .. import _extend

__path__ = _extend(_pkgutil.extend_path(__path__, __name__), "utils")

from .scaler import TorchScaler  # noqa: F401,E402
from .schedulers import ExponentialWarmup  # noqa: F401,E402
from .postprocess import ClassWiseMedianFilter  # noqa: F401,E402


def __getattr__(name):
    if name in ("ManyHotEncoder", "CatManyHotEncoder"):
        from . import encoder          # the reference's desed_task/utils/encoder.py
        return getattr(encoder, name)
    raise AttributeError("module %r has no attribute %r" % (__name__, name))
