"""Hot-path modules (CNN, RNN, CRNN) come from desed_task_b200; anything else resolves to the reference's desed_task/nnet."""
import pkgutil as _pkgutil

from .. import _extend

__path__ = _extend(_pkgutil.extend_path(__path__, __name__), "nnet")
