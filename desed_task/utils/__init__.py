from .scaler import TorchScaler  # noqa: F401
from .schedulers import ExponentialWarmup  # noqa: F401
from .postprocess import ClassWiseMedianFilter  # noqa: F401
