from desed_task_b200.utils.scaler import TorchScaler  # noqa: F401
