from desed_task_b200.utils.schedulers import BaseScheduler, ExponentialWarmup  # noqa: F401
