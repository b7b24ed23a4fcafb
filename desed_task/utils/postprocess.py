from desed_task_b200.utils.postprocess import ClassWiseMedianFilter, median_filter  # noqa: F401
