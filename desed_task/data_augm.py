from desed_task_b200.data_augm import mixup, frame_shift, add_noise  # noqa: F401
