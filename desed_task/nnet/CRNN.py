from desed_task_b200.nnet.CRNN import CRNN  # noqa: F401
