from desed_task_b200.nnet.CNN import CNN, GLU, ContextGating  # noqa: F401
