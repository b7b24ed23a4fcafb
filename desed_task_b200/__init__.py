"""desed_task_b200 - B200-native (sm_100a) implementation of the DCASE-REPO/DESED_task hot path.

The compute path is hand-written CUDA behind a C ABI (include/sedk.h, desed_task_b200/lib/libsedk.so); this package is the
host-side mirror of the reference's Python interface for that path (desed_task.nnet / data_augm / utils.scaler /
utils.postprocess + the SEDTask4 training step).  Importing the package does not load the library; the first op does
and raises if it is missing - there is no CPU or PyTorch fallback.
"""
__version__ = "0.1.0"
