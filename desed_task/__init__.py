"""Import-path shim: `desed_task.*` names of the reference's hot path, served by desed_task_b200 (B200 kernels).

Lets the DCASE recipes keep their imports (`from desed_task.nnet.CRNN import CRNN`, `from desed_task.data_augm import mixup`,
`from desed_task.utils.scaler import TorchScaler`, ...) when this repository precedes the reference on sys.path.  Only the
hot-path modules exist here; dataio / encoder / evaluation stay with the reference package (out of scope)."""
