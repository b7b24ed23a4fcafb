"""One eager (graph-free) training step bracketed by cudaProfilerStart/Stop, for ncu --profile-from-start off."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from desed_task_b200.engine import TrainEngine  # noqa: E402
from desed_task_b200.frontend import MelSpectrogram  # noqa: E402
from desed_task_b200.nnet.CRNN import CRNN  # noqa: E402
from desed_task_b200.optim import FusedAdam  # noqa: E402


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "supervised"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else (48 if workload == "mean_teacher" else 24)
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    student = CRNN(**bench.NET_2023).to(dev)
    student.train()
    teacher = None
    bs = [B // 2, B - B // 2, 0]
    if workload == "mean_teacher":
        import copy
        teacher = copy.deepcopy(student)
        teacher.train()
        bs = [B // 4, B // 4, B // 2]
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1).to(dev)
    eng = TrainEngine(student, mel, bs, bench.L_SAMPLES, opt=FusedAdam(student, 1e-3), teacher=teacher,
                      mixup_type="soft" if teacher is not None else None, use_graph=False)
    a, y = bench.make_batches(3, B, 1, pin=False)
    a = [t.to(dev) for t in a]
    y = [t.to(dev) for t in y]
    for i in range(2):
        eng.step(a[i], y[i])
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    eng.step(a[2], y[2])
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
