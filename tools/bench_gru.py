"""A/B the GRU recurrence variants (single CTA vs 2-CTA cluster) inside one eager training step (per-kernel CUDA events)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from desed_task_b200 import _lib  # noqa: E402
from desed_task_b200.engine import TrainEngine  # noqa: E402
from desed_task_b200.frontend import MelSpectrogram  # noqa: E402
from desed_task_b200.nnet.CRNN import CRNN  # noqa: E402
from desed_task_b200.optim import FusedAdam  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    L = _lib.lib()
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 24
    torch.manual_seed(0)
    student = CRNN(**bench.NET_2023).to(dev)
    student.train()
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1).to(dev)
    a, y = bench.make_batches(2, B, 1, pin=False)
    a = [t.to(dev) for t in a]
    y = [t.to(dev) for t in y]
    for cs in (1, 2, 1, 2):
        L.sedk_set_gru_cluster(cs)
        eng = TrainEngine(student, mel, [B // 2, B - B // 2, 0], bench.L_SAMPLES, opt=FusedAdam(student, 1e-4),
                          use_graph=False)
        for i in range(3):
            eng.step(a[i % 2], y[i % 2])
        torch.cuda.synchronize()
        L.sedk_profile_enable(1)
        for i in range(6):
            eng.step(a[i % 2], y[i % 2])
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(L.sedk_profile_report(buf, len(buf)))
        L.sedk_profile_enable(0)
        prof = {}
        for line in buf.value.decode().strip().splitlines():
            n, c, t = line.split()
            prof[n] = float(t) / 6
        print("cluster=%d  gru_seq_fwd %.4f ms/step  gru_seq_bwd %.4f ms/step  (2 launches each); whole eager step %.3f ms"
              % (cs, prof["gru_seq_fwd"], prof["gru_seq_bwd"], sum(prof.values())))
    L.sedk_set_gru_cluster(1)


if __name__ == "__main__":
    main()
