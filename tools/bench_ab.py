"""A/B kernel variants (sedk_set_option switches) inside one eager training step: per-kernel CUDA-event times.

    python tools/bench_ab.py gru_v2 bnglu_small [--batch 24]
"""
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


def profile(L, student, mel, a, y, B, reps=6):
    eng = TrainEngine(student, mel, [B // 2, B - B // 2, 0], bench.L_SAMPLES, opt=FusedAdam(student, 1e-4), use_graph=False)
    for i in range(3):
        eng.step(a[i % 2], y[i % 2])
    torch.cuda.synchronize()
    L.sedk_profile_enable(1)
    for i in range(reps):
        eng.step(a[i % 2], y[i % 2])
    buf = ctypes.create_string_buffer(1 << 16)
    _lib.check(L.sedk_profile_report(buf, len(buf)))
    L.sedk_profile_enable(0)
    prof = {}
    for line in buf.value.decode().strip().splitlines():
        n, c, t = line.split()
        prof[n] = float(t) / reps
    return prof


def main():
    opts = [a for a in sys.argv[1:] if not a.startswith("--")]
    B = 24
    if "--batch" in sys.argv:
        B = int(sys.argv[sys.argv.index("--batch") + 1])
    dev = torch.device("cuda:0")
    L = _lib.lib()
    torch.manual_seed(0)
    student = CRNN(**bench.NET_2023).to(dev)
    student.train()
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1).to(dev)
    a, y = bench.make_batches(2, B, 1, pin=False)
    a = [t.to(dev) for t in a]
    y = [t.to(dev) for t in y]
    base = profile(L, student, mel, a, y, B)
    print("all defaults: eager step %.3f ms" % sum(base.values()))
    for opt in opts:
        alt = 0
        if "=" in opt:                       # name=value: compare the default against that value
            opt, alt = opt.split("=")
            alt = int(alt)
        dflt = L.sedk_get_option(opt.encode(), 1)
        L.sedk_set_option(opt.encode(), alt)
        off = profile(L, student, mel, a, y, B)
        L.sedk_set_option(opt.encode(), dflt)
        print("--- %s: eager step %.3f ms with the option = %d (%.3f ms at the default %d)"
              % (opt, sum(off.values()), alt, sum(base.values()), dflt))
        for k in sorted(set(base) | set(off)):
            t1, t0 = base.get(k, 0.0), off.get(k, 0.0)
            if abs(t1 - t0) > 0.004:
                print("    %-34s on %.4f ms   off %.4f ms" % (k, t1, t0))
    print("--- breakdown with all defaults")
    for k, v in sorted(base.items(), key=lambda kv: -kv[1])[:30]:
        print("    %-34s %.4f ms" % (k, v))


if __name__ == "__main__":
    main()
