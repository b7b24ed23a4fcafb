"""Micro-benchmark of the fused log-mel kernel: clips/s and algorithmic GB/s (960 512 B/clip) vs the measured HBM peak."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desed_task_b200.frontend import MelSpectrogram, new_minmax  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    mel = MelSpectrogram(16000, 2048, 2048, 256, 0, 8000, n_mels=128, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1).to(dev)
    peak = 6573.2
    try:
        peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    from desed_task_b200._lib import lib
    variants = [("v2", 1), ("v1", 0)] if "--ab" in sys.argv else [("default", None)]
    for tag, opt in variants:
      if opt is not None:
        lib().sedk_set_option(b"logmel_v2", opt)
      for B in (24, 64, 256, 1024):
        wave = torch.randn(B, 160000, device=dev) * 0.1
        mm = new_minmax(B, dev)
        for _ in range(3):
            mel.run(wave, log=True, minmax=mm)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 10
        e0.record()
        for _ in range(n):
            mel.run(wave, log=True, minmax=mm)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        gbs = B * 960512 / ms / 1e6
        print(json.dumps({"kernel": "logmel", "variant": tag, "B": B, "ms": round(ms, 4), "clips_per_s": round(B / ms * 1e3, 1),
                          "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 4)}))


if __name__ == "__main__":
    main()
