"""The reference's own modules on one B200 through PyTorch's library kernels: clips/s per workload."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from baseline import reference_arm  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    out = {"torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "cudnn_allow_tf32": torch.backends.cudnn.allow_tf32, "matmul_allow_tf32": torch.backends.cuda.matmul.allow_tf32}
    for workload, B in (("supervised", 24), ("mean_teacher", 48), ("inference", 64)):
        a, y = bench.make_batches(6, B, 42, pin=False)
        a = [t.to(dev) for t in a]
        y = [t.to(dev) for t in y]
        out[workload] = dict(batch=B, **reference_arm.gpu_library_baseline(dev, workload, B, a, y, steps=20, warmup=5))
        print(workload, out[workload], file=sys.stderr)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
