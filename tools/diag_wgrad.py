"""Debug aid: tcgen05 weight-gradient kernel vs the mma.sync one vs fp64 torch, with error patterns per tap / row / column."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desed_task_b200._lib import check, lib, ptr, stream_ptr  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    L = lib()
    torch.manual_seed(0)
    for (B, T, F, cin, cout) in [(2, 20, 8, 128, 128), (2, 20, 16, 64, 128), (3, 156, 2, 128, 128)]:
        x = torch.randn(B, T, F, cin, device=dev)
        gz = torch.randn(B, T, F, cout, device=dev)
        # reference in fp64: dW[tap][co][ci]
        xp = torch.nn.functional.pad(x.double().permute(0, 3, 1, 2), (1, 1, 1, 1))          # [B,ci,T+2,F+2]
        ref = torch.zeros(9, cout, cin, dtype=torch.float64, device=dev)
        g = gz.double()
        for dy in range(3):
            for dx in range(3):
                xs = xp[:, :, dy:dy + T, dx:dx + F].permute(0, 2, 3, 1)                     # [B,T,F,ci]
                ref[dy * 3 + dx] = torch.einsum("btfo,btfi->oi", g, xs)
        outs = {}
        for on in (0, 1):
            L.sedk_set_tcgen05(on)
            gw = torch.zeros(9, cout, cin, device=dev)
            check(L.sedk_conv_wgrad(ptr(x), ptr(gz), ptr(gw), B, T, F, cin, cout, 0, stream_ptr()), "wgrad")
            torch.cuda.synchronize()
            outs[on] = gw.double()
        L.sedk_set_tcgen05(1)
        sc = ref.abs().max().item()
        print("shape B%d T%d F%d %d->%d: |ref|max %.3f  mma err %.3e  tc5 err %.3e" %
              (B, T, F, cin, cout, sc, (outs[0] - ref).abs().max().item() / sc, (outs[1] - ref).abs().max().item() / sc))
        e = (outs[1] - ref).abs()
        print("  tc5 per-tap max err:", [round(v, 3) for v in (e.amax((1, 2)) / sc).tolist()])
        print("  tc5 per-tap |out|max:", [round(v, 3) for v in (outs[1].abs().amax((1, 2)) / sc).tolist()])
        print("  tc5 err by co block of 32:", [round(v, 3) for v in (e.amax((0, 2)).reshape(-1, 32).amax(1) / sc).tolist()])
        print("  tc5 err by ci block of 32:", [round(v, 3) for v in (e.amax((0, 1)).reshape(-1, 32).amax(1) / sc).tolist()])
        # is the tc5 result a permutation / transpose of the reference?
        t4 = outs[1][4]
        print("  tap4: corr with ref %.3f, with ref^T %.3f" % (
            torch.corrcoef(torch.stack([t4.flatten(), ref[4].flatten()]))[0, 1].item(),
            torch.corrcoef(torch.stack([t4.flatten(), ref[4].t().flatten()]))[0, 1].item() if cin == cout else float("nan")))


if __name__ == "__main__":
    main()
