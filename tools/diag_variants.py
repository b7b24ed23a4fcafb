"""Per-parameter gradient error of the constructor alternates vs the oracle (diagnostic, GPU)."""
import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crnn as ocrnn, frontend as ofe, trainer as otr  # noqa: E402
from tests.test_crnn_gpu import build  # noqa: E402
from tests.util import gen_wave  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    names = sys.argv[1:] or ["relu", "leakyrelu", "cg"]
    x = ofe.features(gen_wave(0, 2))
    g = torch.Generator().manual_seed(11)
    ys = (torch.rand(2, 10, 156, generator=g) < 0.1).float()
    yw = (ys.sum(-1) > 0).float()
    for act in names:
        for precision in (1, 0):
            cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=0.0, activation=act)
            P = ocrnn.init_params(cfg, seed=42, trained_like=True)
            net = build(cfg, P, dev, precision, specaugm_t_p=0.0, specaugm_f_p=0.0, activation=act)
            net.train()
            s, w = net(x.to(dev))
            loss = torch.nn.functional.binary_cross_entropy(s, ys.to(dev)) + torch.nn.functional.binary_cross_entropy(w, yw.to(dev))
            loss.backward()
            Pt = {k: (v.clone().requires_grad_(True) if ocrnn.is_float_param(k) else v.clone()) for k, v in P.items()}
            col = {}
            so, wo = ocrnn.crnn_forward(Pt, x, cfg, True, collect=col)
            (otr.bce(so, ys) + otr.bce(wo, yw)).backward()
            gscale = max(Pt[n].grad.abs().max().item() for n in ocrnn.param_names(P))
            print("== %s precision %d: out diff %.3g, loss %.6f vs %.6f" % (act, precision, (s.cpu() - so).abs().max().item(),
                                                                         loss.item(), (otr.bce(so, ys) + otr.bce(wo, yw)).item()))
            for n, p in net.named_parameters():
                ref = Pt[n].grad
                err = (p.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-2 * gscale)
                if err > 5e-4:
                    print("   %-34s rel err %.3g  (ref max %.3g, gscale %.3g)" % (n, err, ref.abs().max().item(), gscale))
            for i in range(7):
                y = col["conv%d" % i]
                print("   layer %d: pre-BN conv out |.|<1e-5: %d of %d" % (i, int((y.abs() < 1e-5).sum()), y.numel()))


if __name__ == "__main__":
    main()
