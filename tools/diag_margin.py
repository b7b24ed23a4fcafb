"""Prints the actual parity margins of the TF32 production mode against the golden fixtures / CPU oracle (GPU box)."""
import dataclasses
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crnn as ocrnn, frontend as ofe  # noqa: E402
from tests.test_crnn_gpu import aux, build  # noqa: E402
from tests.util import gen_wave, golden  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    feats = ofe.features(gen_wave(0, 2))
    g = golden("crnn")
    for tag in ("2023", "2024"):
        cfg0 = ocrnn.CFG_2023 if tag == "2023" else ocrnn.CFG_2024
        for tl in (0, 1):
            P = ocrnn.init_params(cfg0, seed=42, trained_like=bool(tl))
            for prec in (0, 1):
                net = build(cfg0, P, dev, prec)
                net.eval()
                emb, cm = aux(cfg0, 2)
                with torch.no_grad():
                    s, w = net(feats.to(dev), embeddings=None if emb is None else emb.to(dev),
                               classes_mask=None if cm is None else cm.to(dev))
                key = "%s_tl%d" % (tag, tl)
                es = np.abs(s.cpu().numpy() - g["strong_eval_" + key]).max() if "strong_eval_" + key in g else float("nan")
                print("eval %s precision %d: max|strong - golden| = %.3g" % (key, prec, es))
        cfg = dataclasses.replace(cfg0, dropout=0.0)
        P = ocrnn.init_params(cfg, seed=42, trained_like=True)
        for prec in (0, 1):
            net = build(cfg, P, dev, prec, specaugm_t_p=0.0, specaugm_f_p=0.0)
            net.train()
            emb, cm = aux(cfg, 2)
            s, w = net(feats.to(dev), embeddings=None if emb is None else emb.to(dev),
                       classes_mask=None if cm is None else cm.to(dev))
            key = "%s_tl1" % tag
            print("train %s precision %d: max|strong - golden| = %.3g, max|weak - golden| = %.3g" % (
                key, prec, np.abs(s.detach().cpu().numpy() - g["strong_train_" + key]).max(),
                np.abs(w.detach().cpu().numpy() - g["weak_train_" + key]).max()))


if __name__ == "__main__":
    main()
