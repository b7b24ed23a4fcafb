"""Layer-by-layer forward/backward comparison of the CUDA CRNN against the CPU oracle (debug aid, run on the GPU box)."""
import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crnn as ocrnn, frontend as ofe, trainer as otr  # noqa: E402
from desed_task_b200.nnet import CRNN as crnn_mod  # noqa: E402


def rel(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return (a - b).abs().max().item() / (b.abs().max().item() + 1e-20)


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "2023"
    prec = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    B = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    dev = torch.device("cuda:0")
    cfg = ocrnn.CFG_2023 if tag == "2023" else ocrnn.CFG_2024
    cfg = dataclasses.replace(cfg, dropout=0.0)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    g = torch.Generator().manual_seed(0)
    wave = torch.randn(B, 160000, generator=g) * 0.1
    feats = ofe.features(wave)
    kw = dict(nclass=cfg.nclass, dropout=0.0, n_RNN_cell=cfg.n_RNN_cell, nb_filters=list(cfg.nb_filters),
              pooling=[list(p) for p in cfg.pooling], kernel_size=[3] * 7, padding=[1] * 7, stride=[1] * 7,
              specaugm_t_p=0.0, specaugm_f_p=0.0, activation="glu")
    emb = cm = None
    if cfg.use_embeddings:
        kw.update(use_embeddings=True, embedding_size=768, embedding_type="frame", aggregation_type="pool1d")
        emb = torch.randn(B, 768, 496, generator=g)
        cm = torch.zeros(B, cfg.nclass, dtype=torch.bool)
        cm[:, :10] = True
        cm[-1] = ~cm[-1]
    net = crnn_mod.CRNN(**kw)
    net.load_state_dict(P, strict=True)
    net = net.to(dev)
    net.precision = prec
    net.train()
    Pt = {k: (v.clone().requires_grad_(True) if ocrnn.is_float_param(k) else v.clone()) for k, v in P.items()}
    col = {}
    s_or, w_or = ocrnn.crnn_forward(Pt, feats, cfg, True, embeddings=emb, classes_mask=cm, collect=col)
    for v in col.values():
        v.retain_grad()
    ys = (torch.rand(s_or.shape, generator=g) < 0.1).float()
    yw = (ys.sum(-1) > 0).float()
    loss = otr.bce(s_or, ys) + otr.bce(w_or, yw)
    loss.backward()

    s, w = net(feats.to(dev), embeddings=None if emb is None else emb.to(dev), classes_mask=None if cm is None else cm.to(dev))
    ws = list(net._ws.values())[0]
    print("== forward (precision %d) ==" % prec)
    for i in range(len(cfg.nb_filters)):
        z = ws.conv[i]["z"].permute(0, 3, 1, 2)
        o = ws.conv[i]["out"].permute(0, 3, 1, 2)
        print("conv%d z rel %.3e   pool%d rel %.3e" % (i, rel(z, col["conv%d" % i]), i, rel(o, col["pool%d" % i])))
    if emb is not None:
        print("fused rel %.3e" % rel(ws.fused, col["fused"]))
    for l, d in enumerate(ws.gru):
        print("gru layer %d out absmax %.3e" % (l, d["out"].abs().max().item()))
    print("rnn_out rel %.3e" % rel(ws.gru[-1]["out"], col["rnn_out"]))
    print("strong maxabs %.3e weak maxabs %.3e" % ((s.cpu() - s_or).abs().max().item(), (w.cpu() - w_or).abs().max().item()))
    l2 = torch.nn.functional.binary_cross_entropy(s, ys.to(dev)) + torch.nn.functional.binary_cross_entropy(w, yw.to(dev))
    l2.backward()
    print("== backward ==  loss %.6f vs %.6f" % (l2.item(), loss.item()))
    print("rnn_out grad rel %.3e" % rel(ws.gru[-1]["gout"], col["rnn_out"].grad))
    print("cnn_out grad rel %.3e" % rel(ws.conv[-1]["gout"].reshape(col["cnn_out"].shape), col["cnn_out"].grad))
    for i in reversed(range(len(cfg.nb_filters))):
        go = ws.conv[i]["gout"].permute(0, 3, 1, 2)
        gz = ws.conv[i]["gy"].permute(0, 3, 1, 2)
        print("layer %d: gout rel %.3e   gz rel %.3e" % (i, rel(go, col["pool%d" % i].grad), rel(gz, col["conv%d" % i].grad)))
    gscale = max(Pt[n].grad.abs().max().item() for n in ocrnn.param_names(P))
    for n, p in net.named_parameters():
        ref = Pt[n].grad
        err = (p.grad.cpu() - ref).abs().max().item() / max(ref.abs().max().item(), 1e-2 * gscale)
        print("grad %-40s rel %.3e  (|ref|max %.3e)" % (n, err, ref.abs().max().item()))


if __name__ == "__main__":
    main()
