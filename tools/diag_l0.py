"""Print the A/B deviations of the store-free first block (option l0_fused) instead of asserting (tests/test_layer0_gpu.py)."""
import dataclasses
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import crnn as ocrnn, frontend as ofe  # noqa: E402
from tests.test_crnn_gpu import build  # noqa: E402
from tests.util import gen_wave, maxdiff  # noqa: E402
from desed_task_b200._lib import lib  # noqa: E402

dev = torch.device("cuda:0")
L0 = ("conv0.weight", "conv0.bias", "batchnorm0.weight", "batchnorm0.bias", "glu0.linear.weight", "glu0.linear.bias")


def ab(fn):
    res = {}
    for on in (1, 0):
        lib().sedk_set_option(b"l0_fused", on)
        res[on] = fn()
    lib().sedk_set_option(b"l0_fused", 1)
    return res[1], res[0]


def net_(precision, dropout=0.0, **over):
    cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=dropout)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    kw = dict(specaugm_t_p=0.0, specaugm_f_p=0.0)
    kw.update(over)
    return build(cfg, P, dev, precision, **kw)


x3 = ofe.features(gen_wave(5, 3)).to(dev)
for precision in (1, 0):
    def run():
        net = net_(precision)
        net.eval()
        with torch.no_grad():
            s, w = net(x3)
        ws = list(net._ws.values())[0]
        return s.clone(), w.clone(), ws.conv[0]["out"].clone(), ws.x0.clone()
    a, b = ab(run)
    print("eval precision %d: strong %.3g weak %.3g  block-0 out %.3g  x0 %.3g" % (
        precision, maxdiff(a[0], b[0]), maxdiff(a[1], b[1]), maxdiff(a[2], b[2]), maxdiff(a[3], b[3])))

x5 = ofe.features(gen_wave(9, 5)).to(dev)
wgt = torch.linspace(0.5, 1.5, 156, device=dev)
for mode in ("train", "dropout", "freeze_bn", "specaug"):
    for precision in (1, 0):
        over = {}
        if mode == "freeze_bn":
            over = dict(freeze_bn=True, train_cnn=True)
        if mode == "specaug":
            over = dict(specaugm_t_p=1.0, specaugm_t_l=40, specaugm_f_p=1.0, specaugm_f_l=20)

        def run():
            torch.manual_seed(3)
            net = net_(precision, dropout=0.5 if mode == "dropout" else 0.0, **over)
            net.train()
            net._fwd_count, net._instance = 11, 1
            s, w = net(x5)
            ((s * wgt).mean() + w.mean()).backward()
            sd = net.state_dict()
            ws = list(net._ws.values())[0]
            return (s.detach().clone(), w.detach().clone(),
                    {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None},
                    sd["cnn.cnn.batchnorm0.running_mean"].clone(), sd["cnn.cnn.batchnorm0.running_var"].clone(),
                    ws.conv[0]["out"].clone())
        a, b = ab(run)
        gscale = max(g.abs().max().item() for g in b[2].values())
        errs = {n: (a[2][n] - g).abs().max().item() / max(g.abs().max().item(), 1e-2 * gscale) for n, g in b[2].items()}
        l0 = {n.split("cnn.cnn.")[-1]: "%.2g" % e for n, e in errs.items() if n.endswith(L0)}
        rest = max((e, n) for n, e in errs.items() if not n.endswith(L0))
        print("%s precision %d: strong %.3g weak %.3g rm %.3g rv %.3g out0 %.3g | L0 grads %s | others worst %.2g %s" % (
            mode, precision, maxdiff(a[0], b[0]), maxdiff(a[1], b[1]), maxdiff(a[3], b[3]), maxdiff(a[4], b[4]),
            maxdiff(a[5], b[5]), l0, rest[0], rest[1]))
