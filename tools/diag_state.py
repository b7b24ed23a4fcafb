"""Diagnose cross-run state: does a precision-0 run (tcgen05 on/off) change what a later precision-1 trainer step computes?"""
import dataclasses, random, sys
import numpy as np, torch
sys.path.insert(0, ".")
import tests.test_trainer_gpu as TT, tests.test_crnn_gpu as TC
from oracle import crnn as ocrnn, frontend as ofe
from tests.util import gen_wave
from desed_task_b200._lib import lib
dev = torch.device("cuda:0")

def trainer_grads():
    mod, P, cfg = TT.make(dev)
    audio, labels = TT.data()
    random.seed(0); np.random.seed(0); torch.manual_seed(0)
    loss = mod.training_step((audio.to(dev), labels.to(dev)), 0)
    mod.on_before_zero_grad(); mod.opt.zero_grad(); loss.backward()
    torch.cuda.synchronize()
    ws = list(mod.sed_student._ws.values())[0]
    bufs = dict(x0=ws.x0.clone())
    for i in range(3):
        for k in ("z", "gy", "out", "gout"):
            bufs["%s%d" % (k, i)] = ws.conv[i][k].clone()
    return {n: p.grad.clone() for n, p in mod.sed_student.named_parameters()}, bufs, loss.item()

def ab_step(on, precision=0):
    cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=0.0)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    feats = ofe.features(gen_wave(0, 2))
    lib().sedk_set_tcgen05(on)
    net = TC.build(cfg, P, dev, precision, specaugm_t_p=0.0, specaugm_f_p=0.0)
    net.train()
    s, w = net(feats.to(dev))
    (s.mean() + w.mean()).backward()
    torch.cuda.synchronize()
    lib().sedk_set_tcgen05(1)

def cmp(tag, a, b):
    ga, ba, la = a; gb, bb, lb = b
    print("==", tag, "loss", la, lb)
    for n in ga:
        d = (ga[n] - gb[n]).abs().max().item(); m = ga[n].abs().max().item()
        if d > 1e-4 * max(m, 1e-6):
            print("   grad %-36s maxdiff %.3e  (|g|max %.3e)" % (n, d, m))
    for n in ba:
        d = (ba[n] - bb[n]).abs().max().item(); m = ba[n].abs().max().item()
        if d > 1e-5 * max(m, 1e-6):
            print("   buf  %-36s maxdiff %.3e  (|v|max %.3e)" % (n, d, m))

base = trainer_grads()
again = trainer_grads(); cmp("repeat (no interleaved run)", base, again)
ab_step(1, 1); r = trainer_grads(); cmp("after precision-1 run", base, r)
ab_step(1, 0); r = trainer_grads(); cmp("after precision-0 tcgen05 ON", base, r)
ab_step(0, 0); r = trainer_grads(); cmp("after precision-0 tcgen05 OFF", base, r)
for it in range(4):
    r = trainer_grads(); cmp("one more %d" % it, base, r)
    d = (r[1]["gy1"] - base[1]["gy1"]).abs()
    bad = d > 1e-6
    print("   gy1 shape", tuple(d.shape), "bad elements", int(bad.sum()))
    if bad.any():
        idx = bad.nonzero()
        for ax, nm in enumerate("btfc"):
            u = idx[:, ax].unique()
            print("     axis", nm, "count", len(u), "min", int(u.min()), "max", int(u.max()), u[:24].tolist())
        print("     ratio r/base at bad:", (r[1]["gy1"][bad] / base[1]["gy1"][bad])[:8].tolist())
    d = (r[1]["gy0"] - base[1]["gy0"]).abs(); print("   gy0 bad", int((d > 1e-7).sum()), "of", d.numel())
ab_step(1, 1)
r = trainer_grads(); cmp("after another precision-1 run", base, r)
