"""GPU A/B parity of the store-free first block (csrc/layer0.cu, option l0_fused, default on) against the kernels that move
the first convolution's output through HBM (conv0 + BN/GLU + BN-apply + conv0 weight gradient).

Both sides compute the stencil with the same FMA order, so in the fp32-equivalent mode the FORWARD agrees to an ulp
(measured 6e-8 on the block output, 2.4e-7 on the posteriors; BatchNorm sums are taken in a different order).  In the TF32
mode the pooled block output is rounded to TF32 for the next convolution, where an fp32 ulp can flip a TF32 ulp (2.4e-4
relative): the usual TF32 A/B bounds apply.  The BACKWARD differs by construction: the fused path evaluates the BatchNorm
backward + weight gradient in closed form from fp32/fp64 sums, the unfused one rounds g_z to fp32 (and to TF32 in precision
0) before the weight-gradient GEMM.  Dropout masks are drawn from different counters on the two paths (statistics tested
here; forward/backward mask consistency by the finite-difference test in test_variants_gpu.py, which runs the fused path).
The fused path is also what every oracle-parity test of the suite runs (test_crnn_gpu.py, test_trainer_gpu.py).
Measured deviations: tools/diag_l0.py -> profiles/r2_layer0_ab.txt."""
import dataclasses

import pytest
import torch

from oracle import crnn as ocrnn, frontend as ofe
from tests.test_crnn_gpu import build
from tests.util import gen_wave, maxdiff

pytestmark = pytest.mark.gpu

L0 = ("cnn.cnn.conv0.weight", "cnn.cnn.conv0.bias", "cnn.cnn.batchnorm0.weight", "cnn.cnn.batchnorm0.bias",
      "cnn.cnn.glu0.linear.weight", "cnn.cnn.glu0.linear.bias")


def _ab(dev, fn):
    from desed_task_b200._lib import lib
    res = {}
    for on in (1, 0):
        lib().sedk_set_option(b"l0_fused", on)
        try:
            assert lib().sedk_get_option(b"l0_fused", -1) == on
            res[on] = fn()
        finally:
            lib().sedk_set_option(b"l0_fused", 1)
    return res[1], res[0]


def _net(dev, precision, dropout=0.0, **over):
    cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=dropout)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    kw = dict(specaugm_t_p=0.0, specaugm_f_p=0.0)
    kw.update(over)
    return build(cfg, P, dev, precision, **kw)


@pytest.mark.parametrize("precision,tol", [(1, 5e-7), (0, 5e-4)])
def test_eval_forward_agrees(dev, precision, tol):
    x = ofe.features(gen_wave(5, 3)).to(dev)

    def run():
        net = _net(dev, precision)
        net.eval()
        with torch.no_grad():
            s, w = net(x)
        return s.clone(), w.clone()

    a, b = _ab(dev, run)
    assert maxdiff(a[0], b[0]) < tol and maxdiff(a[1], b[1]) < tol


@pytest.mark.parametrize("precision,tol_out,tol_grad", [(1, 2e-6, 2e-4), (0, 5e-4, 3e-2)])
@pytest.mark.parametrize("mode", ["train", "freeze_bn", "specaug"])
def test_training_step_agrees_with_the_unfused_kernels(dev, precision, tol_out, tol_grad, mode):
    """B = 5 clips (odd: the persistent group loops end on ragged tails), every parameter gradient + BN running statistics."""
    x = ofe.features(gen_wave(9, 5)).to(dev)
    wgt = torch.linspace(0.5, 1.5, 156, device=dev)
    over = {}
    if mode == "freeze_bn":
        over = dict(freeze_bn=True, train_cnn=True)
    if mode == "specaug":
        over = dict(specaugm_t_p=1.0, specaugm_t_l=40, specaugm_f_p=1.0, specaugm_f_l=20)

    def run():
        torch.manual_seed(3)
        net = _net(dev, precision, **over)
        net.train()
        net._fwd_count, net._instance = 11, 1         # pin the dropout / SpecAugment seed across the two runs
        s, w = net(x)
        ((s * wgt).mean() + w.mean()).backward()
        sd = net.state_dict()
        return (s.detach().clone(), w.detach().clone(),
                {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None},
                sd["cnn.cnn.batchnorm0.running_mean"].clone(), sd["cnn.cnn.batchnorm0.running_var"].clone())

    a, b = _ab(dev, run)
    assert maxdiff(a[0], b[0]) < tol_out and maxdiff(a[1], b[1]) < tol_out
    assert maxdiff(a[3], b[3]) < 1e-6 and maxdiff(a[4], b[4]) < 1e-6
    gscale = max(g.abs().max().item() for g in b[2].values())
    worst = ("", 0.0)
    for n, g in b[2].items():
        err = (a[2][n] - g).abs().max().item() / max(g.abs().max().item(), 1e-2 * gscale)
        assert err < tol_grad, (n, err)
        if err > worst[1]:
            worst = (n, err)
    print("layer0 A/B (%s, precision %d): worst gradient deviation %s %.3g" % (mode, precision, worst[0], worst[1]))


def test_first_block_gradients_match_the_oracle_at_the_benchmarked_batch(dev):
    """24 clips (the bench shape: 15 024 x 16 groups of 16 pixels over 148 x k persistent warps), fp32-equivalent mode:
    conv0 / batchnorm0 / glu0 gradients against the CPU oracle's autograd."""
    cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=0.0)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    x = ofe.features(gen_wave(21, 24))
    net = build(cfg, P, dev, 1, specaugm_t_p=0.0, specaugm_f_p=0.0)
    net.train()
    s, w = net(x.to(dev))
    wgt = torch.linspace(0.5, 1.5, 156)
    ((s * wgt.to(dev)).mean() + w.mean()).backward()
    Pt = {k: (v.clone().requires_grad_(True) if ocrnn.is_float_param(k) else v.clone()) for k, v in P.items()}
    so, wo = ocrnn.crnn_forward(Pt, x, cfg, True)
    ((so * wgt).mean() + wo.mean()).backward()
    assert maxdiff(s.detach().cpu(), so.detach()) < 2e-5
    grads = {n: p.grad.cpu() for n, p in net.named_parameters()}
    gscale = max(Pt[n].grad.abs().max().item() for n in grads)
    for n in L0:
        ref = Pt[n].grad
        err = (grads[n] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-2 * gscale)
        assert err < 2e-3, (n, err)


@pytest.mark.parametrize("p,precision", [(0.5, 0), (0.5, 1), (0.3, 0)])
def test_dropout_statistics_of_the_fused_block(dev, p, precision):
    """p = 0.5 draws one fair bit per element (16 steps per Philox call), any other p a 16-bit draw per element: the pooled
    block output is the mean of four kept-or-zeroed elements scaled by 1 / (1 - p) - its expectation is the dropout-free
    output, and a pooled value is exactly zero with probability p^4."""
    x = ofe.features(gen_wave(13, 4)).to(dev)
    outs = {}
    for drop in (0.0, p):
        net = _net(dev, precision, dropout=drop)
        net.train()
        with torch.no_grad():
            net(x)
        outs[drop] = list(net._ws.values())[0].conv[0]["out"].clone()
    ref, got = outs[0.0].double(), outs[p].double()
    big = ref.abs() > 0.05
    assert abs((got[big] / ref[big]).mean().item() - 1.0) < 0.01
    zero = (got[big] == 0).double().mean().item()
    assert abs(zero - p ** 4) < 0.15 * p ** 4, (zero, p ** 4)
    # every kept combination is one of the 16 subset sums: with all four kept the value is ref / (1 - p)
    full = ((got[big] - ref[big] / (1 - p)).abs() < 1e-3 * ref[big].abs()).double().mean().item()
    assert abs(full - (1 - p) ** 4) < 0.15 * (1 - p) ** 4, (full, (1 - p) ** 4)
