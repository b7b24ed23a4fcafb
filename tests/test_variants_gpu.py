"""GPU A/B parity between kernel generations (sedk_set_option switches) and the dropout-mask consistency of the fused
BN+GLU+dropout+pool kernels.  The default variants are also covered against the CPU oracle in test_crnn_gpu.py; these
tests pin the first-generation kernels (still used for other shapes: H = 64/192, large batches, C >= 64) to the same
results, so both code paths stay parity-green."""
import dataclasses

import pytest
import torch

from oracle import crnn as ocrnn, frontend as ofe
from tests.test_crnn_gpu import build
from tests.util import gen_wave, maxdiff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def feats():
    return ofe.features(gen_wave(0, 3))


def _run(dev, feats, precision, dropout=0.0, fwd_count=None):
    cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=dropout)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    net = build(cfg, P, dev, precision, specaugm_t_p=0.0, specaugm_f_p=0.0)
    net.train()
    if fwd_count is not None:
        net._fwd_count = fwd_count
    s, w = net(feats.to(dev))
    ((s * torch.linspace(0.5, 1.5, s.shape[-1], device=dev)).mean() + w.mean()).backward()
    return s.detach().clone(), w.detach().clone(), {n: p.grad.clone() for n, p in net.named_parameters()}


def _compare(a, b, tol_out, tol_grad):
    assert maxdiff(a[0], b[0]) < tol_out and maxdiff(a[1], b[1]) < tol_out
    gscale = max(g.abs().max().item() for g in b[2].values())
    worst = ("", 0.0)
    for n, g in b[2].items():
        err = (a[2][n] - g).abs().max().item() / max(g.abs().max().item(), 1e-2 * gscale)
        if err > worst[1]:
            worst = (n, err)
    assert worst[1] < tol_grad, worst


@pytest.mark.parametrize("option", ["gru_v2", "bnglu_small", "bnglu_tc5", "gemm_tc5", "side_stream", "conv_pair"])
@pytest.mark.parametrize("precision,tol_out,tol_grad", [(1, 2e-5, 2e-3), (0, 5e-4, 3e-2)])
def test_kernel_generations_agree(dev, feats, option, precision, tol_out, tol_grad):
    from desed_task_b200._lib import lib
    default = 0 if option == "conv_pair" else 1         # library defaults (include/sedk.h)
    res = {}
    for on in (1, 0):
        lib().sedk_set_option(option.encode(), on)
        try:
            assert lib().sedk_get_option(option.encode(), -1) == on
            res[on] = _run(dev, feats, precision)
        finally:
            lib().sedk_set_option(option.encode(), default)
    _compare(res[1], res[0], tol_out, tol_grad)


@pytest.mark.parametrize("variant", [1, 2])
@pytest.mark.parametrize("precision,tol_out,tol_grad", [(1, 2e-5, 2e-3), (0, 5e-4, 3e-2)])
def test_gru_third_generation_agrees_with_second(dev, feats, variant, precision, tol_out, tol_grad):
    """csrc/gru3.cu (octet layout, weights in registers; variant 1 = 8 warps, 2 = 16 warps) against the v2 recurrence."""
    from desed_task_b200._lib import lib
    res = {}
    for on in (variant, 0):
        lib().sedk_set_option(b"gru_v3", on)
        try:
            res[on] = _run(dev, feats, precision)
        finally:
            lib().sedk_set_option(b"gru_v3", 1)
    _compare(res[variant], res[0], tol_out, tol_grad)


def test_gru_two_rows_per_cta_path(dev):
    """Batches above 74 rows put two rows on a CTA: that path still runs the first-generation recurrence; it must agree
    with the single-row (v2) result on the same clips."""
    cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=0.0)
    P = ocrnn.init_params(cfg, seed=42, trained_like=True)
    x = ofe.features(gen_wave(3, 2)).to(dev)
    net = build(cfg, P, dev, 1)
    net.eval()
    with torch.no_grad():
        s_small, w_small = net(x)
        s_big, w_big = net(x.repeat(40, 1, 1))
    assert maxdiff(s_big[:2], s_small) < 2e-5 and maxdiff(w_big[78:80], w_small) < 2e-5


@pytest.mark.parametrize("small,precision", [(1, 1), (0, 1), (1, 0)])
def test_dropout_masks_agree_between_forward_and_backward(dev, feats, small, precision):
    """Dropout masks are regenerated in backward from (seed, stream, counter): with the seed pinned, a central finite
    difference of the loss along the gradient direction must reproduce |grad| (a forward/backward mask mismatch in any
    layer would break this by O(1))."""
    from desed_task_b200._lib import lib
    lib().sedk_set_option(b"bnglu_small", small)
    try:
        cfg = dataclasses.replace(ocrnn.CFG_2023, dropout=0.5)
        P = ocrnn.init_params(cfg, seed=42, trained_like=True)
        # precision 0 also exercises the tcgen05 BN+GLU kernels of the 128-channel layers (their own mask mapping)
        net = build(cfg, P, dev, precision, specaugm_t_p=0.0, specaugm_f_p=0.0)
        net.train()
        x = feats.to(dev)
        wgt = torch.linspace(0.5, 1.5, 156, device=dev)

        def loss_at(count):
            net._fwd_count = count
            s, w = net(x)
            return (s * wgt).mean() + w.mean()

        loss = loss_at(100)
        loss.backward()
        names = [n for n, _ in net.named_parameters() if n.startswith("cnn.")]
        params = dict(net.named_parameters())
        grads = {n: params[n].grad.clone() for n in names}
        gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values())).item()
        assert gnorm > 0
        eps = 2e-3 / (gnorm * gnorm)            # loss moves by +-2e-3 along the gradient direction
        vals = []
        with torch.no_grad():
            for sign in (+1.0, -1.0):
                for n in names:
                    params[n].add_(grads[n], alpha=sign * eps)
                vals.append(loss_at(100).item())
                for n in names:
                    params[n].add_(grads[n], alpha=-sign * eps)
        fd = (vals[0] - vals[1]) / (2 * eps)
        assert abs(fd - gnorm * gnorm) / (gnorm * gnorm) < (0.08 if precision else 0.15), (fd, gnorm * gnorm)
        # and a different seed gives different masks
        with torch.no_grad():
            assert abs(loss_at(101).item() - loss.item()) > 1e-6
    finally:
        lib().sedk_set_option(b"bnglu_small", 1)
