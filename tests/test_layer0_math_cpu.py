"""The algebra behind the store-free first block (csrc/layer0.cu), checked on the CPU in fp64 against autograd.

With ONE input channel the first convolution's output z = conv0(x0) is never stored; the backward keeps g_y (gradient at the
BatchNorm output) in registers and only accumulates sums, from which l0_finish evaluates BatchNorm backward + conv0 weight
gradient in closed form:

    g_z = scale (g_y - S1/N - zhat S2/N),        S1 = sum g_y,  S2 = sum g_y zhat,  zhat = (z - mean) invstd
    dW[c][tap] = sum_pix g_z,c x0(pix + tap) = scale ( GX - (S1/N) SX - (S2/N) invstd (ZX - mean SX) )
    GX[c][tap] = sum g_y,c x0(pix + tap),  ZX[c][tap] = sum z_c x0(pix + tap),  SX[tap] = sum x0(pix + tap)

and SX[tap] itself follows from nine border sums of x0 (total, first / last row, first / last column, four corners).
Reference semantics: nn.Conv2d(1, 16, 3, 1, 1) -> nn.BatchNorm2d(16, eps=1e-3) in training mode (desed_task/nnet/CNN.py:66-76)."""
import torch
import torch.nn.functional as F


def _taps(x0):
    """x0 [B, T, F] -> [B, 9, T, F]: x0(pix + tap) with zero padding, tap = 3 dy + dx."""
    xp = F.pad(x0, (1, 1, 1, 1))
    T, Fm = x0.shape[1], x0.shape[2]
    return torch.stack([xp[:, dy:dy + T, dx:dx + Fm] for dy in range(3) for dx in range(3)], 1)


def test_closed_form_weight_gradient_equals_autograd():
    torch.manual_seed(0)
    B, T, Fm, C = 3, 14, 16, 16
    x0 = torch.randn(B, T, Fm, dtype=torch.float64)
    w = torch.randn(C, 1, 3, 3, dtype=torch.float64, requires_grad=True)
    b = torch.randn(C, dtype=torch.float64, requires_grad=True)
    gamma = torch.rand(C, dtype=torch.float64) + 0.5
    beta = torch.randn(C, dtype=torch.float64)
    gamma.requires_grad_(True)
    beta.requires_grad_(True)
    eps = 1e-3
    z = F.conv2d(x0[:, None], w, b, padding=1)                                   # [B, C, T, F]
    y = F.batch_norm(z, None, None, gamma, beta, True, 0.0, eps)
    g_y = torch.randn_like(y)                                                    # any upstream gradient
    (y * g_y).sum().backward()

    N = B * T * Fm
    zc = z.detach()
    mean = zc.mean((0, 2, 3))
    var = zc.var((0, 2, 3), unbiased=False)
    invstd = 1.0 / torch.sqrt(var + eps)
    scale = gamma.detach() * invstd
    zhat = (zc - mean[None, :, None, None]) * invstd[None, :, None, None]
    S1 = g_y.sum((0, 2, 3))
    S2 = (g_y * zhat).sum((0, 2, 3))
    xt = _taps(x0)                                                               # [B, 9, T, F]
    GX = torch.einsum("bctf,bktf->ck", g_y, xt)
    ZX = torch.einsum("bctf,bktf->ck", zc, xt)
    SX = xt.sum((0, 2, 3))
    dW = scale[:, None] * (GX - (S1 / N)[:, None] * SX[None] - (S2 / N * invstd)[:, None] * (ZX - mean[:, None] * SX[None]))
    assert torch.allclose(dW, w.grad.reshape(C, 9), rtol=1e-10, atol=1e-10)
    assert torch.allclose(S1, beta.grad, rtol=1e-12, atol=1e-12)                 # gbeta
    assert torch.allclose(S2, gamma.grad, rtol=1e-10, atol=1e-10)                # ggamma
    assert b.grad.abs().max().item() < 1e-9                                      # the conv bias cancels inside batch-stat BN

    # frozen BatchNorm (freeze_bn / eval-mode autograd): a fixed affine map
    w2 = w.detach().clone().requires_grad_(True)
    b2 = b.detach().clone().requires_grad_(True)
    rm, rv = torch.randn(C, dtype=torch.float64), torch.rand(C, dtype=torch.float64) + 0.5
    z2 = F.conv2d(x0[:, None], w2, b2, padding=1)
    y2 = F.batch_norm(z2, rm, rv, gamma.detach(), beta.detach(), False, 0.0, eps)
    (y2 * g_y).sum().backward()
    sc = gamma.detach() / torch.sqrt(rv + eps)
    assert torch.allclose(sc[:, None] * GX, w2.grad.reshape(C, 9), rtol=1e-10, atol=1e-10)
    assert torch.allclose(sc * S1, b2.grad, rtol=1e-10, atol=1e-10)


def test_tap_sums_follow_from_nine_border_sums():
    """l0_x0 accumulates total / first row / last row / first column / last column / four corners; l0_finish rebuilds SX."""
    torch.manual_seed(1)
    x0 = torch.randn(4, 11, 24, dtype=torch.float64)
    SX = _taps(x0).sum((0, 2, 3))
    tot = x0.sum()
    r0, rl = x0[:, 0].sum(), x0[:, -1].sum()
    c0, cl = x0[:, :, 0].sum(), x0[:, :, -1].sum()
    k00, k0l, kl0, kll = x0[:, 0, 0].sum(), x0[:, 0, -1].sum(), x0[:, -1, 0].sum(), x0[:, -1, -1].sum()
    for tap in range(9):
        dy, dx = divmod(tap, 3)
        s = tot.clone()
        if dy == 0:
            s -= rl            # the window shifted up never sees the last row
        if dy == 2:
            s -= r0
        if dx == 0:
            s -= cl
        if dx == 2:
            s -= c0
        if dy == 0 and dx == 0:
            s += kll
        if dy == 0 and dx == 2:
            s += kl0
        if dy == 2 and dx == 0:
            s += k0l
        if dy == 2 and dx == 2:
            s += k00
        assert abs(s.item() - SX[tap].item()) < 1e-9, tap


def test_plan_constant_matches_the_host_mirror():
    """SEDK_L0_SUMS (include/sedk.h) is the size the workspace reserves behind the BatchNorm statistics."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "sedk.h")).read()
    n = int(re.search(r"#define SEDK_L0_SUMS (\d+)", hdr).group(1))
    src = open(os.path.join(root, "desed_task_b200", "nnet", "CRNN.py")).read()
    assert "n_stats + %d" % n in src
    assert n >= 2 * 144 + 9
