"""mixup / frame_shift / add_noise on libsedk kernels.

Mirror of desed_task/data_augm.py:7-77: same signatures, same RNG consumption order (random.gauss per clip;
np.random.beta then torch.randperm on the CPU generator; torch.rand / torch.randn on the data's device), same return
conventions and the same NotImplementedError for an unknown mixup_label_type.
"""
import random

import numpy as np
import torch

from ._lib import check, lib, ptr, require_cuda, stream_ptr


def _flat(x):
    x = x.float().contiguous()
    B = x.shape[0]
    return x, B, x.numel() // B


def mix_tensors(data, perm_dev, coef_dev):
    x, B, n = _flat(data)
    out = torch.empty_like(x)
    check(lib().sedk_feat_mix_log(ptr(x), ptr(perm_dev), ptr(coef_dev), ptr(out), B, n, 0, 0.0, 0.0, 0.0, None,
                                  stream_ptr()), "sedk_feat_mix_log")
    return out


def mix_labels(target, perm_dev, coef_dev, hard):
    y, B, n = _flat(target)
    out = torch.empty_like(y)
    check(lib().sedk_label_mix(ptr(y), ptr(perm_dev), ptr(coef_dev), ptr(out), B, n, 1 if hard else 0, stream_ptr()),
          "sedk_label_mix")
    return out


def mixup(data, target=None, alpha=0.2, beta=0.2, mixup_label_type="soft"):
    """data_augm.py:19-53."""
    with torch.no_grad():
        require_cuda(data, target)
        batch_size = data.size(0)
        c = np.random.beta(alpha, beta)
        perm = torch.randperm(batch_size)
        perm_dev = perm.to(data.device)
        coef_dev = torch.full((batch_size,), float(c), dtype=torch.float32, device=data.device)
        mixed_data = mix_tensors(data, perm_dev, coef_dev)
        if target is not None:
            if mixup_label_type == "soft":
                mixed_target = mix_labels(target, perm_dev, coef_dev, hard=False)
            elif mixup_label_type == "hard":
                mixed_target = mix_labels(target, perm_dev, coef_dev, hard=True)
            else:
                raise NotImplementedError(
                    f"mixup_label_type: {mixup_label_type} not implemented. choice in "
                    f"{'soft', 'hard'}"
                )
            return mixed_data, mixed_target
        else:
            return mixed_data


def _roll(x, shifts_dev):
    x = x.float().contiguous()
    B = x.shape[0]
    cols = x.shape[-1]
    rows = x.numel() // (B * cols)
    out = torch.empty_like(x)
    check(lib().sedk_roll_last(ptr(x), ptr(out), ptr(shifts_dev), B, rows, cols, stream_ptr()), "sedk_roll_last")
    return out


def frame_shift(mels, labels, net_pooling=4):
    """data_augm.py:7-16."""
    require_cuda(mels, labels)
    bsz, n_bands, frames = mels.shape
    shifts, lshifts = [], []
    for bindx in range(bsz):
        shift = int(random.gauss(0, 90))
        shifts.append(shift)
        lshifts.append(-abs(shift) // net_pooling if shift < 0 else shift // net_pooling)
    sd = torch.tensor(shifts, dtype=torch.int32, device=mels.device)
    ld = torch.tensor(lshifts, dtype=torch.int32, device=mels.device)
    return _roll(mels, sd), _roll(labels, ld)


def add_noise(mels, snrs=(6, 30), dims=(1, 2)):
    """data_augm.py:56-77 (dims must cover every non-batch axis, as in the reference default)."""
    require_cuda(mels)
    if tuple(sorted(d % mels.dim() for d in dims)) != tuple(range(1, mels.dim())):
        raise NotImplementedError("add_noise: dims must cover all non-batch axes (got %s)" % (dims,))
    B = mels.shape[0]
    if isinstance(snrs, (list, tuple)):
        snr = (snrs[0] - snrs[1]) * torch.rand((B,), device=mels.device) + snrs[1]
    else:
        snr = torch.full((B,), float(snrs), device=mels.device)
    x, B, n = _flat(mels)
    stats = torch.empty(B, 2, device=x.device, dtype=torch.float32)
    check(lib().sedk_instance_stats(ptr(x), ptr(stats), B, n, stream_ptr()), "sedk_instance_stats")
    noise = torch.randn(mels.shape, device=mels.device)
    out = torch.empty_like(x)
    check(lib().sedk_add_noise(ptr(x), ptr(noise), ptr(snr.float().contiguous()), ptr(stats), ptr(out), B, n,
                               stream_ptr()), "sedk_add_noise")
    return out
