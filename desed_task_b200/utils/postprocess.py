"""Median-filter post-processing on the GPU (csrc/elementwise.cu median_kernel).

`ClassWiseMedianFilter` mirrors desed_task/utils/postprocess.py:5-17 (numpy [T, C] in, numpy out) but filters whole
batches on the device; `median_filter` is the batched tensor API used by the inference path
(replaces scipy.ndimage.median_filter(c_scores, (k, 1)) at recipes/dcase2023_task4_baseline/local/utils.py:58)."""
import numpy as np
import torch

from .._lib import check, lib, ptr, require_cuda, stream_ptr


def median_filter(strong, win, class_dim=1):
    """strong: cuda fp32 [B, C, T] (class_dim=1, the CRNN output layout) or [B, T, C] (class_dim=2).
    win: int or per-class sequence of window lengths (1..31).  Returns a tensor of the same shape/layout."""
    require_cuda(strong)
    x = strong.float()
    assert x.dim() == 3 and class_dim in (1, 2)
    Cn = x.shape[class_dim]
    Tn = x.shape[3 - class_dim]
    if isinstance(win, torch.Tensor):
        # a ready-made int32 device tensor of per-class windows (validated by its maker; lets the call sit in a CUDA graph)
        if win.dtype != torch.int32 or win.numel() != Cn or win.device != x.device:
            raise ValueError("median_filter: window tensor must be int32 [%d] on %s" % (Cn, x.device))
        w = win
    else:
        if isinstance(win, int):
            win = [win] * Cn
        win = [int(v) for v in win]
        if len(win) != Cn or min(win) < 1 or max(win) > 31:
            raise ValueError("median_filter: need one window in [1, 31] per class (got %s)" % (win,))
        w = torch.tensor(win, dtype=torch.int32, device=x.device)
    out = torch.empty_like(x)
    sc, st = x.stride(class_dim), x.stride(3 - class_dim)
    oc, ot = out.stride(class_dim), out.stride(3 - class_dim)
    check(lib().sedk_median_filter(ptr(x), ptr(out), x.shape[0], Cn, Tn, x.stride(0), sc, st, out.stride(0), oc, ot,
                                   ptr(w), stream_ptr()), "sedk_median_filter")
    return out


class ClassWiseMedianFilter:
    def __init__(self, filter_lens=(1, 1, 1)):
        self.filter_lens = filter_lens

    def __call__(self, x, **kwargs):
        """x: [T, C] numpy array or tensor (one clip, as in the reference) or [B, T, C]."""
        is_np = isinstance(x, np.ndarray)
        t = torch.as_tensor(x, dtype=torch.float32)
        if not t.is_cuda:
            t = t.cuda()
        squeeze = t.dim() == 2
        if squeeze:
            t = t[None]
        out = median_filter(t, list(self.filter_lens)[: t.shape[-1]], class_dim=2)
        if squeeze:
            out = out[0]
        return out.cpu().numpy() if is_np else out
