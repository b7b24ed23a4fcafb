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


def decode_events(scores, thresholds, n_frames=None, class_dim=1, capacity=None):
    """Threshold + run-length event decoding on the device (sedk_decode_events).

    scores: cuda fp32 [B, C, T] (class_dim=1) or [B, T, C] (class_dim=2), normally median-filtered.  thresholds: sequence of
    floats.  n_frames: optional int tensor/sequence [B] of true clip lengths (frames).  Returns (offsets, events) as numpy
    int32 arrays: rows ordered (threshold, clip, class); events[offsets[r]:offsets[r + 1]] = [[onset_frame, offset_frame], ...]
    of row r = (th * B + b) * C + c, with offset_frame exclusive (dcase_util find_contiguous_regions convention)."""
    require_cuda(scores)
    x = scores.float()
    assert x.dim() == 3 and class_dim in (1, 2)
    B, Cn, Tn = x.shape[0], x.shape[class_dim], x.shape[3 - class_dim]
    thr = torch.tensor([float(t) for t in thresholds], dtype=torch.float32, device=x.device)
    nth = thr.numel()
    rows = nth * B * Cn
    nf = None
    if n_frames is not None:
        nf = torch.as_tensor(n_frames, dtype=torch.int32).to(x.device).contiguous()
    if capacity is None:
        capacity = max(1024, rows * 4)
    while True:
        offsets = torch.empty(rows + 1, dtype=torch.int32, device=x.device)
        events = torch.empty(max(capacity, 1), 2, dtype=torch.int32, device=x.device)
        check(lib().sedk_decode_events(ptr(x), B, Cn, Tn, x.stride(0), x.stride(class_dim), x.stride(3 - class_dim), ptr(thr),
                                       nth, ptr(nf), ptr(offsets), ptr(events), int(capacity), stream_ptr()),
              "sedk_decode_events")
        off = offsets.cpu().numpy()                  # one D2H for the whole batch (the reference does one per clip)
        total = int(off[-1])
        if total <= capacity:
            return off, events[:total].cpu().numpy()
        capacity = total                             # rare: more events than the first guess; decode again


def frame_to_time(frames, net_pooling=4, fs=16000, frame_hop=256, audio_len=10.0):
    """ManyHotEncoder._frame_to_time (desed_task/utils/encoder.py:76-78), float64 like the reference."""
    t = np.asarray(frames, dtype=np.float64) * net_pooling / (fs / frame_hop)
    return np.clip(t, a_min=0, a_max=audio_len)


def _score_dataframe(scores, timestamps, event_classes):
    """sed_scores_eval.base_modules.scores.create_score_dataframe when that package is installed; otherwise the same table
    layout (columns onset, offset, then one column per class)."""
    try:
        from sed_scores_eval.base_modules.scores import create_score_dataframe
        return create_score_dataframe(scores=scores, timestamps=timestamps, event_classes=event_classes)
    except ImportError:
        import pandas as pd
        return pd.DataFrame(np.concatenate((timestamps[:-1, None], timestamps[1:, None], scores), axis=1),
                            columns=["onset", "offset", *event_classes])


def batched_decode_preds(strong_preds, filenames, encoder, thresholds=[0.5], median_filter=7, pad_indx=None):  # noqa: B006
    """Drop-in for recipes/dcase2023_task4_baseline/local/utils.py:16-73 (same arguments, same three return values):
    the median filter, the thresholding over all thresholds and the event extraction run batched on the device; the host
    only turns frame indices into seconds (encoder._frame_to_time) and builds the pandas tables.
    `median_filter`: int (2023) or a per-class sequence / ClassWiseMedianFilter.filter_lens (2024)."""
    import pandas as pd
    from pathlib import Path
    require_cuda(strong_preds)
    B, Cn, Tn = strong_preds.shape
    if pad_indx is not None:
        # the reference slices `strong_preds[j][:true_len]` BEFORE transposing, i.e. along the class axis (utils.py:48-50):
        # a no-op whenever true_len >= n_classes, which holds for every padded clip the recipes can produce
        for j in range(B):
            if int(Tn * float(pad_indx[j])) < Cn:
                raise NotImplementedError("pad_indx shorter than the number of classes (reference slices the class axis)")
    wins = getattr(median_filter, "filter_lens", median_filter)
    filt = globals()["median_filter"](strong_preds, wins if isinstance(wins, int) else list(wins)[:Cn], class_dim=1)
    raw = strong_preds.float().transpose(1, 2).cpu().numpy()          # [B, T, C]
    post = filt.transpose(1, 2).cpu().numpy()
    off, ev = decode_events(filt, thresholds, class_dim=1)
    ts = encoder._frame_to_time(np.arange(Tn + 1))
    scores_raw, scores_post, prediction_dfs = {}, {}, {}
    for j in range(B):
        audio_id = Path(filenames[j]).stem
        scores_raw[audio_id] = _score_dataframe(raw[j], ts, encoder.labels)
        scores_post[audio_id] = _score_dataframe(post[j], ts, encoder.labels)
    for ti, th in enumerate(thresholds):
        rows = []
        for j in range(B):
            fname = Path(filenames[j]).stem + ".wav"
            for c in range(Cn):
                r = (ti * B + j) * Cn + c
                for on, of in ev[off[r]:off[r + 1]]:
                    rows.append([encoder.labels[c], encoder._frame_to_time(on), encoder._frame_to_time(of), fname])
        prediction_dfs[th] = pd.DataFrame(rows, columns=["event_label", "onset", "offset", "filename"])
    return scores_raw, scores_post, prediction_dfs
