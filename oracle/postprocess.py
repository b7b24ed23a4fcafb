"""Oracle: median-filter post-processing (CPU, numpy).  Test infrastructure only.

Restates scipy.ndimage.median_filter(scores[T, C], (k, 1)) with the default mode='reflect'
(d c b a | a b c d | d c b a: the edge sample is repeated), as called at
recipes/dcase2023_task4_baseline/local/utils.py:58 (window 7) and per class in
desed_task/utils/postprocess.py:5-17 (2024).  For even k scipy's window origin is k//2
(window covers [i - k//2, i + k - 1 - k//2]) and the median is the element of rank k//2.
"""
import numpy as np


def _reflect(i, n):
    # scipy 'reflect' (half-sample symmetric) extension, valid for any overshoot
    period = 2 * n
    i = np.mod(i, period)
    return np.where(i >= n, period - 1 - i, i)


def median_filter_time(scores, k):
    """scores [T, C] -> [T, C]; window k along time, reflect boundary."""
    scores = np.asarray(scores)
    T = scores.shape[0]
    if k <= 1:
        return scores.copy()
    offs = np.arange(k) - k // 2
    idx = _reflect(np.arange(T)[:, None] + offs[None, :], T)          # [T, k]
    win = scores[idx]                                                  # [T, k, C]
    return np.sort(win, axis=1)[:, k // 2]


def classwise_median_filter(scores, filter_lens):
    """desed_task/utils/postprocess.py:9-17: class c uses window filter_lens[c]. scores [T, C]."""
    out = [median_filter_time(scores[:, c:c + 1], int(filter_lens[c]))[:, 0] for c in range(scores.shape[-1])]
    return np.stack(out, -1)


def find_contiguous_regions(activity):
    """dcase_util.data.DecisionEncoder.find_contiguous_regions (dcase_util 0.2.x, the reference's un-vendored dependency,
    desed_task/utils/encoder.py:5,200) restated from its published algorithm: indices where the boolean activity changes,
    +1, with 0 prepended if it starts active and len appended if it ends active, reshaped to [n, 2] = [onset, offset)."""
    a = np.asarray(activity).astype(bool)
    change = np.logical_xor(a[1:], a[:-1]).nonzero()[0] + 1
    if a.size and a[0]:
        change = np.r_[0, change]
    if a.size and a[-1]:
        change = np.r_[change, a.size]
    return change.reshape((-1, 2))


def frame_to_time(frame, net_pooling=4, fs=16000, frame_hop=256, audio_len=10.0):
    """ManyHotEncoder._frame_to_time, desed_task/utils/encoder.py:76-78."""
    return np.clip(np.asarray(frame, dtype=np.float64) * net_pooling / (fs / frame_hop), a_min=0, a_max=audio_len)


def decode_strong(pred, labels, **enc):
    """ManyHotEncoder.decode_strong, desed_task/utils/encoder.py:189-211.  pred: bool [T, C]."""
    out = []
    for i, col in enumerate(np.asarray(pred).T):
        for on, of in find_contiguous_regions(col):
            out.append([labels[i], float(frame_to_time(on, **enc)), float(frame_to_time(of, **enc))])
    return out


def batched_decode(strong_preds, labels, thresholds=(0.5,), median_filter=7, **enc):
    """The numeric part of batched_decode_preds (recipes/dcase2023_task4_baseline/local/utils.py:45-71) without pandas:
    strong_preds [B, C, T] -> (filtered scores [B, T, C], {threshold: [(clip, label, onset_s, offset_s), ...]}) in the
    reference's append order (clip, then class, then time)."""
    strong_preds = np.asarray(strong_preds)
    post, preds = [], {th: [] for th in thresholds}
    for j in range(strong_preds.shape[0]):
        c_scores = strong_preds[j].T
        c_scores = median_filter_time(c_scores, median_filter) if np.isscalar(median_filter) else \
            classwise_median_filter(c_scores, median_filter)
        post.append(c_scores)
        for th in thresholds:
            for lab, on, of in decode_strong(c_scores > th, labels, **enc):
                preds[th].append((j, lab, on, of))
    return np.stack(post), preds


def time_to_frame(time, fs=16000, frame_hop=256, net_pooling=4, n_frames=156):
    """ManyHotEncoder._time_to_frame, desed_task/utils/encoder.py:71-74."""
    return np.clip(np.asarray(time, dtype=np.float64) * fs / frame_hop / net_pooling, a_min=0, a_max=n_frames)


def encode_strong(events, labels, n_frames=156, **enc):
    """ManyHotEncoder.encode_strong_df for the list-of-lists input ([label, onset, offset(, confidence)], encoder.py:139-159):
    y[int(t2f(onset)) : int(ceil(t2f(offset))), class] = confidence or 1, events applied in order.  Returns [n_frames, C]."""
    y = np.zeros((n_frames, len(labels)))
    for e in events:
        if e[0] == "":
            continue
        i = labels.index(e[0])
        onset = int(time_to_frame(e[1], n_frames=n_frames, **enc))
        offset = int(np.ceil(time_to_frame(e[2], n_frames=n_frames, **enc)))
        y[onset:offset, i] = e[3] if len(e) == 4 else 1
    return y
