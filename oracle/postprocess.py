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
