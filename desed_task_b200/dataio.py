"""Input-pipeline pieces of the hot path on the device: batched strong-label encoding.

The reference encodes labels per item in its DataLoader workers with a pandas row loop
(ManyHotEncoder.encode_strong_df, desed_task/utils/encoder.py:80-171, called from desed_task/dataio/datasets.py:187-237)
and ships dense [T', C] arrays; here the events of a whole batch are turned into frame spans with the reference's own
float64 formulas (vectorised numpy, a few hundred events per batch) and rasterised by one kernel (sedk_encode_strong) into
the [B, C, T'] tensor the trainers consume.  16-bit PCM audio input is the other half of f2: frontend.MelSpectrogram.run and
the engines accept int16 waveforms directly (sedk_logmel_fwd_i16).
"""
import numpy as np
import torch

from ._lib import check, lib, ptr, stream_ptr


def time_to_frame(time, fs=16000, frame_hop=256, net_pooling=4, n_frames=156):
    """ManyHotEncoder._time_to_frame (encoder.py:71-74), float64."""
    frame = np.asarray(time, dtype=np.float64) * fs / frame_hop
    return np.clip(frame / net_pooling, a_min=0, a_max=n_frames)


def encode_strong_batch(events_per_clip, labels, device, n_frames=156, fs=16000, frame_hop=256, net_pooling=4, out=None):
    """events_per_clip: one list per clip of [label, onset_s, offset_s] or [label, onset_s, offset_s, confidence] (the
    list-of-lists form of encode_strong_df; `label` a class name from `labels`, "" entries are skipped like upstream).
    Returns a cuda fp32 tensor [B, C, n_frames] (the layout after the datasets' transpose, datasets.py:66-71)."""
    B, C = len(events_per_clip), len(labels)
    index = {l: i for i, l in enumerate(labels)}
    clip, cls, on, off, val, offs = [], [], [], [], [], [0]
    for b, evs in enumerate(events_per_clip):
        for e in evs:
            if e[0] == "":
                continue
            clip.append(b)
            cls.append(index[e[0]])
            on.append(e[1])
            off.append(e[2])
            val.append(e[3] if len(e) == 4 else 1.0)
        offs.append(len(clip))
    n = len(clip)
    kw = dict(fs=fs, frame_hop=frame_hop, net_pooling=net_pooling, n_frames=n_frames)
    on_f = time_to_frame(np.asarray(on, np.float64), **kw).astype(np.int64)                    # int(...) truncates
    off_f = np.ceil(time_to_frame(np.asarray(off, np.float64), **kw)).astype(np.int64)         # int(np.ceil(...))
    ev = np.stack([np.asarray(clip, np.int64), np.asarray(cls, np.int64), on_f, off_f], 1).astype(np.int32) if n else \
        np.zeros((1, 4), np.int32)
    ev_d = torch.from_numpy(ev).to(device)
    val_d = torch.tensor(val if n else [0.0], dtype=torch.float32, device=device)
    off_d = torch.tensor(offs, dtype=torch.int32, device=device)
    if out is None:
        out = torch.empty(B, C, n_frames, dtype=torch.float32, device=device)
    check(lib().sedk_encode_strong(ptr(ev_d), ptr(val_d), ptr(off_d), ptr(out), B, C, n_frames, stream_ptr()),
          "sedk_encode_strong")
    return out
