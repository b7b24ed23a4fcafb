"""Log-mel front end on the fused sm_100a kernel (csrc/logmel.cu).

Host-side mirror of what the reference builds at recipes/dcase2023_task4_baseline/local/sed_trainer.py:79-91
(`torchaudio.transforms.MelSpectrogram(...)`) and :253-264 (`take_log`): same constructor arguments, same module tree
(`spectrogram.window`, `mel_scale.fb` buffers -> same state_dict keys), same output layout [..., n_mels, frames].
The arithmetic itself lives in libsedk (C ABI, include/sedk.h); this file only prepares constant tables.
"""
import math

import numpy as np
import torch
from torch import nn

from . import _lib
from ._lib import MelTables, check, lib, ptr, require_cuda, stream_ptr

N_FFT = 2048


def melscale_fbanks(n_freqs, f_min, f_max, n_mels, sample_rate):
    """HTK triangular filterbank [n_freqs, n_mels] fp32; same op order as torchaudio functional.py:563-573,507-513
    (norm=None) so the values are bit-identical to the reference's `mel_scale.fb` buffer."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + (f_min / 700.0))
    m_max = 2595.0 * math.log10(1.0 + (f_max / 700.0))
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.max(torch.zeros(1), torch.min(down, up))


def _twiddles():
    k = np.arange(1024, dtype=np.float64)
    ang = -2.0 * np.pi * k / 2048.0
    tw2048 = np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32)          # exp(-2 pi i k / 2048)
    k1 = np.arange(32, dtype=np.float64)[:, None]
    n2 = np.arange(32, dtype=np.float64)[None, :]
    ang = -2.0 * np.pi * (k1 * n2) / 1024.0
    tw32 = np.stack([np.cos(ang), np.sin(ang)], -1).astype(np.float32)            # [k1][n2]
    return torch.from_numpy(tw2048.reshape(-1)), torch.from_numpy(tw32.reshape(-1))


def sparse_filterbank(fb):
    """Column m of fb [n_freqs, n_mels] -> (start, len, offset, weights) of its first..last non-zero run."""
    fbn = fb.detach().cpu().numpy()
    starts, lens, offs, ws = [], [], [], []
    off = 0
    for m in range(fbn.shape[1]):
        nz = np.nonzero(fbn[:, m])[0]
        if len(nz) == 0:
            s, n = 0, 0
        else:
            s, n = int(nz[0]), int(nz[-1] - nz[0] + 1)
        starts.append(s)
        lens.append(n)
        offs.append(off)
        ws.append(fbn[s:s + n, m])
        off += n
    w = np.concatenate(ws).astype(np.float32) if off > 0 else np.zeros(1, np.float32)
    return (torch.tensor(starts, dtype=torch.int32), torch.tensor(lens, dtype=torch.int32),
            torch.tensor(offs, dtype=torch.int32), torch.from_numpy(w))


class _DeviceTables:
    """Constant tables of the kernel on one device (kept alive by the owning module)."""

    def __init__(self, window, fb, hop, device):
        tw2048, tw32 = _twiddles()
        st, ln, of, w = sparse_filterbank(fb)
        self.tensors = [t.to(device).contiguous() for t in
                        (window.detach().float(), tw2048, tw32, st, ln, of, w)]
        t = self.tensors
        self.struct = MelTables(ptr(t[0]), ptr(t[1]), ptr(t[2]), ptr(t[3]), ptr(t[4]), ptr(t[5]), ptr(t[6]),
                                int(fb.shape[1]), int(hop))


class Spectrogram(nn.Module):
    """Holder of the `window` buffer (state_dict key `spectrogram.window`, as torchaudio's Spectrogram)."""

    def __init__(self, n_fft, win_length, hop_length, window_fn, wkwargs):
        super().__init__()
        self.n_fft, self.win_length, self.hop_length = n_fft, win_length, hop_length
        window = window_fn(win_length) if wkwargs is None else window_fn(win_length, **wkwargs)
        self.register_buffer("window", window.float())


class MelScale(nn.Module):
    """Holder of the `fb` buffer (state_dict key `mel_scale.fb`, as torchaudio's MelScale)."""

    def __init__(self, n_mels, sample_rate, f_min, f_max, n_stft):
        super().__init__()
        self.register_buffer("fb", melscale_fbanks(n_stft, f_min, f_max, n_mels, sample_rate))


class MelSpectrogram(nn.Module):
    """Drop-in for torchaudio.transforms.MelSpectrogram on the configuration the recipes use.

    Supported: n_fft == win_length == 2048, power == 1, center=True, pad_mode='reflect', onesided, HTK scale,
    norm=None (sed_trainer.py:79-91).  Anything else raises NotImplementedError - there is no fallback path.
    """

    def __init__(self, sample_rate=16000, n_fft=400, win_length=None, hop_length=None, f_min=0.0, f_max=None,
                 pad=0, n_mels=128, window_fn=torch.hann_window, power=2.0, normalized=False, wkwargs=None,
                 center=True, pad_mode="reflect", onesided=None, norm=None, mel_scale="htk"):
        super().__init__()
        win_length = win_length if win_length is not None else n_fft
        hop_length = hop_length if hop_length is not None else win_length // 2
        f_max = float(f_max) if f_max is not None else float(sample_rate // 2)
        unsupported = []
        if n_fft != N_FFT or win_length != N_FFT:
            unsupported.append("n_fft/win_length must be 2048 (got %s/%s)" % (n_fft, win_length))
        if power != 1:
            unsupported.append("power must be 1 (got %s)" % (power,))
        if pad != 0 or normalized or not center or pad_mode != "reflect" or onesided is False:
            unsupported.append("only pad=0, normalized=False, center=True, pad_mode='reflect', onesided")
        if norm is not None or mel_scale != "htk":
            unsupported.append("only norm=None, mel_scale='htk'")
        if unsupported:
            raise NotImplementedError("desed_task_b200.MelSpectrogram: " + "; ".join(unsupported))
        self.sample_rate, self.n_fft, self.win_length, self.hop_length = sample_rate, n_fft, win_length, hop_length
        self.n_mels, self.f_min, self.f_max, self.power = n_mels, f_min, f_max, power
        self.spectrogram = Spectrogram(n_fft, win_length, hop_length, window_fn, wkwargs)
        self.mel_scale = MelScale(n_mels, sample_rate, f_min, f_max, n_fft // 2 + 1)
        self._tables = {}

    def __getstate__(self):
        # the table cache holds ctypes structures with device pointers: rebuilt on first use after a copy / unpickle
        state = dict(self.__dict__)
        state["_tables"] = {}
        return state

    def tables(self, device):
        key = (str(device), self.spectrogram.window._version, self.mel_scale.fb._version,
               self.spectrogram.window.data_ptr(), self.mel_scale.fb.data_ptr())
        tab = self._tables.get("tab")
        if tab is None or self._tables.get("key") != key:
            tab = _DeviceTables(self.spectrogram.window, self.mel_scale.fb, self.hop_length, device)
            self._tables = {"tab": tab, "key": key}
        return tab

    def n_frames(self, n_samples):
        return 1 + n_samples // self.hop_length

    def run(self, waveform, log=False, amin=1e-5, db_range=(-50.0, 80.0), minmax=None, time_major=False):
        """waveform [..., L] cuda fp32 - or int16 PCM, used as x / 32768 like torchaudio.load's normalisation (bit-identical
        result, half the bytes) - -> mel [..., n_mels, T] (a transposed view of a [.., T, n_mels] buffer when time_major).
        log=True fuses take_log; minmax (uint32 [B,2], initialised) receives per-clip min/max."""
        require_cuda(waveform)
        lead = waveform.shape[:-1]
        L = waveform.shape[-1]
        pcm16 = waveform.dtype == torch.int16
        w = waveform.reshape(-1, L).contiguous() if pcm16 else waveform.reshape(-1, L).float().contiguous()
        B = w.shape[0]
        T = self.n_frames(L)
        tab = self.tables(w.device)
        if time_major:
            buf = torch.empty(B, T, self.n_mels, device=w.device, dtype=torch.float32)
            out = buf.transpose(1, 2)
        else:
            out = torch.empty(B, self.n_mels, T, device=w.device, dtype=torch.float32)
        fn = lib().sedk_logmel_fwd_i16 if pcm16 else lib().sedk_logmel_fwd
        check(fn(ptr(w), B, L, tab.struct, ptr(out), out.stride(0), out.stride(1), out.stride(2),
                 1 if log else 0, amin, db_range[0], db_range[1], ptr(minmax), stream_ptr()), "sedk_logmel_fwd")
        return out.reshape(lead + out.shape[1:]) if len(lead) != 1 else out

    def forward(self, waveform):
        return self.run(waveform, log=False)


def new_minmax(B, device):
    mm = torch.empty(B, 2, dtype=torch.int32, device=device)
    check(lib().sedk_minmax_init(ptr(mm), B, stream_ptr()), "sedk_minmax_init")
    return mm


def decode_minmax(mm):
    out = torch.empty(mm.shape[0], 2, dtype=torch.float32, device=mm.device)
    check(lib().sedk_minmax_decode(ptr(mm), ptr(out), mm.shape[0], stream_ptr()), "sedk_minmax_decode")
    return out


def take_log(mels, amin=1e-5, db_range=(-50.0, 80.0), perm=None, coef=None, minmax=None, log=True):
    """SEDTask4.take_log (sed_trainer.py:253-264): AmplitudeToDB('amplitude', amin=1e-5) + clamp(-50, 80), optionally
    fused with mixup on the linear mel (perm int64 [B], coef fp32 [B]) and the per-clip min/max reduction."""
    require_cuda(mels)
    x = mels.float().contiguous()
    B = x.shape[0] if x.dim() > 1 else 1
    n = x.numel() // B
    out = torch.empty_like(x)
    check(lib().sedk_feat_mix_log(ptr(x), ptr(perm), ptr(coef), ptr(out), B, n, 1 if log else 0, amin, db_range[0],
                                  db_range[1], ptr(minmax), stream_ptr()), "sedk_feat_mix_log")
    return out


class AmplitudeToDB(nn.Module):
    """Subset of torchaudio.transforms.AmplitudeToDB used by the recipes: stype='amplitude', top_db=None."""

    def __init__(self, stype="power", top_db=None):
        super().__init__()
        if stype != "amplitude" or top_db is not None:
            raise NotImplementedError("desed_task_b200.AmplitudeToDB supports stype='amplitude', top_db=None")
        self.amin = 1e-10
        self.multiplier = 20.0

    def forward(self, x):
        return take_log(x, amin=self.amin, db_range=(-float("inf"), float("inf")))
