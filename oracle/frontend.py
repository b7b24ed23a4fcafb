"""Oracle: log-mel front end + scaler (CPU, fp32).  Test infrastructure only.

Restates, in plain torch ops:
  * torchaudio.transforms.MelSpectrogram as constructed at
    recipes/dcase2023_task4_baseline/local/sed_trainer.py:79-91
    (torchaudio 2.11.0 functional.py:54-145 `spectrogram`, :492-587 `melscale_fbanks`,
    transforms/_transforms.py:407-417 `MelScale.forward`);
  * SEDTask4.take_log, sed_trainer.py:253-264;
  * desed_task/utils/scaler.py:90-120 TorchScaler.forward.
"""
import math

import torch

SAMPLE_RATE = 16000
N_FFT = 2048
HOP = 256
N_MELS = 128
F_MIN = 0.0
F_MAX = 8000.0


def hamming_window(n=N_FFT):
    """torch.hamming_window(n, periodic=False) (sed_trainer.py:88-89): 0.54-0.46cos(2 pi k/(n-1))."""
    return torch.hamming_window(n, periodic=False, dtype=torch.float32)


def _hz_to_mel_htk(f):
    # torchaudio functional.py:439-440
    return 2595.0 * math.log10(1.0 + (f / 700.0))


def melscale_fbanks(n_freqs=N_FFT // 2 + 1, f_min=F_MIN, f_max=F_MAX, n_mels=N_MELS,
                    sample_rate=SAMPLE_RATE):
    """HTK triangular filterbank, norm=None, [n_freqs, n_mels] fp32.

    Same op sequence as torchaudio functional.py:563-573 + :507-513 so the result is
    bit-identical to `MelSpectrogram(...).mel_scale.fb`.
    """
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = _hz_to_mel_htk(f_min)
    m_max = _hz_to_mel_htk(f_max)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10.0 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    zero = torch.zeros(1)
    down_slopes = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up_slopes = slopes[:, 2:] / f_diff[1:]
    return torch.max(zero, torch.min(down_slopes, up_slopes))


def spectrogram(wave, n_fft=N_FFT, hop=HOP, window=None):
    """|STFT| with center=True reflect padding, onesided, power=1 (functional.py:107-145)."""
    if window is None:
        window = hamming_window(n_fft)
    shape = wave.shape
    w2 = wave.reshape(-1, shape[-1])
    spec = torch.stft(w2, n_fft=n_fft, hop_length=hop, win_length=n_fft, window=window,
                      center=True, pad_mode="reflect", normalized=False, onesided=True,
                      return_complex=True)
    spec = spec.reshape(shape[:-1] + spec.shape[-2:])
    return spec.abs()


def mel_spectrogram(wave, fb=None, window=None):
    """wave [..., L] f32 -> linear-amplitude mel [..., 128, 1+L//256] (sed_trainer.py:282)."""
    if fb is None:
        fb = melscale_fbanks()
    spec = spectrogram(wave, window=window)
    # MelScale.forward, _transforms.py:417: matmul(spec^T, fb)^T
    return torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)


def take_log(mels):
    """AmplitudeToDB('amplitude') with amin=1e-5, ref 1.0, then clamp(-50, 80).
    sed_trainer.py:253-264; torchaudio functional.py `amplitude_to_DB`: 20*log10(clamp(x, amin)) - 20*0."""
    x_db = 20.0 * torch.log10(torch.clamp(mels, min=1e-5))
    x_db = x_db - 20.0 * 0.0
    return x_db.clamp(min=-50, max=80)


def scaler(tensor, statistic="instance", normtype="minmax", dims=(1, 2), eps=1e-8,
           mean=None, mean_squared=None):
    """TorchScaler.forward, desed_task/utils/scaler.py:90-120."""
    if statistic is None or normtype is None:
        return tensor
    if statistic == "dataset":
        if normtype == "mean":
            return tensor - mean
        if normtype == "standard":
            std = torch.sqrt(mean_squared - mean ** 2)
            return (tensor - mean) / (std + eps)
        raise NotImplementedError
    if normtype == "mean":
        return tensor - torch.mean(tensor, dims, keepdim=True)
    if normtype == "standard":
        return (tensor - torch.mean(tensor, dims, keepdim=True)) / (
            torch.std(tensor, dims, keepdim=True) + eps)
    if normtype == "minmax":
        mn = torch.amin(tensor, dim=dims, keepdim=True)
        mx = torch.amax(tensor, dim=dims, keepdim=True)
        return (tensor - mn) / (mx - mn + eps) * 2 - 1
    raise NotImplementedError


def features(wave):
    """waveform -> scaled log-mel, the input of CRNN.forward (sed_trainer.py:266-267,282)."""
    return scaler(take_log(mel_spectrogram(wave)))
