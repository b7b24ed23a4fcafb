"""Host side of the input pipeline over include/sedk_io.h / lib/libsedkio.so:

* `read_audio_batch`  - the reference's `read_audio` (desed_task/dataio/datasets.py:57-74: torchaudio.load -> to_mono ->
  pad_audio -> float()) for a whole batch of 16-bit PCM WAV files on a thread pool, straight into one (pinned) host tensor.
  Mono / picked-channel clips come back as int16 (what the engines take with audio_dtype=torch.int16: the front end divides
  by 32768 in its load path, bit-identical to the normalised fp32 waveform); channel means as fp32.  The Python `random` /
  `np.random` draws of pad_audio / to_mono are made here, file by file in order, so a loop over the reference's read_audio
  and one call of this function consume the same random streams.
* `write_pcm16_shard` / `Pcm16Shard` - pre-decoded int16 shards ("pre-decoded int16 shards / pinned-memory ring" of the
  survey): one mmap-ed file per few thousand clips, batches gathered by index with pad_audio's pad / cut rule.

The reference quirks are kept: `padded_indx` is computed AFTER padding / cutting and is therefore always [1.0]
(datasets.py:30,41); `random_channel` draws `np.random.randint(0, channels - 1)` and so never picks the last channel and
raises for a mono file (datasets.py:19).
"""
import ctypes as C
import os
import random

import numpy as np
import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "lib", "libsedkio.so")
_lib = None


class SedkIoError(RuntimeError):
    pass


class WavInfo(C.Structure):
    _fields_ = [("sample_rate", C.c_int32), ("channels", C.c_int32), ("bits_per_sample", C.c_int32), ("frames", C.c_int64),
                ("data_offset", C.c_int64)]


_SIGS = {
    "sedkio_last_error": (C.c_char_p, []),
    "sedkio_wav_probe": (C.c_int, [C.c_char_p, C.POINTER(WavInfo)]),
    "sedkio_read_audio_batch": (C.c_int, [C.POINTER(C.c_char_p), C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.POINTER(WavInfo), C.c_void_p, C.c_int]),
    "sedkio_shard_write": (C.c_int, [C.c_char_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int, C.c_int32]),
    "sedkio_shard_open": (C.c_int, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "sedkio_shard_close": (None, [C.c_void_p]),
    "sedkio_shard_clips": (C.c_int64, [C.c_void_p]),
    "sedkio_shard_sample_rate": (C.c_int32, [C.c_void_p]),
    "sedkio_shard_length": (C.c_int64, [C.c_void_p, C.c_int64]),
    "sedkio_shard_gather": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_int]),
}


def lib():
    """Loads lib/libsedkio.so (built by `python -m desed_task_b200.build`); no fallback to a Python decoder."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            raise SedkIoError("libsedkio.so is not built: run `python -m desed_task_b200.build`")
        h = C.CDLL(_LIB)
        for name, (res, args) in _SIGS.items():
            fn = getattr(h, name)
            fn.restype, fn.argtypes = res, args
        _lib = h
    return _lib


def _check(rc, what):
    if rc != 0:
        raise SedkIoError("%s failed (%d): %s" % (what, rc, lib().sedkio_last_error().decode(errors="replace")))


def wav_info(path):
    info = WavInfo()
    _check(lib().sedkio_wav_probe(os.fsencode(path), C.byref(info)), "sedkio_wav_probe")
    return info


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def read_audio_batch(files, pad_to, test=False, random_channel=False, out=None, n_threads=0):
    """`[read_audio(f, False, random_channel, pad_to, test) for f in files]` in one call.

    Returns (audio [n, pad_to] host tensor - int16 when every clip is one channel or a picked channel, fp32 otherwise or
    when `out` is fp32 -, onset_s list, offset_s list, padded_indx list), the last three as pad_audio returns them."""
    files = [os.fspath(f) for f in files]
    n = len(files)
    L = lib()
    infos = (WavInfo * max(n, 1))()
    for i, f in enumerate(files):
        _check(L.sedkio_wav_probe(os.fsencode(f), C.byref(infos[i])), "sedkio_wav_probe")
    # the reference's draws, per file in order: to_mono first (np.random), then pad_audio (random)
    channel = np.full(n, -1, np.int32)
    onset = np.zeros(n, np.int64)
    onset_s, offset_s, padded = [], [], []
    for i in range(n):
        ch, frames, fs = infos[i].channels, infos[i].frames, infos[i].sample_rate
        if random_channel:
            channel[i] = np.random.randint(0, ch - 1)                    # datasets.py:19 (never the last channel)
        if frames > pad_to and not test:
            onset[i] = random.randint(0, frames - pad_to)                # datasets.py:36
        o_s = round(onset[i] / fs, 3) if frames > pad_to else 0.000
        onset_s.append(o_s)
        offset_s.append(round(o_s + (pad_to / fs), 3))
        padded.append([1.0])                                             # computed after the pad / cut upstream: always 1
    need_f32 = any(infos[i].channels > 1 and channel[i] < 0 for i in range(n))
    if out is None:
        out = torch.empty(n, pad_to, dtype=torch.float32 if need_f32 else torch.int16)
    if out.dtype not in (torch.int16, torch.float32) or tuple(out.shape) != (n, pad_to) or not out.is_contiguous() \
            or out.is_cuda:
        raise ValueError("out must be a contiguous host int16 / fp32 tensor [n, pad_to]")
    if need_f32 and out.dtype != torch.float32:
        raise ValueError("a channel mean is not an int16 signal: pass an fp32 `out` (or random_channel=True)")
    if n == 0:
        return out, onset_s, offset_s, padded
    paths = (C.c_char_p * max(n, 1))(*[os.fsencode(f) for f in files])
    status = np.zeros(max(n, 1), np.int32)
    p = C.c_void_p(out.data_ptr())
    _check(L.sedkio_read_audio_batch(paths, n, pad_to, _vp(onset), _vp(channel), p if out.dtype == torch.int16 else None,
                                     p if out.dtype == torch.float32 else None, infos, _vp(status), n_threads),
           "sedkio_read_audio_batch")
    return out, onset_s, offset_s, padded


def write_pcm16_shard(path, clips, sample_rate=16000):
    """clips: list of 1-D int16 tensors / arrays (any lengths) -> one SEDKPCM1 shard."""
    arrs = [np.ascontiguousarray(c.numpy() if isinstance(c, torch.Tensor) else c, dtype=np.int16).reshape(-1) for c in clips]
    n = len(arrs)
    stride = max([a.size for a in arrs] + [1])
    buf = np.zeros((max(n, 1), stride), np.int16)
    lengths = np.zeros(max(n, 1), np.int64)
    for i, a in enumerate(arrs):
        buf[i, :a.size] = a
        lengths[i] = a.size
    _check(lib().sedkio_shard_write(os.fsencode(path), _vp(buf), stride, _vp(lengths), n, sample_rate), "sedkio_shard_write")
    return path


class Pcm16Shard:
    """mmap-ed SEDKPCM1 shard.  `read_batch(indices, pad_to)` = pad_audio's rule on every clip, gathered by a thread pool
    into one host tensor (pass a pinned `out` to feed the engines' H2D ring)."""

    def __init__(self, path):
        self._h = C.c_void_p()
        _check(lib().sedkio_shard_open(os.fsencode(path), C.byref(self._h)), "sedkio_shard_open")
        self.path = path
        self.sample_rate = int(lib().sedkio_shard_sample_rate(self._h))

    def __len__(self):
        return int(lib().sedkio_shard_clips(self._h))

    def length(self, i):
        n = int(lib().sedkio_shard_length(self._h, int(i)))
        if n < 0:
            raise IndexError(i)
        return n

    def read_batch(self, indices, pad_to, test=False, out=None, n_threads=0):
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        n = idx.size
        onset = np.zeros(max(n, 1), np.int64)
        onset_s, offset_s = [], []
        for i in range(n):
            frames = self.length(idx[i])
            if frames > pad_to and not test:
                onset[i] = random.randint(0, frames - pad_to)            # datasets.py:36
            o_s = round(onset[i] / self.sample_rate, 3) if frames > pad_to else 0.000
            onset_s.append(o_s)
            offset_s.append(round(o_s + (pad_to / self.sample_rate), 3))
        if out is None:
            out = torch.empty(n, pad_to, dtype=torch.int16)
        if out.dtype != torch.int16 or tuple(out.shape) != (n, pad_to) or not out.is_contiguous() or out.is_cuda:
            raise ValueError("out must be a contiguous host int16 tensor [n, pad_to]")
        if n:
            _check(lib().sedkio_shard_gather(self._h, _vp(idx), n, pad_to, _vp(onset), C.c_void_p(out.data_ptr()),
                                             n_threads), "sedkio_shard_gather")
        return out, onset_s, offset_s, [[1.0]] * n

    def close(self):
        if self._h:
            lib().sedkio_shard_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:      # noqa: BLE001 - interpreter shutdown
            pass
