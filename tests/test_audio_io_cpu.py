"""Host-side input path (include/sedk_io.h, desed_task_b200/audio_io.py): PCM16 WAV decoding with the reference's
read_audio semantics (desed_task/dataio/datasets.py:14-74) and pre-decoded int16 shards.

torchaudio.load itself cannot run in this image (it needs torchcodec), so the comparison stands torchaudio's documented
int16 normalisation (sample / 32768, float32, [channels, frames]) in for the decode and runs the reference's OWN `to_mono` /
`pad_audio` source on top of it when the checkout is present (extracted with ast from /root/reference: the module itself
imports h5py, which is not installed); the restatement below is what runs on a box without the checkout, and is held equal
to the reference's functions whenever both exist."""
import ast
import os
import random
import re
import struct
import wave

import numpy as np
import pytest
import torch

REF = "/root/reference/desed_task/dataio/datasets.py"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def aio():
    from desed_task_b200 import build
    build.build_io()
    from desed_task_b200 import audio_io
    return audio_io


def write_wav(path, data, fs=16000, extensible=False, junk=None):
    """data int16 [frames] or [frames, channels]"""
    data = np.asarray(data, np.int16)
    ch = 1 if data.ndim == 1 else data.shape[1]
    if not extensible and junk is None:
        with wave.open(str(path), "wb") as w:
            w.setnchannels(ch)
            w.setsampwidth(2)
            w.setframerate(fs)
            w.writeframes(data.tobytes())
        return path
    raw = data.tobytes()
    if extensible:
        fmt = struct.pack("<HHIIHHHHIH14s", 0xFFFE, ch, fs, fs * ch * 2, ch * 2, 16, 22, 16, 0, 1,
                          bytes.fromhex("000000001000800000aa00389b71"))
    else:
        fmt = struct.pack("<HHIIHH", 1, ch, fs, fs * ch * 2, ch * 2, 16)
    chunks = b"fmt " + struct.pack("<I", len(fmt)) + fmt
    if junk is not None:
        chunks += b"LIST" + struct.pack("<I", len(junk)) + junk + (b"\0" if len(junk) & 1 else b"")
    chunks += b"data" + struct.pack("<I", len(raw)) + raw
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 4 + len(chunks)) + b"WAVE" + chunks)
    return path


def load_like_torchaudio(path):
    """torchaudio.load(normalize=True) on 16-bit PCM: float32 [channels, frames] = sample / 32768."""
    with wave.open(str(path), "rb") as w:
        ch, fs, n = w.getnchannels(), w.getframerate(), w.getnframes()
        x = np.frombuffer(w.readframes(n), np.int16).reshape(n, ch).T.astype(np.float32) / 32768.0
    return torch.from_numpy(x.copy()), fs


# ---- restatement of datasets.py:14-47 (what runs without the checkout)
def to_mono(mixture, random_ch=False):
    if mixture.ndim > 1:
        if not random_ch:
            mixture = torch.mean(mixture, 0)
        else:
            indx = np.random.randint(0, mixture.shape[0] - 1)
            mixture = mixture[indx]
    return mixture


def pad_audio(audio, target_len, fs, test=False):
    if audio.shape[-1] < target_len:
        audio = torch.nn.functional.pad(audio, (0, target_len - audio.shape[-1]), mode="constant")
        padded_indx = [target_len / len(audio)]
        onset_s = 0.000
    elif len(audio) > target_len:
        clip_onset = 0 if test else random.randint(0, len(audio) - target_len)
        audio = audio[clip_onset:clip_onset + target_len]
        onset_s = round(clip_onset / fs, 3)
        padded_indx = [target_len / len(audio)]
    else:
        onset_s = 0.000
        padded_indx = [1.0]
    offset_s = round(onset_s + (target_len / fs), 3)
    return audio, onset_s, offset_s, padded_indx


def reference_functions():
    """The reference's own to_mono / pad_audio, compiled from its source file (None without the checkout)."""
    if not os.path.isfile(REF):
        return None
    tree = ast.parse(open(REF).read())
    keep = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("to_mono", "pad_audio")]
    ns = {"torch": torch, "np": np, "random": random}
    exec(compile(ast.Module(body=keep, type_ignores=[]), REF, "exec"), ns)
    return ns["to_mono"], ns["pad_audio"]


def read_audio_ref(fns, file, random_channel, pad_to, test):
    """datasets.py:57-74 with the decode stood in (see module docstring)."""
    tm, pa = fns
    mixture, fs = load_like_torchaudio(file)
    mixture = tm(mixture, random_channel)
    mixture, onset_s, offset_s, padded_indx = pa(mixture, pad_to, fs, test=test)
    return mixture.float(), onset_s, offset_s, padded_indx


@pytest.fixture(scope="module")
def files(tmp_path_factory):
    d = tmp_path_factory.mktemp("wav")
    rng = np.random.RandomState(0)
    mk = lambda n, c=None: rng.randint(-32768, 32768, size=(n,) if c is None else (n, c)).astype(np.int16)  # noqa: E731
    out = {
        "short": write_wav(d / "short.wav", mk(3001)),
        "exact": write_wav(d / "exact.wav", mk(8000)),
        "long": write_wav(d / "long.wav", mk(20011)),
        "long2": write_wav(d / "long2.wav", mk(9999)),
        "stereo": write_wav(d / "stereo.wav", mk(12000, 2)),
        "three": write_wav(d / "three.wav", mk(5000, 3)),
        "ext": write_wav(d / "ext.wav", mk(7000), extensible=True),
        "junk": write_wav(d / "junk.wav", mk(6000), junk=b"INFOodd"),
        "rate": write_wav(d / "rate.wav", mk(50000), fs=44100),
    }
    return {k: str(v) for k, v in out.items()}


@pytest.mark.parametrize("use_reference", [False, True])
@pytest.mark.parametrize("test_mode", [True, False])
def test_batch_equals_a_loop_over_read_audio(aio, files, use_reference, test_mode):
    fns = reference_functions() if use_reference else (to_mono, pad_audio)
    if fns is None:
        pytest.skip("reference checkout not present")
    names = ["short", "exact", "long", "long2", "ext", "junk", "rate", "long"]
    paths = [files[k] for k in names]
    pad_to = 8000
    random.seed(5)
    np.random.seed(5)
    want = [read_audio_ref(fns, p, False, pad_to, test_mode) for p in paths]
    state_ref = random.getstate()
    random.seed(5)
    np.random.seed(5)
    audio, onset_s, offset_s, padded = aio.read_audio_batch(paths, pad_to, test=test_mode, n_threads=3)
    assert random.getstate() == state_ref                      # the same draws were consumed
    assert audio.dtype == torch.int16 and tuple(audio.shape) == (len(paths), pad_to)
    for i, (w, o, f, p) in enumerate(want):
        # x / 32768 in fp32 is exact: the int16 batch IS the reference's float waveform
        assert torch.equal(audio[i].float() / 32768.0, w), names[i]
        assert onset_s[i] == o and offset_s[i] == f and padded[i] == p, names[i]
    # and the fp32 output path gives the same values directly
    random.seed(5)
    out32 = torch.empty(len(paths), pad_to, dtype=torch.float32)
    aio.read_audio_batch(paths, pad_to, test=test_mode, out=out32)
    assert torch.equal(out32, torch.stack([w for w, *_ in want]))


@pytest.mark.parametrize("use_reference", [False, True])
def test_multichannel_mean_and_random_channel(aio, files, use_reference):
    fns = reference_functions() if use_reference else (to_mono, pad_audio)
    if fns is None:
        pytest.skip("reference checkout not present")
    paths = [files["stereo"], files["three"], files["stereo"]]
    random.seed(9)
    np.random.seed(9)
    want = [read_audio_ref(fns, p, False, 6000, False) for p in paths]
    random.seed(9)
    np.random.seed(9)
    audio, onset_s, offset_s, _ = aio.read_audio_batch(paths, 6000)
    assert audio.dtype == torch.float32
    for i, (w, o, f, _) in enumerate(want):
        assert torch.equal(audio[i], w) and onset_s[i] == o and offset_s[i] == f
    with pytest.raises(ValueError):                            # a channel mean is not an int16 signal
        aio.read_audio_batch(paths, 6000, out=torch.empty(3, 6000, dtype=torch.int16))
    # random_channel: np.random.randint(0, channels - 1) - never the last channel (datasets.py:19)
    random.seed(2)
    np.random.seed(2)
    want = [read_audio_ref(fns, p, True, 6000, True) for p in paths]
    random.seed(2)
    np.random.seed(2)
    audio, *_ = aio.read_audio_batch(paths, 6000, test=True, random_channel=True)
    assert audio.dtype == torch.int16
    for i, (w, *_r) in enumerate(want):
        assert torch.equal(audio[i].float() / 32768.0, w)
    with pytest.raises(ValueError):                            # mono + random channel: randint(0, 0) raises upstream too
        aio.read_audio_batch([files["short"]], 6000, random_channel=True)


def test_restatement_equals_the_reference_functions():
    fns = reference_functions()
    if fns is None:
        pytest.skip("reference checkout not present")
    g = torch.Generator().manual_seed(0)
    for shape, tl in (((2, 900), 500), ((1, 300), 500), ((3, 500), 500)):
        x = torch.randn(*shape, generator=g)
        for rc in (False, True):
            if rc and shape[0] == 1:
                continue
            np.random.seed(1)
            a = to_mono(x, rc)
            np.random.seed(1)
            b = fns[0](x, rc)
            assert torch.equal(a, b)
            random.seed(3)
            ra = pad_audio(a, tl, 16000)
            random.seed(3)
            rb = fns[1](b, tl, 16000)
            assert torch.equal(ra[0], rb[0]) and ra[1:] == rb[1:]


def test_probe_and_errors(aio, files, tmp_path):
    info = aio.wav_info(files["three"])
    assert (info.sample_rate, info.channels, info.bits_per_sample, info.frames) == (16000, 3, 16, 5000)
    assert aio.wav_info(files["rate"]).sample_rate == 44100
    assert aio.wav_info(files["junk"]).frames == 6000
    with pytest.raises(aio.SedkIoError, match="No such file"):
        aio.read_audio_batch([str(tmp_path / "missing.wav")], 100)
    bad = tmp_path / "bad.wav"
    bad.write_bytes(b"not a wave file at all, definitely")
    with pytest.raises(aio.SedkIoError, match="RIFF"):
        aio.read_audio_batch([files["short"], str(bad)], 100)
    eight = tmp_path / "eight.wav"
    with wave.open(str(eight), "wb") as w:
        w.setnchannels(1)
        w.setsampwidth(1)
        w.setframerate(16000)
        w.writeframes(bytes(100))
    with pytest.raises(aio.SedkIoError, match="16-bit"):
        aio.wav_info(str(eight))
    assert tuple(aio.read_audio_batch([], 100)[0].shape) == (0, 100)


def test_shard_round_trip(aio, files, tmp_path):
    clips = [torch.from_numpy(np.random.RandomState(i).randint(-32768, 32768, size=n).astype(np.int16))
             for i, n in enumerate((100, 8000, 12345, 0, 8001))]
    path = aio.write_pcm16_shard(str(tmp_path / "a.shard"), clips, sample_rate=16000)
    sh = aio.Pcm16Shard(path)
    assert len(sh) == 5 and sh.sample_rate == 16000 and [sh.length(i) for i in range(5)] == [100, 8000, 12345, 0, 8001]
    idx = [2, 0, 4, 3, 1, 2]
    random.seed(11)
    want = [pad_audio(clips[i].float(), 8000, 16000) for i in idx]
    state = random.getstate()
    random.seed(11)
    out = torch.empty(len(idx), 8000, dtype=torch.int16).pin_memory() if torch.cuda.is_available() else None
    audio, onset_s, offset_s, padded = sh.read_batch(idx, 8000, out=out, n_threads=2)
    assert random.getstate() == state
    for i, (w, o, f, p) in enumerate(want):
        assert torch.equal(audio[i].float(), w) and onset_s[i] == o and offset_s[i] == f and padded[i] == p
    audio_t, onset_t, *_ = sh.read_batch(idx, 8000, test=True)
    assert all(o == 0 for o in onset_t) and torch.equal(audio_t[0], clips[2][:8000])
    with pytest.raises(IndexError):
        sh.read_batch([5], 10)
    with pytest.raises(IndexError):
        sh.length(9)
    sh.close()
    # a shard written from decoded files serves the same batches as decoding them
    wavs = [files["short"], files["long"], files["exact"]]
    dec, *_ = aio.read_audio_batch(wavs, 30000, test=True)               # longer than any clip: nothing is cut
    lens = [aio.wav_info(f).frames for f in wavs]
    sh2 = aio.Pcm16Shard(aio.write_pcm16_shard(str(tmp_path / "b.shard"), [dec[i, :lens[i]] for i in range(3)]))
    a, *_ = sh2.read_batch([0, 1, 2], 8000, test=True)
    b, *_ = aio.read_audio_batch(wavs, 8000, test=True)
    assert torch.equal(a, b)
    sh2.close()
    trunc = tmp_path / "trunc.shard"
    trunc.write_bytes(open(path, "rb").read()[:200])
    with pytest.raises(aio.SedkIoError, match="shard"):
        aio.Pcm16Shard(str(trunc))


def test_header_symbols_are_exported_and_bound(aio):
    hdr = open(os.path.join(ROOT, "include", "sedk_io.h")).read()
    names = set(re.findall(r"SEDKIO_API\s+[\w\s\*]+?\b(sedkio_\w+)\s*\(", hdr))
    assert len(names) >= 10
    L = aio.lib()
    for n in names:
        assert hasattr(L, n), n
        assert n in aio._SIGS, n
