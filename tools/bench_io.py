"""CPU throughput of the host-side input path (include/sedk_io.h): batches of 10-s / 16-kHz / 16-bit clips decoded from WAV
files and gathered from a pre-decoded shard, against a per-item Python decode (wave + numpy -> float32 tensor, what a
DataLoader worker does per clip with torchaudio.load; torchaudio.load itself needs torchcodec, absent from this image).
Files live in the page cache: this times decode + copy, not the disk.

    python tools/bench_io.py [--clips 512] [--batch 64]
"""
import argparse
import os
import sys
import tempfile
import time
import wave

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from desed_task_b200 import audio_io, build  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--clips", type=int, default=512)
    ap.add_argument("--batch", type=int, default=64)
    args = ap.parse_args()
    build.build_io()
    L = 160000
    d = tempfile.mkdtemp(prefix="sedkio_")
    rng = np.random.RandomState(0)
    files = []
    for i in range(args.clips):
        p = os.path.join(d, "%05d.wav" % i)
        with wave.open(p, "wb") as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(rng.randint(-3000, 3000, size=L).astype(np.int16).tobytes())
        files.append(p)
    B = args.batch
    batches = [files[i:i + B] for i in range(0, len(files), B)]
    mb = L * 2 / 1e6

    def report(name, secs, n):
        print("%-58s %9.0f clips/s  %7.2f GB/s of PCM16" % (name, n / secs, n * mb / 1e3 / secs))

    # per-item Python decode (one worker)
    t = time.perf_counter()
    for f in files[:128]:
        with wave.open(f, "rb") as w:
            x = np.frombuffer(w.readframes(w.getnframes()), np.int16).astype(np.float32) / 32768.0
        torch.from_numpy(x)
    report("python wave + numpy -> fp32 tensor, per item, 1 process", time.perf_counter() - t, 128)
    out = torch.empty(B, L, dtype=torch.int16)
    for nt in (1, 4, 16):
        audio_io.read_audio_batch(batches[0], L, test=True, out=out, n_threads=nt)
        t = time.perf_counter()
        for b in batches:
            audio_io.read_audio_batch(b, L, test=True, out=out[:len(b)], n_threads=nt)
        report("sedkio_read_audio_batch (WAV -> int16 batch), %2d threads" % nt, time.perf_counter() - t, len(files))
    dec = torch.empty(len(files), L, dtype=torch.int16)
    for i, b in enumerate(batches):
        audio_io.read_audio_batch(b, L, test=True, out=dec[i * B:i * B + len(b)])
    shard = os.path.join(d, "all.shard")
    audio_io.write_pcm16_shard(shard, [dec[i] for i in range(len(files))])
    sh = audio_io.Pcm16Shard(shard)
    perm = np.random.RandomState(1).permutation(len(files))
    for nt in (1, 4, 16):
        sh.read_batch(perm[:B], L, test=True, out=out, n_threads=nt)
        t = time.perf_counter()
        for i in range(0, len(files), B):
            idx = perm[i:i + B]
            sh.read_batch(idx, L, test=True, out=out[:len(idx)], n_threads=nt)
        report("sedkio_shard_gather (mmap shard -> int16 batch), %2d threads" % nt, time.perf_counter() - t, len(files))
    sh.close()
    print("host: %d cores; %d clips of 10 s @16 kHz (%.0f MB), batch %d" % (os.cpu_count(), len(files), len(files) * mb, B))
    for f in files + [shard]:
        os.remove(f)
    os.rmdir(d)


if __name__ == "__main__":
    main()
