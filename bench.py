#!/usr/bin/env python
"""bench.py - clips/s of the SED hot path (10-s 16-kHz clip -> 128-mel -> CRNN) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload supervised|mean_teacher|dcase2024|inference] [--batch B]

Workloads (BASELINE.json `configs`; the metric is quoted on `supervised` = configs[1], the default):
  supervised    CRNN supervised training, batch 24/GPU ([12 strong, 12 weak]), dcase2023 CRNN, train-mode BN, dropout 0.5,
                SpecAugment, BCE strong + BCE weak, Adam.
  mean_teacher  configs[2]: batch 48/GPU [12 strong, 12 weak, 24 unlabelled], mixup, student fwd+bwd + teacher fwd, BCE + MSE
                consistency, EMA + Adam.
  dcase2024     configs[3]: 2024 CRNN (27 classes, 192-unit BiGRU) + synthetic frame embeddings [B, 768, 496], five-way batch
                split, mixup on features and embeddings, masked losses, gradient clipping.
  inference     configs[4]: mel -> eval CRNN -> median filter (k = 7), batch 64/GPU, clips sharded, no collective.
A step = one pass of the whole hot path over one batch.  `value` times K steps with inputs resident in HBM; `e2e` times the
same K steps through the public engine call with pinned HOST batches (H2D of the inputs and a D2H read of the loss / the
scores inside the timed region).  `--impl reference` times the reference's own CPU path (its unmodified modules from
baseline/_ref, else the oracle port) on the host cores.  Rank 0 prints ONE JSON line.
"""
import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line: NCCL's own banner / debug output ("NCCL version ...") goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

# ... and so does everything else a library may print (PyTorch's own "NCCL version" banner goes to stdout): fd 1 is pointed
# at stderr for the whole run and the JSON line is written to the saved descriptor at the end
_STDOUT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line):
    sys.stdout.flush()
    os.write(_STDOUT_FD, (line + "\n").encode())


import torch  # noqa: E402

L_SAMPLES = 160000
N_MELS, HOP = 128, 256
NET_2023 = dict(dropout=0.5, rnn_layers=2, n_in_channel=1, nclass=10, attention=True, n_RNN_cell=128, activation="glu",
                rnn_type="BGRU", kernel_size=[3] * 7, padding=[1] * 7, stride=[1] * 7,
                nb_filters=[16, 32, 64, 128, 128, 128, 128],
                pooling=[[2, 2], [2, 2], [1, 2], [1, 2], [1, 2], [1, 2], [1, 2]], dropout_recurrent=0,
                use_embeddings=False)
# recipes/dcase2024_task4_baseline/confs/pretrained.yaml:86-110
NET_2024 = dict(NET_2023, dropout=0.2, rnn_layers=1, nclass=27, n_RNN_cell=192, use_embeddings=True, embedding_size=768,
                embedding_type="frame", aggregation_type="pool1d", specaugm_t_p=0.0, specaugm_f_p=0.0,
                dropstep_recurrent=0.0, dropstep_recurrent_len=16)
EMB_SHAPE = (768, 496)
METRIC = "clips/sec (10s 16kHz 128-mel CRNN fwd+bwd)"
DEFAULT_BATCH = {"supervised": 24, "mean_teacher": 48, "dcase2024": 24, "inference": 64}


def batch_split(workload, B):
    if workload == "mean_teacher":
        return [B // 4, B // 4, B // 2]
    if workload == "dcase2024":
        # shipped [12 maestro, 6 synth, 6 strong, 12 weak, 24 unlabelled] = 60 (pretrained.yaml:8), kept in proportion
        if B == 60:
            return [12, 6, 6, 12, 24]
        base = [B * 12 // 60, B * 6 // 60, B * 6 // 60, B * 12 // 60]
        base = [max(1, v) for v in base]
        return base + [B - sum(base)]
    if workload == "inference":
        return [B]
    return [B // 2, B - B // 2, 0]


def workload_config(workload, B, world):
    """The `config` object, identical for both arms (the driver compares them)."""
    desc = {
        "supervised": "dcase2023 CRNN supervised training step (mel + fwd + BCE + bwd + Adam), %d clips/GPU x 10 s @16 kHz, "
                      "128 mel, train-mode BN, dropout 0.5, SpecAugment" % B,
        "mean_teacher": "dcase2023 CRNN mean-teacher step (mel, mixup, student fwd+bwd, teacher fwd, BCE + MSE consistency, "
                        "EMA, Adam), %d clips/GPU x 10 s @16 kHz, 128 mel" % B,
        "dcase2024": "dcase2024 CRNN (27 classes, BiGRU-192) + frame embeddings [768, 496] mean-teacher step (5-way split, "
                     "mixup on features and embeddings, class-masked losses, grad clip 5.0), %d clips/GPU x 10 s" % B,
        "inference": "dcase2023 CRNN inference (mel + eval forward + median filter k=7), %d clips/GPU x 10 s @16 kHz, clips "
                     "sharded over ranks" % B,
    }[workload]
    return {"workload": desc, "name": workload, "global_batch": B * world, "batch_split": batch_split(workload, B),
            "parallelism": "dp%d" % world}


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nme, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# =====================================================================================================================
# algorithmic work per launch
def layer_geometry(B, net=NET_2023):
    T, F, cin = 1 + L_SAMPLES // HOP, N_MELS, 1
    out = []
    for cout, (pt, pf) in zip(net["nb_filters"], net["pooling"]):
        out.append(dict(cin=cin, cout=cout, T=T, F=F, pt=pt, pf=pf, pix=B * T * F))
        T, F, cin = T // pt, F // pf, cout
    return out


def kernel_work(name, B, net=NET_2023):
    """Algorithmic FLOPs and minimal HBM bytes of ONE launch of a kernel family."""
    geo = layer_geometry(B, net)
    m = re.match(r"conv(?:3x3|_wgrad)(?:_tc5)?(?:_pair)?_(\d+)to(\d+)_F(\d+)$", name)
    if m:
        a, b, F = int(m.group(1)), int(m.group(2)), int(m.group(3))
        cands = [x for x in geo if x["F"] == F] or [x for x in geo if x["F"] == 2 * F]     # paired view: F / 2, 2 C
        pix = cands[0]["pix"] if [x for x in geo if x["F"] == F] else cands[0]["pix"] / 2
        return 2.0 * 9 * a * b * pix, 4.0 * pix * (a + b)
    m = re.match(r"gemm(?:_tc5)?_[NT]{2}_(\d+)x(\d+)x(\d+)_[xk](\d+)$", name)
    if m:
        M, N, K, n = [int(v) for v in m.groups()]
        return 2.0 * M * N * K * n, 4.0 * n * (M * K + K * N) + 4.0 * M * N * (n if "_x" in name else 1)
    if name.startswith("glu_wgrad_tc5_c"):
        C = int(name.rsplit("_c", 1)[1])
        pix = sum(x["pix"] for x in geo if x["cout"] in (64, 128)) / 5.0     # mean over the five launches of a step
        return 2.0 * C * C * pix, 4.0 * pix * C * 2
    if name.startswith("bnglu_pool_fwd_c") or name.startswith("bnglu_pool_bwd_c") or name.startswith("bn_bwd_apply_c"):
        C = int(name.rsplit("_c", 1)[1])
        cands = [x for x in geo if x["cout"] == C]
        pix = sum(x["pix"] for x in cands) / len(cands)
        pool = sum(x["pt"] * x["pf"] for x in cands) / len(cands)
        if name.startswith("bnglu_pool_fwd"):
            return 2.0 * C * C * pix, 4.0 * pix * C * (1 + 1 / pool)
        if name.startswith("bnglu_pool_bwd"):
            return 6.0 * C * C * pix, 4.0 * pix * C * (2 + 1 / pool)
        return 8.0 * pix * C, 4.0 * pix * C * 3
    if name.startswith("bnglu_tc5_fwd_c") or name.startswith("bnglu_tc5_bwd_c"):
        C = int(name.rsplit("_c", 1)[1])
        cands = [x for x in geo if x["cout"] == C]
        pix = sum(x["pix"] for x in cands) / len(cands)
        if "_fwd_" in name:       # z in, lin out (saved for backward), pooled out
            return 2.0 * C * C * pix, 4.0 * pix * C * 2.5
        return 2.0 * C * C * pix, 4.0 * pix * C * 4.5          # z, lin, gout/2 in; g_lin, gy out
    if name == "conv0_fwd":
        g = geo[0]
        return 2.0 * 9 * g["cout"] * g["pix"], 4.0 * g["pix"] * (2 + g["cout"])
    if name == "conv0_wgrad":
        g = geo[0]
        return 2.0 * 9 * g["cout"] * g["pix"], 4.0 * g["pix"] * (1 + g["cout"])
    # store-free first block (csrc/layer0.cu): the convolution output never reaches HBM; algorithmic bytes are the
    # one-channel input + the pooled output (forward) / its gradient (backward)
    if name == "l0_prep":
        g = geo[0]
        return 2.0 * 2 * 9 * g["cout"] * g["pix"], 4.0 * g["pix"] * 2
    if name == "l0_fwd":
        g = geo[0]
        return 2.0 * (9 + g["cout"]) * g["cout"] * g["pix"], 4.0 * g["pix"] * (1 + g["cout"] / 4.0)
    if name == "l0_bwd":
        g = geo[0]
        return 2.0 * (2 * 9 + 3 * g["cout"]) * g["cout"] * g["pix"], 4.0 * g["pix"] * (1 + g["cout"] / 4.0)
    if name == "logmel":
        return 626 * 70e3 * B, 960512.0 * B
    if name.startswith("gru_seq"):
        H = net["n_RNN_cell"]
        return 2.0 * 3 * H * H * B * 156 * 2, 4.0 * B * 156 * 2 * (3 * H + 6 * H)
    if name == "adam_ema":
        n = 1112420 if net["nclass"] == 10 else 1785094
        return 12.0 * n, 40.0 * n
    return 0.0, 0.0


FAMILIES = [
    ("front end (logmel)", r"^logmel$"),
    ("conv0 (1->16, CUDA cores)", r"^conv0_"),
    ("first block, store-free (stencil + BN + GLU + pool recomputed)", r"^l0_"),
    ("conv3x3 tcgen05 fwd/dgrad", r"^conv3x3_tc5"),
    ("conv3x3 mma.sync fwd/dgrad", r"^conv3x3_\d"),
    ("conv wgrad tcgen05", r"^conv_wgrad_tc5"),
    ("conv wgrad mma.sync", r"^conv_wgrad_\d"),
    ("BN+GLU+pool C<=32 (registers)", r"^bnglu_pool"),
    ("BN+GLU+pool C>=64 (tcgen05)", r"^bnglu_tc5|^glu_wgrad_tc5"),
    ("BN backward apply", r"^bn_bwd_apply"),
    ("GRU recurrence", r"^gru_seq"),
    ("GRU / fusion GEMMs", r"^gemm"),
    ("heads + loss", r"^heads|^sed_loss|^dropout|^emb_"),
    ("optimizer (Adam + EMA)", r"^adam_ema|^sumsq"),
    ("glue (packs, BN finalize, fix-ups)", r"."),
]


def family_table(prof, NP, B, net, pk):
    """Per-family: ms/step, algorithmic FLOPs / bytes per step, achieved rates and the fraction of the bounding peak."""
    ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        traffic = {}
    rows = {}
    for name, (cnt, tot) in prof.items():
        fam = next(f for f, pat in FAMILIES if re.search(pat, name))
        fl, by = kernel_work(name, B, net)
        r = rows.setdefault(fam, dict(ms=0.0, launches=0.0, flops=0.0, bytes=0.0, dram=0.0, dram_known=True))
        r["ms"] += tot / NP
        r["launches"] += cnt / NP
        r["flops"] += fl * cnt / NP
        r["bytes"] += by * cnt / NP
        if name in traffic and by > 0:
            r["dram"] += traffic[name] * cnt / NP
        elif by > 0:
            r["dram_known"] = False
    out = []
    for fam, _ in FAMILIES:
        if fam not in rows:
            continue
        r = rows[fam]
        e = {"family": fam, "ms_per_step": round(r["ms"], 4), "launches_per_step": round(r["launches"], 1)}
        if r["bytes"] > 0 and r["ms"] > 0:
            gbs, tfs = r["bytes"] / r["ms"] / 1e6, r["flops"] / r["ms"] / 1e9
            e.update({"algorithmic_MB": round(r["bytes"] / 1e6, 2), "algorithmic_GFLOP": round(r["flops"] / 1e9, 3),
                      "GBps": round(gbs, 1), "TFLOPs": round(tfs, 2)})
            if r["flops"] / r["bytes"] < ridge:
                e.update({"bound": "hbm", "frac": round(gbs / pk["hbm_gbs"], 4)})
            else:
                e.update({"bound": "tensor", "frac": round(tfs / pk["bf16_tflops_sustained"], 4)})
            if r["dram"] > 0 and r["dram_known"]:
                e["dram_over_algorithmic"] = round(r["dram"] / r["bytes"], 2)
        out.append(e)
    return out


def make_batches(nbuf, B, seed, pin, nclass=10, emb=False, pcm16=False, emb_pooled=False):
    g = torch.Generator().manual_seed(seed)
    n_s = B // 2
    audio, labels, embs = [], [], []
    for _ in range(nbuf):
        a = torch.randn(B, L_SAMPLES, generator=g) * 0.1
        if pcm16:                                              # the stored format of the datasets: 16-bit PCM
            a = (a * 32768.0).round().clamp(-32768, 32767).to(torch.int16)
        y = (torch.rand(B, nclass, 156, generator=g) < 0.1).float()
        y[n_s:, :, 1:] = 0.0                                   # weak clips: clip-level tags live in frame 0
        if pin:
            a, y = a.pin_memory(), y.pin_memory()
        audio.append(a)
        labels.append(y)
        if emb:
            # reference storage: fp32 [768, 496]; --emb-pooled: pre-pooled to the CRNN's 156 frames, bf16 (embeddings.py)
            e = torch.randn(B, EMB_SHAPE[0], 156, generator=g).bfloat16() if emb_pooled else torch.randn(B, *EMB_SHAPE, generator=g)
            embs.append(e.pin_memory() if pin else e)
    if emb:
        return audio, labels, embs
    return audio, labels


def class_masks_2024(split):
    """valid_class_mask per row: MAESTRO rows use the 17 MAESTRO classes, DESED rows the 10 DESED classes
    (recipes/dcase2024_task4_baseline/train_pretrained.py:190-193 / local/classes_dict.py)."""
    B = sum(split)
    cm = torch.zeros(B, 27, dtype=torch.bool)
    cm[:split[0], 10:] = True
    cm[split[0]:, :10] = True
    return cm


# =====================================================================================================================
def run_ours(args):
    import torch.distributed as dist
    from desed_task_b200 import _lib
    from desed_task_b200.engine import InferEngine, TrainEngine
    from desed_task_b200.frontend import MelSpectrogram
    from desed_task_b200.nnet.CRNN import CRNN
    from desed_task_b200.optim import FusedAdam
    from desed_task_b200.utils.schedulers import ExponentialWarmup

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d does not match WORLD_SIZE %d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("launch multi-GPU runs with torch.distributed.run --nproc-per-node %d" % args.gpus)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    L = _lib.lib()
    if L.sedk_device_cc() < 100:
        raise SystemExit("bench.py needs an sm_100 device (got cc %d); there is no fallback path" % L.sedk_device_cc())

    wl, B = args.workload, args.batch
    split = batch_split(wl, B)
    is_2024, is_inf = wl == "dcase2024", wl == "inference"
    net_cfg = NET_2024 if is_2024 else NET_2023
    nclass = net_cfg["nclass"]
    torch.manual_seed(42)
    student = CRNN(**net_cfg).to(dev)
    teacher = None
    if wl in ("mean_teacher", "dcase2024"):
        import copy
        teacher = copy.deepcopy(student)
        for p in teacher.parameters():
            p.detach_()
        teacher.train()
    mel = MelSpectrogram(16000, 2048, 2048, HOP, 0, 8000, n_mels=N_MELS, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1).to(dev)
    opt = sched = None
    if is_inf:
        student.eval()
    else:
        student.train()
        opt = FusedAdam(student, 1e-3, betas=(0.9, 0.999))
        sched = ExponentialWarmup(opt, 1e-3, 50 * 250)

    def new_engine(use_graph, distributed=True):
        adt = torch.int16 if args.pcm16 else torch.float32
        if is_inf:
            return InferEngine(student, mel, B, L_SAMPLES, median_window=7, use_graph=use_graph, audio_dtype=adt)
        if is_2024:
            return TrainEngine(student, mel, split, L_SAMPLES, opt=opt, scheduler=sched, teacher=teacher,
                               mixup_type="soft", use_graph=use_graph, distributed=distributed, grad_clip=5.0,
                               emb_shape=(EMB_SHAPE[0], 156) if args.emb_pooled else EMB_SHAPE,
                               emb_dtype=torch.bfloat16 if args.emb_pooled else torch.float32,
                               class_masks=class_masks_2024(split).to(dev), recipe="2024", audio_dtype=adt)
        return TrainEngine(student, mel, split, L_SAMPLES, opt=opt, scheduler=sched, teacher=teacher,
                           mixup_type="soft" if teacher is not None else None, use_graph=use_graph,
                           distributed=distributed, audio_dtype=adt)

    eng = new_engine(True)
    NBUF = max(4, -(-192 // B) + 1) if not is_2024 else 4          # distinct batches: > 126 MB of audio (+ embeddings) > L2
    mk = make_batches(NBUF, B, 42 + rank, pin=True, nclass=nclass, emb=is_2024, pcm16=args.pcm16,
                      emb_pooled=args.emb_pooled)
    host_a, host_y = mk[0], mk[1]
    host_e = mk[2] if is_2024 else [None] * NBUF
    dev_a = [a.to(dev) for a in host_a]
    dev_y = [y.to(dev) for y in host_y]
    dev_e = [e.to(dev) if e is not None else None for e in host_e]

    def one(e, src, i):
        a, y, em = src
        if is_inf:
            return e.step(a[i % NBUF], inputs_ready=True)
        if is_2024:
            return e.step(a[i % NBUF], y[i % NBUF], em[i % NBUF], inputs_ready=True)
        return e.step(a[i % NBUF], y[i % NBUF], inputs_ready=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(use_host, steps, warm, sample=True):
        src = (host_a, host_y, host_e) if use_host else (dev_a, dev_y, dev_e)
        # the clock sampler (an nvidia-smi child polling every 100 ms) is started BEFORE the warm-up: its start-up (NVML
        # initialisation, driver locks) perturbs the first ~50 ms, which would otherwise be the whole timed region
        sampler = ClockSampler(local)
        if rank == 0 and sample:
            sampler.start()
            t_wait = time.time()
            while not sampler.rows and time.time() - t_wait < 3.0:
                time.sleep(0.02)
        for i in range(warm):
            one(eng, src, i)
        barrier()
        launches0, replays0 = L.sedk_launch_count(), eng.replays
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            one(eng, src, warm + i)
        e1.record()
        barrier()
        if use_host:
            eng.read() if is_inf else eng.read_losses()          # the per-step D2H results have all landed
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if (rank == 0 and sample) else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        launches = (L.sedk_launch_count() - launches0) + (eng.replays - replays0) * eng.graph_kernels
        return ms, clocks, launches

    W = max(args.warmup, 3)
    # settle: graph capture / instantiation and the CPU-heavy set-up leave the clocks and the first replays off steady
    # state; a fixed number of extra untimed steps precedes the W warm-up + K timed steps of each measurement
    for i in range(args.settle):
        one(eng, (dev_a, dev_y, dev_e), i)
    ms_dev, clocks, launches = timed(False, args.steps, W)
    ms_e2e, clocks_e2e, _ = timed(True, args.steps, W)
    # spread: the same K-step block repeated (device-resident inputs), min / median reported next to the headline block
    reps = sorted(timed(False, args.steps, 1, sample=False)[0] / args.steps for _ in range(args.repeats))
    final = None if is_inf else eng.read_losses()
    total_clips = B * world * args.steps
    value = total_clips / ms_dev * 1e3
    e2e = total_clips / ms_e2e * 1e3
    pk, pk_kind = peaks()

    out = None
    if rank == 0:
        # ---- per-kernel device times (eager, outside any timed region) -> dominant kernel + roofline + family table
        prof = {}
        eng2 = new_engine(False, distributed=False)       # rank-0-only pass: no collective in it
        src = (dev_a, dev_y, dev_e)
        for i in range(3):
            one(eng2, src, i)
        torch.cuda.synchronize(dev)
        L.sedk_profile_enable(1)
        NP = 5
        for i in range(NP):
            one(eng2, src, 3 + i)
        torch.cuda.synchronize(dev)
        buf = ctypes.create_string_buffer(1 << 16)
        _lib.check(L.sedk_profile_report(buf, len(buf)), "sedk_profile_report")
        L.sedk_profile_enable(0)
        for line in buf.value.decode().strip().splitlines():
            nme, cnt, tot = line.split()
            prof[nme] = (int(cnt), float(tot))
        step_ms_eager = sum(t for _, t in prof.values()) / NP
        top = sorted(prof.items(), key=lambda kv: -kv[1][1])
        dom, (dcnt, dtot) = top[0]
        avg_ms = dtot / dcnt
        flops, nbytes = kernel_work(dom, B, net_cfg)
        ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        if nbytes > 0 and flops / nbytes < ridge:
            roof = {"bound": "hbm", "achieved": round(nbytes / avg_ms / 1e6, 1), "peak": pk["hbm_gbs"], "unit": "GB/s"}
        else:
            roof = {"bound": "tensor", "achieved": round(flops / avg_ms / 1e9, 2), "peak": pk["bf16_tflops_sustained"],
                    "unit": "TFLOP/s"}
        roof["frac"] = round(roof["achieved"] / roof["peak"], 4) if roof["peak"] else None
        us_per_step = {k: round(v[1] / v[0] * 1e3 / 156.0, 3) for k, v in prof.items() if k.startswith("gru_seq")}
        if dom.startswith("gru_seq"):
            roof["note"] = ("latency-bound persistent recurrence: 156 dependent time steps per launch, %.2f us per time step; "
                            "neither HBM nor the tensor pipe can bound it (DESIGN.md section 4)" % (avg_ms * 1e3 / 156.0))
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom)
        except Exception:
            pass
        roof["traffic"] = traffic
        roof.update({"kernel": dom, "avg_launch_ms": round(avg_ms, 5), "launches_per_step": dcnt / NP,
                     "share_of_step": round(dtot / NP / step_ms_eager, 4), "peak_source": pk_kind + " (MEASURED_PEAKS.json"
                     " bf16 sustained / hbm copy; TF32 nominal is half the bf16 rate)",
                     "algorithmic_flops": flops, "algorithmic_bytes": nbytes})
        breakdown = [{"kernel": k, "ms_per_step": round(v[1] / NP, 4), "launches_per_step": v[0] / NP} for k, v in top]
        # ---- front-end bandwidth (the second headline: mel GB/s vs HBM peak) and its fp32 rate
        lm_cnt, lm_tot = prof.get("logmel", (1, 0.0))
        mel_gbs = B * 960512 / (lm_tot / lm_cnt) / 1e6 if lm_tot > 0 else None
        mel_tfs = B * 626 * 70e3 / (lm_tot / lm_cnt) / 1e9 if lm_tot > 0 else None
        cpu = lib_bar = None
        if world == 1 and not args.quick:
            cpu = cpu_baseline(wl, B, budget_s=20.0)
            lib_bar = library_bar(dev, wl, B, dev_a, dev_y)
        h2d = B * L_SAMPLES * (2 if args.pcm16 else 4) + (0 if is_inf else B * nclass * 156 * 4 + 64) + ((B * EMB_SHAPE[0] * 156 * 2 if args.emb_pooled else B * EMB_SHAPE[0] * EMB_SHAPE[1] * 4) if is_2024 else 0)
        d2h = (B * nclass * 156 * 4 + B * nclass * 4) if is_inf else 64
        cfg = workload_config(wl, B, world)             # identical in both arms; arm-specific detail goes to `implementation`
        impl = {"precision": "front end fp32; CRNN GEMMs TF32 (fp32 storage / accumulate); recurrence, BN statistics, "
                             "heads, losses, optimizer fp32",
                "l2": "inputs rotate over %d distinct batches (%.0f MB > L2); per-step activations ~0.6 GB"
                      % (NBUF, NBUF * B * L_SAMPLES * 4 / 1e6),
                "cuda_graph": "forward + loss + backward (+ fused EMA/Adam at N = 1): one graph replay per step; at N > 1 "
                              + ("the gradient all-reduce is this library's own NVLink kernel fused with EMA + Adam "
                                 "(csrc/nvls.cu, multimem.ld_reduce / multimem.st over symmetric memory), inside the same graph"
                                 if getattr(eng, "ar_mode", "") == "nvls" else
                                 "the NCCL all-reduce and the optimizer kernel follow the graph as eager launches"),
                "first_block": "store-free (csrc/layer0.cu): conv0 output recomputed from the one-channel input, BN backward + "
                               "conv0 weight gradient in closed form" if L.sedk_get_option(b"l0_fused", 1) else "conv0 output through HBM",
                "embeddings": ("pre-pooled bf16 [768, 156]" if args.emb_pooled else "fp32 [768, 496]") if is_2024 else None,
                "audio_input": "int16 PCM (x / 32768 in the front end)" if args.pcm16 else "fp32 waveform",
                "settle_steps": args.settle, "allreduce": getattr(eng, "ar_mode", None) if world > 1 else None,
                "overlap": "front end of step k+1 runs on its own stream concurrently with step k's graph (ping-pong "
                           "log-mel buffers); weight-gradient GEMMs on a side branch of the graph"}
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic", "config": cfg, "implementation": impl,
            "clocks": clocks,
            "e2e": {"value": round(e2e, 1), "unit": "clips/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "clocks": clocks_e2e},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "gpu_library_baseline": lib_bar,
            "repeats": {"blocks": args.repeats, "steps_per_block": args.steps,
                        "ms_per_step_min": round(reps[0], 4) if reps else None,
                        "ms_per_step_median": round(reps[len(reps) // 2], 4) if reps else None,
                        "ms_per_step_max": round(reps[-1], 4) if reps else None},
            "kernel_families": family_table(prof, NP, B, net_cfg, pk),
            "kernel_breakdown_ms": breakdown,
            "eager_step_ms": round(step_ms_eager, 4),
            "gru_us_per_time_step": us_per_step,
            "frontend": {"mel_GBps": None if mel_gbs is None else round(mel_gbs, 1),
                         "frac_of_hbm_peak": None if mel_gbs is None else round(mel_gbs / pk["hbm_gbs"], 4),
                         "fp32_TFLOPs": None if mel_tfs is None else round(mel_tfs, 2),
                         "algorithmic_bytes_per_clip": 960512, "fp32_flop_per_clip": 626 * 70e3},
            "final_loss": final,
        }
        if lib_bar and isinstance(lib_bar.get("best"), (int, float)) and lib_bar["best"] > 0:
            out["vs_gpu_library"] = round(value / lib_bar["best"], 2)
        emit(json.dumps(out))
    if world > 1:
        # no collective teardown: every rank leaves as soon as its own work is done (rank 0's profiling pass above is local);
        # destroying a process group whose collectives were captured into CUDA graphs was seen to hang
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return out


# =====================================================================================================================
def library_bar(dev, workload, B, dev_a, dev_y):
    """the reference's UNMODIFIED modules on the same B200 through PyTorch's library kernels."""
    if workload == "dcase2024":
        return {"unavailable": "the 2024 library leg is not wired (needs the Lightning trainer's batch plumbing)"}
    try:
        from baseline import reference_arm
        if dev_a[0].dtype == torch.int16:               # the reference consumes the normalised fp32 waveform
            dev_a = [a.float() / 32768.0 for a in dev_a]
        res = reference_arm.gpu_library_baseline(dev, workload, B, dev_a, dev_y, steps=15, warmup=4)
    except Exception as e:                                           # noqa: BLE001
        return {"unavailable": str(e).splitlines()[0][:200]}
    nums = [v for v in res.values() if isinstance(v, (int, float))]
    res.update({"unit": "clips/s", "best": max(nums) if nums else None,
                "what": "reference desed_task.nnet.CRNN + torchaudio MelSpectrogram/AmplitudeToDB + reference TorchScaler + "
                        "torch.optim.Adam on cuda:0 (torch %s, cuDNN %s; cudnn TF32 on, matmul fp32 = PyTorch defaults), "
                        "inputs resident" % (torch.__version__, torch.backends.cudnn.version())})
    return res


def reference_cpu_step(workload, B, threads):
    """One step of the reference's own CPU path.  Returns (callable, kind): the unmodified reference modules when
    baseline/_ref (or the checkout) is present (`kind` "reference"), else the oracle port (`kind` "port")."""
    torch.set_num_threads(threads)
    split = batch_split(workload, B)
    if workload != "dcase2024":
        try:
            from baseline import reference_arm
            audio, labels = make_batches(1, B, 7, pin=False)
            audio, labels = audio[0], labels[0]
            path = reference_arm.ReferencePath("cpu", teacher=workload == "mean_teacher")
            if workload == "supervised":
                return (lambda: float(path.supervised_step(audio, labels, split[0]).detach())), "reference"
            if workload == "mean_teacher":
                return (lambda: float(path.mean_teacher_step(audio, labels, split).detach())), "reference"
            return (lambda: len(path.inference(audio))), "reference"
        except ImportError:
            pass
    from oracle import crnn as ocrnn, trainer as otr
    cfg = ocrnn.CFG_2024 if workload == "dcase2024" else ocrnn.CFG_2023
    P = ocrnn.init_params(cfg, seed=42)
    names = ocrnn.param_names(P)
    for k in names:
        P[k].requires_grad_(True)
    state = {}
    mk = make_batches(1, B, 7, pin=False, nclass=cfg.nclass, emb=workload == "dcase2024")
    audio, labels = mk[0][0], mk[1][0]
    if workload == "dcase2024":
        Pt = {k: v.detach().clone() for k, v in P.items()}
        emb, cm = mk[2][0], class_masks_2024(split)

        def step24():
            out = otr.mean_teacher_step_2024(P, Pt, audio, labels, emb, cm, split, 1, 12500, cfg, mix=None)
            grads = torch.autograd.grad(out["tot_loss"], [P[k] for k in names])
            with torch.no_grad():
                otr.update_ema(0.999, 1, P, Pt, names)
                otr.adam_step({k: P[k] for k in names}, dict(zip(names, grads)), state, names, 1e-3)
            return float(out["tot_loss"].detach())
        return step24, "port"
    n_s = split[0]

    def step():
        spec = ocrnn.draw_specaugment(B, N_MELS, 626)
        loss, _, _ = otr.supervised_step(P, audio, labels, n_s, B - n_s, cfg, True, fwd_kw=dict(specaug=spec),
                                         gru_impl="aten")
        grads = torch.autograd.grad(loss, [P[k] for k in names])
        with torch.no_grad():
            otr.adam_step({k: P[k] for k in names}, dict(zip(names, grads)), state, names, 1e-3)
        return float(loss.detach())
    return step, "port"


def best_threads(workload):
    """The reference path is many small torch CPU ops: more threads is not always faster.  Time one small step at a few
    thread counts and keep the fastest (reported as `cores`)."""
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
    best, best_t = cands[0], None
    for t in cands:
        step, _ = reference_cpu_step(workload, 4 if workload != "dcase2024" else 10, t)
        step()
        t0 = time.time()
        step()
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best, best_t = t, dt
    return best


def cpu_baseline(workload, B, budget_s=20.0):
    prev = torch.get_num_threads()
    threads = best_threads(workload)
    step, kind = reference_cpu_step(workload, B, threads)
    step()                                                       # warm-up
    t0 = time.time()
    n = 0
    while n < 2 or (time.time() - t0 < budget_s and n < 20):
        step()
        n += 1
    dt = time.time() - t0
    torch.set_num_threads(prev)
    return {"value": round(B * n / dt, 2), "unit": "clips/s", "cores": threads, "kind": kind,
            "sample": "%d full %s steps of %d clips (fp32, torch CPU ops, %d threads), %.1f s" % (n, workload, B, threads, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wl, B = args.workload, args.batch
    threads = best_threads(wl)
    step, kind = reference_cpu_step(wl, B, threads)
    W = max(args.warmup, 1)
    for _ in range(W):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    v = round(B * args.steps / dt, 2)
    what = ("the reference's unmodified modules (baseline/_ref: desed_task.nnet.CRNN, data_augm, TorchScaler; torchaudio "
            "front end; torch.optim.Adam)" if kind == "reference" else "oracle port of the reference's torch/torchaudio path")
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": W, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(wl, B, max(1, args.gpus)),
        "implementation": "host CPU (rank 0 only), fp32, %s" % what,
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": threads, "kind": kind,
                         "sample": "%d steps of %d clips, %d of %d host threads (fastest of a sweep)"
                                   % (args.steps, B, threads, os.cpu_count() or 1)},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="supervised", choices=sorted(DEFAULT_BATCH))
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU (default: 24 / 48 / 24 / 64 by workload)")
    ap.add_argument("--repeats", type=int, default=5, help="extra timed K-step blocks for the min / median spread")
    ap.add_argument("--pcm16", action="store_true", help="feed 16-bit PCM audio (bit-identical front end, half the H2D bytes)")
    ap.add_argument("--emb-pooled", action="store_true", dest="emb_pooled",
                    help="dcase2024: feed embeddings in the pre-pooled bf16 storage format (240 KB instead of 1.52 MB per clip)")
    ap.add_argument("--settle", type=int, default=20, help="extra untimed steps before the warm-up (after graph capture)")
    ap.add_argument("--quick", action="store_true", help="development runs: skip the CPU / GPU-library baseline legs")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = DEFAULT_BATCH[args.workload]
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
