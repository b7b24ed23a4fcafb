#!/usr/bin/env python
"""bench.py - clips/s of the SED hot path (10-s 16-kHz clip -> 128-mel -> CRNN fwd+bwd+Adam) on N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload supervised|mean_teacher]

Workload (BASELINE.json configs[1], the one the metric is quoted on): supervised CRNN training, batch 24 per GPU
([12 strong, 12 weak]), dcase2023 CRNN, 10-s clips, train-mode BN, dropout 0.5, SpecAugment, BCE strong + BCE weak, Adam.
A step = one pass of the whole hot path over one batch.  `value` times K steps with inputs resident in HBM; `e2e` times
the same K steps through the public engine call with pinned HOST batches (H2D of audio+labels and a D2H read of the loss
inside the timed region).  `--impl reference` times the reference's CPU path (oracle port of the reference modules) on
the host cores.  One JSON line is printed by rank 0.
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

import torch  # noqa: E402

L_SAMPLES = 160000
N_MELS, HOP = 128, 256
NET_2023 = dict(dropout=0.5, rnn_layers=2, n_in_channel=1, nclass=10, attention=True, n_RNN_cell=128, activation="glu",
                rnn_type="BGRU", kernel_size=[3] * 7, padding=[1] * 7, stride=[1] * 7,
                nb_filters=[16, 32, 64, 128, 128, 128, 128],
                pooling=[[2, 2], [2, 2], [1, 2], [1, 2], [1, 2], [1, 2], [1, 2]], dropout_recurrent=0,
                use_embeddings=False)
METRIC = "clips/sec (10s 16kHz 128-mel CRNN fwd+bwd)"


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


def layer_geometry(B):
    T, F, cin = 1 + L_SAMPLES // HOP, N_MELS, 1
    out = []
    for cout, (pt, pf) in zip(NET_2023["nb_filters"], NET_2023["pooling"]):
        out.append(dict(cin=cin, cout=cout, T=T, F=F, pt=pt, pf=pf, pix=B * T * F))
        T, F, cin = T // pt, F // pf, cout
    return out


def kernel_work(name, B):
    """Algorithmic FLOPs and minimal HBM bytes of ONE launch of a kernel family."""
    geo = layer_geometry(B)
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
    if name == "logmel":
        return 626 * 70e3 * B, 960512.0 * B
    if name.startswith("gru_seq"):
        H = NET_2023["n_RNN_cell"]
        return 2.0 * 3 * H * H * B * 156 * 2, 4.0 * B * 156 * 2 * (3 * H + 6 * H)
    return 0.0, 0.0


def make_batches(nbuf, B, seed, pin):
    g = torch.Generator().manual_seed(seed)
    n_s = B // 2
    audio, labels = [], []
    for _ in range(nbuf):
        a = torch.randn(B, L_SAMPLES, generator=g) * 0.1
        y = (torch.rand(B, 10, 156, generator=g) < 0.1).float()
        y[n_s:, :, 1:] = 0.0                                   # weak clips: clip-level tags live in frame 0
        if pin:
            a, y = a.pin_memory(), y.pin_memory()
        audio.append(a)
        labels.append(y)
    return audio, labels


# =====================================================================================================================
def run_ours(args):
    import torch.distributed as dist
    from desed_task_b200 import _lib
    from desed_task_b200.engine import TrainEngine
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

    B = args.batch
    mean_teacher = args.workload == "mean_teacher"
    batch_sizes = [B // 4, B // 4, B // 2] if mean_teacher else [B // 2, B - B // 2, 0]
    torch.manual_seed(42)
    student = CRNN(**NET_2023).to(dev)
    student.train()
    teacher = None
    if mean_teacher:
        import copy
        teacher = copy.deepcopy(student)
        for p in teacher.parameters():
            p.detach_()
        teacher.train()
    mel = MelSpectrogram(16000, 2048, 2048, HOP, 0, 8000, n_mels=N_MELS, window_fn=torch.hamming_window,
                         wkwargs={"periodic": False}, power=1).to(dev)
    opt = FusedAdam(student, 1e-3, betas=(0.9, 0.999))
    sched = ExponentialWarmup(opt, 1e-3, 50 * 250)

    def new_engine(use_graph, distributed=True):
        return TrainEngine(student, mel, batch_sizes, L_SAMPLES, opt=opt, scheduler=sched, teacher=teacher,
                           mixup_type="soft" if mean_teacher else None, use_graph=use_graph, distributed=distributed)

    eng = new_engine(True)
    NBUF = 12                                                    # 12 x 15.4 MB of audio > 126 MB L2
    host_a, host_y = make_batches(NBUF, B, 42 + rank, pin=True)
    dev_a = [a.to(dev) for a in host_a]
    dev_y = [y.to(dev) for y in host_y]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(use_host, steps, warm):
        src_a, src_y = (host_a, host_y) if use_host else (dev_a, dev_y)
        for i in range(warm):
            eng.step(src_a[i % NBUF], src_y[i % NBUF], inputs_ready=True)
        barrier()
        launches0, replays0 = L.sedk_launch_count(), eng.replays
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            r = eng.step(src_a[(warm + i) % NBUF], src_y[(warm + i) % NBUF], inputs_ready=True)
            if use_host:
                pass                                             # the D2H loss read is issued inside step(); drained below
        e1.record()
        barrier()
        if use_host:
            eng.read_losses()
        ms = e0.elapsed_time(e1)
        clocks = sampler.stop() if rank == 0 else None
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        launches = (L.sedk_launch_count() - launches0) + (eng.replays - replays0) * eng.graph_kernels
        return ms, clocks, launches

    W = max(args.warmup, 3)
    ms_dev, clocks, launches = timed(False, args.steps, W)
    ms_e2e, clocks_e2e, _ = timed(True, args.steps, W)
    loss_now = eng.read_losses()
    total_clips = B * world * args.steps
    value = total_clips / ms_dev * 1e3
    e2e = total_clips / ms_e2e * 1e3
    pk, pk_kind = peaks()

    out = None
    if rank == 0:
        # ---- per-kernel device times (eager, outside any timed region) -> dominant kernel + roofline
        prof = {}
        eng2 = new_engine(False, distributed=False)       # rank-0-only pass: no collective in it
        for i in range(3):
            eng2.step(dev_a[i], dev_y[i])
        torch.cuda.synchronize(dev)
        L.sedk_profile_enable(1)
        NP = 5
        for i in range(NP):
            eng2.step(dev_a[(3 + i) % NBUF], dev_y[(3 + i) % NBUF])
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
        flops, nbytes = kernel_work(dom, B)
        ridge = pk["bf16_tflops_sustained"] * 1e12 / (pk["hbm_gbs"] * 1e9)
        if nbytes > 0 and flops / nbytes < ridge:
            roof = {"bound": "hbm", "achieved": round(nbytes / avg_ms / 1e6, 1), "peak": pk["hbm_gbs"], "unit": "GB/s"}
        else:
            roof = {"bound": "tensor", "achieved": round(flops / avg_ms / 1e9, 2), "peak": pk["bf16_tflops_sustained"],
                    "unit": "TFLOP/s"}
        roof["frac"] = round(roof["achieved"] / roof["peak"], 4)
        if dom.startswith("gru_seq"):
            roof["note"] = ("latency-bound persistent recurrence: 156 dependent time steps per launch, %.2f us per step on "
                            "%d of the SMs; neither HBM nor the tensor pipe can bound it (DESIGN.md section 4)"
                            % (avg_ms * 1e3 / 156.0, 2 * B))
        # measured DRAM traffic of that kernel (dram__bytes_read.sum + dram__bytes_write.sum per launch) from the committed
        # ncu full-set capture, when there is one (profiles/traffic.json: {kernel family: bytes per launch})
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
        top5 = [{"kernel": k, "ms_per_step": round(v[1] / NP, 4), "launches_per_step": v[0] / NP} for k, v in top]
        # ---- front-end bandwidth (the second headline: mel GB/s vs HBM peak)
        lm_cnt, lm_tot = prof.get("logmel", (1, 0.0))
        mel_gbs = B * 960512 / (lm_tot / lm_cnt) / 1e6 if lm_tot > 0 else None
        cpu = cpu_baseline(B, budget_s=20.0) if (world == 1 and not args.quick) else None
        out = {
            "metric": METRIC, "value": round(value, 1), "unit": "clips/s", "n_gpus": world, "steps": args.steps,
            "warmup": W, "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "tf32", "data": "synthetic",
            "config": {"workload": "dcase2023 CRNN %s training step, %d clips/GPU x 10 s @16 kHz, 128 mel; front end fp32, "
                                   "CRNN GEMMs TF32 (fp32 storage/accumulate), train-mode BN, dropout 0.5, SpecAugment, "
                                   "Adam" % (args.workload, B),
                       "global_batch": B * world, "batch_split": batch_sizes, "parallelism": "dp%d" % world,
                       "l2": "inputs rotate over %d distinct batches (%.0f MB > L2); per-step activations ~0.6 GB"
                             % (NBUF, NBUF * B * L_SAMPLES * 4 / 1e6),
                       "cuda_graph": True,
                       "overlap": "front end of step k+1 runs on its own stream concurrently with step k's graph "
                                  "(ping-pong log-mel buffers); weight-gradient GEMMs on a side branch of the graph"},
            "clocks": clocks,
            "e2e": {"value": round(e2e, 1), "unit": "clips/s", "ms_per_step": round(ms_e2e / args.steps, 4),
                    "h2d_bytes_per_step": B * L_SAMPLES * 4 + B * 10 * 156 * 4 + 64, "d2h_bytes_per_step": 64,
                    "clocks": clocks_e2e},
            "gpu_launches": int(launches),
            "roofline": roof,
            "cpu_baseline": cpu,
            "kernel_breakdown_ms": top5,
            "eager_step_ms": round(step_ms_eager, 4),
            "frontend": {"mel_GBps": None if mel_gbs is None else round(mel_gbs, 1),
                         "frac_of_hbm_peak": None if mel_gbs is None else round(mel_gbs / pk["hbm_gbs"], 4),
                         "algorithmic_bytes_per_clip": 960512},
            "final_loss": loss_now,
        }
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return out


# =====================================================================================================================
def cpu_step_fn(B, threads):
    """The reference's own CPU path for this workload: oracle port (plain-torch restatement of the reference modules,
    pinned against the live reference in oracle/make_golden.py).  Returns a callable running one full training step."""
    import dataclasses
    from oracle import crnn as ocrnn, trainer as otr
    torch.set_num_threads(threads)
    cfg = ocrnn.CFG_2023
    P = ocrnn.init_params(cfg, seed=42)
    names = ocrnn.param_names(P)
    for k in names:
        P[k].requires_grad_(True)
    state = {}
    audio, labels = make_batches(1, B, 7, pin=False)
    audio, labels = audio[0], labels[0]
    n_s = B // 2

    def step():
        spec = ocrnn.draw_specaugment(B, N_MELS, 626)
        loss, _, _ = otr.supervised_step(P, audio, labels, n_s, B - n_s, cfg, True, fwd_kw=dict(specaug=spec),
                                         gru_impl="aten")
        grads = torch.autograd.grad(loss, [P[k] for k in names])
        with torch.no_grad():
            otr.adam_step({k: P[k] for k in names}, dict(zip(names, grads)), state, names, 1e-3)
        return float(loss.detach())
    return step


def best_threads():
    """The reference path is many small torch CPU ops: more threads is not faster.  Time one small step at a few thread
    counts and keep the fastest (reported as `cores`)."""
    ncpu = os.cpu_count() or 1
    cands = sorted({t for t in (8, 16, 32, 64, ncpu) if t <= ncpu})
    best, best_t = cands[0], None
    for t in cands:
        step = cpu_step_fn(4, t)
        step()
        t0 = time.time()
        step()
        dt = time.time() - t0
        if best_t is None or dt < best_t:
            best, best_t = t, dt
    return best


def cpu_baseline(B, budget_s=20.0):
    threads = best_threads()
    prev = torch.get_num_threads()
    step = cpu_step_fn(B, threads)
    step()                                                       # warm-up
    t0 = time.time()
    n = 0
    while n < 2 or (time.time() - t0 < budget_s and n < 20):
        step()
        n += 1
    dt = time.time() - t0
    torch.set_num_threads(prev)
    return {"value": round(B * n / dt, 2), "unit": "clips/s", "cores": threads, "kind": "port",
            "sample": "%d full training steps of %d clips (fp32, torch CPU ops, %d threads), %.1f s" % (n, B, threads, dt)}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = args.batch
    threads = best_threads()
    step = cpu_step_fn(B, threads)
    W = max(1, min(args.warmup, 2))
    for _ in range(W):
        step()
    t0 = time.time()
    for _ in range(args.steps):
        step()
    dt = time.time() - t0
    v = round(B * args.steps / dt, 2)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": W, "ms_per_step": round(dt / args.steps * 1e3, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "dcase2023 CRNN %s training step on the host CPU (oracle port of the reference's torch/"
                               "torchaudio path), %d clips x 10 s per step" % (args.workload, B), "global_batch": B},
        "cpu_baseline": {"value": v, "unit": "clips/s", "cores": threads, "kind": "port",
                         "sample": "%d steps of %d clips, %d of %d host threads (fastest of a sweep)"
                                   % (args.steps, B, threads, os.cpu_count() or 1)},
        "e2e": {"value": v, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="supervised", choices=["supervised", "mean_teacher"])
    ap.add_argument("--batch", type=int, default=None, help="clips per GPU (default 24 supervised / 48 mean-teacher)")
    ap.add_argument("--quick", action="store_true", help="development runs: skip the CPU baseline leg")
    args = ap.parse_args()
    if args.batch is None:
        args.batch = 48 if args.workload == "mean_teacher" else 24
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
