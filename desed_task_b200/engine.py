"""Fused training step of the SED hot path: H2D -> log-mel -> (mixup) -> student fwd -> teacher fwd -> losses -> student
bwd -> [NCCL all-reduce of the flat gradient] -> fused EMA + Adam.

This is the autograd-free restatement of SEDTask4.training_step + on_before_zero_grad + optimizer step
(recipes/dcase2023_task4_baseline/local/sed_trainer.py:269-365, train_sed.py:199-206) that `SEDTask4.fit_step` and
`bench.py` drive.  Everything between the front end and the optimizer is captured once in a CUDA graph (fixed shapes);
per-step scalars (dropout seed counter, mixup coefficients / permutations, consistency weight, Adam bias corrections, EMA
alpha, gradient scale) live in device memory so that a replay sees fresh values.  Host randomness is consumed in the
reference's order: `random.random()` (mixup yes/no), then for the weak and the strong sub-batch `np.random.beta` +
`torch.randperm` (data_augm.py:33-35).
"""
import random

import numpy as np
import torch

from . import data_augm, ddp
from ._lib import check, lib, ptr, stream_ptr
from .frontend import new_minmax
from .optim import FusedAdam, ema_alpha, flatten_parameters


class TrainEngine:
    def __init__(self, student, mel_spec, batch_sizes, n_samples, opt=None, scheduler=None, teacher=None,
                 ema_factor=0.999, const_max=2.0, mixup_type=None, use_graph=True, process_group=None,
                 grad_clip=0.0, emb_shape=None, class_masks=None, distributed=True):
        self.student, self.teacher, self.mel_spec = student, teacher, mel_spec
        self.batch_sizes = list(batch_sizes)
        self.n_s, self.n_w = self.batch_sizes[0], self.batch_sizes[1]
        self.B = int(sum(self.batch_sizes))
        self.L = n_samples
        self.dev = next(student.parameters()).device
        self.opt = opt if opt is not None else FusedAdam(student, 1e-3)
        if not isinstance(self.opt, FusedAdam):
            raise TypeError("TrainEngine drives desed_task_b200.optim.FusedAdam (same maths as torch.optim.Adam)")
        self.scheduler = scheduler
        self.ema_factor, self.const_max = ema_factor, const_max
        self.mixup_type = mixup_type
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if distributed and (process_group is not None or
                            (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        self.grad_clip = grad_clip
        self.emb_shape = emb_shape
        self.class_masks = class_masks
        dev, B = self.dev, self.B
        self.C = student.nclass
        self.T = mel_spec.n_frames(n_samples)
        self.audio_dev = [torch.empty(B, n_samples, device=dev), torch.empty(B, n_samples, device=dev)]
        # the front end runs on its own stream into ping-pong buffers, so that step k+1's log-mel overlaps step k's graph
        self.mel_bufs = [torch.empty(B, mel_spec.n_mels, self.T, device=dev) for _ in range(2)]
        self.mel_buf = self.mel_bufs[0]
        self.logmel = torch.empty_like(self.mel_buf)
        self.Tp = None
        self.labels_dev = None
        self.emb_dev = torch.empty(B, *emb_shape, device=dev) if emb_shape else None
        self.minmaxs = [torch.empty(B, 2, dtype=torch.int32, device=dev) for _ in range(2)]
        self.minmax = self.minmaxs[0]
        self.perm = torch.arange(B, dtype=torch.int64, device=dev)
        self.coef = torch.ones(B, device=dev)
        self.perm_s = torch.arange(max(self.n_s, 1), dtype=torch.int64, device=dev)
        self.perm_w = torch.arange(max(self.n_w, 1), dtype=torch.int64, device=dev)
        self.coef_s = torch.ones(max(self.n_s, 1), device=dev)
        self.coef_w = torch.ones(max(self.n_w, 1), device=dev)
        self.hyper = torch.zeros(4, device=dev)
        self.cw = torch.zeros(1, device=dev)
        self.seed_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        self.losses = torch.zeros(16, device=dev)
        self.gstrong = self.gweak = None
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.ring = 4
        self.host_scal = [torch.zeros(8, dtype=torch.float32).pin_memory() for _ in range(self.ring)]
        self.host_perm = [torch.zeros(B + self.n_s + self.n_w + 2, dtype=torch.int64).pin_memory() for _ in range(self.ring)]
        self.host_coef = [torch.zeros(B + self.n_s + self.n_w + 2, dtype=torch.float32).pin_memory() for _ in range(self.ring)]
        self.host_loss = [torch.zeros(16, dtype=torch.float32).pin_memory() for _ in range(self.ring)]
        self.ring_ev = [None] * self.ring
        self.slot_ev = [None, None]
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.fe_stream = torch.cuda.Stream(device=dev)
        self.teacher_stream = torch.cuda.Stream(device=dev)
        self.buf_free_ev = [None, None]
        self.graphs = [None, None]      # one captured graph per ping-pong buffer
        self.keeps = [None, None]
        self.graph = None
        self.graph_kernels = 0          # kernels captured in the graph (one replay launches all of them)
        self.replays = 0
        self.step_idx = 0
        self.p_flat = flatten_parameters(student)
        self.t_flat = flatten_parameters(teacher) if teacher is not None else None
        student.seed_dev = self.seed_ctr
        if teacher is not None:
            teacher.seed_dev = self.seed_ctr
        self.ws = None

    # ------------------------------------------------------------------------------------------------------------
    def _device_part(self, do_mix, slot=0):
        """Everything between the front end and the optimiser (graph-capturable), reading ping-pong buffer `slot`."""
        s = stream_ptr()
        L = lib()
        n_s, n_w, B = self.n_s, self.n_w, self.B
        self.mel_buf, self.minmax = self.mel_bufs[slot], self.minmaxs[slot]
        check(L.sedk_bump_counter(ptr(self.seed_ctr), 1, s), "sedk_bump_counter")
        if do_mix:
            check(L.sedk_minmax_init(ptr(self.minmax), B, s), "sedk_minmax_init")
        n = self.mel_buf[0].numel()
        labels = self.labels_dev
        labels_weak = (torch.sum(labels[n_s:n_s + n_w], -1) > 0).float() if n_w > 0 else None
        labels_strong = labels[:n_s]
        if do_mix:
            # mixup on the LINEAR mel (sed_trainer.py:296-301), fused with take_log + per-clip min/max
            check(L.sedk_feat_mix_log(ptr(self.mel_buf), ptr(self.perm), ptr(self.coef), ptr(self.logmel), B, n, 1, 1e-5,
                                      -50.0, 80.0, ptr(self.minmax), s), "sedk_feat_mix_log")
            hard = self.mixup_type == "hard"
            if n_w > 0:
                labels_weak = data_augm.mix_labels(labels_weak, self.perm_w, self.coef_w, hard)
            if n_s > 0:
                labels_strong = data_augm.mix_labels(labels_strong.contiguous(), self.perm_s, self.coef_s, hard)
            feats = self.logmel
        else:
            feats = self.mel_buf        # the front-end kernel already wrote the log-mel and the min/max
        emb = self.emb_dev
        cm = self.class_masks
        t_strong = t_weak = None
        if self.teacher is not None:
            # the teacher's (no-grad) forward is independent of the student's: fork it onto its own stream; inside a
            # capture the event pair turns it into a parallel branch of the graph
            cur = torch.cuda.current_stream(self.dev)
            ev_fork = torch.cuda.Event()
            ev_fork.record(cur)
            self.teacher_stream.wait_event(ev_fork)
            with torch.cuda.stream(self.teacher_stream):
                t_strong, t_weak, _ = self.teacher.forward_direct(feats, self.minmax, emb, cm)
                ev_join = torch.cuda.Event()
                ev_join.record(self.teacher_stream)
        strong, weak, ws = self.student.forward_direct(feats, self.minmax, emb, cm)
        self.ws = ws
        if self.teacher is not None:
            torch.cuda.current_stream(self.dev).wait_event(ev_join)
        if self.gstrong is None:
            self.gstrong, self.gweak = torch.empty_like(strong), torch.empty_like(weak)
        self.keeps[slot] = (labels_weak, labels_strong, strong, weak, t_strong, t_weak)
        check(L.sedk_sed_loss_dev(ptr(strong), ptr(weak), ptr(t_strong), ptr(t_weak),
                                  ptr(labels_strong.contiguous()) if n_s > 0 else None,
                                  ptr(labels_weak) if n_w > 0 else None, B, self.C, strong.shape[2], n_s, n_w,
                                  ptr(self.cw), ptr(self.losses), ptr(self.gstrong), ptr(self.gweak), s),
              "sedk_sed_loss_dev")
        self.student.backward_direct(ws, self.gstrong, self.gweak)

    def _optimizer_part(self):
        ws = self.ws
        if self.world > 1:
            ddp.allreduce_sum_(ws.gflat, self.pg)
        if self.grad_clip and self.grad_clip > 0:
            # 2024 recipe: gradient_clip 5.0 (pretrained.yaml:17) - norm of the (averaged) gradient, on device
            check(lib().sedk_sumsq(ptr(ws.gflat), ws.gflat.numel(), ptr(self.sumsq), stream_ptr()), "sedk_sumsq")
            norm = torch.sqrt(self.sumsq.float()) * self.hyper[3]
            self.hyper[3:4] = self.hyper[3:4] * torch.clamp(self.grad_clip / (norm + 1e-6), max=1.0)
        self.opt.step_flat(ws.gflat, ema_flat=self.t_flat, hyper_dev=self.hyper)

    # ------------------------------------------------------------------------------------------------------------
    def _host_scalars(self):
        """Per-step host decisions, in the reference's RNG order; returns (do_mix, ring slot)."""
        r = self.step_idx % self.ring
        if self.ring_ev[r] is not None:
            self.ring_ev[r].synchronize()
        n_s, n_w, B = self.n_s, self.n_w, self.B
        hp, hc, hs = self.host_perm[r], self.host_coef[r], self.host_scal[r]
        do_mix = False
        hp[:B] = torch.arange(B)
        hc[:B] = 1.0
        hp[B:B + n_s] = torch.arange(n_s)
        hp[B + n_s:B + n_s + n_w] = torch.arange(n_w)
        hc[B:] = 1.0
        if self.mixup_type is not None and 0.5 > random.random():
            do_mix = True
            if n_w > 0:
                c = float(np.random.beta(0.2, 0.2))
                pw = torch.randperm(n_w)
                hp[n_s:n_s + n_w] = pw + n_s
                hc[n_s:n_s + n_w] = c
                hp[B + n_s:B + n_s + n_w] = pw
                hc[B + n_s:B + n_s + n_w] = c
            if n_s > 0:
                c = float(np.random.beta(0.2, 0.2))
                ps = torch.randperm(n_s)
                hp[:n_s] = ps
                hc[:n_s] = c
                hp[B:B + n_s] = ps
                hc[B:B + n_s] = c
        step_num = self.scheduler.step_num if self.scheduler is not None else self.opt.step_count + 1
        scale = self.scheduler._get_scaling_factor() if self.scheduler is not None else 1.0
        a = ema_alpha(self.ema_factor, step_num) if self.teacher is not None else 0.0
        self.opt.step_count += 1
        h = self.opt.hyper(self.opt.step_count, a, 1.0 / self.world)
        hs[0], hs[1], hs[2], hs[3] = h
        hs[4] = self.const_max * scale if self.teacher is not None else 0.0
        return do_mix, r

    def step(self, audio_host, labels_host, emb_host=None, inputs_ready=False):
        """One optimisation step.  audio [B, L] / labels [B, C, T'] fp32: pinned host tensors (copied on a copy stream) or
        device tensors.  The front end runs on its own stream into ping-pong buffers, so the log-mel of step k+1 overlaps
        the forward/backward graph of step k.  For DEVICE inputs the front end waits for the caller's current stream unless
        `inputs_ready=True` (the caller guarantees the batch was complete before this call)."""
        dev, B = self.dev, self.B
        k = self.step_idx
        slot = k % 2
        cur = torch.cuda.current_stream(dev)
        resident = audio_host.is_cuda
        fe = self.fe_stream
        if not resident:
            # ---- H2D of this step's inputs on the copy stream (overlaps the previous step's compute)
            with torch.cuda.stream(self.copy_stream):
                if self.slot_ev[slot] is not None:
                    self.copy_stream.wait_event(self.slot_ev[slot])
                self.audio_dev[slot].copy_(audio_host, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.copy_stream)
            fe.wait_event(ev_in)
        elif not inputs_ready:
            fe.wait_stream(cur)
        if self.labels_dev is None:
            self.labels_dev = torch.empty(labels_host.shape, device=dev)
            fe.wait_stream(cur)                                   # first step: allocations / table uploads on `cur`
        do_mix, r = self._host_scalars()
        mixing_graph = self.mixup_type is not None
        audio_in = audio_host if resident else self.audio_dev[slot]
        # ---- front end on its own stream (eager: reads the double-buffered audio slot, writes ping-pong buffer `slot`)
        if self.buf_free_ev[slot] is not None:
            fe.wait_event(self.buf_free_ev[slot])                 # the graph that last read this buffer has finished
        with torch.cuda.stream(fe):
            self.mel_buf, self.minmax = self.mel_bufs[slot], self.minmaxs[slot]
            if mixing_graph:
                self.mel_spec_run(audio_in, log=False)
            else:
                check(lib().sedk_minmax_init(ptr(self.minmax), B, stream_ptr()), "sedk_minmax_init")
                self.mel_spec_run(audio_in, log=True)
            ev_fe = torch.cuda.Event()
            ev_fe.record(fe)
        if not resident:
            self.slot_ev[slot] = ev_fe
        # ---- per-step device scalars and labels (current stream)
        self.labels_dev.copy_(labels_host, non_blocking=True)
        if emb_host is not None:
            self.emb_dev.copy_(emb_host, non_blocking=True)
        self.hyper.copy_(self.host_scal[r][:4], non_blocking=True)
        self.cw.copy_(self.host_scal[r][4:5], non_blocking=True)
        if mixing_graph:
            n_s, n_w = self.n_s, self.n_w
            self.perm.copy_(self.host_perm[r][:B], non_blocking=True)
            self.coef.copy_(self.host_coef[r][:B], non_blocking=True)
            if n_s > 0:
                self.perm_s.copy_(self.host_perm[r][B:B + n_s], non_blocking=True)
                self.coef_s.copy_(self.host_coef[r][B:B + n_s], non_blocking=True)
            if n_w > 0:
                self.perm_w.copy_(self.host_perm[r][B + n_s:B + n_s + n_w], non_blocking=True)
                self.coef_w.copy_(self.host_coef[r][B + n_s:B + n_s + n_w], non_blocking=True)
        cur.wait_event(ev_fe)
        # ---- forward / loss / backward
        if not self.use_graph:
            self._device_part(mixing_graph, slot)
        elif self.graphs[slot] is None:
            # warm-up (loads kernels, opts into shared memory, allocates workspaces), then capture
            snap = self._snapshot()
            self._device_part(mixing_graph, slot)
            torch.cuda.synchronize(dev)
            self._restore(snap)
            g = torch.cuda.CUDAGraph()
            n0 = lib().sedk_launch_count()
            # thread_local: other threads (e.g. the NCCL watchdog polling events) must not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._device_part(mixing_graph, slot)
            self.graph_kernels = int(lib().sedk_launch_count() - n0)
            self.graphs[slot] = g
            self.graph = g
            self._restore(snap)
            g.replay()
            self.replays += 1
        else:
            self.graphs[slot].replay()
            self.replays += 1
        ev_free = torch.cuda.Event()
        ev_free.record(cur)
        self.buf_free_ev[slot] = ev_free
        self._optimizer_part()
        self.host_loss[r].copy_(self.losses, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.ring_ev[r] = ev
        if self.scheduler is not None:
            self.scheduler.step()
        self.step_idx += 1
        return r

    def mel_spec_run(self, audio, log):
        tab = self.mel_spec.tables(audio.device)
        out = self.mel_buf
        check(lib().sedk_logmel_fwd(ptr(audio), self.B, self.L, tab.struct, ptr(out), out.stride(0), out.stride(1),
                                    out.stride(2), 1 if log else 0, 1e-5, -50.0, 80.0,
                                    ptr(self.minmax) if log else None, stream_ptr()), "sedk_logmel_fwd")

    def _snapshot(self):
        """BN running statistics + the seed counter are mutated by a forward: the warm-up / capture passes must not
        count as training steps."""
        mods = [self.student] + ([self.teacher] if self.teacher is not None else [])
        bufs = [b for m in mods for b in m.buffers()]
        return [(b, b.clone()) for b in bufs] + [(self.seed_ctr, self.seed_ctr.clone())]

    def _restore(self, snap):
        for b, c in snap:
            b.copy_(c)

    def read_losses(self, r=None):
        """Losses of ring slot r (default: last step): {total, bce_strong, bce_weak, mse_strong, mse_weak, ...}."""
        if r is None:
            r = (self.step_idx - 1) % self.ring
        self.ring_ev[r].synchronize()
        h = self.host_loss[r]
        keys = ["total", "bce_strong", "bce_weak", "mse_strong", "mse_weak", "bce_strong_teacher", "bce_weak_teacher",
                "cons_weight"]
        return {k: float(h[i]) for i, k in enumerate(keys)}
