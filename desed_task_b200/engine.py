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
import os
import random

import numpy as np
import torch

from . import data_augm, ddp
from ._lib import check, lib, ptr, stream_ptr
from .frontend import new_minmax
from .optim import FusedAdam, ema_alpha, flatten_parameters


class TrainEngine:
    """recipe="2023": sub-batches [n_strong, n_weak, n_unlabelled] (sed_trainer.py:286-289).
    recipe="2024": sub-batches [maestro, synth, strong, weak, unlabelled] (sed_trainer_pretrained.py:335-346): strong rows
    [0, indx_strong), weak rows [indx_strong, indx_weak), consistency on rows [indx_maestro, B); mixup inside the three label
    groups with independent (c, perm) for features and embeddings, labels mixed by both (:282-301,350-363); weak labels
    derived from the mixed labels and everything masked by the per-row `class_masks` (:365-370)."""

    def __init__(self, student, mel_spec, batch_sizes, n_samples, opt=None, scheduler=None, teacher=None,
                 ema_factor=0.999, const_max=2.0, mixup_type=None, use_graph=True, process_group=None,
                 grad_clip=0.0, emb_shape=None, class_masks=None, distributed=True, recipe="2023", mixup_prob=0.5,
                 self_sup_loss="mse", graph_optimizer=True, audio_dtype=torch.float32, emb_dtype=torch.float32):
        self.student, self.teacher, self.mel_spec = student, teacher, mel_spec
        if audio_dtype not in (torch.float32, torch.int16):
            raise ValueError("audio_dtype must be torch.float32 or torch.int16 (16-bit PCM, used as x / 32768)")
        self.batch_sizes = list(batch_sizes)
        self.recipe = str(recipe)
        if self.recipe == "2024":
            if len(self.batch_sizes) != 5:
                raise ValueError("recipe='2024' needs batch_sizes [maestro, synth, strong, weak, unlabelled]")
            cs = np.cumsum(self.batch_sizes)
            self.n_s, self.n_w, self.cons_row0 = int(cs[2]), int(cs[3] - cs[2]), int(cs[0])
            # mixup groups in the reference's call order: weak, synth + strong, maestro
            self.groups = [(int(cs[2]), int(cs[3])), (int(cs[0]), int(cs[2])), (0, int(cs[0]))]
        else:
            self.n_s, self.n_w, self.cons_row0 = self.batch_sizes[0], self.batch_sizes[1], 0
            self.groups = None
        self.B = int(sum(self.batch_sizes))
        self.L = n_samples
        self.dev = next(student.parameters()).device
        self.opt = opt if opt is not None else FusedAdam(student, 1e-3)
        if not isinstance(self.opt, FusedAdam):
            raise TypeError("TrainEngine drives desed_task_b200.optim.FusedAdam (same maths as torch.optim.Adam)")
        self.scheduler = scheduler
        self.ema_factor, self.const_max = ema_factor, const_max
        self.mixup_type, self.mixup_prob = mixup_type, mixup_prob
        if self_sup_loss not in ("mse", "bce"):
            raise NotImplementedError("self_sup_loss=%r" % (self_sup_loss,))
        self.cons_kind = 1 if self_sup_loss == "bce" else 0
        self.use_const_weight = False           # 2024: current_epoch >= epoch_decay (sed_trainer_pretrained.py:402-405)
        self.use_graph = use_graph
        self.pg = process_group
        self.world = 1
        if distributed and (process_group is not None or
                            (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        # the gradient all-reduce and the fused EMA + Adam kernel are captured in the same CUDA graph as forward / backward
        self.graph_optimizer = bool(graph_optimizer) and use_graph
        # data-parallel all-reduce schedule (env SEDK_AR_MODE).  "eager": the graph ends with the backward, one NCCL
        # all-reduce of the flat gradient and the fused EMA + Adam kernel follow as eager launches (issued while the graph
        # still runs).  "split": three slices reduced inside the graph underneath the rest of the backward
        # (sedk_crnn_backward_phase); "single": one in-graph all-reduce after the backward.  Measured on 2 x B200
        # (profiles/r2_allreduce_modes.txt): eager +24 us per step over N = 1, split +42 us, single +48 us - NCCL kernels
        # captured into the graph cost more than they hide on a 2.2 ms step, so the overlapped schedules stay optional.
        # "nvls" (default): no NCCL on the step - the flat gradient lives in NVLink-symmetric memory and ONE kernel of this library
        # (csrc/nvls.cu) reduces it through the switch (multimem.ld_reduce / multimem.st) and applies EMA + Adam, inside the
        # same CUDA graph as forward / backward, as at N = 1.  Falls back to "eager" where symmetric memory is unavailable.
        import os
        self.ar_mode = os.environ.get("SEDK_AR_MODE", "nvls")
        if self.ar_mode not in ("eager", "split", "single", "nvls"):
            raise ValueError("SEDK_AR_MODE must be nvls, eager, split or single")
        self.nvls = None
        if self.ar_mode == "nvls" and self.world <= 1:
            self.ar_mode = "eager"
        if self.world > 1 and self.ar_mode == "eager":
            self.graph_optimizer = False
        self.grad_clip = grad_clip
        self.emb_shape = emb_shape
        self.class_masks = class_masks
        self.cm_float = class_masks.float() if class_masks is not None else None
        dev, B = self.dev, self.B
        self.C = student.nclass
        self.T = mel_spec.n_frames(n_samples)
        self.audio_dev = [torch.empty(B, n_samples, device=dev, dtype=audio_dtype) for _ in range(2)]
        # the front end runs on its own stream into ping-pong buffers, so that step k+1's log-mel overlaps step k's graph
        self.mel_bufs = [torch.empty(B, mel_spec.n_mels, self.T, device=dev) for _ in range(2)]
        self.mel_buf = self.mel_bufs[0]
        self.logmel = torch.empty_like(self.mel_buf)
        self.labels_dev = None
        # embeddings are the largest per-step input (1.52 MB per clip): double-buffered like the audio, copied on the copy stream
        self.emb_devs = [torch.empty(B, *emb_shape, device=dev) for _ in range(2)] if emb_shape else None
        # bf16 storage format (desed_task_b200.embeddings.pool_embeddings): staged as bf16, upcast on the copy stream
        self.emb_stage = [torch.empty(B, *emb_shape, device=dev, dtype=torch.bfloat16) for _ in range(2)] \
            if (emb_shape and emb_dtype == torch.bfloat16) else None
        self.emb_dev = self.emb_devs[0] if emb_shape else None
        self.emb_mixed = torch.empty_like(self.emb_dev) if (emb_shape and self.recipe == "2024" and mixup_type) else None
        self.minmaxs = [torch.empty(B, 2, dtype=torch.int32, device=dev) for _ in range(2)]
        self.minmax = self.minmaxs[0]
        # ---- every per-step scalar lives in ONE device block that is refreshed by ONE pinned-host copy per step:
        #      int64 perm[NP] | fp32 coef[NP] | fp32 hyper[4] | fp32 cw[1] | pad
        #      2023: perm/coef [0,B) features, [B, B+n_s) strong labels, [B+n_s, B+n_s+n_w) weak labels
        #      2024: perm/coef [0,B) features (+ first label mix), [B, 2B) embeddings (+ second label mix)
        NP = self.NP = 2 * B + 2
        nbytes = 8 * NP + 4 * NP + 4 * 8
        self.scal_dev = torch.zeros(nbytes, dtype=torch.uint8, device=dev)
        self.perm_all = self.scal_dev[:8 * NP].view(torch.int64)
        fl = self.scal_dev[8 * NP:].view(torch.float32)
        self.coef_all, self.hyper, self.cw = fl[:NP], fl[NP:NP + 4], fl[NP + 4:NP + 5]
        n_s, n_w = self.n_s, self.n_w
        self.perm, self.coef = self.perm_all[:B], self.coef_all[:B]
        if self.recipe == "2024":
            self.perm_e, self.coef_e = self.perm_all[B:2 * B], self.coef_all[B:2 * B]
        else:
            self.perm_s, self.coef_s = self.perm_all[B:B + max(n_s, 1)], self.coef_all[B:B + max(n_s, 1)]
            self.perm_w = self.perm_all[B + n_s:B + n_s + max(n_w, 1)]
            self.coef_w = self.coef_all[B + n_s:B + n_s + max(n_w, 1)]
        self.seed_ctr = torch.zeros(1, dtype=torch.int64, device=dev)
        self.losses = torch.zeros(16, device=dev)
        self.gstrong = self.gweak = None
        self.sumsq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.ring = 4
        self.host_scal = [torch.zeros(nbytes, dtype=torch.uint8).pin_memory() for _ in range(self.ring)]
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
        if self.world > 1 and self.ar_mode == "nvls":
            self._setup_nvls(process_group)

    def _setup_nvls(self, process_group):
        """Collective: allocate the student's gradient region in NVLink-symmetric memory, exchange the handles, agree over the
        group that every rank succeeded - otherwise every rank falls back to the NCCL schedule ("eager")."""
        from . import nvls
        student, dev = self.student, self.dev
        obj = nvls.try_create(process_group, dev)
        ok, why = obj is not None, ""
        if ok:
            try:
                student.grad_alloc = obj.alloc
                ws = student._workspace(self.B, self.mel_spec.n_mels, self.T, dev,
                                        tuple(self.emb_shape) if self.emb_shape else None)
                if ws.zero_bwd.data_ptr() != obj.buf.data_ptr():
                    raise RuntimeError("the student's gradient region is not the symmetric allocation")
                # measured (profiles/r2_nvls_check.txt): with two ranks plain peer loads / stores beat the multicast round
                # trip through the switch (29.8 vs 31.8 us); at eight the switch reduction wins (37.0 vs 44.0 us)
                mc = os.environ.get("SEDK_NVLS_MULTICAST", "auto")
                obj.connect(use_multicast=(self.world > 2) if mc == "auto" else mc != "0")
            except Exception as e:      # noqa: BLE001 - any failure means "use the NCCL path"
                ok, why = False, "%s: %s" % (type(e).__name__, e)
        flag = torch.tensor([1 if ok else 0], device=dev)
        torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN, group=process_group)
        if int(flag.item()) == 1:
            self.nvls = obj
            return
        if why:
            import sys
            print("[desed_task_b200] fused NVLink all-reduce unavailable (%s); using NCCL" % why, file=sys.stderr)
        self.nvls, self.ar_mode, self.graph_optimizer = None, "eager", False
        student.grad_alloc = None

    def _host_views(self, r):
        NP = self.NP
        h = self.host_scal[r]
        perm = h[:8 * NP].view(torch.int64)
        fl = h[8 * NP:].view(torch.float32)
        return perm, fl[:NP], fl[NP:NP + 5]

    # ------------------------------------------------------------------------------------------------------------
    def _labels_2023(self, do_mix):
        n_s, n_w = self.n_s, self.n_w
        labels = self.labels_dev
        labels_weak = (torch.sum(labels[n_s:n_s + n_w], -1) > 0).float() if n_w > 0 else None
        labels_strong = labels[:n_s]
        if do_mix:
            hard = self.mixup_type == "hard"
            if n_w > 0:
                labels_weak = data_augm.mix_labels(labels_weak, self.perm_w, self.coef_w, hard)
            if n_s > 0:
                labels_strong = data_augm.mix_labels(labels_strong.contiguous(), self.perm_s, self.coef_s, hard)
        return labels_strong, labels_weak

    def _labels_2024(self, do_mix):
        n_s, n_w = self.n_s, self.n_w
        labels = self.labels_dev
        if do_mix:
            hard = self.mixup_type == "hard"
            labels = data_augm.mix_labels(labels, self.perm, self.coef, hard)           # with the features' (c, perm)
            if self.emb_dev is not None:
                labels = data_augm.mix_labels(labels, self.perm_e, self.coef_e, hard)   # again with the embeddings'
        labels_weak = (torch.sum(labels[n_s:n_s + n_w], -1) > 0).float() if n_w > 0 else None
        if self.cm_float is not None:
            labels = labels * self.cm_float[:, :, None]
            if n_w > 0:
                labels_weak = labels_weak * self.cm_float[n_s:n_s + n_w]
        return labels[:n_s], labels_weak

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
        emb = self.emb_dev = self.emb_devs[slot] if self.emb_devs is not None else None
        if do_mix:
            # mixup on the LINEAR mel (sed_trainer.py:296-301), fused with take_log + per-clip min/max
            check(L.sedk_feat_mix_log(ptr(self.mel_buf), ptr(self.perm), ptr(self.coef), ptr(self.logmel), B, n, 1, 1e-5,
                                      -50.0, 80.0, ptr(self.minmax), s), "sedk_feat_mix_log")
            feats = self.logmel
            if self.recipe == "2024" and emb is not None:
                # the embeddings are mixed with their own (c, perm) (sed_trainer_pretrained.py:293-299)
                check(L.sedk_feat_mix_log(ptr(emb), ptr(self.perm_e), ptr(self.coef_e), ptr(self.emb_mixed), B,
                                          emb[0].numel(), 0, 0.0, 0.0, 0.0, None, s), "sedk_feat_mix_log(emb)")
                emb = self.emb_mixed
        else:
            feats = self.mel_buf        # the front-end kernel already wrote the log-mel and the min/max
        labels_strong, labels_weak = (self._labels_2024 if self.recipe == "2024" else self._labels_2023)(do_mix)
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
        labels_strong = labels_strong.contiguous() if n_s > 0 else None
        self.keeps[slot] = (labels_weak, labels_strong, strong, weak, t_strong, t_weak, emb, feats)
        check(L.sedk_sed_loss_ex(ptr(strong), ptr(weak), ptr(t_strong), ptr(t_weak), ptr(labels_strong),
                                 ptr(labels_weak) if n_w > 0 else None, B, self.C, strong.shape[2], n_s, n_w,
                                 self.cons_row0, self.cons_kind, 0.0, ptr(self.cw), ptr(self.losses), ptr(self.gstrong),
                                 ptr(self.gweak), s), "sedk_sed_loss_ex")
        if self.world > 1 and self.graph_optimizer and self.ar_mode == "split":
            # data parallel: the RNN / head half of the flat gradient is final after phase 1 - its all-reduce runs on a
            # side stream (a parallel branch of the graph) underneath the CNN backward; the CNN half follows at the end
            # (2 MB), then the 128-channel conv layers (98 % of the CNN parameters) under the backward of the three
            # HBM-bound bottom layers; only the last ~0.1 MB slice is reduced after the backward has finished
            n_cnn, n_low = self.student.cnn_param_count(), self.student.cnn_lower_param_count(3)
            cur = torch.cuda.current_stream(self.dev)

            def reduce_on_side(flat_slice):
                ev = torch.cuda.Event()
                ev.record(cur)
                self.teacher_stream.wait_event(ev)
                with torch.cuda.stream(self.teacher_stream):
                    ddp.allreduce_sum_(flat_slice, self.pg)
            self.student.backward_direct(ws, self.gstrong, self.gweak, phases=1)
            reduce_on_side(ws.gflat[n_cnn:])
            self.student.backward_direct(ws, self.gstrong, self.gweak, phases=4)
            reduce_on_side(ws.gflat[n_low:n_cnn])
            self.student.backward_direct(ws, self.gstrong, self.gweak, phases=8)
            ddp.allreduce_sum_(ws.gflat[:n_low], self.pg)
            ev_ar = torch.cuda.Event()
            ev_ar.record(self.teacher_stream)
            cur.wait_event(ev_ar)
            self._reduced = True
        else:
            self.student.backward_direct(ws, self.gstrong, self.gweak)
            self._reduced = False

    def _optimizer_part(self):
        ws = self.ws
        if self.nvls is not None:
            clip = bool(self.grad_clip and self.grad_clip > 0)
            # the whole flat gradient in one launch; with clipping the update follows the norm (all-reduce only here)
            self.nvls.step(self.opt, ws.gflat.numel(), self.t_flat, self.hyper, do_adam=not clip)
            if not clip:
                return
        elif self.world > 1 and not getattr(self, "_reduced", False):
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
        hp, hc, hs = self._host_views(r)
        do_mix = False
        hp[:B] = torch.arange(B)
        hc[:] = 1.0
        if self.recipe == "2024":
            hp[B:2 * B] = torch.arange(B)
            if self.mixup_type is not None and self.mixup_prob > random.random():
                do_mix = True
                for lo, hi in self.groups:
                    c = float(np.random.beta(0.2, 0.2))                  # features (data_augm.py:33-35)
                    pm = torch.randperm(hi - lo)
                    hp[lo:hi] = pm + lo
                    hc[lo:hi] = c
                    if self.emb_dev is not None:
                        c = float(np.random.beta(0.2, 0.2))              # embeddings: an independent draw
                        pm = torch.randperm(hi - lo)
                        hp[B + lo:B + hi] = pm + lo
                        hc[B + lo:B + hi] = c
        else:
            hp[B:B + n_s] = torch.arange(n_s)
            hp[B + n_s:B + n_s + n_w] = torch.arange(n_w)
            if self.mixup_type is not None and self.mixup_prob > random.random():
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
        if self.use_const_weight:
            scale = 1.0
        a = ema_alpha(self.ema_factor, step_num) if self.teacher is not None else 0.0
        self.opt.step_count += 1
        h = self.opt.hyper(self.opt.step_count, a, 1.0 / self.world)
        hs[0], hs[1], hs[2], hs[3] = h
        hs[4] = self.const_max * scale if self.teacher is not None else 0.0
        return do_mix, r

    def step(self, audio_host, labels_host, emb_host=None, inputs_ready=False):
        """One optimisation step.  audio [B, L] / labels [B, C, T'] fp32 (/ embeddings [B, E, Te]): pinned host tensors
        (copied on a copy stream) or device tensors.  The front end runs on its own stream into ping-pong buffers, so the
        log-mel of step k+1 overlaps the forward/backward graph of step k.  For DEVICE inputs the front end waits for the
        caller's current stream unless `inputs_ready=True` (the caller guarantees the batch was complete before this call)."""
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
        ev_emb = None
        if emb_host is not None:
            bf16 = emb_host.dtype == torch.bfloat16
            if bf16 and self.emb_stage is None:
                raise ValueError("bf16 embeddings need TrainEngine(emb_dtype=torch.bfloat16)")
            if emb_host.is_cuda:
                if bf16:
                    check(lib().sedk_bf16_to_f32(ptr(emb_host.contiguous()), ptr(self.emb_devs[slot]), emb_host.numel(),
                                                 stream_ptr()), "sedk_bf16_to_f32")
                else:
                    self.emb_devs[slot].copy_(emb_host, non_blocking=True)
            else:
                with torch.cuda.stream(self.copy_stream):
                    if self.buf_free_ev[slot] is not None:
                        self.copy_stream.wait_event(self.buf_free_ev[slot])     # the graph that last read this slot is done
                    if bf16:
                        self.emb_stage[slot].copy_(emb_host, non_blocking=True)
                        check(lib().sedk_bf16_to_f32(ptr(self.emb_stage[slot]), ptr(self.emb_devs[slot]),
                                                     emb_host.numel(), stream_ptr()), "sedk_bf16_to_f32")
                    else:
                        self.emb_devs[slot].copy_(emb_host, non_blocking=True)
                    ev_emb = torch.cuda.Event()
                    ev_emb.record(self.copy_stream)
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
        # ---- per-step device scalars (ONE copy of the pinned block) and labels / embeddings (current stream)
        self.labels_dev.copy_(labels_host, non_blocking=True)
        self.scal_dev.copy_(self.host_scal[r], non_blocking=True)
        cur.wait_event(ev_fe)
        if ev_emb is not None:
            cur.wait_event(ev_emb)
        # ---- forward / loss / backward (/ all-reduce / optimiser)
        in_graph_opt = self.graph_optimizer
        if not self.use_graph:
            self._device_part(mixing_graph, slot)
        elif self.graphs[slot] is None:
            # warm-up (loads kernels, opts into shared memory, allocates workspaces), then capture
            snap = self._snapshot()
            self._device_part(mixing_graph, slot)
            if in_graph_opt:
                self._optimizer_part()
            torch.cuda.synchronize(dev)
            self._restore(snap)
            g = torch.cuda.CUDAGraph()
            n0 = lib().sedk_launch_count()
            # thread_local: other threads (e.g. the NCCL watchdog polling events) must not invalidate the capture
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                self._device_part(mixing_graph, slot)
                if in_graph_opt:
                    self._optimizer_part()
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
        if not (self.use_graph and in_graph_opt):
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
        fn = lib().sedk_logmel_fwd_i16 if audio.dtype == torch.int16 else lib().sedk_logmel_fwd
        check(fn(ptr(audio), self.B, self.L, tab.struct, ptr(out), out.stride(0), out.stride(1), out.stride(2),
                 1 if log else 0, 1e-5, -50.0, 80.0, ptr(self.minmax) if log else None, stream_ptr()), "sedk_logmel_fwd")

    def _snapshot(self):
        """Everything a step mutates besides its own outputs - BN running statistics, the seed counter, and (when the
        optimiser is part of the graph) parameters, Adam moments, the teacher and the per-step scalar block: the warm-up /
        capture passes must not count as training steps."""
        mods = [self.student] + ([self.teacher] if self.teacher is not None else [])
        bufs = [b for m in mods for b in m.buffers()]
        extra = [self.seed_ctr, self.scal_dev]
        if self.graph_optimizer:
            self.opt._ensure()
            extra += [self.p_flat, self.opt.m, self.opt.v] + ([self.t_flat] if self.t_flat is not None else [])
        return [(b, b.clone()) for b in bufs + extra]

    def _restore(self, snap):
        for b, c in snap:
            b.copy_(c)

    def _check_nvls(self):
        # the query synchronises the device: every 64th read is enough to turn a lost peer into an error instead of garbage
        self._nvls_reads = getattr(self, "_nvls_reads", 0) + 1
        if self.nvls is not None and self._nvls_reads % 64 == 1 and lib().sedk_nvls_fault() != 0:
            raise RuntimeError("fused NVLink all-reduce: a peer did not reach the barrier within 10 s (rank lost?); the "
                               "optimizer state of this process is no longer valid")

    def read_losses(self, r=None):
        self._check_nvls()
        """Losses of ring slot r (default: last step): {total, bce_strong, bce_weak, mse_strong, mse_weak, ...}."""
        if r is None:
            r = (self.step_idx - 1) % self.ring
        self.ring_ev[r].synchronize()
        h = self.host_loss[r]
        keys = ["total", "bce_strong", "bce_weak", "mse_strong", "mse_weak", "bce_strong_teacher", "bce_weak_teacher",
                "cons_weight"]
        return {k: float(h[i]) for i, k in enumerate(keys)}


class InferEngine:
    """Inference hot path of validation_step / test_step (recipes/dcase2023_task4_baseline/local/sed_trainer.py:367-390,
    608-640 + local/utils.py:45-63): waveform -> log-mel (+ per-clip min/max) -> eval-mode CRNN (BatchNorm on running
    statistics, no dropout) -> per-class median filter, for fixed-size batches of clips.

    Pipelined over three streams like TrainEngine: H2D of batch k+1 (copy stream) and its log-mel (front-end stream, ping-pong
    buffers) overlap the forward graph of batch k; the filtered scores [B, C, T'] and the clip-level posteriors [B, C] are
    read back into a ring of pinned host buffers.  Clips shard over ranks with no collective (ddp.shard_clip_range)."""

    def __init__(self, model, mel_spec, batch, n_samples, median_window=7, use_graph=True, emb_shape=None, class_masks=None,
                 audio_dtype=torch.float32):
        from .utils.postprocess import median_filter
        self._median = median_filter
        self.model, self.mel_spec = model, mel_spec
        self.B, self.L = int(batch), int(n_samples)
        self.dev = dev = next(model.parameters()).device
        self.use_graph = use_graph
        self.T = mel_spec.n_frames(n_samples)
        self.C = model.nclass
        wins = [int(median_window)] * self.C if isinstance(median_window, int) else [int(v) for v in median_window]
        if len(wins) != self.C or min(wins) < 1 or max(wins) > 31:
            raise ValueError("InferEngine: need one median window in [1, 31] per class")
        self.win = torch.tensor(wins, dtype=torch.int32, device=dev)
        self.emb_shape, self.class_masks = emb_shape, class_masks
        B = self.B
        self.audio_dev = [torch.empty(B, n_samples, device=dev, dtype=audio_dtype) for _ in range(2)]
        self.mel_bufs = [torch.empty(B, mel_spec.n_mels, self.T, device=dev) for _ in range(2)]
        self.minmaxs = [torch.empty(B, 2, dtype=torch.int32, device=dev) for _ in range(2)]
        self.emb_dev = torch.empty(B, *emb_shape, device=dev) if emb_shape else None
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.fe_stream = torch.cuda.Stream(device=dev)
        self.graphs, self.outs = [None, None], [None, None]
        self.buf_free_ev, self.slot_ev = [None, None], [None, None]
        self.ring = 4
        self.host_scores, self.host_weak, self.ring_ev = None, None, [None] * self.ring
        self.graph_kernels = 0
        self.replays = 0
        self.step_idx = 0

    def _device_part(self, slot):
        strong, weak, _ = self.model.forward_direct(self.mel_bufs[slot], self.minmaxs[slot], self.emb_dev, self.class_masks)
        return strong, weak, self._median(strong, self.win, class_dim=1)

    def step(self, audio, emb=None, inputs_ready=False):
        """One batch: audio [B, L] fp32, pinned host (copied on the copy stream) or device.  Returns the ring slot whose
        pinned host buffers receive (median-filtered strong scores [B, C, T'], weak [B, C]); `read(r)` waits for them."""
        if self.model.training:
            raise RuntimeError("InferEngine needs model.eval() (BatchNorm on running statistics, no dropout)")
        dev, B = self.dev, self.B
        k = self.step_idx
        slot, r = k % 2, k % self.ring
        cur = torch.cuda.current_stream(dev)
        fe = self.fe_stream
        resident = audio.is_cuda
        if not resident:
            with torch.cuda.stream(self.copy_stream):
                if self.slot_ev[slot] is not None:
                    self.copy_stream.wait_event(self.slot_ev[slot])
                self.audio_dev[slot].copy_(audio, non_blocking=True)
                ev_in = torch.cuda.Event()
                ev_in.record(self.copy_stream)
            fe.wait_event(ev_in)
        elif not inputs_ready:
            fe.wait_stream(cur)
        if k == 0:
            fe.wait_stream(cur)
        if self.buf_free_ev[slot] is not None:
            fe.wait_event(self.buf_free_ev[slot])
        with torch.cuda.stream(fe):
            tab = self.mel_spec.tables(dev)
            out, mm = self.mel_bufs[slot], self.minmaxs[slot]
            check(lib().sedk_minmax_init(ptr(mm), B, stream_ptr()), "sedk_minmax_init")
            src = audio if resident else self.audio_dev[slot]
            fn = lib().sedk_logmel_fwd_i16 if src.dtype == torch.int16 else lib().sedk_logmel_fwd
            check(fn(ptr(src), B, self.L, tab.struct, ptr(out), out.stride(0), out.stride(1), out.stride(2), 1, 1e-5, -50.0,
                     80.0, ptr(mm), stream_ptr()), "sedk_logmel_fwd")
            ev_fe = torch.cuda.Event()
            ev_fe.record(fe)
        if not resident:
            self.slot_ev[slot] = ev_fe
        if emb is not None:
            self.emb_dev.copy_(emb, non_blocking=True)
        cur.wait_event(ev_fe)
        with torch.no_grad():
            if not self.use_graph:
                self.outs[slot] = self._device_part(slot)
            elif self.graphs[slot] is None:
                self._device_part(slot)                            # warm-up: kernels loaded, workspaces allocated
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                n0 = lib().sedk_launch_count()
                with torch.cuda.graph(g, capture_error_mode="thread_local"):
                    self.outs[slot] = self._device_part(slot)
                self.graph_kernels = int(lib().sedk_launch_count() - n0)
                self.graphs[slot] = g
                g.replay()
                self.replays += 1
            else:
                self.graphs[slot].replay()
                self.replays += 1
        ev_free = torch.cuda.Event()
        ev_free.record(cur)
        self.buf_free_ev[slot] = ev_free
        strong, weak, med = self.outs[slot]
        if self.host_scores is None:
            self.host_scores = [torch.empty(med.shape, dtype=torch.float32).pin_memory() for _ in range(self.ring)]
            self.host_weak = [torch.empty(weak.shape, dtype=torch.float32).pin_memory() for _ in range(self.ring)]
        if self.ring_ev[r] is not None:
            self.ring_ev[r].synchronize()
        self.host_scores[r].copy_(med, non_blocking=True)
        self.host_weak[r].copy_(weak, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.ring_ev[r] = ev
        self.step_idx += 1
        return r

    def read(self, r=None):
        if r is None:
            r = (self.step_idx - 1) % self.ring
        self.ring_ev[r].synchronize()
        return self.host_scores[r], self.host_weak[r]
