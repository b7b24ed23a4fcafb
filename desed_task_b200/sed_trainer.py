"""SEDTask4 - the mean-teacher training module of the DCASE Task 4 recipes on the sm_100a hot path.

Host-side mirror of recipes/dcase2023_task4_baseline/local/sed_trainer.py:24-365 for the hot-path methods: same
constructor signature, same attributes (`mel_spec`, `scaler`, `sed_student`, `sed_teacher`, `supervised_loss`, ...), same
`take_log`, `detect`, `update_ema`, `training_step`, `on_before_zero_grad`, `lr_scheduler_step`, `configure_optimizers`.
It subclasses pytorch_lightning.LightningModule when Lightning is installed and a minimal stand-in otherwise (this image
has no Lightning).  Metrics / pandas / PSDS plumbing of the reference module (validation/test epoch ends) is out of scope.

Two ways to train:
  * `training_step(batch, idx)` - the reference's composition, autograd-connected (works under a Lightning Trainer or a
    plain `loss.backward(); opt.step()` loop);
  * `fit_step(batch)` - the fused engine (desed_task_b200/engine.py): one CUDA graph for forward/loss/backward plus one
    fused EMA+Adam kernel; what bench.py measures.
"""
import random
from copy import deepcopy

import torch

from . import _lib
from ._lib import check, lib, ptr, require_cuda, stream_ptr
from .data_augm import mixup
from .frontend import MelSpectrogram, new_minmax, take_log as _take_log
from .optim import FusedAdam, update_ema as _update_ema
from .utils.postprocess import median_filter
from .utils.scaler import TorchScaler

try:                                            # pragma: no cover - Lightning is absent in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
except Exception:                               # noqa: BLE001
    class _Base(torch.nn.Module):
        """Just enough of LightningModule for the hot-path methods: `hparams` dict and a `log` sink."""

        def __init__(self):
            super().__init__()
            self.hparams = {}
            self.logged = {}

        def log(self, name, value, **kwargs):
            self.logged[name] = value


class _SedLoss(torch.autograd.Function):
    """BCE(strong rows) + BCE(weak rows) + w * (MSE strong + MSE weak) in one kernel pair (sedk_sed_loss)."""

    @staticmethod
    def forward(ctx, strong, weak, t_strong, t_weak, labels, labels_weak, n_strong, n_weak, cons_weight, cons_row0,
                cons_kind):
        B, C, T = strong.shape
        losses = torch.zeros(16, device=strong.device)
        gs, gw = torch.empty_like(strong), torch.empty_like(weak)
        check(lib().sedk_sed_loss_ex(ptr(strong.contiguous()), ptr(weak.contiguous()), ptr(t_strong), ptr(t_weak),
                                     ptr(labels), ptr(labels_weak), B, C, T, n_strong, n_weak, int(cons_row0),
                                     int(cons_kind), float(cons_weight), None, ptr(losses), ptr(gs), ptr(gw),
                                     stream_ptr()), "sedk_sed_loss_ex")
        ctx.save_for_backward(gs, gw)
        return losses[0], losses[:8]

    @staticmethod
    def backward(ctx, g_total, g_parts):
        gs, gw = ctx.saved_tensors
        return gs * g_total, gw * g_total, None, None, None, None, None, None, None, None, None


def sed_loss(strong, weak, labels_strong, labels_weak, t_strong=None, t_weak=None, cons_weight=0.0, cons_row0=0,
             self_sup_loss="mse"):
    """rows [0, n_strong) of `strong` against labels_strong [n_strong,C,T]; rows [n_strong, n_strong+n_weak) of `weak`
    against labels_weak [n_weak,C]; consistency (MSELoss or BCELoss, sed_trainer.py:96-100) against the teacher on rows
    [cons_row0, B).  Returns (total, parts[8])."""
    n_s = 0 if labels_strong is None else labels_strong.shape[0]
    n_w = 0 if labels_weak is None else labels_weak.shape[0]
    ls = labels_strong.float().contiguous() if n_s else None
    lw = labels_weak.float().contiguous() if n_w else None
    ts = t_strong.detach().float().contiguous() if t_strong is not None else None
    tw = t_weak.detach().float().contiguous() if t_weak is not None else None
    return _SedLoss.apply(strong, weak, ts, tw, ls, lw, n_s, n_w, cons_weight, cons_row0,
                          1 if self_sup_loss == "bce" else 0)


class SEDTask4(_Base):
    def __init__(self, hparams, encoder, sed_student, opt=None, train_data=None, valid_data=None, test_data=None,
                 train_sampler=None, scheduler=None, fast_dev_run=False, evaluation=False, sed_teacher=None):
        super(SEDTask4, self).__init__()
        self.hparams.update(hparams)
        self.encoder = encoder
        self.sed_student = sed_student
        if sed_teacher is None:
            self.sed_teacher = deepcopy(sed_student)
        else:
            self.sed_teacher = sed_teacher
        self.opt = opt
        self.train_data, self.valid_data, self.test_data = train_data, valid_data, test_data
        self.train_sampler = train_sampler
        self.scheduler = scheduler
        self.fast_dev_run = fast_dev_run
        self.evaluation = evaluation
        self.num_workers = 1 if fast_dev_run else self.hparams["training"].get("num_workers", 1)
        feat_params = self.hparams["feats"]
        self.mel_spec = MelSpectrogram(
            sample_rate=feat_params["sample_rate"], n_fft=feat_params["n_window"], win_length=feat_params["n_window"],
            hop_length=feat_params["hop_length"], f_min=feat_params["f_min"], f_max=feat_params["f_max"],
            n_mels=feat_params["n_mels"], window_fn=torch.hamming_window, wkwargs={"periodic": False}, power=1)
        for param in self.sed_teacher.parameters():
            param.detach_()
        self.supervised_loss = torch.nn.BCELoss()
        if hparams["training"]["self_sup_loss"] == "mse":
            self.selfsup_loss = torch.nn.MSELoss()
        elif hparams["training"]["self_sup_loss"] == "bce":
            self.selfsup_loss = torch.nn.BCELoss()
        else:
            raise NotImplementedError
        self.scaler = self._init_scaler()
        self._engine = None

    # ---- scaler / features --------------------------------------------------------------------------------------
    def _init_scaler(self):
        """sed_trainer.py:201-251 (the dataset-statistics branch needs a dataloader and is fitted by the caller)."""
        sc = self.hparams["scaler"]
        if sc["statistic"] == "instance":
            return TorchScaler("instance", sc["normtype"], sc["dims"])
        elif sc["statistic"] == "dataset":
            return TorchScaler("dataset", sc["normtype"], sc["dims"])
        raise NotImplementedError

    def take_log(self, mels):
        """sed_trainer.py:253-264: AmplitudeToDB('amplitude', amin=1e-5) then clamp(-50, 80)."""
        return _take_log(mels, amin=1e-5, db_range=(-50.0, 80.0))

    def _fusable_scaler(self):
        s = self.scaler
        return s.statistic == "instance" and s.normtype == "minmax" and tuple(s.dims) == (1, 2)

    def detect(self, mel_feats, model, **kwargs):
        """sed_trainer.py:266-267: model(scaler(take_log(mel))).  With the shipped instance/minmax scaler the log, the
        per-clip min/max and the scaling are fused (min/max in the log kernel, the affine map in the first conv load)."""
        if self._fusable_scaler() and hasattr(model, "run"):
            mm = new_minmax(mel_feats.shape[0], mel_feats.device)
            logmel = _take_log(mel_feats, amin=1e-5, db_range=(-50.0, 80.0), minmax=mm)
            return model.run(logmel, minmax=mm, **kwargs)
        return model(self.scaler(self.take_log(mel_feats)), **kwargs)

    # ---- mean teacher -------------------------------------------------------------------------------------------
    def update_ema(self, alpha, global_step, model, ema_model):
        """sed_trainer.py:187-199 (one fused kernel over flat parameter buffers)."""
        _update_ema(alpha, global_step, model, ema_model)

    def lr_scheduler_step(self, scheduler, optimizer_idx, metric):
        scheduler.step()

    def configure_optimizers(self):
        return [self.opt], [self.scheduler]

    def on_before_zero_grad(self, *args, **kwargs):
        self.update_ema(self.hparams["training"]["ema_factor"], self.scheduler["scheduler"].step_num, self.sed_student,
                        self.sed_teacher)

    def training_step(self, batch, batch_indx):
        """sed_trainer.py:269-356, same order of operations and RNG consumption; returns the autograd-connected loss."""
        audio, labels = batch[0], batch[1]
        require_cuda(audio, labels)
        indx_synth, indx_weak, indx_unlabelled = self.hparams["training"]["batch_size"]
        features = self.mel_spec(audio)
        batch_num = features.shape[0]
        strong_mask = torch.zeros(batch_num).to(features).bool()
        weak_mask = torch.zeros(batch_num).to(features).bool()
        strong_mask[:indx_synth] = 1
        weak_mask[indx_synth: indx_weak + indx_synth] = 1
        labels_weak = (torch.sum(labels[weak_mask], -1) > 0).float()
        mixup_type = self.hparams["training"].get("mixup")
        if mixup_type is not None and 0.5 > random.random():
            features[weak_mask], labels_weak = mixup(features[weak_mask], labels_weak, mixup_label_type=mixup_type)
            features[strong_mask], labels[strong_mask] = mixup(features[strong_mask], labels[strong_mask],
                                                               mixup_label_type=mixup_type)
        strong_preds_student, weak_preds_student = self.detect(features, self.sed_student)
        with torch.no_grad():
            strong_preds_teacher, weak_preds_teacher = self.detect(features, self.sed_teacher)
        weight = self.hparams["training"]["const_max"] * self.scheduler["scheduler"]._get_scaling_factor()
        tot_loss, parts = sed_loss(strong_preds_student, weak_preds_student, labels[:indx_synth], labels_weak,
                                   strong_preds_teacher, weak_preds_teacher, weight,
                                   self_sup_loss=self.hparams["training"]["self_sup_loss"])
        self.log("train/student/loss_strong", parts[1])
        self.log("train/student/loss_weak", parts[2])
        self.log("train/teacher/loss_strong", parts[5])
        self.log("train/teacher/loss_weak", parts[6])
        self.log("train/step", self.scheduler["scheduler"].step_num, prog_bar=True)
        self.log("train/student/tot_self_loss", (parts[3] + parts[4]) * weight, prog_bar=True)
        self.log("train/weight", weight)
        self.log("train/student/tot_supervised", parts[3], prog_bar=True)
        self.log("train/student/weak_self_sup_loss", parts[4])
        self.log("train/student/strong_self_sup_loss", parts[3])
        if self.opt is not None:
            self.log("train/lr", self.opt.param_groups[-1]["lr"], prog_bar=True)
        return tot_loss

    # ---- fused engine -------------------------------------------------------------------------------------------
    def engine(self, n_samples, use_graph=True, process_group=None, teacher=True):
        if self._engine is None:
            from .engine import TrainEngine
            tr = self.hparams["training"]
            sched = self.scheduler["scheduler"] if isinstance(self.scheduler, dict) else self.scheduler
            opt = self.opt if isinstance(self.opt, FusedAdam) else FusedAdam(
                self.sed_student, self.hparams["opt"]["lr"] if "opt" in self.hparams else 1e-3)
            if opt is not self.opt and sched is not None:
                sched.optimizer = opt
            self.opt = opt
            self._engine = TrainEngine(
                self.sed_student, self.mel_spec, tr["batch_size"], n_samples, opt=opt, scheduler=sched,
                teacher=self.sed_teacher if teacher else None, ema_factor=tr.get("ema_factor", 0.999),
                const_max=tr.get("const_max", 2.0), mixup_type=tr.get("mixup"), use_graph=use_graph,
                process_group=process_group, grad_clip=tr.get("gradient_clip", 0.0) or 0.0,
                recipe="2024" if len(tr["batch_size"]) == 5 else "2023", mixup_prob=tr.get("mixup_prob", 0.5),
                self_sup_loss=tr.get("self_sup_loss", "mse"))
        return self._engine

    def fit_step(self, batch, use_graph=True, process_group=None):
        """One fused optimisation step on a HOST batch (pinned audio [B, L] fp32, labels [B, C, T'] fp32)."""
        audio, labels = batch[0], batch[1]
        eng = self.engine(audio.shape[-1], use_graph=use_graph, process_group=process_group)
        return eng.step(audio, labels)

    # ---- inference ----------------------------------------------------------------------------------------------
    @torch.no_grad()
    def predict(self, audio, teacher=False, median_window=None):
        """validation_step / test_step hot part (sed_trainer.py:367-390,608-640): mel -> detect -> median filter."""
        model = self.sed_teacher if teacher else self.sed_student
        mel = self.mel_spec(audio)
        strong, weak = self.detect(mel, model)
        if median_window is None:
            median_window = self.hparams["training"].get("median_window", 7)
        return strong, weak, median_filter(strong, median_window, class_dim=1)
