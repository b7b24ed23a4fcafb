"""Oracle: augmentation, losses, mean-teacher step, EMA, Adam, warm-up (CPU fp32).  Test infrastructure only.

Restates desed_task/data_augm.py:7-77, recipes/dcase2023_task4_baseline/local/sed_trainer.py:187-199
(update_ema), :269-356 (training_step), recipes/dcase2023_task4_baseline/train_sed.py:199-206 (Adam +
ExponentialWarmup) and desed_task/utils/schedulers.py:85-104.  Random draws are explicit arguments.
"""
import math
import random

import numpy as np
import torch
import torch.nn.functional as F

from . import crnn as ocrnn
from . import frontend as ofe


# ----------------------------------------------------------------------------- augmentation
def draw_mixup(batch_size, alpha=0.2, beta=0.2):
    """Consume RNG exactly like data_augm.py:33-35: np.random.beta then torch.randperm (CPU generator)."""
    c = float(np.random.beta(alpha, beta))
    perm = torch.randperm(batch_size)
    return c, perm


def mixup(data, target=None, c=None, perm=None, mixup_label_type="soft"):
    """data_augm.py:19-53 with (c, perm) injected."""
    with torch.no_grad():
        mixed = c * data + (1 - c) * data[perm, :]
        if target is None:
            return mixed
        if mixup_label_type == "soft":
            mt = torch.clamp(c * target + (1 - c) * target[perm, :], min=0, max=1)
        elif mixup_label_type == "hard":
            mt = torch.clamp(target + target[perm, :], min=0, max=1)
        else:
            raise NotImplementedError(mixup_label_type)
        return mixed, mt


def draw_frame_shift(bsz):
    """data_augm.py:12: one random.gauss(0, 90) per clip."""
    return [int(random.gauss(0, 90)) for _ in range(bsz)]


def frame_shift(mels, labels, shifts, net_pooling=4):
    """data_augm.py:7-16 with the shifts injected."""
    sm, sl = [], []
    for b, shift in enumerate(shifts):
        sm.append(torch.roll(mels[b], shift, dims=-1))
        ls = -abs(shift) // net_pooling if shift < 0 else shift // net_pooling
        sl.append(torch.roll(labels[b], ls, dims=-1))
    return torch.stack(sm), torch.stack(sl)


def add_noise(mels, snr_db, noise, dims=(1, 2)):
    """data_augm.py:56-77 with snr [B] (dB) and unit-normal noise injected."""
    snr = 10 ** (snr_db.reshape(-1, 1, 1) / 20)
    sigma = torch.std(mels, dim=dims, keepdim=True) / snr
    return mels + noise * sigma


# ----------------------------------------------------------------------------- schedule / optimiser
def warmup_scale(step_num, rampup_len, exponent=-5.0):
    """ExponentialWarmup._get_scaling_factor, schedulers.py:85-92 (no annealing)."""
    if rampup_len == 0:
        return 1.0
    current = float(np.clip(step_num, 0.0, rampup_len))
    phase = 1.0 - current / rampup_len
    return float(np.exp(exponent * phase * phase))


def update_ema(alpha, global_step, student, teacher, names):
    """sed_trainer.py:187-199: alpha=min(1-1/(step+1), alpha); ema = ema*alpha + (1-alpha)*p."""
    alpha = min(1 - 1 / (global_step + 1), alpha)
    for k in names:
        teacher[k].mul_(alpha).add_(student[k], alpha=1 - alpha)
    return alpha


def adam_step(params, grads, state, names, lr, betas=(0.9, 0.999), eps=1e-8, step=None):
    """torch.optim.Adam (no weight decay / amsgrad), single-tensor formula as in
    torch/optim/adam.py `_single_tensor_adam`: denom = sqrt(v)/sqrt(bc2) + eps; p -= lr/bc1 * m/denom."""
    b1, b2 = betas
    for k in names:
        st = state.setdefault(k, {"step": 0, "m": torch.zeros_like(params[k]), "v": torch.zeros_like(params[k])})
        st["step"] += 1
        t = st["step"]
        g = grads[k]
        st["m"].lerp_(g, 1 - b1)
        st["v"].mul_(b2).addcmul_(g, g, value=1 - b2)
        bc1 = 1 - b1 ** t
        bc2 = 1 - b2 ** t
        denom = (st["v"].sqrt() / math.sqrt(bc2)).add_(eps)
        params[k].addcdiv_(st["m"], denom, value=-lr / bc1)


# ----------------------------------------------------------------------------- losses
def bce(p, y):
    """torch.nn.BCELoss (mean; log clamped at -100)."""
    return F.binary_cross_entropy(p, y)


def supervised_losses(strong, weak, labels, n_strong, n_weak):
    """sed_trainer.py:286-292,309-315: rows [0,n_strong) strong-labelled, [n_strong, n_strong+n_weak) weak."""
    labels_weak = (torch.sum(labels[n_strong:n_strong + n_weak], -1) > 0).float()
    loss_strong = bce(strong[:n_strong], labels[:n_strong])
    loss_weak = bce(weak[n_strong:n_strong + n_weak], labels_weak)
    return loss_strong, loss_weak


# ----------------------------------------------------------------------------- whole steps
def detect(mel, P, cfg, training, **kw):
    """sed_trainer.py:266-267: model(scaler(take_log(mel)))."""
    return ocrnn.crnn_forward(P, ofe.scaler(ofe.take_log(mel)), cfg, training, **kw)


def supervised_step(P, audio, labels, n_strong, n_weak, cfg=ocrnn.CFG_2023, training=True, fwd_kw=None,
                    gru_impl="aten"):
    """BASELINE config 2: mel -> log -> scaler -> student fwd -> BCE strong + BCE weak.
    Returns (loss, strong, weak); caller runs autograd + adam_step."""
    mel = ofe.mel_spectrogram(audio)
    strong, weak = detect(mel, P, cfg, training, gru_impl=gru_impl, **(fwd_kw or {}))
    ls, lw = supervised_losses(strong, weak, labels, n_strong, n_weak)
    return ls + lw, strong, weak


def mean_teacher_step(student, teacher, audio, labels, batch_sizes, step_num, rampup_len, cfg=ocrnn.CFG_2023,
                      const_max=2.0, mix=None, mixup_type="soft", student_kw=None, teacher_kw=None,
                      gru_impl="aten"):
    """sed_trainer.py:269-356 (2023 recipe).  batch_sizes=[n_synth, n_weak, n_unlabelled].
    mix = None or dict(weak=(c,perm), strong=(c,perm)) (already drawn: sed_trainer.py:295 `0.5 > random.random()`).
    Returns dict of losses + predictions; tot_loss carries the autograd graph of the student."""
    n_s, n_w, _ = batch_sizes
    features = ofe.mel_spectrogram(audio)
    labels = labels.clone()
    labels_weak = (torch.sum(labels[n_s:n_s + n_w], -1) > 0).float()
    if mix is not None:
        features = features.clone()
        fw, labels_weak = mixup(features[n_s:n_s + n_w], labels_weak, *mix["weak"], mixup_label_type=mixup_type)
        fs, ls = mixup(features[:n_s], labels[:n_s], *mix["strong"], mixup_label_type=mixup_type)
        features[n_s:n_s + n_w] = fw
        features[:n_s] = fs
        labels[:n_s] = ls
    strong_s, weak_s = detect(features, student, cfg, True, gru_impl=gru_impl, **(student_kw or {}))
    loss_strong = bce(strong_s[:n_s], labels[:n_s])
    loss_weak = bce(weak_s[n_s:n_s + n_w], labels_weak)
    with torch.no_grad():
        strong_t, weak_t = detect(features, teacher, cfg, True, gru_impl=gru_impl, **(teacher_kw or {}))
        loss_strong_t = bce(strong_t[:n_s], labels[:n_s])
        loss_weak_t = bce(weak_t[n_s:n_s + n_w], labels_weak)
    weight = const_max * warmup_scale(step_num, rampup_len)
    strong_ss = F.mse_loss(strong_s, strong_t.detach())
    weak_ss = F.mse_loss(weak_s, weak_t.detach())
    tot_self = (strong_ss + weak_ss) * weight
    tot = loss_strong + loss_weak + tot_self
    return dict(tot_loss=tot, loss_strong=loss_strong, loss_weak=loss_weak, loss_strong_teacher=loss_strong_t,
                loss_weak_teacher=loss_weak_t, weight=weight, strong_self_sup=strong_ss, weak_self_sup=weak_ss,
                strong_student=strong_s, weak_student=weak_s, strong_teacher=strong_t, weak_teacher=weak_t,
                features=features, labels=labels, labels_weak=labels_weak)


def draw_mixup_2024(batch_sizes, with_embeddings=True):
    """RNG consumption of the three `apply_mixup` calls of the 2024 step (sed_trainer_pretrained.py:350-363 ->
    :282-301): groups in the order weak [indx_strong, indx_weak), synth+strong [indx_maestro, indx_strong),
    maestro [0, indx_maestro); inside a group first the features' (c, perm), then the embeddings' (c, perm)."""
    cs = np.cumsum(batch_sizes)
    groups = [(int(cs[2]), int(cs[3])), (int(cs[0]), int(cs[2])), (0, int(cs[0]))]
    out = []
    for lo, hi in groups:
        feat = draw_mixup(hi - lo)
        emb = draw_mixup(hi - lo) if with_embeddings else None
        out.append(((lo, hi), feat, emb))
    return out


def mean_teacher_step_2024(student, teacher, audio, labels, embeddings, valid_class_mask, batch_sizes, step_num,
                           rampup_len, cfg=ocrnn.CFG_2024, const_max=2.0, mix=None, mixup_type="soft", student_kw=None,
                           teacher_kw=None, gru_impl="aten", use_const_weight=False):
    """recipes/dcase2024_task4_baseline/local/sed_trainer_pretrained.py:318-430.

    batch_sizes = [maestro, synth, strong, weak, unlabelled] (:335-337).  Masks (:339-346): strong rows [0, indx_strong),
    weak rows [indx_strong, indx_weak), consistency rows `mask_unlabeled` = [indx_maestro, B).  `mix` = None or the list
    draw_mixup_2024 returns (the `mixup_prob > random.random()` draw is the caller's): per group the FEATURES are mixed
    with one (c, perm) and the EMBEDDINGS with an independent one, and the group's labels are mixed by both in turn
    (:282-301).  Weak labels are derived from the MIXED labels (:366), then labels / weak labels are masked by
    valid_class_mask (:367-370).  `use_const_weight`: current_epoch >= epoch_decay (:402-405)."""
    cs = np.cumsum(batch_sizes)
    i_maestro, i_strong, i_weak = int(cs[0]), int(cs[2]), int(cs[3])
    features = ofe.mel_spectrogram(audio)
    labels = labels.clone()
    if mix is not None:
        features = features.clone()
        embeddings = embeddings.clone()
        for (lo, hi), (c1, p1), emb_draw in mix:
            f, y = mixup(features[lo:hi], labels[lo:hi], c1, p1, mixup_label_type=mixup_type)
            features[lo:hi], labels[lo:hi] = f, y
            if emb_draw is not None:
                c2, p2 = emb_draw
                e, y = mixup(embeddings[lo:hi], labels[lo:hi], c2, p2, mixup_label_type=mixup_type)
                embeddings[lo:hi], labels[lo:hi] = e, y
    labels_weak = (torch.sum(labels[i_strong:i_weak], -1) > 0).float()
    labels = labels.masked_fill(~valid_class_mask[:, :, None].expand_as(labels), 0.0)
    labels_weak = labels_weak.masked_fill(~valid_class_mask[i_strong:i_weak], 0.0)
    kw = dict(embeddings=embeddings, classes_mask=valid_class_mask, gru_impl=gru_impl)
    strong_s, weak_s = detect(features, student, cfg, True, **kw, **(student_kw or {}))
    loss_strong = bce(strong_s[:i_strong], labels[:i_strong])
    loss_weak = bce(weak_s[i_strong:i_weak], labels_weak)
    with torch.no_grad():
        strong_t, weak_t = detect(features, teacher, cfg, True, **kw, **(teacher_kw or {}))
    weight = const_max if use_const_weight else const_max * warmup_scale(step_num, rampup_len)
    strong_ss = F.mse_loss(strong_s[i_maestro:], strong_t.detach()[i_maestro:])
    weak_ss = F.mse_loss(weak_s[i_maestro:], weak_t.detach()[i_maestro:])
    tot = loss_strong + loss_weak + (strong_ss + weak_ss) * weight
    return dict(tot_loss=tot, loss_strong=loss_strong, loss_weak=loss_weak, weight=weight, strong_self_sup=strong_ss,
                weak_self_sup=weak_ss, strong_student=strong_s, weak_student=weak_s, strong_teacher=strong_t,
                weak_teacher=weak_t, features=features, embeddings=embeddings, labels=labels, labels_weak=labels_weak)
