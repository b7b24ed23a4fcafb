"""The reference's OWN hot path, built from its unmodified modules: bench.py's `--impl reference` arm (host CPU) and the
`gpu_library_baseline` leg (the same modules moved to cuda:0, i.e. cuFFT / cuBLAS / cuDNN through PyTorch).

Benchmark infrastructure only - nothing here is imported by desed_task_b200.  Modules come from baseline.refload (the
checkout in the build container, the baseline/_ref install on the GPU box): `desed_task.nnet.CRNN.CRNN`,
`desed_task.data_augm.mixup`, `desed_task.utils.scaler.TorchScaler`, plus torchaudio's MelSpectrogram / AmplitudeToDB with
the constructor arguments of recipes/dcase2023_task4_baseline/local/sed_trainer.py:79-91,253-264.  The Lightning module
itself cannot be imported (pytorch_lightning / codecarbon / sed_scores_eval are absent), so the step composition restates
sed_trainer.py:269-356 (training_step), :187-199 (update_ema), :358-365 (hook order) and train_sed.py:199-201 (Adam) around
those modules; 2024: sed_trainer_pretrained.py:282-301,318-430.
"""
import copy
import random

import numpy as np
import torch

from . import refload

NET_2023 = dict(dropout=0.5, rnn_layers=2, n_in_channel=1, nclass=10, attention=True, n_RNN_cell=128, activation="glu",
                rnn_type="BGRU", kernel_size=[3] * 7, padding=[1] * 7, stride=[1] * 7,
                nb_filters=[16, 32, 64, 128, 128, 128, 128],
                pooling=[[2, 2], [2, 2], [1, 2], [1, 2], [1, 2], [1, 2], [1, 2]], dropout_recurrent=0,
                use_embeddings=False)
# recipes/dcase2024_task4_baseline/confs/pretrained.yaml:86-110
NET_2024 = dict(NET_2023, dropout=0.2, rnn_layers=1, nclass=27, n_RNN_cell=192, use_embeddings=True, embedding_size=768,
                embedding_type="frame", aggregation_type="pool1d", specaugm_t_p=0.0, specaugm_f_p=0.0,
                dropstep_recurrent=0.0, dropstep_recurrent_len=16)


class ReferencePath:
    """mel -> take_log -> scaler -> CRNN (+ teacher) with the reference's modules on `device`."""

    def __init__(self, device, net_cfg=None, teacher=False, seed=42, lr=1e-3):
        from torchaudio.transforms import AmplitudeToDB, MelSpectrogram
        R = refload.load()
        self.R, self.device = R, torch.device(device)
        torch.manual_seed(seed)
        self.mel_spec = MelSpectrogram(sample_rate=16000, n_fft=2048, win_length=2048, hop_length=256, f_min=0,
                                       f_max=8000, n_mels=128, window_fn=torch.hamming_window,
                                       wkwargs={"periodic": False}, power=1).to(self.device)
        self.amp_to_db = AmplitudeToDB(stype="amplitude")
        self.amp_to_db.amin = 1e-5
        self.scaler = R.TorchScaler("instance", "minmax", [1, 2])
        self.student = R.CRNN(**(net_cfg or NET_2023)).to(self.device)
        self.teacher = None
        if teacher:
            self.teacher = copy.deepcopy(self.student)
            for p in self.teacher.parameters():
                p.detach_()
        self.bce = torch.nn.BCELoss()
        self.mse = torch.nn.MSELoss()
        self.lr = lr
        self.opt = torch.optim.Adam(self.student.parameters(), lr, betas=(0.9, 0.999))
        self.step_num = 1

    def take_log(self, mels):
        return self.amp_to_db(mels).clamp(min=-50, max=80)

    def detect(self, mel, model, **kw):
        return model(self.scaler(self.take_log(mel)), **kw)

    # ---- BASELINE config 2: supervised step ([n_s strong | n_w weak] rows)
    def supervised_step(self, audio, labels, n_s, autocast=False):
        self.student.train()
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=autocast):
            feats = self.mel_spec(audio)
            strong, weak = self.detect(feats, self.student)
        labels_weak = (torch.sum(labels[n_s:], -1) > 0).float()
        loss = self.bce(strong[:n_s].float(), labels[:n_s]) + self.bce(weak[n_s:].float(), labels_weak)
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        self.opt.step()
        return loss

    # ---- BASELINE config 3: mean-teacher step, sed_trainer.py:269-365
    def mean_teacher_step(self, audio, labels, batch_sizes, mixup_type="soft", autocast=False, const_max=2.0,
                          ema_factor=0.999, rampup=12500):
        n_s, n_w, _ = batch_sizes
        self.student.train()
        self.teacher.train()
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=autocast):
            features = self.mel_spec(audio)
        B = features.shape[0]
        strong_mask = torch.zeros(B, device=features.device).bool()
        weak_mask = torch.zeros(B, device=features.device).bool()
        strong_mask[:n_s] = 1
        weak_mask[n_s:n_s + n_w] = 1
        labels = labels.clone()
        labels_weak = (torch.sum(labels[weak_mask], -1) > 0).float()
        if mixup_type is not None and 0.5 > random.random():
            features[weak_mask], labels_weak = self.R.data_augm.mixup(features[weak_mask], labels_weak,
                                                                     mixup_label_type=mixup_type)
            features[strong_mask], labels[strong_mask] = self.R.data_augm.mixup(features[strong_mask], labels[strong_mask],
                                                                               mixup_label_type=mixup_type)
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=autocast):
            strong_s, weak_s = self.detect(features, self.student)
            with torch.no_grad():
                strong_t, weak_t = self.detect(features, self.teacher)
        strong_s, weak_s = strong_s.float(), weak_s.float()
        loss = self.bce(strong_s[strong_mask], labels[strong_mask]) + self.bce(weak_s[weak_mask], labels_weak)
        phase = 1.0 - min(self.step_num, rampup) / rampup
        weight = const_max * float(np.exp(-5.0 * phase * phase))
        loss = loss + (self.mse(strong_s, strong_t.float().detach()) + self.mse(weak_s, weak_t.float().detach())) * weight
        # PL 1.9 automatic optimisation: on_before_zero_grad (EMA) -> zero_grad -> backward -> optimizer.step
        alpha = min(1 - 1 / (self.step_num + 1), ema_factor)
        for ema_p, p in zip(self.teacher.parameters(), self.student.parameters()):
            ema_p.mul_(alpha).add_(p.detach(), alpha=1 - alpha)
        self.opt.zero_grad(set_to_none=False)
        loss.backward()
        self.opt.step()
        self.step_num += 1
        return loss

    # ---- BASELINE config 5: inference, sed_trainer.py:608-640 + local/utils.py:45-63 (median filter on the host, per clip)
    @torch.no_grad()
    def inference(self, audio, median_window=7, autocast=False):
        import scipy.ndimage
        self.student.eval()
        with torch.autocast(self.device.type, dtype=torch.bfloat16, enabled=autocast):
            strong, weak = self.detect(self.mel_spec(audio), self.student)
        out = []
        for c in strong.float():
            c = c.transpose(0, 1).detach().cpu().numpy()
            out.append(scipy.ndimage.median_filter(c, (median_window, 1)))
        return out


def timed_gpu(fn, steps, warmup):
    """clips-agnostic device timing of `fn(i)` (CUDA events on the current stream, sync on both sides)."""
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def gpu_library_baseline(dev, workload, B, audio, labels, steps=20, warmup=5):
    """The bar SURVEY.md 8(d)(ii) sets: the reference's modules on the same B200 through PyTorch's library kernels
    (fp32 as shipped - cuDNN convs/RNN in TF32 by PyTorch's defaults - and bf16 autocast), eager and, where it captures,
    as one torch.cuda.CUDAGraph per step.  `audio` / `labels`: lists of DEVICE batches rotated over the steps.
    Returns {variant: clips/s}."""
    out = {}
    n = len(audio)
    for tag, autocast in (("fp32", False), ("bf16_autocast", True)):
        try:
            random.seed(0); np.random.seed(0)
            if workload == "inference":
                path = ReferencePath(dev)
                fn = lambda i: path.inference(audio[i % n], autocast=autocast)                      # noqa: E731
            elif workload == "mean_teacher":
                path = ReferencePath(dev, teacher=True)
                bs = [B // 4, B // 4, B // 2]
                fn = lambda i: path.mean_teacher_step(audio[i % n], labels[i % n], bs, autocast=autocast)   # noqa: E731
            else:
                path = ReferencePath(dev)
                fn = lambda i: path.supervised_step(audio[i % n], labels[i % n], B // 2, autocast=autocast)  # noqa: E731
            ms = timed_gpu(fn, steps, warmup)
            out[tag + "_eager"] = round(B / ms * 1e3, 1)
        except Exception as e:                                                                   # noqa: BLE001
            out[tag + "_eager"] = "failed: %s" % (str(e).splitlines()[0][:160],)
            continue
        if workload != "supervised":
            continue
        # whole-step CUDA graph (static input buffers, capturable Adam): the best the library path can do about launch latency
        try:
            path = ReferencePath(dev)
            path.opt = torch.optim.Adam(path.student.parameters(), 1e-3, betas=(0.9, 0.999), capturable=True)
            sa, sl = audio[0].clone(), labels[0].clone()
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    path.supervised_step(sa, sl, B // 2, autocast=autocast)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                path.supervised_step(sa, sl, B // 2, autocast=autocast)

            def replay(i):
                sa.copy_(audio[i % n], non_blocking=True)
                sl.copy_(labels[i % n], non_blocking=True)
                g.replay()
            ms = timed_gpu(replay, steps, warmup)
            out[tag + "_cuda_graph"] = round(B / ms * 1e3, 1)
        except Exception as e:                                                                   # noqa: BLE001
            out[tag + "_cuda_graph"] = "failed: %s" % (str(e).splitlines()[0][:160],)
            torch.cuda.synchronize()
    return out
