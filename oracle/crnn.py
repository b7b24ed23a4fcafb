"""Oracle: functional CRNN forward (CPU, fp32, autograd-differentiable).  Test infrastructure only.

Restates desed_task/nnet/CRNN.py:221-306 (forward), :152-178 (heads), :207-219 (specaugment),
desed_task/nnet/CNN.py:5-30,66-98 (GLU / ContextGating / conv block) and
desed_task/nnet/RNN.py:19-30 (BiGRU) over a *parameter dict* that uses the reference's own
state_dict key names, so a reference checkpoint can be evaluated without the reference code.
Random draws (dropout masks, specaugment / dropstep spans) are INJECTED, never drawn here, so
that the CUDA path and the oracle can be fed identical randomness.
"""
from dataclasses import dataclass, field
from typing import Optional, Sequence

import torch
import torch.nn.functional as F


@dataclass
class CRNNConfig:
    """The subset of CRNN(**config['net']) / CNN(**kwargs) arguments that changes the math
    (CRNN.py:12-38, CNN.py:34-46).  Defaults = recipes/dcase2023_task4_baseline/confs/default.yaml:72-84."""
    n_in_channel: int = 1
    nclass: int = 10
    attention: bool = True
    activation: str = "glu"
    dropout: float = 0.5
    n_RNN_cell: int = 128
    n_layers_RNN: int = 2
    kernel_size: Sequence[int] = (3, 3, 3, 3, 3, 3, 3)
    padding: Sequence[int] = (1, 1, 1, 1, 1, 1, 1)
    stride: Sequence[int] = (1, 1, 1, 1, 1, 1, 1)
    nb_filters: Sequence[int] = (16, 32, 64, 128, 128, 128, 128)
    pooling: Sequence[Sequence[int]] = ((2, 2), (2, 2), (1, 2), (1, 2), (1, 2), (1, 2), (1, 2))
    normalization: str = "batch"
    use_embeddings: bool = False
    embedding_size: int = 768
    aggregation_type: str = "pool1d"
    dropstep_recurrent: float = 0.0
    bn_eps: float = 1e-3          # CNN.py:76
    bn_momentum: float = 0.99     # CNN.py:76


CFG_2023 = CRNNConfig()
# recipes/dcase2024_task4_baseline/confs/pretrained.yaml:86-110 (rnn_layers is ignored -> 2 layers)
CFG_2024 = CRNNConfig(nclass=27, dropout=0.2, n_RNN_cell=192, use_embeddings=True,
                      embedding_size=768, aggregation_type="pool1d")


def gru_loop(x, params, prefix, hidden, num_layers):
    """Explicit BiGRU (PyTorch nn.GRU semantics, gate order r,z,n; RNN.py:19-30).
    r = s(W_ir x + b_ir + W_hr h + b_hr); z likewise; n = tanh(W_in x + b_in + r*(W_hn h + b_hn));
    h' = (1-z)*n + z*h."""
    B, T, _ = x.shape
    inp = x
    for layer in range(num_layers):
        outs = []
        for suffix in ("", "_reverse"):
            w_ih = params[f"{prefix}weight_ih_l{layer}{suffix}"]
            w_hh = params[f"{prefix}weight_hh_l{layer}{suffix}"]
            b_ih = params[f"{prefix}bias_ih_l{layer}{suffix}"]
            b_hh = params[f"{prefix}bias_hh_l{layer}{suffix}"]
            gi = inp @ w_ih.t() + b_ih                     # [B,T,3H]
            h = x.new_zeros(B, hidden)
            hs = [None] * T
            order = range(T) if suffix == "" else range(T - 1, -1, -1)
            for t in order:
                gh = h @ w_hh.t() + b_hh
                i_r, i_z, i_n = gi[:, t].chunk(3, -1)
                h_r, h_z, h_n = gh.chunk(3, -1)
                r = torch.sigmoid(i_r + h_r)
                z = torch.sigmoid(i_z + h_z)
                n = torch.tanh(i_n + r * h_n)
                h = (1 - z) * n + z * h
                hs[t] = h
            outs.append(torch.stack(hs, 1))
        inp = torch.cat(outs, -1)
    return inp


def gru_aten(x, params, prefix, hidden, num_layers):
    """Same math through ATen's fused GRU (what nn.GRU calls) - used for the CPU baseline timing."""
    flat = []
    for layer in range(num_layers):
        for suffix in ("", "_reverse"):
            for name in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                flat.append(params[f"{prefix}{name}_l{layer}{suffix}"])
    h0 = x.new_zeros(2 * num_layers, x.shape[0], hidden)
    out, _ = torch._VF.gru(x, h0, flat, True, num_layers, 0.0, False, True, True)
    return out


def span_mask(n, start, end, device=None):
    """mask[b, i] = start[b] <= i < end[b] (torchaudio functional.py:868-878)."""
    idx = torch.arange(n, device=device)[None, :]
    return (idx >= start[:, None]) & (idx < end[:, None])


def draw_specaugment(B, n_f, n_t, f_l=10, f_p=0.2, t_l=5, t_p=0.2, device=None):
    """Draw the mask spans exactly like CRNN.apply_specaugment (CRNN.py:207-219) does through torchaudio
    functional.py:857-869: per mask value = rand(B) * param, min_value = rand(B) * (size - value); the 'freq' mask
    is drawn first.  param = l if p == 1 else min(l, int(size * p))."""
    def one(size, l, p):
        param = l if p == 1.0 else min(l, int(size * p))
        if param < 1:
            return None
        value = torch.rand(B, device=device) * param
        min_value = torch.rand(B, device=device) * (size - value)
        return min_value.long(), min_value.long() + value.long()
    spec = {}
    f = one(n_f, f_l, f_p)
    if f is not None:
        spec["f_start"], spec["f_end"] = f
    t = one(n_t, t_l, t_p)
    if t is not None:
        spec["t_start"], spec["t_end"] = t
    return spec


def draw_dropstep(B, frames, length, p, with_embeddings, device=None):
    """The span draws of the `dropstep` TimeMasking(length, iid_masks=True, p) in CRNN.forward (CRNN.py:288-301), in the
    reference's RNG order: the span for x first, then (embedding branch only) the span for the embeddings.
    Returns dict(x_start, x_end[, e_start, e_end]) or None when the mask parameter rounds to < 1 (no draw happens)."""
    param = length if p == 1.0 else min(length, int(frames * p))
    if param < 1:
        return None
    out = {}
    for tag in ("x", "e") if with_embeddings else ("x",):
        value = torch.rand(B, device=device) * param
        min_value = torch.rand(B, device=device) * (frames - value)
        out[tag + "_start"], out[tag + "_end"] = min_value.long(), min_value.long() + value.long()
    return out


def apply_specaugment(x, spec):
    """CRNN.py:207-219: 'freq' mask (TimeMasking on the transposed tensor) then time mask, fill 0.0.
    spec = dict(f_start, f_end, t_start, t_end) int64 [B] (already drawn)."""
    if spec is None:
        return x
    B, Fm, T = x.shape
    if "f_start" in spec:
        fm = span_mask(Fm, spec["f_start"], spec["f_end"])          # [B,F]
        x = x.masked_fill(fm[:, :, None], 0.0)
    if "t_start" in spec:
        tm = span_mask(T, spec["t_start"], spec["t_end"])           # [B,T]
        x = x.masked_fill(tm[:, None, :], 0.0)
    return x


def cnn_forward(x, params, cfg, training, bn_state=None, drop_masks=None, collect=None, bn_eval=False):
    """CNN.py:66-98: conv -> BN -> act -> dropout -> avgpool, per layer.  x: (B, Cin, T, F).
    bn_eval: BatchNorm on its running statistics although `training` (freeze_bn, CRNN.py:308-323)."""
    for i, nout in enumerate(cfg.nb_filters):
        p = "cnn.cnn."
        x = F.conv2d(x, params[f"{p}conv{i}.weight"], params[f"{p}conv{i}.bias"],
                     stride=cfg.stride[i], padding=cfg.padding[i])
        if collect is not None:
            collect[f"conv{i}"] = x
        if cfg.normalization == "batch":
            rm = params[f"{p}batchnorm{i}.running_mean"]
            rv = params[f"{p}batchnorm{i}.running_var"]
            bn_train = training and not bn_eval
            if bn_train:
                rm, rv = rm.clone(), rv.clone()
            x = F.batch_norm(x, rm, rv, params[f"{p}batchnorm{i}.weight"], params[f"{p}batchnorm{i}.bias"],
                             bn_train, cfg.bn_momentum, cfg.bn_eps)
            if bn_train and bn_state is not None:
                bn_state[f"{p}batchnorm{i}.running_mean"] = rm
                bn_state[f"{p}batchnorm{i}.running_var"] = rv
        elif cfg.normalization == "layer":
            x = F.group_norm(x, 1, params[f"{p}layernorm{i}.weight"], params[f"{p}layernorm{i}.bias"])
        act = cfg.activation.lower()
        if act == "glu":          # CNN.py:11-16
            lin = F.linear(x.permute(0, 2, 3, 1), params[f"{p}glu{i}.linear.weight"],
                           params[f"{p}glu{i}.linear.bias"]).permute(0, 3, 1, 2)
            x = lin * torch.sigmoid(x)
        elif act == "cg":         # CNN.py:25-30
            lin = F.linear(x.permute(0, 2, 3, 1), params[f"{p}cg{i}.linear.weight"],
                           params[f"{p}cg{i}.linear.bias"]).permute(0, 3, 1, 2)
            x = x * torch.sigmoid(lin)
        elif act == "relu":
            x = F.relu(x)
        elif act == "leakyrelu":
            x = F.leaky_relu(x, 0.2)
        if training and cfg.dropout > 0:
            if drop_masks is not None:
                x = x * drop_masks[i] / (1.0 - cfg.dropout)
            else:
                x = F.dropout(x, cfg.dropout, True)
        if collect is not None:
            collect[f"act{i}"] = x
        x = F.avg_pool2d(x, tuple(cfg.pooling[i]))
        if collect is not None:
            collect[f"pool{i}"] = x
    return x


def heads(x, params, cfg, classes_mask=None):
    """CRNN._get_logits_one_head, CRNN.py:152-178 (pad_mask is None at every call site)."""
    strong = torch.sigmoid(F.linear(x, params["dense.weight"], params["dense.bias"]))
    cm = None
    if classes_mask is not None:
        cm = ~classes_mask[:, None].expand_as(strong)
    if cfg.attention:
        sof = F.linear(x, params["dense_softmax.weight"], params["dense_softmax.bias"])
        if cm is not None:
            sof = sof.masked_fill(cm, -1e30)
        sof = torch.softmax(sof, dim=-1)           # over CLASSES (CRNN.py:124,166)
        sof = torch.clamp(sof, min=1e-7, max=1)
        weak = (strong * sof).sum(1) / sof.sum(1)
    else:
        weak = strong.mean(1)
    if cm is not None:
        strong = strong.masked_fill(cm, 0.0)
        weak = weak.masked_fill(cm[:, 0], 0.0)
    return strong.transpose(1, 2), weak


def crnn_forward(params, x, cfg=CFG_2023, training=False, embeddings=None, classes_mask=None,
                 specaug=None, drop_masks=None, rnn_drop_mask=None, emb_drop_mask=None,
                 dropstep=None, bn_state=None, collect=None, gru_impl="loop", bn_eval=False):
    """CRNN.forward (CRNN.py:221-306).  x: [B, n_mels, T] scaled log-mel.

    drop_masks: list of 0/1 keep masks (B,C,T,F) per conv layer; rnn_drop_mask: [B,T',2H] keep mask for
    the post-RNN dropout (CRNN.py:304); emb_drop_mask: [B,T',nb+emb] keep mask (CRNN.py:294);
    dropstep: dict(x_start,x_end,e_start,e_end) frame spans (CRNN.py:288-293).
    Returns (strong [B,C,T'], weak [B,C])."""
    if training:
        x = apply_specaugment(x, specaug)
    x = x.transpose(1, 2).unsqueeze(1)                     # (B,1,T,F)  CRNN.py:224
    x = cnn_forward(x, params, cfg, training, bn_state, drop_masks, collect, bn_eval)
    bs, chan, frames, freq = x.shape
    assert freq == 1, "oracle covers the shipped configs (freq pooled to 1)"
    x = x.squeeze(-1).permute(0, 2, 1)                     # [B,T',C]  CRNN.py:244-245
    if collect is not None:
        collect["cnn_out"] = x
    if cfg.use_embeddings:
        if cfg.aggregation_type == "interpolate":                           # CRNN.py:270-278
            emb = F.interpolate(embeddings.unsqueeze(1), size=(embeddings.shape[1], frames),
                                mode="nearest-exact").squeeze(1).transpose(1, 2)
        else:
            assert cfg.aggregation_type == "pool1d"
            emb = F.adaptive_avg_pool1d(embeddings, frames).transpose(1, 2)   # CRNN.py:280-283
        if training and cfg.dropstep_recurrent and dropstep is not None:
            xm = span_mask(frames, dropstep["x_start"], dropstep["x_end"])
            em = span_mask(frames, dropstep["e_start"], dropstep["e_end"])
            x = x.masked_fill(xm[:, :, None], 0.0)
            emb = emb.masked_fill(em[:, :, None], 0.0)
        cat = torch.cat((x, emb), -1)
        if training and cfg.dropout > 0:
            if emb_drop_mask is not None:
                cat = cat * emb_drop_mask / (1.0 - cfg.dropout)
            else:
                cat = F.dropout(cat, cfg.dropout, True)
        x = F.linear(cat, params["cat_tf.weight"], params["cat_tf.bias"])  # CRNN.py:294
        if collect is not None:
            collect["fused"] = x
    elif training and cfg.dropstep_recurrent and dropstep is not None:     # CRNN.py:295-301 (no embeddings)
        xm = span_mask(frames, dropstep["x_start"], dropstep["x_end"])
        x = x.masked_fill(xm[:, :, None], 0.0)
        if cfg.dropout > 0:
            x = x * emb_drop_mask / (1.0 - cfg.dropout) if emb_drop_mask is not None else F.dropout(x, cfg.dropout, True)
    gru = gru_loop if gru_impl == "loop" else gru_aten
    x = gru(x, params, "rnn.rnn.", cfg.n_RNN_cell, cfg.n_layers_RNN)       # CRNN.py:303
    if collect is not None:
        collect["rnn_out"] = x
    if training and cfg.dropout > 0:                                       # CRNN.py:304
        if rnn_drop_mask is not None:
            x = x * rnn_drop_mask / (1.0 - cfg.dropout)
        else:
            x = F.dropout(x, cfg.dropout, True)
    return heads(x, params, cfg, classes_mask)


def init_params(cfg=CFG_2023, seed=42, trained_like=False):
    """Build a parameter dict with the reference's key names/shapes and PyTorch's default inits
    (same init *distributions* as nn.Conv2d / nn.Linear / nn.GRU / nn.BatchNorm2d; seeded).
    trained_like=True perturbs BN affine + running stats and scales the heads so posteriors move away
    from 0.5."""
    g = torch.Generator().manual_seed(seed)

    def uni(shape, bound):
        return (torch.rand(shape, generator=g) * 2 - 1) * bound

    P = {}
    cin = cfg.n_in_channel
    for i, cout in enumerate(cfg.nb_filters):
        k = cfg.kernel_size[i]
        bound = 1.0 / (cin * k * k) ** 0.5
        P[f"cnn.cnn.conv{i}.weight"] = uni((cout, cin, k, k), bound)
        P[f"cnn.cnn.conv{i}.bias"] = uni((cout,), bound)
        P[f"cnn.cnn.batchnorm{i}.weight"] = torch.ones(cout)
        P[f"cnn.cnn.batchnorm{i}.bias"] = torch.zeros(cout)
        P[f"cnn.cnn.batchnorm{i}.running_mean"] = torch.zeros(cout)
        P[f"cnn.cnn.batchnorm{i}.running_var"] = torch.ones(cout)
        P[f"cnn.cnn.batchnorm{i}.num_batches_tracked"] = torch.zeros((), dtype=torch.long)
        act = cfg.activation.lower()
        if act in ("glu", "cg"):
            b = 1.0 / cout ** 0.5
            P[f"cnn.cnn.{act}{i}.linear.weight"] = uni((cout, cout), b)
            P[f"cnn.cnn.{act}{i}.linear.bias"] = uni((cout,), b)
        if trained_like:
            P[f"cnn.cnn.batchnorm{i}.weight"] = 1.0 + uni((cout,), 0.5)
            P[f"cnn.cnn.batchnorm{i}.bias"] = uni((cout,), 0.5)
            P[f"cnn.cnn.batchnorm{i}.running_mean"] = uni((cout,), 0.2)
            P[f"cnn.cnn.batchnorm{i}.running_var"] = 0.5 + torch.rand(cout, generator=g)
        cin = cout
    H = cfg.n_RNN_cell
    nb_in = cfg.nb_filters[-1]
    if cfg.use_embeddings:
        b = 1.0 / (nb_in + cfg.embedding_size) ** 0.5
        P["cat_tf.weight"] = uni((nb_in, nb_in + cfg.embedding_size), b)
        P["cat_tf.bias"] = uni((nb_in,), b)
    for layer in range(cfg.n_layers_RNN):
        n_in = nb_in if layer == 0 else 2 * H
        b = 1.0 / H ** 0.5
        for suffix in ("", "_reverse"):
            P[f"rnn.rnn.weight_ih_l{layer}{suffix}"] = uni((3 * H, n_in), b)
            P[f"rnn.rnn.weight_hh_l{layer}{suffix}"] = uni((3 * H, H), b)
            P[f"rnn.rnn.bias_ih_l{layer}{suffix}"] = uni((3 * H,), b)
            P[f"rnn.rnn.bias_hh_l{layer}{suffix}"] = uni((3 * H,), b)
    b = 1.0 / (2 * H) ** 0.5
    scale = 6.0 if trained_like else 1.0
    P["dense.weight"] = uni((cfg.nclass, 2 * H), b) * scale
    P["dense.bias"] = uni((cfg.nclass,), b) * scale
    if cfg.attention:
        P["dense_softmax.weight"] = uni((cfg.nclass, 2 * H), b) * scale
        P["dense_softmax.bias"] = uni((cfg.nclass,), b) * scale
    return P


def is_float_param(name):
    return not name.endswith(("running_mean", "running_var", "num_batches_tracked"))


def param_names(P):
    """Trainable tensors, in nn.Module.parameters() order of the reference CRNN
    (cnn -> rnn -> dense -> dense_softmax -> cat_tf; CRNN.py:80-150)."""
    order = []
    for k in P:
        if k.startswith("cnn.") and is_float_param(k):
            order.append(k)
    for k in P:
        if k.startswith("rnn."):
            order.append(k)
    for k in ("dense.weight", "dense.bias", "dense_softmax.weight", "dense_softmax.bias",
              "cat_tf.weight", "cat_tf.bias"):
        if k in P:
            order.append(k)
    return order
