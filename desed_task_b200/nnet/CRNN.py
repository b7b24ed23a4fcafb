"""CRNN on the sm_100a kernels.  Host-side mirror of desed_task/nnet/CRNN.py:10-323.

Same constructor signature (unknown keys fall through **kwargs into CNN and are dropped, CRNN.py:80-82 / CNN.py:45), same
module tree / parameter order / state_dict keys (so `CRNN(**config["net"])`, `deepcopy`, `zip(ema.parameters(),
model.parameters())` and published checkpoints work unchanged), same `forward(x, pad_mask, embeddings, classes_mask) ->
(strong [B,C,T//4], weak [B,C])`, same quirks (`train()` returns None; `rnn_layers` is ignored).

The arithmetic is ONE call into libsedk per forward and one per backward (sedk_crnn_forward / sedk_crnn_backward,
include/sedk.h) on a cached plan; there is no PyTorch fallback: configurations the kernels do not cover raise
NotImplementedError when they are used.
"""
import ctypes
import warnings

import torch
import torch.nn as nn

from .. import _lib
from .._lib import ConvLayer, CrnnPlan, GruLayer, check, lib, ptr, require_cuda, stream_ptr
from .CNN import CNN
from .RNN import BidirectionalGRU

PRECISION_TF32 = 0      # tensor-core TF32 multiplies, fp32 accumulate (what cuDNN does for the reference on GPU)
PRECISION_FP32 = 1      # 3xTF32 error-compensated: fp32-equivalent, used by the strict parity tests

_default_precision = PRECISION_TF32


def set_default_precision(p):
    global _default_precision
    assert p in (PRECISION_TF32, PRECISION_FP32)
    _default_precision = p


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class _Workspace:
    """All device buffers + the C plan for one (batch, frames, device) shape of one model."""

    def __init__(self, model, B, n_mels, n_frames, device, emb_shape):
        self.key = (B, n_mels, n_frames, str(device), emb_shape)
        self.generation = 0
        self.B = B
        f32 = dict(device=device, dtype=torch.float32)
        self.bufs = []

        def new(*shape, dtype=torch.float32, zero=False):
            t = (torch.zeros if zero else torch.empty)(*shape, device=device, dtype=dtype)
            self.bufs.append(t)
            return t

        params = list(model.named_parameters())
        self.pnames = [n for n, _ in params]
        sizes = [p.numel() for _, p in params]
        cnn = model.cnn
        n_conv = len(cnn.nb_filters)
        # One contiguous region holds every buffer the backward accumulates into (the flat gradient + the packed conv
        # weight-gradient accumulators) and one holds every BatchNorm statistics array: sedk_crnn_plan.zero_bwd / zero_fwd
        # let the library clear each with a single memset.
        chans = [1] + list(cnn.nb_filters)
        gw_sizes = [0] + [9 * chans[i + 1] * chans[i] for i in range(1, n_conv)]
        from ..optim import flat_layout
        g_offs, n_grad = flat_layout(sizes)       # the same 16-byte-aligned layout as the flat parameter buffer
        # data parallel: the engine may supply the region from NVLink-symmetric memory (desed_task_b200/nvls.py)
        galloc = getattr(model, "grad_alloc", None)
        if galloc is not None:
            self.zero_bwd = galloc(n_grad + sum(gw_sizes))
            self.bufs.append(self.zero_bwd)
        else:
            self.zero_bwd = new(n_grad + sum(gw_sizes), zero=True)
        self.gflat = self.zero_bwd[:n_grad]
        self.g_offs = g_offs
        gw_off = [n_grad + sum(gw_sizes[:i]) for i in range(n_conv)]
        # + the 304 sums of the store-free first block (sedk_crnn_plan.l0_sums), cleared by the same memset
        n_stats = sum(4 * c for c in cnn.nb_filters)
        self.zero_fwd = new(n_stats + 304, dtype=torch.float64, zero=True)
        st_off = [sum(4 * c for c in cnn.nb_filters[:i]) for i in range(n_conv)]
        self.gviews = {}
        for (n, p), sz, off in zip(params, sizes, g_offs):
            self.gviews[n] = self.gflat[off:off + sz].view(p.shape)

        plan = CrnnPlan()
        plan.B, plan.n_mels, plan.n_frames = B, n_mels, n_frames
        plan.n_conv = n_conv
        plan.zero_bwd, plan.zero_bwd_bytes = _vp(self.zero_bwd), self.zero_bwd.numel() * 4
        plan.zero_fwd, plan.zero_fwd_bytes = _vp(self.zero_fwd), self.zero_fwd.numel() * 8
        plan.l0_sums = _vp(self.zero_fwd[n_stats:])
        plan.n_gru = model.rnn.num_layers
        plan.nclass = model.nclass
        plan.bn_eps, plan.bn_momentum = 1e-3, 0.99
        plan.dropout_p = float(model.dropout.p)
        plan.scaler_eps = 1e-8
        T, F, cin = n_frames, n_mels, 1
        self.x0 = new(B, T, F)
        plan.x0 = _vp(self.x0)
        self.conv = []
        for i in range(n_conv):
            C = cnn.nb_filters[i]
            pt, pf = cnn.pooling[i]
            L = plan.conv[i]
            L.cin, L.cout, L.T, L.F, L.pt, L.pf = cin, C, T, F, pt, pf
            d = dict(
                # 16-channel layers: room for the paired-pixel tcgen05 packs (4x the plain pack, kernels.h conv_wpack_floats)
                wpack=new(2 * 9 * C * cin * (4 if 16 in (C, cin) else 1)) if i > 0 else None,
                gwpack=self.zero_bwd[gw_off[i]:gw_off[i] + gw_sizes[i]] if i > 0 else None,
                z=new(B, T, F, C), gy=new(B, T, F, C),
                out=new(B, T // pt, F // pf, C), gout=new(B, T // pt, F // pf, C),
                stats=self.zero_fwd[st_off[i]:st_off[i] + 4 * C], bn=new(4 * C, zero=True),
                # tcgen05 BN+GLU path of the 128-channel layers (include/sedk.h: glu_pack, lin)
                glu_pack=new(3 * 128 * 128 + 128, zero=True) if C in (64, 128) else None,
                lin=new(B, T, F, C) if C in (64, 128) else None)
            for k, v in d.items():
                setattr(L, k, _vp(v))
            pre = "cnn.cnn."
            L.gw = _vp(self.gviews[pre + "conv%d.weight" % i])
            L.gb = _vp(self.gviews[pre + "conv%d.bias" % i])
            L.ggamma = _vp(self.gviews[pre + "batchnorm%d.weight" % i])
            L.gbeta = _vp(self.gviews[pre + "batchnorm%d.bias" % i])
            gate = cnn.gate_name()            # "glu" / "cg" (ContextGating) or None for the plain (leaky) ReLU activations
            if gate is not None:
                L.gglu_w = _vp(self.gviews[pre + "%s%d.linear.weight" % (gate, i)])
                L.gglu_b = _vp(self.gviews[pre + "%s%d.linear.bias" % (gate, i)])
            self.conv.append(d)
            T, F, cin = T // pt, F // pf, C
        if F != 1:
            raise NotImplementedError("CRNN kernels need the CNN to pool the frequency axis to 1 (got %d); the "
                                      "reference only warns and flattens (CRNN.py:237-242)" % F)
        self.Tp, self.nb = T, cin
        in_dim = cin
        if emb_shape is None and model.dropstep_recurrent:
            # CRNN.py:295-301: dropstep mask + dropout on the CNN output, no embeddings
            self.cat_in = new(B, T, cin)
            self.gfused = new(B, T, cin)
            plan.cat_in, plan.gfused = _vp(self.cat_in), _vp(self.gfused)
        if emb_shape is not None:
            E, Te = emb_shape
            plan.emb_dim, plan.emb_T = E, Te
            plan.emb_mode = 1 if model.aggregation_type == "interpolate" else 0
            self.cat_in = new(B, T, cin + E)
            self.fused = new(B, T, cin)
            self.gfused = new(B, T, cin)
            plan.cat_in, plan.fused, plan.gfused = _vp(self.cat_in), _vp(self.fused), _vp(self.gfused)
            plan.gcat_w, plan.gcat_b = _vp(self.gviews["cat_tf.weight"]), _vp(self.gviews["cat_tf.bias"])
        H = model.rnn.n_hidden
        self.gru = []
        for l in range(model.rnn.num_layers):
            G = plan.gru[l]
            G.in_dim, G.hidden = in_dim, H
            d = dict(out=new(B, T, 2 * H), gout=new(B, T, 2 * H))
            G.out, G.gout = _vp(d["out"]), _vp(d["gout"])
            for di, suf in enumerate(("", "_reverse")):
                for nm, shape in (("gi", (B, T, 3 * H)), ("gates", (B, T, 4 * H)), ("hprev", (B, T, H)),
                                  ("dghn", (B, T, H))):
                    t = new(*shape)
                    d[nm + suf] = t
                    getattr(G, nm)[di] = t.data_ptr()
                for nm in ("w_ih", "w_hh", "b_ih", "b_hh"):
                    key = "rnn.rnn.%s_l%d%s" % (nm.replace("w_", "weight_").replace("b_", "bias_"), l, suf)
                    getattr(G, "g" + nm)[di] = self.gviews[key].data_ptr()
            self.gru.append(d)
            in_dim = 2 * H
        self.specaug_buf = new(B, 4, dtype=torch.int32, zero=True)
        self.dropstep_buf = new(B, 4, dtype=torch.int32, zero=True)
        self.rnn_drop = new(B, T, in_dim)
        self.grnn_drop = new(B, T, in_dim)
        self.sof = new(B, T, model.nclass)
        self.hsum = new(B, 2, model.nclass)
        plan.hsum = _vp(self.hsum)
        plan.rnn_drop, plan.grnn_drop, plan.sof = _vp(self.rnn_drop), _vp(self.grnn_drop), _vp(self.sof)
        plan.gdense_w, plan.gdense_b = _vp(self.gviews["dense.weight"]), _vp(self.gviews["dense.bias"])
        plan.gsoft_w, plan.gsoft_b = _vp(self.gviews["dense_softmax.weight"]), _vp(self.gviews["dense_softmax.bias"])
        self.plan = plan
        self.param_sig = None

    def bind_params(self, model):
        sig = tuple(p.data_ptr() for p in model.parameters()) + tuple(b.data_ptr() for b in model.buffers())
        if sig == self.param_sig:
            return
        plan = self.plan
        sd = dict(model.named_parameters())
        sd.update(dict(model.named_buffers()))
        for i in range(plan.n_conv):
            L = plan.conv[i]
            pre = "cnn.cnn."
            L.w, L.b = _vp(sd[pre + "conv%d.weight" % i]), _vp(sd[pre + "conv%d.bias" % i])
            L.gamma, L.beta = _vp(sd[pre + "batchnorm%d.weight" % i]), _vp(sd[pre + "batchnorm%d.bias" % i])
            L.running_mean = _vp(sd[pre + "batchnorm%d.running_mean" % i])
            L.running_var = _vp(sd[pre + "batchnorm%d.running_var" % i])
            L.num_batches = _vp(sd[pre + "batchnorm%d.num_batches_tracked" % i])
            gate = model.cnn.gate_name()
            if gate is not None:
                L.glu_w = _vp(sd[pre + "%s%d.linear.weight" % (gate, i)])
                L.glu_b = _vp(sd[pre + "%s%d.linear.bias" % (gate, i)])
        for l in range(plan.n_gru):
            G = plan.gru[l]
            for di, suf in enumerate(("", "_reverse")):
                G.w_ih[di] = sd["rnn.rnn.weight_ih_l%d%s" % (l, suf)].data_ptr()
                G.w_hh[di] = sd["rnn.rnn.weight_hh_l%d%s" % (l, suf)].data_ptr()
                G.b_ih[di] = sd["rnn.rnn.bias_ih_l%d%s" % (l, suf)].data_ptr()
                G.b_hh[di] = sd["rnn.rnn.bias_hh_l%d%s" % (l, suf)].data_ptr()
        plan.dense_w, plan.dense_b = _vp(sd["dense.weight"]), _vp(sd["dense.bias"])
        plan.soft_w, plan.soft_b = _vp(sd["dense_softmax.weight"]), _vp(sd["dense_softmax.bias"])
        if "cat_tf.weight" in sd:
            plan.cat_w, plan.cat_b = _vp(sd["cat_tf.weight"]), _vp(sd["cat_tf.bias"])
        for p in model.parameters():
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _lib.SedkError("CRNN parameters must be contiguous fp32 tensors")
        self.param_sig = sig


class _CRNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, ws, args, *params):
        strong, weak = model._launch_forward(ws, *args)
        ctx.model, ctx.ws, ctx.gen = model, ws, ws.generation
        return strong, weak

    @staticmethod
    def backward(ctx, gstrong, gweak):
        ws, model = ctx.ws, ctx.model
        if ws.generation != ctx.gen:
            raise RuntimeError("CRNN backward: the saved activations were overwritten by a later forward of the same "
                               "module and batch shape; run backward before the next forward")
        grads = model._launch_backward(ws, gstrong, gweak)
        return (None, None, None) + grads


class CRNN(nn.Module):
    def __init__(
        self,
        n_in_channel=1,
        nclass=10,
        attention=True,
        activation="glu",
        dropout=0.5,
        train_cnn=True,
        rnn_type="BGRU",
        n_RNN_cell=128,
        n_layers_RNN=2,
        dropout_recurrent=0,
        cnn_integration=False,
        freeze_bn=False,
        use_embeddings=False,
        embedding_size=527,
        embedding_type="global",
        frame_emb_enc_dim=512,
        aggregation_type="global",
        specaugm_t_p=0.2,
        specaugm_t_l=5,
        specaugm_f_p=0.2,
        specaugm_f_l=10,
        dropstep_recurrent=0.0,
        dropstep_recurrent_len=5,
        **kwargs,
    ):
        super(CRNN, self).__init__()
        self.n_in_channel = n_in_channel
        self.attention = attention
        self.cnn_integration = cnn_integration
        self.freeze_bn = freeze_bn
        self.use_embeddings = use_embeddings
        self.embedding_type = embedding_type
        self.aggregation_type = aggregation_type
        self.nclass = nclass
        self.dropstep_recurrent = dropstep_recurrent
        self.dropstep_recurrent_len = dropstep_recurrent_len
        self.specaugm_t_p = specaugm_t_p
        self.specaugm_t_l = specaugm_t_l
        self.specaugm_f_p = specaugm_f_p
        self.specaugm_f_l = specaugm_f_l
        self.precision = None           # None -> module default (set_default_precision)

        n_in_cnn = n_in_channel
        if cnn_integration:
            n_in_cnn = 1
        self.cnn = CNN(n_in_channel=n_in_cnn, activation=activation, conv_dropout=dropout, **kwargs)
        self.train_cnn = train_cnn
        if not train_cnn:
            for param in self.cnn.parameters():
                param.requires_grad = False
        self.rnn_type = rnn_type
        if rnn_type == "BGRU":
            nb_in = self.cnn.nb_filters[-1]
            if self.cnn_integration:
                nb_in = nb_in * n_in_channel
            self.rnn = BidirectionalGRU(n_in=nb_in, n_hidden=n_RNN_cell, dropout=dropout_recurrent,
                                        num_layers=n_layers_RNN)
        else:
            NotImplementedError("Only BGRU supported for CRNN for now")     # (sic) the reference does not raise here
        self.dropout = nn.Dropout(dropout)
        if isinstance(self.nclass, (tuple, list)) and len(self.nclass) > 1:
            # multiple heads: the reference itself fails here when attention is on (CRNN.py:113) - keep that behaviour
            self.dense = torch.nn.ModuleList([])
            self.sigmoid = nn.Sigmoid()
            self.softmax = nn.Softmax(dim=-1)
            for current_classes in self.nclass:
                self.dense.append(nn.Linear(n_RNN_cell * 2, current_classes))
                if self.attention:
                    self.dense_softmax.append(nn.Linear(n_RNN_cell * 2, current_classes))
        else:
            if isinstance(self.nclass, (tuple, list)):
                self.nclass = self.nclass[0]
            self.dense = nn.Linear(n_RNN_cell * 2, self.nclass)
            self.sigmoid = nn.Sigmoid()
            if self.attention:
                self.dense_softmax = nn.Linear(n_RNN_cell * 2, self.nclass)
                self.softmax = nn.Softmax(dim=-1)
        if self.use_embeddings:
            if self.aggregation_type == "frame":
                self.frame_embs_encoder = nn.GRU(batch_first=True, input_size=embedding_size, hidden_size=512,
                                                 bidirectional=True)
                self.shrink_emb = torch.nn.Sequential(torch.nn.Linear(2 * frame_emb_enc_dim, nb_in),
                                                      torch.nn.LayerNorm(nb_in))
                self.cat_tf = torch.nn.Linear(2 * nb_in, nb_in)
            elif self.aggregation_type == "global":
                self.shrink_emb = torch.nn.Sequential(torch.nn.Linear(embedding_size, nb_in), torch.nn.LayerNorm(nb_in))
                self.cat_tf = torch.nn.Linear(2 * nb_in, nb_in)
            elif self.aggregation_type == "interpolate":
                self.cat_tf = torch.nn.Linear(nb_in + embedding_size, nb_in)
            elif self.aggregation_type == "pool1d":
                self.cat_tf = torch.nn.Linear(nb_in + embedding_size, nb_in)
            else:
                self.cat_tf = torch.nn.Linear(2 * nb_in, nb_in)
        self.embedding_size = embedding_size
        self._ws = {}
        self._fwd_count = 0
        # deterministic per-instance salt of the dropout / SpecAugment seed (construction order; a deepcopy gets a new one):
        # two runs under the same seed_everything() draw the same masks, two models in one process draw different ones
        self._instance = CRNN._next_instance()
        self.seed_dev = None            # optional device uint64 counter added to the dropout seed (CUDA-graph replays)

    _instances = 0

    @staticmethod
    def _next_instance():
        CRNN._instances += 1
        return CRNN._instances

    def __getstate__(self):
        # workspaces hold ctypes structures full of device pointers: per-process scratch, never pickled / deep-copied
        # (torch.save(module), Lightning ddp_spawn); they are rebuilt on the first forward
        state = dict(self.__dict__)
        state["_ws"] = {}
        state.pop("grad_alloc", None)          # process-local symmetric-memory allocator (desed_task_b200/nvls.py)
        return state

    # ------------------------------------------------------------------------------------------------------------
    def _unsupported(self):
        r = self.cnn.unsupported_reason()
        if r:
            return r
        if self.rnn_type != "BGRU" or not hasattr(self, "rnn"):
            return "rnn_type=%r" % (self.rnn_type,)
        if self.rnn.n_hidden not in (64, 128, 192):
            return "n_RNN_cell=%d (kernels: 64, 128, 192)" % self.rnn.n_hidden
        if self.rnn.dropout:
            return "dropout_recurrent != 0"
        if self.cnn_integration or self.n_in_channel != 1:
            return "cnn_integration / multi-channel input"
        if self.attention not in (True, "legacy"):
            return "attention=False"
        if isinstance(self.nclass, (tuple, list)):
            return "multi-head nclass"
        if self.nclass > 32:
            return "nclass > 32"
        if self.use_embeddings and self.aggregation_type not in ("pool1d", "interpolate"):
            return "aggregation_type=%r (kernels implement 'pool1d' and 'interpolate')" % self.aggregation_type
        return None

    def _precision(self):
        return _default_precision if self.precision is None else self.precision

    def _workspace(self, B, n_mels, n_frames, device, emb_shape):
        key = (B, n_mels, n_frames, str(device), emb_shape)
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 4:
                self._ws.clear()
            ws = _Workspace(self, B, n_mels, n_frames, device, emb_shape)
            self._ws[key] = ws
        return ws

    def __deepcopy__(self, memo):
        # workspaces are per-module scratch: never copied (SEDTask4 deep-copies the student into the teacher)
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        import copy
        for k, v in self.__dict__.items():
            if k == "grad_alloc":
                continue
            new.__dict__[k] = {} if k == "_ws" else copy.deepcopy(v, memo)
        new._instance = CRNN._next_instance()
        return new

    @staticmethod
    def _span_param(size, mask_param, p):
        """torchaudio mask_along_axis_iid (functional.py:857-859): the effective mask parameter; < 1 disables the mask."""
        return mask_param if p == 1.0 else min(mask_param, int(size * p))

    def _spans(self, buf, B, size_a, param_a, size_b, param_b, seed, stream_id):
        """One kernel draws both spans of every example (sedk_mask_spans); None when both masks are disabled."""
        if param_a < 1 and param_b < 1:
            return None
        check(lib().sedk_mask_spans(ptr(buf), B, size_a, param_a, size_b, param_b, seed,
                                    _vp(getattr(self, "seed_dev", None)), stream_id, stream_ptr()), "sedk_mask_spans")
        return buf

    def _specaug_spans(self, ws, B, n_mels, n_frames, seed):
        """CRNN.apply_specaugment (CRNN.py:207-219): 'freq' mask first, then time mask."""
        return self._spans(ws.specaug_buf, B, n_mels, self._span_param(n_mels, self.specaugm_f_l, self.specaugm_f_p),
                           n_frames, self._span_param(n_frames, self.specaugm_t_l, self.specaugm_t_p), seed, 300)

    def _dropstep_spans(self, ws, B, frames, seed):
        """CRNN.py:288-293: TimeMasking(dropstep_len, iid, p) on x, then on the embeddings."""
        prm = self._span_param(frames, self.dropstep_recurrent_len, self.dropstep_recurrent)
        return self._spans(ws.dropstep_buf, B, frames, prm, frames, prm, seed, 301)

    # ------------------------------------------------------------------------------------------------------------
    def _launch_forward(self, ws, x, minmax, embeddings, classes_mask, specaug, dropstep, training, seed, save=None):
        plan = ws.plan
        ws.bind_params(self)
        ws.generation += 1
        B = ws.B
        # plan.training = "save what the backward needs"; dropout / SpecAugment follow the module's mode, BatchNorm uses
        # the running statistics in eval mode and under freeze_bn (CRNN.py:308-323: BatchNorm2d.eval() inside train())
        save = training if save is None else save
        plan.training = 1 if (training or save) else 0
        plan.bn_eval = 1 if (not training or self.freeze_bn) else 0
        plan.precision = self._precision()
        plan.activation = self.cnn.activation_code()
        plan.seed = seed
        plan.seed_dev = _vp(getattr(self, "seed_dev", None))
        plan.dropout_p = float(self.dropout.p) if training else 0.0
        plan.x = _vp(x)
        plan.x_sb, plan.x_sm, plan.x_st = x.stride(0), x.stride(1), x.stride(2)
        plan.minmax = _vp(minmax)
        plan.specaug = _vp(specaug)
        plan.dropstep = _vp(dropstep)
        plan.emb = _vp(embeddings)
        plan.classes_mask = _vp(classes_mask)
        strong = torch.empty(B, self.nclass, ws.Tp, device=x.device, dtype=torch.float32)
        weak = torch.empty(B, self.nclass, device=x.device, dtype=torch.float32)
        plan.strong, plan.weak = _vp(strong), _vp(weak)
        ws.keep = (x, minmax, embeddings, classes_mask, specaug, dropstep, strong, weak)
        check(lib().sedk_crnn_forward(ctypes.byref(plan), stream_ptr()), "sedk_crnn_forward")
        return strong, weak

    def _launch_backward(self, ws, gstrong, gweak):
        plan = ws.plan
        # parameters may have been re-pointed since the forward (flatten_parameters on the first EMA/optimizer call
        # copies them into one flat buffer and frees the old storages): never read weights through stale pointers
        ws.bind_params(self)
        gs = gstrong.float().contiguous() if gstrong is not None else None
        gw = gweak.float().contiguous() if gweak is not None else None
        plan.gstrong, plan.gweak = _vp(gs), _vp(gw)
        check(lib().sedk_crnn_backward(ctypes.byref(plan), stream_ptr()), "sedk_crnn_backward")
        out = []
        for (n, p) in self.named_parameters():
            out.append(ws.gviews[n].clone() if p.requires_grad else None)
        return tuple(out)

    def run(self, x, minmax=None, embeddings=None, classes_mask=None, pad_mask=None, autograd=None):
        """x: [B, n_mels, T] cuda fp32 (any strides).  With `minmax` (uint32 [B,2] from the front end) x is the
        UN-scaled log-mel and the instance min-max scaler is applied inside the first conv kernel."""
        reason = self._unsupported()
        if reason:
            raise NotImplementedError("desed_task_b200 CRNN kernels do not cover this configuration: " + reason)
        if pad_mask is not None:
            raise NotImplementedError("pad_mask is not supported (every reference call site passes None)")
        require_cuda(x, embeddings, classes_mask)
        if x.dim() != 3:
            raise ValueError("CRNN expects [batch, n_mels, frames] input, got %s" % (tuple(x.shape),))
        x = x.float()
        B, n_mels, n_frames = x.shape
        emb_shape = None
        if self.use_embeddings:
            if embeddings is None:
                raise ValueError("use_embeddings=True but no embeddings were given")
            if embeddings.dtype == torch.bfloat16:          # the pre-pooled bf16 storage format (desed_task_b200.embeddings)
                from ..embeddings import upcast
                embeddings = upcast(embeddings)
            embeddings = embeddings.float().contiguous()
            emb_shape = (embeddings.shape[1], embeddings.shape[2])
        else:
            embeddings = None
        if classes_mask is not None:
            classes_mask = classes_mask.to(torch.bool).contiguous()
        ws = self._workspace(B, n_mels, n_frames, x.device, emb_shape)
        training = self.training
        self._fwd_count += 1
        seed = (torch.initial_seed() * 1000003 + self._fwd_count * 7919 + self._instance * 104729) & 0xFFFFFFFFFFFFFFFF
        specaug = self._specaug_spans(ws, B, n_mels, n_frames, seed) if training else None
        dropstep = None
        if training and self.dropstep_recurrent:
            # with embeddings: spans for x and for the embeddings (CRNN.py:288-293); without: the x span only (:295-301)
            dropstep = self._dropstep_spans(ws, B, ws.Tp, seed)
        # autograd works in both modes, as in the reference: an eval-mode forward with grad enabled (BN-frozen fine-tuning,
        # saliency, THOP-style profiling) saves its activations and backpropagates through running-statistics BatchNorm
        want_grad = torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if autograd is not None:
            want_grad = autograd
        args = (x, minmax, embeddings, classes_mask, specaug, dropstep, training, seed, training or want_grad)
        if want_grad:
            return _CRNNFunction.apply(self, ws, args, *list(self.parameters()))
        return self._launch_forward(ws, *args)

    def forward(self, x, pad_mask=None, embeddings=None, classes_mask=None):
        return self.run(x, None, embeddings, classes_mask, pad_mask)

    # ---- autograd-free path used by the fused trainer (gradients stay in the workspace's flat buffer)
    def forward_direct(self, x, minmax=None, embeddings=None, classes_mask=None):
        strong, weak = self.run(x, minmax, embeddings, classes_mask, autograd=False)
        key = (x.shape[0], x.shape[1], x.shape[2], str(x.device),
               None if not self.use_embeddings else (embeddings.shape[1], embeddings.shape[2]))
        return strong, weak, self._ws[key]

    def backward_direct(self, ws, gstrong, gweak, phases=3):
        """Runs sedk_crnn_backward (phases = 3) or one half of it (1: heads + BiGRU + fusion, 2: CNN; include/sedk.h);
        returns the flat gradient buffer (parameters() order, overwritten each call)."""
        plan = ws.plan
        ws.bind_params(self)
        plan.gstrong, plan.gweak = _vp(gstrong), _vp(gweak)
        ws.keep_g = (gstrong, gweak)
        check(lib().sedk_crnn_backward_phase(ctypes.byref(plan), int(phases), stream_ptr()), "sedk_crnn_backward_phase")
        return ws.gflat

    def cnn_param_count(self):
        """Number of leading entries of the flat parameter / gradient buffers (optim.flat_layout) that belong to the CNN
        (parameters() order: cnn, rnn, dense, dense_softmax, [cat_tf])."""
        from ..optim import flat_layout
        return flat_layout([p.numel() for p in self.cnn.parameters()])[1]

    def cnn_lower_param_count(self, n_layers=3):
        """Flat-buffer entries of the first `n_layers` conv blocks (phase 8 of sedk_crnn_backward_phase)."""
        from ..optim import flat_layout
        names = ("conv%d.", "batchnorm%d.", "glu%d.", "cg%d.", "layernorm%d.")
        sizes = [p.numel() for n, p in self.cnn.named_parameters()]
        keep = [any(("cnn." + k % i) in n for i in range(n_layers) for k in names) for n, _ in self.cnn.named_parameters()]
        # parameters() order is layer by layer, so the lower layers are a prefix
        assert keep == sorted(keep, reverse=True)
        return flat_layout(sizes[:sum(keep)])[1]

    def train(self, mode=True):
        """Override the default train() to freeze the BN parameters (CRNN.py:308-323; returns None like the reference)."""
        super(CRNN, self).train(mode)
        if self.freeze_bn:
            print("Freezing Mean/Var of BatchNorm2D.")
            if self.freeze_bn:
                print("Freezing Weight/Bias of BatchNorm2D.")
        if self.freeze_bn:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
                    if self.freeze_bn:
                        m.weight.requires_grad = False
                        m.bias.requires_grad = False
