"""Parameter containers mirroring desed_task/nnet/CNN.py:5-114 (same constructor arguments, module names and therefore
state_dict keys: cnn.conv{i}, cnn.batchnorm{i}, cnn.glu{i}.linear, ...).  The torch.nn modules here only HOLD the
parameters / buffers with the reference's initialisation; the arithmetic runs in libsedk (csrc/conv.cu, csrc/bnglu.cu)
through desed_task_b200.nnet.CRNN.  Calling these containers directly is not supported (there is no PyTorch fallback)."""
import torch.nn as nn


class _KernelOnly(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(
            "%s is a parameter container: its arithmetic is fused into the sm_100a CRNN kernels and is reachable through "
            "desed_task_b200.nnet.CRNN.CRNN only (no stand-alone PyTorch fallback)" % type(self).__name__)


class GLU(_KernelOnly):
    def __init__(self, input_num):
        super(GLU, self).__init__()
        self.sigmoid = nn.Sigmoid()
        self.linear = nn.Linear(input_num, input_num)


class ContextGating(_KernelOnly):
    def __init__(self, input_num):
        super(ContextGating, self).__init__()
        self.sigmoid = nn.Sigmoid()
        self.linear = nn.Linear(input_num, input_num)


class CNN(_KernelOnly):
    def __init__(self, n_in_channel, activation="Relu", conv_dropout=0, kernel_size=[3, 3, 3], padding=[1, 1, 1],
                 stride=[1, 1, 1], nb_filters=[64, 64, 64], pooling=[(1, 4), (1, 4), (1, 4)], normalization="batch",
                 **transformer_kwargs):
        super(CNN, self).__init__()
        self.nb_filters = nb_filters
        self.n_in_channel = n_in_channel
        self.activation = activation
        self.conv_dropout = conv_dropout
        self.kernel_size, self.padding, self.stride = list(kernel_size), list(padding), list(stride)
        self.pooling = [tuple(p) if isinstance(p, (list, tuple)) else (p, p) for p in pooling]
        self.normalization = normalization
        cnn = nn.Sequential()

        def conv(i, normalization="batch", dropout=None, activ="relu"):
            nIn = n_in_channel if i == 0 else nb_filters[i - 1]
            nOut = nb_filters[i]
            cnn.add_module("conv{0}".format(i), nn.Conv2d(nIn, nOut, kernel_size[i], stride[i], padding[i]))
            if normalization == "batch":
                cnn.add_module("batchnorm{0}".format(i), nn.BatchNorm2d(nOut, eps=0.001, momentum=0.99))
            elif normalization == "layer":
                cnn.add_module("layernorm{0}".format(i), nn.GroupNorm(1, nOut))
            if activ.lower() == "leakyrelu":
                cnn.add_module("relu{0}".format(i), nn.LeakyReLU(0.2))
            elif activ.lower() == "relu":
                cnn.add_module("relu{0}".format(i), nn.ReLU())
            elif activ.lower() == "glu":
                cnn.add_module("glu{0}".format(i), GLU(nOut))
            elif activ.lower() == "cg":
                cnn.add_module("cg{0}".format(i), ContextGating(nOut))
            if dropout is not None:
                cnn.add_module("dropout{0}".format(i), nn.Dropout(dropout))

        for i in range(len(nb_filters)):
            conv(i, normalization=normalization, dropout=conv_dropout, activ=activation)
            cnn.add_module("pooling{0}".format(i), nn.AvgPool2d(pooling[i]))
        self.cnn = cnn

    _ACT = {"glu": 0, "cg": 1, "relu": 2, "leakyrelu": 3}

    def activation_code(self):
        """sedk_crnn_plan.activation (CNN.py:81-88); an unknown name adds no activation module in the reference either."""
        return self._ACT.get(self.activation.lower(), -1)

    def gate_name(self):
        a = self.activation.lower()
        return a if a in ("glu", "cg") else None

    def unsupported_reason(self):
        """None if the sm_100a kernels cover this configuration, else a human-readable reason."""
        if self.activation.lower() not in self._ACT:
            return "activation=%r (kernels implement 'glu', 'cg', 'relu', 'leakyrelu')" % self.activation
        if self.normalization != "batch":
            return "normalization=%r (kernels implement 'batch')" % self.normalization
        if self.n_in_channel != 1:
            return "n_in_channel=%d (kernels implement 1)" % self.n_in_channel
        n = len(self.nb_filters)
        if n > 8:
            return "more than 8 conv layers"
        if any(k != 3 for k in self.kernel_size[:n]) or any(p != 1 for p in self.padding[:n]) or \
                any(s != 1 for s in self.stride[:n]):
            return "kernel_size/padding/stride other than 3/1/1"
        if self.nb_filters[0] not in (16, 32, 64):
            return "first layer width %d not in {16,32,64}" % self.nb_filters[0]
        ok_pairs = {(16, 32), (32, 64), (64, 128), (128, 128)}
        for a, b in zip(self.nb_filters[:-1], self.nb_filters[1:]):
            if (a, b) not in ok_pairs:
                return "channel step %d->%d has no kernel instantiation (supported: 16->32, 32->64, 64->128, 128->128)" % (a, b)
        for p in self.pooling[:n]:
            if p[0] not in (1, 2) or p[1] not in (1, 2):
                return "pooling %s (kernels implement factors 1 and 2)" % (p,)
        return None
