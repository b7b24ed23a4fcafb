"""Parameter containers mirroring desed_task/nnet/RNN.py:7-53 (nn.GRU / nn.LSTM hold the weights under the reference's
state_dict keys `rnn.weight_ih_l0[_reverse]`, ...).  The recurrence itself runs in csrc/gru.cu."""
from torch import nn as nn

from .CNN import _KernelOnly


class BidirectionalGRU(_KernelOnly):
    def __init__(self, n_in, n_hidden, dropout=0, num_layers=1):
        super(BidirectionalGRU, self).__init__()
        self.n_in, self.n_hidden, self.dropout, self.num_layers = n_in, n_hidden, dropout, num_layers
        self.rnn = nn.GRU(n_in, n_hidden, bidirectional=True, dropout=dropout, batch_first=True, num_layers=num_layers)


class BidirectionalLSTM(_KernelOnly):
    def __init__(self, nIn, nHidden, nOut, dropout=0, num_layers=1):
        super(BidirectionalLSTM, self).__init__()
        self.rnn = nn.LSTM(nIn, nHidden // 2, bidirectional=True, batch_first=True, dropout=dropout,
                           num_layers=num_layers)
        self.embedding = nn.Linear(nHidden * 2, nOut)
