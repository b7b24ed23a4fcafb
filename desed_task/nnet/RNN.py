from desed_task_b200.nnet.RNN import BidirectionalGRU, BidirectionalLSTM  # noqa: F401
