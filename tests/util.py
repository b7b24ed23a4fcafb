import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def gen_wave(seed, B, L=160000, sigma=0.1):
    """Same generator as oracle/make_golden.py::gen_wave."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, L, generator=g) * sigma


def maxdiff(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()
