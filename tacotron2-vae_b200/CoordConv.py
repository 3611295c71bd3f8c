"""CoordConv parameter holder (reference CoordConv.py:142-161).  The coordinate planes (CoordConv.py:37-74, rank 2,
with_r) are generated inside the t2v im2col kernel; this class only reproduces the parameter layout, including the
dead Conv2d weight/bias the reference class owns by subclassing nn.Conv2d (quirk Q6)."""
import math

import torch
from torch import nn


class Conv2dParams(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=(3, 3)):
        super().__init__()
        kh, kw = kernel_size
        bound = 1.0 / math.sqrt(in_channels * kh * kw)
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels, kh, kw).uniform_(-bound, bound))
        self.bias = nn.Parameter(torch.empty(out_channels).uniform_(-bound, bound))


class CoordConv2d(Conv2dParams):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, with_r=False):
        super().__init__(in_channels, out_channels, kernel_size)            # dead parameters, kept for checkpoints
        self.rank = 2
        self.with_r = with_r
        self.conv = Conv2dParams(in_channels + 2 + int(with_r), out_channels, kernel_size)
