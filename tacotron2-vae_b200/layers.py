"""Parameter holders + thin callables with the reference's names (reference layers.py:7-92).

`LinearNorm` / `ConvNorm` keep the attribute paths that define the checkpoint keys (`...linear_layer.weight`,
`...conv.weight`) and the reference's Xavier initialisation, but own no torch compute: calling them runs the
t2v kernels.  `TacotronSTFT` is the mel front-end (layers.py:54-92) on the t2v STFT/mel kernels."""
import math

import numpy as np
import torch
from torch import nn

from t2v import frontend as _frontend
from t2v import infer as _infer

_GAINS = {"linear": 1.0, "sigmoid": 1.0, "tanh": 5.0 / 3.0, "relu": math.sqrt(2.0)}


def _xavier_uniform_(w, gain_name):
    fan_out = w.shape[0] * int(np.prod(w.shape[2:])) if w.dim() > 2 else w.shape[0]
    fan_in = int(np.prod(w.shape[1:]))
    bound = _GAINS[gain_name] * math.sqrt(6.0 / (fan_in + fan_out))
    with torch.no_grad():
        w.uniform_(-bound, bound)


class _Affine(nn.Module):
    """weight [+ bias] holder named like nn.Linear / nn.Conv1d (default bias init U(+-1/sqrt(fan_in)))."""

    def __init__(self, w_shape, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.empty(*w_shape))
        fan_in = int(np.prod(w_shape[1:]))
        bound = 1.0 / math.sqrt(fan_in)
        with torch.no_grad():
            self.weight.uniform_(-bound, bound)
        if bias:
            self.bias = nn.Parameter(torch.empty(w_shape[0]).uniform_(-bound, bound))
        else:
            self.register_parameter("bias", None)

    def forward(self, x):                     # nn.Linear semantics on the last dim (used for vae_gst.fc3 etc.)
        return _infer.linear(x, self.weight, self.bias)


class LinearNorm(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, w_init_gain="linear"):
        super().__init__()
        self.linear_layer = _Affine((out_dim, in_dim), bias)
        _xavier_uniform_(self.linear_layer.weight, w_init_gain)

    def forward(self, x):
        return self.linear_layer(x)


class ConvNorm(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=1, stride=1, padding=None, dilation=1, bias=True,
                 w_init_gain="linear"):
        super().__init__()
        if padding is None:
            assert kernel_size % 2 == 1
            padding = dilation * (kernel_size - 1) // 2
        assert stride == 1 and dilation == 1, "only the stride-1 convolutions of the hot path are built"
        self.kernel_size, self.padding = kernel_size, padding
        self.conv = _Affine((out_channels, in_channels, kernel_size), bias)
        _xavier_uniform_(self.conv.weight, w_init_gain)

    def forward(self, signal):
        raise RuntimeError("ConvNorm is a parameter holder; convolutions run inside the fused t2v stacks")


class BatchNormParams(nn.Module):
    """state of nn.BatchNorm1d/2d (eps 1e-5, momentum .1): the t2v kernels read/update these buffers in place."""

    def __init__(self, num_features):
        super().__init__()
        self.num_features = num_features
        self.weight = nn.Parameter(torch.ones(num_features))
        self.bias = nn.Parameter(torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class TacotronSTFT(nn.Module):
    """wav [B,S] in [-1,1] -> log-mel [B,n_mel,S//hop+1] (reference layers.py:54-92, stft.py:77-105)."""

    def __init__(self, filter_length=1024, hop_length=256, win_length=1024, n_mel_channels=80, sampling_rate=22050,
                 mel_fmin=0.0, mel_fmax=8000.0):
        super().__init__()
        assert win_length == filter_length, "the reference recipe uses win_length == filter_length"
        self.n_mel_channels = n_mel_channels
        self.sampling_rate = sampling_rate
        self.filter_length, self.hop_length, self.win_length = filter_length, hop_length, win_length
        self.stft_fn = _frontend.STFT(filter_length, hop_length, win_length)
        mel_basis = _frontend.slaney_mel_filterbank(sampling_rate, filter_length, n_mel_channels, mel_fmin, mel_fmax)
        self.register_buffer("mel_basis", torch.from_numpy(mel_basis).float())

    def mel_spectrogram(self, y, ref_level_db=20, magnitude_power=1.5):
        assert float(y.min()) >= -1 and float(y.max()) <= 1
        return _frontend.mel_spectrogram(y, self.stft_fn, self.mel_basis)
