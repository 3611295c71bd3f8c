"""VAE / GST reference encoder surface (reference modules.py:8-85) on the t2v engine."""
import torch
from torch import nn

from CoordConv import Conv2dParams, CoordConv2d
from layers import BatchNormParams, _Affine
from t2v import infer as _infer


class _GRUParams(nn.Module):
    def __init__(self, input_size, hidden_size):
        super().__init__()
        k = 1.0 / (hidden_size ** 0.5)
        for name, shape in (("weight_ih_l0", (3 * hidden_size, input_size)), ("weight_hh_l0", (3 * hidden_size, hidden_size)),
                            ("bias_ih_l0", (3 * hidden_size,)), ("bias_hh_l0", (3 * hidden_size,))):
            setattr(self, name, nn.Parameter(torch.empty(*shape).uniform_(-k, k)))


class ReferenceEncoder(nn.Module):
    """[N, 80, T] mel -> [N, ref_enc_gru_size]; 6 x (3x3 stride-2 conv, BN, ReLU) + GRU (modules.py:34-85)."""

    def __init__(self, hparams):
        super().__init__()
        f = [1] + list(hparams.ref_enc_filters)
        K = len(hparams.ref_enc_filters)
        convs = [CoordConv2d(f[0], f[1], (3, 3), stride=(2, 2), padding=(1, 1), with_r=True)]
        convs += [Conv2dParams(f[i], f[i + 1], (3, 3)) for i in range(1, K)]
        self.convs = nn.ModuleList(convs)
        self.bns = nn.ModuleList([BatchNormParams(c) for c in hparams.ref_enc_filters])
        width = hparams.n_mel_channels
        for _ in range(K):
            width = (width - 3 + 2) // 2 + 1
        self.gru = _GRUParams(hparams.ref_enc_filters[-1] * width, hparams.E // 2)
        self.n_mels = hparams.n_mel_channels


class VAE_GST(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.ref_encoder = ReferenceEncoder(hparams)
        self.fc1 = _Affine((hparams.z_latent_dim, hparams.ref_enc_gru_size))
        self.fc2 = _Affine((hparams.z_latent_dim, hparams.ref_enc_gru_size))
        self.fc3 = _Affine((hparams.E, hparams.z_latent_dim))
        self._root = None          # set by Tacotron2 (the engine addresses parameters by their full names)

    def forward(self, inputs):
        """-> (style_embed [N,E], mu, logvar, z); z = mu in eval mode (modules.py:16-31)."""
        root = self._root()
        return _infer.vae_gst(root._ops(), root._state(), inputs, self.training)
