"""Tacotron2Loss_VAE (reference loss_function.py:6-45) on the t2v loss kernels."""
import numpy as np
from torch import nn

from t2v.functions import VaeLossFunction


class Tacotron2Loss_VAE(nn.Module):
    def __init__(self, hparams):
        super().__init__()
        self.anneal_function = hparams.anneal_function
        self.lag, self.k, self.x0, self.upper = hparams.anneal_lag, hparams.anneal_k, hparams.anneal_x0, hparams.anneal_upper

    def kl_anneal_function(self, anneal_function, lag, step, k, x0, upper):
        if anneal_function == "logistic":
            return float(upper / (upper + np.exp(-k * (step - x0))))
        if anneal_function == "linear":
            return min(upper, step / x0) if step > lag else 0
        if anneal_function == "constant":
            return 0.001

    def forward(self, model_output, targets, step):
        mel_target, gate_target = targets[0], targets[1]
        mel_out, mel_out_postnet, gate_out, _, mu, logvar, _, _ = model_output
        kl_weight = self.kl_anneal_function(self.anneal_function, self.lag, step, self.k, self.x0, self.upper)
        total, recon, kl = VaeLossFunction.apply(mel_out, mel_out_postnet, gate_out, mu, logvar, mel_target.detach(),
                                                 gate_target.detach(), kl_weight)
        return total, recon, kl, kl_weight
