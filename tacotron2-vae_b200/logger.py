"""Tacotron2Logger with the reference's scalar/image names (reference logger.py:8-56) on torch.utils.tensorboard
(tensorboardX / matplotlib are not part of this environment; images are logged only when matplotlib is importable)."""
import random

import torch

try:
    from torch.utils.tensorboard import SummaryWriter
except Exception:       # noqa: BLE001
    class SummaryWriter(object):
        def __init__(self, *a, **k):
            self.scalars = []

        def add_scalar(self, tag, value, step):
            self.scalars.append((tag, float(value), step))

        def add_histogram(self, *a, **k):
            pass

        def add_image(self, *a, **k):
            pass


class Tacotron2Logger(SummaryWriter):
    def __init__(self, logdir):
        super().__init__(logdir)

    def log_training(self, reduced_loss, grad_norm, learning_rate, duration, recon_loss, kl_div, kl_weight, iteration):
        for tag, v in (("training.loss", reduced_loss), ("grad.norm", grad_norm), ("learning.rate", learning_rate),
                       ("duration", duration), ("kl_div", kl_div), ("kl_weight", kl_weight), ("recon_loss", recon_loss)):
            self.add_scalar(tag, float(v), iteration)

    def log_validation(self, reduced_loss, model, y, y_pred, iteration):
        self.add_scalar("validation.loss", reduced_loss, iteration)
        _, mel_outputs, gate_outputs, alignments, mus, _, _, emotions = y_pred       # order pinned by model.forward
        for tag, value in model.named_parameters():
            self.add_histogram(tag.replace(".", "/"), value.detach().float().cpu().numpy(), iteration)
        try:
            from plotting_utils import (plot_alignment_to_numpy, plot_gate_outputs_to_numpy, plot_scatter,
                                        plot_spectrogram_to_numpy)
        except Exception:   # noqa: BLE001
            return
        mel_targets, gate_targets = y
        idx = random.randint(0, alignments.size(0) - 1)
        self.add_image("alignment", plot_alignment_to_numpy(alignments[idx].detach().cpu().numpy().T), iteration)
        self.add_image("mel_target", plot_spectrogram_to_numpy(mel_targets[idx].detach().cpu().numpy()), iteration)
        self.add_image("mel_predicted", plot_spectrogram_to_numpy(mel_outputs[idx].detach().cpu().numpy()), iteration)
        self.add_image("gate", plot_gate_outputs_to_numpy(gate_targets[idx].detach().cpu().numpy(),
                                                          torch.sigmoid(gate_outputs[idx]).detach().cpu().numpy()), iteration)
        self.add_image("latent_dim", plot_scatter(mus, emotions), iteration)
