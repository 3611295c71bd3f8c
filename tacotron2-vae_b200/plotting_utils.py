"""Figure helpers for the validation log (reference plotting_utils.py); need matplotlib, which is optional here."""
import matplotlib
matplotlib.use("Agg")
import matplotlib.pylab as plt      # noqa: E402
import numpy as np                  # noqa: E402


def _to_numpy(fig):
    fig.canvas.draw()
    data = np.asarray(fig.canvas.buffer_rgba())[..., :3]
    plt.close(fig)
    return data.transpose(2, 0, 1)


def plot_alignment_to_numpy(alignment, info=None):
    fig, ax = plt.subplots(figsize=(6, 4))
    im = ax.imshow(alignment, aspect="auto", origin="lower", interpolation="none")
    fig.colorbar(im, ax=ax)
    ax.set_xlabel("Decoder timestep" + ("\n\n" + info if info else ""))
    ax.set_ylabel("Encoder timestep")
    return _to_numpy(fig)


def plot_spectrogram_to_numpy(spectrogram):
    fig, ax = plt.subplots(figsize=(12, 3))
    im = ax.imshow(spectrogram, aspect="auto", origin="lower", interpolation="none")
    fig.colorbar(im, ax=ax)
    ax.set_xlabel("Frames")
    ax.set_ylabel("Channels")
    return _to_numpy(fig)


def plot_gate_outputs_to_numpy(gate_targets, gate_outputs):
    fig, ax = plt.subplots(figsize=(12, 3))
    ax.scatter(range(len(gate_targets)), gate_targets, alpha=0.5, color="green", marker="+", s=1, label="target")
    ax.scatter(range(len(gate_outputs)), gate_outputs, alpha=0.5, color="red", marker=".", s=1, label="predicted")
    ax.set_xlabel("Frames (Green target, Red predicted)")
    ax.set_ylabel("Gate State")
    return _to_numpy(fig)


def plot_scatter(mus, y):
    """first two latent means of the batch, one colour per emotion class (reference plotting_utils.py:63-83 -> "latent_dim")"""
    colors, labels = ("r", "b", "g", "y"), ("neu", "sad", "ang", "hap")
    mus = mus.detach().cpu().numpy()
    cls = np.argmax(y.detach().cpu().numpy(), 1)
    fig, ax = plt.subplots(figsize=(12, 12))
    for i, (c, label) in enumerate(zip(colors, labels)):
        ax.scatter(mus[cls == i, 0], mus[cls == i, 1], c=c, label=label, alpha=0.5)
    ax.legend(loc="upper left")
    return _to_numpy(fig)
