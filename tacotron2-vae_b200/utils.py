"""Host-side helpers with the reference's names (reference utils.py:9-44)."""
import os

import numpy as np
import torch

max_wav_value = 32768.0


def get_mask_from_lengths(lengths):
    """True where position < length.  The reference builds a uint8 mask on torch.cuda.LongTensor and inverts it
    with `~` (utils.py:9-13, model.py:412,511), which is logical-not only on PyTorch<=1.1; this is the bool form."""
    max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device, dtype=lengths.dtype)
    return ids.unsqueeze(0) < lengths.unsqueeze(1)


def load_wav_to_torch(full_path):
    from scipy.io.wavfile import read
    sampling_rate, data = read(full_path)
    return torch.from_numpy(data.astype(np.float32)), sampling_rate


def load_filepaths_and_text(filename, split="|"):
    with open(filename, encoding="utf-8") as f:
        return [line.strip().split(split) for line in f]


def to_gpu(x):
    x = x.contiguous()
    if torch.cuda.is_available():
        x = x.cuda(non_blocking=True)
    return x


def str2bool(v):
    return v.lower() in ("true", "1")


def makedirs(path):
    if not os.path.exists(path):
        os.makedirs(path)


def add_postfix(path, postfix):
    stem, ext = path.rsplit(".", 1)
    return "{}.{}.{}".format(stem, postfix, ext)
