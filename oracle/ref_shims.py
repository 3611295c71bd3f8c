"""TEST INFRASTRUCTURE ONLY -- loads the *unmodified* reference from /root/reference.

Nothing under ``oracle/`` is imported by the product path (``tacotron2-vae_b200/``); only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may use it.  This file exists only in the build container: ``/root/reference`` does not
travel to the GPU box, so everything here is used to (a) validate ``oracle/port.py`` against
the real reference and (b) generate the committed fixtures under ``tests/golden/``.

The reference targets PyTorch 1.0 / TF-contrib / librosa 0.6 (requirements.txt).  To import
its modules unmodified on torch 2.11 we install ``sys.modules`` shims for the absent
third-party packages and three runtime patches for PyTorch-1.0-isms:

  * ``tensorflow.contrib.training.HParams``      -> the HParams stand-in of this repo
  * ``librosa.filters.mel`` / ``librosa.util.*`` -> Slaney filterbank restatement (librosa 0.6.0
    defaults: htk=False, norm=1 -- parity unpinned at this boundary, no matrix is stored
    in the reference)
  * ``jamo``, ``unidecode``, ``inflect``, ``nltk`` -> minimal stand-ins (text path)
  * ``Tensor.cuda`` / ``Module.cuda`` -> identity on CPU (CoordConv.py:62-65 calls .cuda()
    unconditionally); ``get_mask_from_lengths`` -> bool mask on CPU (utils.py:9-13 allocates
    torch.cuda.LongTensor and returns uint8, which ``~`` no longer treats as logical-not)
"""
import importlib
import importlib.machinery
import os
import sys
import types

import numpy as np
import torch

REFERENCE_ROOT = os.environ.get("T2V_REFERENCE_ROOT", "/root/reference")
_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_PKG = os.path.join(_REPO, "tacotron2-vae_b200")


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "model.py"))


# --------------------------------------------------------------------------- librosa
def slaney_mel_filterbank(sr, n_fft, n_mels=128, fmin=0.0, fmax=None):
    """librosa 0.6.0 ``filters.mel`` defaults (htk=False, norm=1); called positionally at
    reference layers.py:62-63."""
    if fmax is None:
        fmax = float(sr) / 2

    def hz_to_mel(f):
        f = np.asanyarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asanyarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        min_log_hz = 1000.0
        min_log_mel = min_log_hz / f_sp
        logstep = np.log(6.4) / 27.0
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    n_mels = int(n_mels)
    weights = np.zeros((n_mels, 1 + n_fft // 2))
    fftfreqs = np.linspace(0, float(sr) / 2, 1 + n_fft // 2, endpoint=True)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = np.subtract.outer(mel_f, fftfreqs)
    for i in range(n_mels):
        lower = -ramps[i] / fdiff[i]
        upper = ramps[i + 2] / fdiff[i + 1]
        weights[i] = np.maximum(0, np.minimum(lower, upper))
    enorm = 2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels])
    weights *= enorm[:, np.newaxis]
    return weights


def _pad_center(data, size, axis=-1, **kw):
    n = data.shape[axis]
    lpad = int((size - n) // 2)
    lengths = [(0, 0)] * data.ndim
    lengths[axis] = (lpad, int(size - n - lpad))
    return np.pad(data, lengths, mode="constant")


def _module(name):
    """stub module with a __spec__ (importlib.util.find_spec raises on modules without one; torch._dynamo probes 'tensorflow')"""
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    return m


def _install_shims():
    if "tensorflow" not in sys.modules:
        sys.path.insert(0, _PKG)
        hp_mod = importlib.import_module("hparams")  # this repo's HParams stand-in
        sys.path.remove(_PKG)
        del sys.modules["hparams"]
        tf = _module("tensorflow")
        tf.contrib = types.SimpleNamespace(training=types.SimpleNamespace(HParams=hp_mod.HParams))
        tf.logging = types.SimpleNamespace(info=lambda *a, **k: None)
        sys.modules["tensorflow"] = tf
    if "librosa" not in sys.modules:
        lib = _module("librosa")
        filt = _module("librosa.filters")
        filt.mel = slaney_mel_filterbank
        util = _module("librosa.util")
        util.pad_center = _pad_center
        util.tiny = lambda x: np.finfo(np.asarray(x).dtype if np.issubdtype(np.asarray(x).dtype, np.floating) else np.float32).tiny
        util.normalize = lambda S, norm=np.inf, **kw: S if norm is None else S / np.max(np.abs(S))
        lib.filters, lib.util = filt, util
        sys.modules.update({"librosa": lib, "librosa.filters": filt, "librosa.util": util})
    for name in ("jamo", "jamo.jamo", "unidecode", "inflect", "nltk"):
        if name not in sys.modules:
            sys.modules[name] = _module(name)
    j = sys.modules["jamo"]
    if not hasattr(j, "h2j"):
        def _h2j(s):
            out = []
            for ch in s:
                o = ord(ch)
                if 0xAC00 <= o <= 0xD7A3:
                    o -= 0xAC00
                    out.append(chr(0x1100 + o // 588))
                    out.append(chr(0x1161 + (o % 588) // 28))
                    if o % 28:
                        out.append(chr(0x11A7 + o % 28))
                else:
                    out.append(ch)
            return "".join(out)
        j.h2j = _h2j
        j.hangul_to_jamo = lambda s: iter(_h2j(s))
        j.j2h = lambda *a: "".join(a)
        j.hcj_to_jamo = lambda c, position="vowel": c
        j.is_hcj = lambda c: 0x3131 <= ord(c) <= 0x318E
        j.jamo = sys.modules["jamo.jamo"]
        sys.modules["jamo.jamo"]._jamo_char_to_hcj = lambda c: c
    sys.modules["unidecode"].unidecode = lambda s: s
    sys.modules["inflect"].engine = lambda: types.SimpleNamespace(number_to_words=lambda n, **k: str(n))
    sys.modules["nltk"].sent_tokenize = lambda s: [s]


def _bool_mask_from_lengths(lengths):
    max_len = int(torch.max(lengths).item())
    ids = torch.arange(0, max_len, device=lengths.device)
    return ids < lengths.unsqueeze(1)


_REF_MODULES = ("hparams", "layers", "stft", "audio_processing", "utils", "CoordConv", "modules",
                "model", "loss_function", "fp16_optimizer", "loss_scaler", "distributed")
_loaded = {}


def load_reference():
    """Import the reference's flat modules under the shims; returns a namespace of modules.
    The reference modules are removed from ``sys.modules``/``sys.path`` afterwards so that the
    product's same-named modules (tacotron2-vae_b200/model.py ...) can be imported next to them."""
    if _loaded:
        return types.SimpleNamespace(**_loaded)
    if not reference_available():
        raise RuntimeError("reference tree not found at %s" % REFERENCE_ROOT)
    _install_shims()
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k in _REF_MODULES}
    sys.path.insert(0, REFERENCE_ROOT)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        for name in ("hparams", "audio_processing", "stft", "layers", "utils", "CoordConv", "modules",
                     "fp16_optimizer", "model", "loss_function"):
            _loaded[name] = importlib.import_module(name)
    finally:
        sys.path.remove(REFERENCE_ROOT)
        for name in _REF_MODULES:
            sys.modules.pop(name, None)
        sys.modules.update(saved)
    _loaded["utils"].get_mask_from_lengths = _bool_mask_from_lengths
    _loaded["model"].get_mask_from_lengths = _bool_mask_from_lengths
    return types.SimpleNamespace(**_loaded)


class InjectedRandomness(object):
    """Context manager that makes the reference deterministic: every ``F.dropout`` call pops the
    next keep-mask from ``masks`` (float 0/1 tensors, reference call order) and every
    ``torch.randn_like`` returns ``eps``.  With ``masks=None`` dropout becomes identity.
    Records the shapes/p of the calls in ``self.calls``."""

    def __init__(self, masks=None, eps=None, record=False):
        self.masks = list(masks) if masks is not None else None
        self.eps = eps
        self.record = record
        self.calls = []

    def __enter__(self):
        import torch.nn.functional as F
        self._F, self._drop, self._randn = F, F.dropout, torch.randn_like

        def dropout(x, p=0.5, training=True, inplace=False):
            self.calls.append((tuple(x.shape), float(p), bool(training)))
            if not training or p == 0.0:
                return x
            if self.masks is None:
                return x
            m = self.masks.pop(0)
            assert tuple(m.shape) == tuple(x.shape), (m.shape, x.shape)
            return x * m.to(x.dtype) / (1.0 - p)

        def randn_like(x, **kw):
            if self.eps is None:
                return torch.zeros_like(x)
            return self.eps.to(x.dtype).reshape(x.shape)

        F.dropout = dropout
        torch.randn_like = randn_like
        return self

    def __exit__(self, *exc):
        self._F.dropout = self._drop
        torch.randn_like = self._randn
        return False
