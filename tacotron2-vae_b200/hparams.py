"""Hyper-parameter surface of the hot path.

Drop-in for the reference's ``hparams.create_hparams`` (reference hparams.py:3-126), which
builds a ``tf.contrib.training.HParams``.  TensorFlow-contrib no longer exists, so this
module ships a small attribute bag with the three things the callers use: attribute
access/mutation, ``.parse("a=1,b=[x,y]")`` with coercion to the default's type, and
``.values()``.  The defaults table below is the contract (reference hparams.py:6-117).
"""
import ast

# (name, default) in the reference's own order -- hparams.py:9-116
_DEFAULTS = (
    # experiment
    ("epochs", 300), ("iters_per_checkpoint", 500), ("seed", 1234),
    ("dynamic_loss_scaling", True), ("fp16_run", False), ("distributed_run", False),
    ("dist_backend", "nccl"), ("dist_url", "tcp://localhost:54321"),
    ("cudnn_enabled", True), ("cudnn_benchmark", True),
    # data
    ("load_mel_from_disk", False),
    ("training_files", "filelists/ms_kor_train.txt"),
    ("validation_files", "filelists/ms_kor_val.txt"),
    ("text_cleaners", ["korean_cleaners"]), ("sort_by_length", False),
    # audio
    ("max_wav_value", 32768.0), ("sampling_rate", 16000), ("filter_length", 1024),
    ("hop_length", 256), ("win_length", 1024), ("n_mel_channels", 80),
    ("mel_fmin", 0.0), ("mel_fmax", 8000.0),
    # model
    ("n_symbols", 80), ("symbols_embedding_dim", 512),
    ("encoder_kernel_size", 5), ("encoder_n_convolutions", 3), ("encoder_embedding_dim", 512),
    ("n_speakers", 1), ("speaker_embedding_dim", 16),
    ("n_emotions", 4), ("emotion_embedding_dim", 16),
    ("E", 512), ("ref_enc_filters", [32, 32, 64, 64, 128, 128]), ("ref_enc_size", [3, 3]),
    ("ref_enc_strides", [2, 2]), ("ref_enc_pad", [1, 1]), ("ref_enc_gru_size", 512 // 2),
    ("z_latent_dim", 32), ("anneal_function", "logistic"), ("anneal_k", 0.0025),
    ("anneal_x0", 10000), ("anneal_upper", 0.2), ("anneal_lag", 50000),
    ("prosody_n_convolutions", 6), ("prosody_conv_dim_in", [1, 32, 32, 64, 64, 128]),
    ("prosody_conv_dim_out", [32, 32, 64, 64, 128, 128]), ("prosody_conv_kernel", 3),
    ("prosody_conv_stride", 2), ("prosody_embedding_dim", 128),
    ("n_frames_per_step", 1), ("decoder_rnn_dim", 1024), ("prenet_dim", 256),
    ("max_decoder_steps", 1000), ("gate_threshold", 0.5),
    ("p_attention_dropout", 0.1), ("p_decoder_dropout", 0.1),
    ("attention_rnn_dim", 1024), ("attention_dim", 128),
    ("attention_location_n_filters", 32), ("attention_location_kernel_size", 31),
    ("postnet_embedding_dim", 512), ("postnet_kernel_size", 5), ("postnet_n_convolutions", 5),
    # optimisation
    ("use_saved_learning_rate", False), ("learning_rate", 1e-3), ("weight_decay", 1e-6),
    ("grad_clip_thresh", 1.0), ("batch_size", 64), ("mask_padding", True),
)


def _split_top_level(s):
    """Split on commas that are not inside brackets/quotes."""
    out, depth, cur, quote = [], 0, [], None
    for ch in s:
        if quote:
            cur.append(ch)
            if ch == quote:
                quote = None
            continue
        if ch in "\"'":
            quote = ch
            cur.append(ch)
        elif ch in "[(":
            depth += 1
            cur.append(ch)
        elif ch in "])":
            depth -= 1
            cur.append(ch)
        elif ch == "," and depth == 0:
            out.append("".join(cur))
            cur = []
        else:
            cur.append(ch)
    if cur:
        out.append("".join(cur))
    return [p.strip() for p in out if p.strip()]


def _coerce(raw, like):
    raw = raw.strip()
    if isinstance(like, bool):
        return raw.lower() in ("true", "1", "t", "yes")
    if isinstance(like, int):
        return int(float(raw)) if raw.lower() not in ("true", "false") else int(raw.lower() == "true")
    if isinstance(like, float):
        return float(raw)
    if isinstance(like, list):
        try:
            v = ast.literal_eval(raw)
        except (ValueError, SyntaxError):
            v = [x.strip().strip("'\"") for x in raw.strip("[]").split(",") if x.strip()]
        v = list(v) if isinstance(v, (list, tuple)) else [v]
        if like and v:
            v = [_coerce(str(x), like[0]) if not isinstance(x, type(like[0])) else x for x in v]
        return v
    return raw.strip("'\"")


class HParams(object):
    """Attribute bag with ``parse``/``values`` (stand-in for tf.contrib.training.HParams)."""

    def __init__(self, **kw):
        object.__setattr__(self, "_names", [])
        for k, v in kw.items():
            self.add_hparam(k, v)

    def add_hparam(self, name, value):
        if name in self._names:
            raise ValueError("hyperparameter %r already exists" % name)
        self._names.append(name)
        object.__setattr__(self, name, value)

    def set_hparam(self, name, value):
        if name not in self._names:
            raise KeyError(name)
        object.__setattr__(self, name, value)

    def parse(self, spec):
        for item in _split_top_level(spec or ""):
            if "=" not in item:
                raise ValueError("could not parse hparam %r (expected name=value)" % item)
            name, raw = item.split("=", 1)
            name = name.strip()
            if name not in self._names:
                raise ValueError("unknown hyperparameter %r" % name)
            object.__setattr__(self, name, _coerce(raw, getattr(self, name)))
        return self

    def values(self):
        return {k: getattr(self, k) for k in self._names}

    def __contains__(self, name):
        return name in self._names

    def __repr__(self):
        return "HParams(%s)" % ", ".join("%s=%r" % kv for kv in self.values().items())


def create_hparams(hparams_string=None, verbose=False):
    """Same call surface as reference hparams.py:3 (callers: train.py:273, synthesizer.py:49)."""
    hp = HParams(**{k: (list(v) if isinstance(v, list) else v) for k, v in _DEFAULTS})
    if hparams_string:
        hp.parse(hparams_string)
    if verbose:
        print("Final parsed hparams: %s" % hp.values())
    return hp
