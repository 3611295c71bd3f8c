"""TEST INFRASTRUCTURE ONLY -- CPU restatement ("port") of the reference hot path.

This is the oracle the CUDA path is checked against on the GPU box, where /root/reference
does not exist.  It is a plain-PyTorch-fp32, single-file, functional restatement of the
algorithm in the reference's model.py / modules.py / CoordConv.py / loss_function.py /
stft.py / layers.py, written against a flat ``state_dict`` (reference key names).  It is
pinned two ways (tests/test_oracle_cpu.py):
  * against the real reference imported in the build container (oracle/ref_shims.py), and
  * against the committed fixtures in tests/golden/ generated from the real reference by
    oracle/gen_golden.py.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
import this module; the product (tacotron2-vae_b200/) never does.

Randomness is explicit: dropout keep-masks and the VAE eps are inputs (``Rand``).  Layouts
follow the reference ([B, C, T] activations) so fixtures are directly comparable.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


class Rand(object):
    """Explicit randomness of one training forward (all float 0/1 keep masks, reference layout).

    enc[i]   [B,512,Ti]      p=.5  (model.py:177)       i = 0..2
    prenet[i][To+1,B,256]    p=.5  (model.py:101)       i = 0..1   (always on, even in eval)
    dec      [To,4,B,1024]   p=.1  (model.py:361-381)   order per step: att_h, att_c, dec_h, dec_c
    post[i]  [B,512|80,To]   p=.5  (model.py:145-146)   i = 0..4
    eps      [B,32]                (modules.py:19)
    """

    def __init__(self, enc=None, prenet=None, dec=None, post=None, eps=None):
        self.enc, self.prenet, self.dec, self.post, self.eps = enc, prenet, dec, post, eps

    @staticmethod
    def draw(B, Ti, To, seed=0, hp=None):
        g = torch.Generator().manual_seed(seed)

        def keep(shape, p):
            return (torch.rand(shape, generator=g) >= p).float()
        return Rand(enc=[keep((B, 512, Ti), 0.5) for _ in range(3)],
                    prenet=[keep((To + 1, B, 256), 0.5) for _ in range(2)],
                    dec=keep((To, 4, B, 1024), 0.1),
                    post=[keep((B, 512, To), 0.5) for _ in range(4)] + [keep((B, 80, To), 0.5)],
                    eps=torch.randn((B, 32), generator=g))

    def as_reference_call_list(self):
        """Masks in the order the reference calls F.dropout during Tacotron2.forward."""
        out = list(self.enc) + list(self.prenet)
        for t in range(self.dec.shape[0]):
            out += [self.dec[t, 0], self.dec[t, 1], self.dec[t, 2], self.dec[t, 3]]
        return out + list(self.post)


def _drop(x, mask, p):
    return x if mask is None else x * mask * (1.0 / (1.0 - p))


def _bn(x, P, pre, training, stats_out):
    """BatchNorm over all dims but channel 1 (nn.BatchNorm1d/2d semantics, eps 1e-5, momentum .1).
    Batch statistics include padded positions (reference quirk Q4)."""
    w, b = P[pre + ".weight"], P[pre + ".bias"]
    rm, rv = P[pre + ".running_mean"], P[pre + ".running_var"]
    red = [d for d in range(x.dim()) if d != 1]
    shp = [1, -1] + [1] * (x.dim() - 2)
    if training:
        mean = x.mean(red)
        var = x.var(red, unbiased=False)
        n = x.numel() // x.shape[1]
        if stats_out is not None:
            stats_out[pre + ".running_mean"] = 0.9 * rm + 0.1 * mean.detach()
            stats_out[pre + ".running_var"] = 0.9 * rv + 0.1 * var.detach() * (n / max(n - 1, 1))
            stats_out[pre + ".num_batches_tracked"] = P[pre + ".num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    return (x - mean.view(shp)) * torch.rsqrt(var.view(shp) + 1e-5) * w.view(shp) + b.view(shp)


def lstm_cell(x, h, c, w_ih, w_hh, b_ih, b_hh):
    """nn.LSTMCell arithmetic; gate order i, f, g, o."""
    g = x @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
    H = h.shape[1]
    i, f, gg, o = torch.sigmoid(g[:, :H]), torch.sigmoid(g[:, H:2 * H]), torch.tanh(g[:, 2 * H:3 * H]), torch.sigmoid(g[:, 3 * H:])
    c2 = f * c + i * gg
    return o * torch.tanh(c2), c2


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    """nn.GRU arithmetic; gate order r, z, n."""
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    H = h.shape[1]
    r = torch.sigmoid(gi[:, :H] + gh[:, :H])
    z = torch.sigmoid(gi[:, H:2 * H] + gh[:, H:2 * H])
    n = torch.tanh(gi[:, 2 * H:] + r * gh[:, 2 * H:])
    return (1 - z) * n + z * h


# ------------------------------------------------------------------------------- encoder
def encoder(P, x, in_len, training, masks=None, stats_out=None):
    """model.py:151-203.  x [B,512,Ti] -> [B,Ti,512].  ``in_len`` None => Encoder.inference
    (no packing).  Packed semantics: each row runs over its own length, padded outputs are 0."""
    for i in range(3):
        pre = "encoder.convolutions.%d" % i
        x = F.conv1d(x, P[pre + ".0.conv.weight"], P[pre + ".0.conv.bias"], padding=2)
        x = torch.relu(_bn(x, P, pre + ".1", training, stats_out))
        if training:
            x = _drop(x, None if masks is None else masks[i], 0.5)
    x = x.transpose(1, 2)                                    # [B,Ti,512]
    B, Ti, _ = x.shape
    H = 256
    lens = [Ti] * B if in_len is None else [int(v) for v in in_len]
    out = x.new_zeros(B, Ti, 2 * H)
    outs = [[None] * Ti for _ in range(2)]
    for d, sfx in enumerate(("", "_reverse")):
        w_ih, w_hh = P["encoder.lstm.weight_ih_l0" + sfx], P["encoder.lstm.weight_hh_l0" + sfx]
        b_ih, b_hh = P["encoder.lstm.bias_ih_l0" + sfx], P["encoder.lstm.bias_hh_l0" + sfx]
        h = x.new_zeros(B, H)
        c = x.new_zeros(B, H)
        act = torch.tensor(lens)
        order = range(Ti) if d == 0 else range(Ti - 1, -1, -1)
        for t in order:
            live = (act > t).float().unsqueeze(1)           # rows whose sequence covers t
            h2, c2 = lstm_cell(x[:, t], h, c, w_ih, w_hh, b_ih, b_hh)
            h = live * h2 + (1 - live) * h                  # dead rows: state untouched (stays 0 in reverse)
            c = live * c2 + (1 - live) * c
            outs[d][t] = live * h2
    out = torch.cat([torch.stack(outs[0], 1), torch.stack(outs[1], 1)], dim=2)
    return out


# ---------------------------------------------------------------------- reference encoder
def add_coords(x):
    """CoordConv.py:37-74 (rank 2, with_r): x [N,1,H,W] -> [N,4,H,W]."""
    N, _, Hh, Ww = x.shape
    xx = (torch.arange(Hh, dtype=torch.int32).float() / (Hh - 1)) * 2 - 1      # varies along H (time)
    yy = (torch.arange(Ww, dtype=torch.int32).float() / (Ww - 1)) * 2 - 1      # varies along W (mel)
    xx = xx.view(1, 1, Hh, 1).expand(N, 1, Hh, Ww)
    yy = yy.view(1, 1, 1, Ww).expand(N, 1, Hh, Ww)
    rr = torch.sqrt((xx - 0.5) ** 2 + (yy - 0.5) ** 2)
    return torch.cat([x, xx, yy, rr], dim=1)


def ref_encoder(P, mel, training, stats_out=None):
    """modules.py:65-80.  mel [N,80,T] is *reinterpreted* (view, not transpose -- quirk Q1) as
    [N,1,T,80]; 6 x (conv3x3 s2 p1 -> BN2d -> ReLU); GRU over T' frames; last hidden [N,256]."""
    N = mel.shape[0]
    x = mel.contiguous().view(N, 1, -1, 80)
    pre = "vae_gst.ref_encoder."
    for i in range(6):
        if i == 0:
            x = add_coords(x)
            w, b = P[pre + "convs.0.conv.weight"], P[pre + "convs.0.conv.bias"]
        else:
            w, b = P[pre + "convs.%d.weight" % i], P[pre + "convs.%d.bias" % i]
        x = F.conv2d(x, w, b, stride=2, padding=1)
        x = torch.relu(_bn(x, P, pre + "bns.%d" % i, training, stats_out))
    x = x.transpose(1, 2).contiguous()                       # [N,T',128,W']
    Tp = x.shape[1]
    x = x.view(N, Tp, -1)
    h = x.new_zeros(N, 256)
    for t in range(Tp):
        h = gru_cell(x[:, t], h, P[pre + "gru.weight_ih_l0"], P[pre + "gru.weight_hh_l0"],
                     P[pre + "gru.bias_ih_l0"], P[pre + "gru.bias_hh_l0"])
    return h


def vae_gst(P, mel, training, eps=None, stats_out=None):
    """modules.py:16-31 -> (style [N,512], mu, logvar, z)."""
    h = ref_encoder(P, mel, training, stats_out)
    mu = h @ P["vae_gst.fc1.weight"].t() + P["vae_gst.fc1.bias"]
    logvar = h @ P["vae_gst.fc2.weight"].t() + P["vae_gst.fc2.bias"]
    if training:
        e = torch.zeros_like(mu) if eps is None else eps
        z = e * torch.exp(0.5 * logvar) + mu
    else:
        z = mu
    style = z @ P["vae_gst.fc3.weight"].t() + P["vae_gst.fc3.bias"]
    return style, mu, logvar, z


# ------------------------------------------------------------------------------- decoder
def prenet(P, x, masks=None):
    """model.py:91-102; dropout is always on (quirk Q2); masks = [m0, m1] or None (identity)."""
    for i in range(2):
        x = torch.relu(x @ P["decoder.prenet.layers.%d.linear_layer.weight" % i].t())
        x = _drop(x, None if masks is None else masks[i], 0.5)
    return x


class DecoderState(object):
    """The per-utterance state the reference keeps on the Decoder object (model.py:260-291)."""

    def __init__(self, P, memory, mask):
        B, Ti, _ = memory.shape
        z = memory.new_zeros
        self.ah, self.ac, self.dh, self.dc = z(B, 1024), z(B, 1024), z(B, 1024), z(B, 1024)
        self.w, self.wcum, self.ctx = z(B, Ti), z(B, Ti), z(B, 512)
        self.memory = memory
        self.pmem = memory @ P["decoder.attention_layer.memory_layer.linear_layer.weight"].t()
        self.mask = mask                                      # True at padded text positions


def decode_step(P, S, x, training, m4=None, mask_value=-float("inf")):
    """Decoder.decode (model.py:346-389) on explicit state ``S``; x [B,256] (prenet output).
    m4 = (att_h, att_c, dec_h, dec_c) keep masks or None.  Returns mel [B,80], gate [B,1], w [B,Ti]."""
    a = "decoder.attention_rnn."
    h, c = lstm_cell(torch.cat((x, S.ctx), -1), S.ah, S.ac, P[a + "weight_ih"], P[a + "weight_hh"],
                     P[a + "bias_ih"], P[a + "bias_hh"])
    if training:
        h = _drop(h, None if m4 is None else m4[0], 0.1)
        c = _drop(c, None if m4 is None else m4[1], 0.1)
    S.ah, S.ac = h, c
    L = "decoder.attention_layer."
    q = h @ P[L + "query_layer.linear_layer.weight"].t()                                   # [B,128]
    wcat = torch.stack((S.w, S.wcum), dim=1)                                              # [B,2,Ti]
    loc = F.conv1d(wcat, P[L + "location_layer.location_conv.conv.weight"], padding=15)   # [B,32,Ti]
    loc = loc.transpose(1, 2) @ P[L + "location_layer.location_dense.linear_layer.weight"].t()
    e = (torch.tanh(q.unsqueeze(1) + loc + S.pmem) @ P[L + "v.linear_layer.weight"].t()).squeeze(-1)
    if S.mask is not None:
        e = e.masked_fill(S.mask, mask_value)
    w = torch.softmax(e, dim=1)
    S.ctx = torch.bmm(w.unsqueeze(1), S.memory).squeeze(1)
    S.w = w
    S.wcum = S.wcum + w
    d = "decoder.decoder_rnn."
    h2, c2 = lstm_cell(torch.cat((S.ah, S.ctx), -1), S.dh, S.dc, P[d + "weight_ih"], P[d + "weight_hh"],
                       P[d + "bias_ih"], P[d + "bias_hh"])
    if training:
        h2 = _drop(h2, None if m4 is None else m4[2], 0.1)
        c2 = _drop(c2, None if m4 is None else m4[3], 0.1)
    S.dh, S.dc = h2, c2
    hc = torch.cat((h2, S.ctx), dim=1)
    mel = hc @ P["decoder.linear_projection.linear_layer.weight"].t() + P["decoder.linear_projection.linear_layer.bias"]
    gate = hc @ P["decoder.gate_layer.linear_layer.weight"].t() + P["decoder.gate_layer.linear_layer.bias"]
    return mel, gate, w


def decoder_teacher_forced(P, memory, targets, in_len, training, rand=None):
    """Decoder.forward (model.py:391-426): mel [B,80,To], gate [B,To], align [B,To,Ti]."""
    B, _, To = targets.shape
    frames = torch.cat((targets.new_zeros(1, B, 80), targets.permute(2, 0, 1)), dim=0)    # go frame + To frames
    pre = prenet(P, frames, None if rand is None else rand.prenet)
    Ti = memory.shape[1]
    mask = ~(torch.arange(Ti).unsqueeze(0) < torch.as_tensor(in_len).unsqueeze(1))
    S = DecoderState(P, memory, mask)
    mels, gates, aligns = [], [], []
    for t in range(To):
        m4 = None if rand is None or rand.dec is None else rand.dec[t]
        mel, gate, w = decode_step(P, S, pre[t], training, m4)
        mels.append(mel)
        gates.append(gate.squeeze(1))
        aligns.append(w)
    return torch.stack(mels, 2), torch.stack(gates, 1), torch.stack(aligns, 1)


def decoder_free_running(P, memory, n_steps, prenet_masks=None, training=False):
    """Decoder.inference (model.py:428-464) for a fixed number of steps (config 5 semantics: the
    stop test is evaluated by the caller).  prenet_masks [n_steps,2,B,256] or None."""
    B = memory.shape[0]
    S = DecoderState(P, memory, None)
    x = memory.new_zeros(B, 80)
    mels, gates, aligns = [], [], []
    for t in range(n_steps):
        p = prenet(P, x, None if prenet_masks is None else prenet_masks[t])
        mel, gate, w = decode_step(P, S, p, training)
        mels.append(mel); gates.append(gate); aligns.append(w)
        x = mel
    return torch.stack(mels, 2), torch.stack(gates, 1), torch.stack(aligns, 1)


# ------------------------------------------------------------------------------- postnet
def postnet(P, x, training, masks=None, stats_out=None):
    """model.py:105-148 (residual added by the caller)."""
    for i in range(5):
        pre = "postnet.convolutions.%d" % i
        x = F.conv1d(x, P[pre + ".0.conv.weight"], P[pre + ".0.conv.bias"], padding=2)
        x = _bn(x, P, pre + ".1", training, stats_out)
        if i < 4:
            x = torch.tanh(x)
        if training:
            x = _drop(x, None if masks is None else masks[i], 0.5)
    return x


# ------------------------------------------------------------------------------- model
def tacotron2_forward(P, text, in_len, mel_tgt, out_len, training=True, rand=None, stats_out=None,
                      mask_padding=True):
    """Tacotron2.forward (model.py:522-547) + parse_output (509-520).
    Returns [mel, mel_post, gate, align, mu, logvar, z] (the 8th element, emotions, is a pass-through)."""
    emb = F.embedding(text, P["transcript_embedding.weight"]).transpose(1, 2)              # [B,512,Ti]
    enc = encoder(P, emb, in_len, training, None if rand is None else rand.enc, stats_out)
    style, mu, logvar, z = vae_gst(P, mel_tgt, training, None if rand is None else rand.eps, stats_out)
    memory = enc + style.unsqueeze(1)                                                      # incl. padded rows
    mel, gate, align = decoder_teacher_forced(P, memory, mel_tgt, in_len, training, rand)
    post = postnet(P, mel, training, None if rand is None else rand.post, stats_out)
    mel_post = mel + post
    if mask_padding:
        To = mel.shape[2]
        pad = ~(torch.arange(To).unsqueeze(0) < torch.as_tensor(out_len).unsqueeze(1))     # [B,To]
        # in-place on .data exactly like the reference: autograd does not see it, and the tensor
        # saved as Postnet conv-0's input is the *masked* one (quirk Q10)
        mel.data.masked_fill_(pad.unsqueeze(1), 0.0)
        mel_post.data.masked_fill_(pad.unsqueeze(1), 0.0)
        gate.data.masked_fill_(pad, 1e3)
    return [mel, mel_post, gate, align, mu, logvar, z]


def kl_weight(anneal_function, step, lag=50000, k=0.0025, x0=10000, upper=0.2):
    """loss_function.py:15-24."""
    if anneal_function == "logistic":
        return float(upper / (upper + np.exp(-k * (step - x0))))
    if anneal_function == "linear":
        return min(upper, step / x0) if step > lag else 0
    if anneal_function == "constant":
        return 0.001
    return None


def vae_loss(outputs, mel_tgt, gate_tgt, klw):
    """loss_function.py:27-45 -> (total, recon, kl)."""
    mel, mel_post, gate, _, mu, logvar = outputs[:6]
    mel_loss = ((mel - mel_tgt) ** 2).mean() + ((mel_post - mel_tgt) ** 2).mean()
    gate_loss = F.binary_cross_entropy_with_logits(gate.reshape(-1, 1), gate_tgt.reshape(-1, 1))
    kl = -0.5 * torch.sum(1 + logvar - mu.pow(2) - logvar.exp())
    recon = mel_loss + gate_loss
    return recon + klw * kl, recon, kl


# ------------------------------------------------------------------------------- STFT / mel
def hann_periodic(n):
    return 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)


def mel_filterbank(sr=16000, n_fft=1024, n_mels=80, fmin=0.0, fmax=8000.0):
    """librosa 0.6.0 filters.mel defaults (Slaney scale, area normalised) as float32 [80,513]."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def h2m(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def m2h(m):
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)
    pts = m2h(np.linspace(h2m(fmin), h2m(fmax), n_mels + 2))
    bins = np.linspace(0, sr / 2.0, n_fft // 2 + 1)
    W = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lo = (bins - pts[i]) / (pts[i + 1] - pts[i])
        hi = (pts[i + 2] - bins) / (pts[i + 2] - pts[i + 1])
        W[i] = np.maximum(0, np.minimum(lo, hi)) * (2.0 / (pts[i + 2] - pts[i]))
    return W.astype(np.float32)


def mel_spectrogram(wav, n_fft=1024, hop=256, sr=16000, n_mels=80, fmin=0.0, fmax=8000.0):
    """TacotronSTFT.mel_spectrogram (layers.py:75-92, stft.py:77-105, audio_processing.py:77-83):
    reflect-pad n_fft/2, hann-windowed DFT every ``hop`` samples, magnitude, mel, log(clamp 1e-5).
    wav [B,S] in [-1,1] -> [B,80,S//hop+1]."""
    k = np.arange(n_fft // 2 + 1)[:, None] * np.arange(n_fft)[None, :] * (2.0 * np.pi / n_fft)
    win = hann_periodic(n_fft)
    basis = np.concatenate([np.cos(k) * win, -np.sin(k) * win], 0).astype(np.float32)    # [1026,1024]
    x = F.pad(wav.unsqueeze(1), (n_fft // 2, n_fft // 2), mode="reflect")
    ft = F.conv1d(x, torch.from_numpy(basis).unsqueeze(1), stride=hop)
    c = n_fft // 2 + 1
    mag = torch.sqrt(ft[:, :c] ** 2 + ft[:, c:] ** 2)
    mel = torch.matmul(torch.from_numpy(mel_filterbank(sr, n_fft, n_mels, fmin, fmax)), mag)
    return torch.log(torch.clamp(mel, min=1e-5))


# ------------------------------------------------------------------------------- helpers
def synthetic_batch(B, Ti, To, seed=0):
    """SURVEY.md section 8(d) synthetic inputs (collate layout, data_utils.py:82-137)."""
    g = torch.Generator().manual_seed(seed)
    in_len = torch.randint(max(Ti // 2, 1), Ti + 1, (B,), generator=g)
    in_len[0] = Ti
    in_len, _ = torch.sort(in_len, descending=True)
    out_len = torch.randint(max(To // 2, 1), To + 1, (B,), generator=g)
    out_len[int(torch.randint(0, B, (1,), generator=g))] = To
    text = torch.randint(1, 80, (B, Ti), generator=g)
    mel = torch.randn(B, 80, To, generator=g) * 2.0 - 5.0
    gate = torch.zeros(B, To)
    for b in range(B):
        text[b, int(in_len[b]):] = 0
        mel[b, :, int(out_len[b]):] = 0
        gate[b, int(out_len[b]) - 1:] = 1
    speakers = torch.zeros(B, 1, dtype=torch.long)
    emotions = torch.zeros(B, 4, dtype=torch.long)
    emotions[torch.arange(B), torch.randint(0, 4, (B,), generator=g)] = 1
    return text, in_len, mel, gate, out_len, speakers, emotions


# ------------------------------------------------------------------------------- parameters
def param_shapes():
    """state_dict keys/shapes of reference model.Tacotron2 at hparams defaults."""
    S = {"transcript_embedding.weight": (80, 512)}
    for n, (o, i) in (("speaker_embedding", (16, 1)), ("emotion_embedding", (16, 4))):
        S[n + ".linear_layer.weight"] = (o, i)
        S[n + ".linear_layer.bias"] = (o,)

    def bn(pre, c):
        S[pre + ".weight"] = (c,); S[pre + ".bias"] = (c,)
        S[pre + ".running_mean"] = (c,); S[pre + ".running_var"] = (c,)
        S[pre + ".num_batches_tracked"] = ()
    for i in range(3):
        S["encoder.convolutions.%d.0.conv.weight" % i] = (512, 512, 5)
        S["encoder.convolutions.%d.0.conv.bias" % i] = (512,)
        bn("encoder.convolutions.%d.1" % i, 512)
    for sfx in ("", "_reverse"):
        S["encoder.lstm.weight_ih_l0" + sfx] = (1024, 512)
        S["encoder.lstm.weight_hh_l0" + sfx] = (1024, 256)
        S["encoder.lstm.bias_ih_l0" + sfx] = (1024,)
        S["encoder.lstm.bias_hh_l0" + sfx] = (1024,)
    S["decoder.prenet.layers.0.linear_layer.weight"] = (256, 80)
    S["decoder.prenet.layers.1.linear_layer.weight"] = (256, 256)
    for n, k in (("attention_rnn", 768), ("decoder_rnn", 1536)):
        S["decoder.%s.weight_ih" % n] = (4096, k)
        S["decoder.%s.weight_hh" % n] = (4096, 1024)
        S["decoder.%s.bias_ih" % n] = (4096,)
        S["decoder.%s.bias_hh" % n] = (4096,)
    L = "decoder.attention_layer."
    S[L + "query_layer.linear_layer.weight"] = (128, 1024)
    S[L + "memory_layer.linear_layer.weight"] = (128, 512)
    S[L + "v.linear_layer.weight"] = (1, 128)
    S[L + "location_layer.location_conv.conv.weight"] = (32, 2, 31)
    S[L + "location_layer.location_dense.linear_layer.weight"] = (128, 32)
    S["decoder.linear_projection.linear_layer.weight"] = (80, 1536)
    S["decoder.linear_projection.linear_layer.bias"] = (80,)
    S["decoder.gate_layer.linear_layer.weight"] = (1, 1536)
    S["decoder.gate_layer.linear_layer.bias"] = (1,)
    chans = [80, 512, 512, 512, 512, 80]
    for i in range(5):
        S["postnet.convolutions.%d.0.conv.weight" % i] = (chans[i + 1], chans[i], 5)
        S["postnet.convolutions.%d.0.conv.bias" % i] = (chans[i + 1],)
        bn("postnet.convolutions.%d.1" % i, chans[i + 1])
    R = "vae_gst.ref_encoder."
    S[R + "convs.0.weight"] = (32, 1, 3, 3); S[R + "convs.0.bias"] = (32,)          # dead (quirk Q6)
    S[R + "convs.0.conv.weight"] = (32, 4, 3, 3); S[R + "convs.0.conv.bias"] = (32,)
    f = [32, 32, 64, 64, 128, 128]
    for i in range(1, 6):
        S[R + "convs.%d.weight" % i] = (f[i], f[i - 1], 3, 3); S[R + "convs.%d.bias" % i] = (f[i],)
    for i in range(6):
        bn(R + "bns.%d" % i, f[i])
    S[R + "gru.weight_ih_l0"] = (768, 256); S[R + "gru.weight_hh_l0"] = (768, 256)
    S[R + "gru.bias_ih_l0"] = (768,); S[R + "gru.bias_hh_l0"] = (768,)
    S["vae_gst.fc1.weight"] = (32, 256); S["vae_gst.fc1.bias"] = (32,)
    S["vae_gst.fc2.weight"] = (32, 256); S["vae_gst.fc2.bias"] = (32,)
    S["vae_gst.fc3.weight"] = (512, 32); S["vae_gst.fc3.bias"] = (512,)
    return S


def init_params(seed=1234):
    """Deterministic synthetic weights in the reference state_dict layout (CPU generator, so the
    same tensors are produced here and on the GPU box).  Not the reference's init stream -- the
    fixtures are made by loading THESE tensors into the real reference (oracle/gen_golden.py)."""
    g = torch.Generator().manual_seed(seed)
    P = {}
    for k, shp in param_shapes().items():
        if k.endswith("num_batches_tracked"):
            P[k] = torch.zeros((), dtype=torch.long)
        elif k.endswith("running_var"):
            P[k] = torch.rand(shp, generator=g) + 0.5
        elif k.endswith("running_mean"):
            P[k] = (torch.rand(shp, generator=g) - 0.5) * 0.2
        elif len(shp) == 1:
            is_bn_w = k.endswith(".weight")
            P[k] = torch.rand(shp, generator=g) + 0.5 if is_bn_w else (torch.rand(shp, generator=g) - 0.5) * 0.2
        else:
            fan_out = shp[0] * int(np.prod(shp[2:])) if len(shp) > 2 else shp[0]
            fan_in = int(np.prod(shp[1:]))
            if "lstm" in k or "_rnn" in k or "gru" in k:
                a = 1.0 / math.sqrt(shp[0] // (3 if "gru" in k else 4))
            else:
                a = math.sqrt(6.0 / (fan_in + fan_out))
            P[k] = (torch.rand(shp, generator=g) * 2 - 1) * a
    return P


def grad_probe(name, shape):
    """Deterministic +-1 probe vector used to fingerprint a gradient tensor in the fixtures."""
    n = int(np.prod(shape)) if len(shape) else 1
    h = (np.arange(n, dtype=np.uint64) * np.uint64(2654435761) + np.uint64(sum(map(ord, name)))) >> np.uint64(7)
    return torch.from_numpy(((h & np.uint64(1)).astype(np.float32) * 2 - 1)).reshape(shape)
