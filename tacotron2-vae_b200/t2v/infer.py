"""Inference-surface helpers: the sub-module calls inference.ipynb / synthesizer.py make on the model
(model.transcript_embedding(ids), model.encoder.inference(x), model.vae_gst(mel), model.vae_gst.fc3(z),
decoder.prenet / decode, model.postnet(mel)) routed to the t2v kernels.  No autograd here."""
import os

import torch

from . import _lib, engine
from ._lib import call as L


def default_precision():
    return os.environ.get("T2V_PRECISION", "fp16")


def _need_cuda(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise RuntimeError("%s: the t2v engine runs on CUDA tensors only (libt2v_b200.so); move the model/input to a "
                           "B200 -- there is no CPU fallback" % what)


def linear(x, weight, bias=None):
    _need_cuda(x, "linear")
    K = x.shape[-1]
    N = weight.shape[0]
    x2 = x.reshape(-1, K).contiguous().float()
    out = torch.empty(x2.shape[0], N, device=x.device)
    engine.Ops.gemm(x2, K, 1, weight.detach().contiguous(), K, 1, out, N, x2.shape[0], N, K, 1.0, 0.0,
                    None if bias is None else bias.detach())
    return out.view(*x.shape[:-1], N)


def state_tensors(module):
    """name -> fp32 CUDA tensor for every parameter and buffer of the model (no copies)."""
    P = {}
    for k, v in module.named_parameters():
        P[k] = v.detach()
    for k, v in module.named_buffers():
        P[k] = v
    for k, v in P.items():
        _need_cuda(v, "parameter %s" % k)
        if v.dtype.is_floating_point and (v.dtype != torch.float32 or not v.is_contiguous()):
            raise RuntimeError("parameter %s must be contiguous fp32 (bf16/tf32 operands are derived inside the kernels)" % k)
    return P


def embedding(P, ids):
    _need_cuda(ids, "transcript_embedding")
    ids = ids.long().contiguous()
    B, Ti = ids.shape
    X0 = engine.embedding_forward(P, ids, ids.device)
    out = torch.empty(B, Ti, 512, device=ids.device)
    L("t2v_unpad_add", X0, None, out, B, Ti, 512, 0)
    return out


def encoder_inference(ops, P, x_bct, training):
    """Encoder.inference (model.py:194-203): x [B,512,Ti] -> [B,Ti,512], no packing."""
    _need_cuda(x_bct, "encoder.inference")
    x = x_bct.contiguous().float()
    B, C, Ti = x.shape
    dev = x.device
    X0 = torch.zeros(B * (Ti + 4), C, device=dev)
    L("t2v_bct_to_padded", x, X0, B, C, Ti, 0.0)
    HoutP, _ = engine.encoder_forward(ops, P, None, None, training, None, 0, dev, packed=False, X0=X0, shape=(B, Ti))
    out = torch.empty(B, Ti, 512, device=dev)
    L("t2v_unpad_add", HoutP, None, out, B, Ti, 512, 0)
    return out


def vae_gst(ops, P, mel, training, eps=None):
    _need_cuda(mel, "vae_gst")
    mel = mel.contiguous().float()
    dev = mel.device
    if training and eps is None:
        eps = torch.randn(mel.shape[0], P["vae_gst.fc1.weight"].shape[0], device=dev)
    style, mulv, z, _ = engine.vae_forward(ops, P, mel, training, eps, dev)
    Z = z.shape[1]
    return style, mulv[:, :Z].contiguous(), mulv[:, Z:].contiguous(), z


def postnet(ops, P, mel_bct, training, seed=0):
    _need_cuda(mel_bct, "postnet")
    x = mel_bct.contiguous().float()
    B, C, T = x.shape
    dev = x.device
    X0 = torch.zeros(B * (T + 4), C, device=dev)
    L("t2v_bct_to_padded", x, X0, B, C, T, 0.0)
    Y5, _ = engine.postnet_forward(ops, P, X0, B, T, training, None, seed, dev)
    out = torch.empty(B, C, T, device=dev)
    L("t2v_padded_to_bct", Y5, None, out, B, C, T, None, 0.0)
    return out


def prenet(P, x, masks=None, seed=0, base=0):
    """Prenet.forward (model.py:91-102) on [..., 80] rows; dropout .5 always on (explicit masks or RNG)."""
    _need_cuda(x, "prenet")
    lead = x.shape[:-1]
    x2 = x.reshape(-1, x.shape[-1]).contiguous().float()
    n = x2.shape[0]
    dev = x.device
    h = x2
    for i in range(2):
        W = P["decoder.prenet.layers.%d.linear_layer.weight" % i]
        pre = torch.empty(n, W.shape[0], device=dev)
        engine.Ops.gemm(h, h.shape[1], 1, W, W.shape[1], 1, pre, W.shape[0], n, W.shape[0], W.shape[1])
        out = torch.empty_like(pre)
        m = None if masks is None else masks[i].contiguous().float()
        L("t2v_relu_drop_fwd", pre, out, W.shape[0], n, W.shape[0], m, seed, engine.SITE_PRENET + i, 0.5, base, 0)
        h = out
    return h.view(*lead, h.shape[1])


class DecoderSession(object):
    """The state Decoder.initialize_decoder_states keeps on the module (model.py:260-291) as t2v sequence buffers,
    advanced one step per `step()` (Decoder.decode, model.py:346-389)."""

    def __init__(self, ops, P, memory, in_len, max_steps, training=False, seed=0, mask_value=-float("inf")):
        _need_cuda(memory, "decoder memory")
        memory = memory.contiguous().float()
        self.ops, self.P = ops, P
        self.B, self.Ti, _ = memory.shape
        self.To = int(max_steps)
        dev = memory.device
        self.dev = dev
        self.W = {}
        self.W["Wa"], self.W["Wd"], self.W["Wpg"], self.W["bpg"] = engine.pack_decoder_weights(P, dev, ops.R)
        self.W["Wq"] = ops.wr(P["decoder.attention_layer.query_layer.linear_layer.weight"])
        self.W["WconvT"] = engine.conv_weight_T(P, dev)
        engine.pack_step_weights(ops, self.W, dev)       # tile-contiguous copies for the persistent loop kernel
        engine.pack_projection_lo(ops, P, self.W, dev)
        self.memory = memory
        self.pmem = torch.empty(self.B * self.Ti, 128, device=dev)
        ops.linear(memory, 512, ops.wr(P["decoder.attention_layer.memory_layer.linear_layer.weight"]), 512, self.pmem, 128,
                   self.B * self.Ti, 128, 512)
        self.buf = engine.alloc_decoder_buffers(self.B, self.Ti, self.To, dev, save=False, op16=ops.op16)
        self.O = torch.zeros(self.To * self.B, 84, device=dev)
        self.S = _lib.T2VDecoderSeq()
        self.in_len = None if in_len is None else in_len.long().contiguous()
        engine._fill_seq_struct(self.S, ops, P, self.W, self.B, self.Ti, self.To, training, seed, None, mask_value,
                                self.in_len, self.memory, self.pmem, self.buf)
        self.t = 0

    def step(self, prenet_out):
        if self.t >= self.To:
            raise RuntimeError("decoder session exhausted (%d steps)" % self.To)
        B, t = self.B, self.t
        x = prenet_out.contiguous().float()
        L("t2v_copy2d", x, 256, 1, engine._p(self.buf["XA"], t * B * 1792), 1792, B, 256, 0.0, self.ops.RX)
        L("t2v_decoder_fwd_steps", self.S, t, t + 1)
        if self.ops.op16:      # single steps run the per-step launches (fp32 rows): keep the 16-bit operand copies in step
            for k, w in (("XA", 1792), ("XD", 2560)):
                L("t2v_cvt16_2d", engine._p(self.buf[k], t * B * w), w, self.buf[k + "16"].data_ptr() + 2 * t * B * w, w,
                  2 * B, w, self.ops.op16)
        o = engine._p(self.O, t * B * 84)
        XD = self.buf["XD"]
        self.ops.linear(engine._p(XD, (t + 1) * B * 2560 + 1536), 2560, self.W["Wpg"], 1536, o, 84, B, 81, 1024,
                        bias=self.W["bpg"])
        self.ops.linear(engine._p(XD, t * B * 2560 + 1024), 2560, engine._p(self.W["Wpg"], 1024), 1536, o, 84, B, 81, 512,
                        accumulate=True)
        row = self.O[t * B:(t + 1) * B]
        self.t += 1
        return row[:, :80], row[:, 80:81], self.buf["align"][:, t, :]

    # ---- batched free-running decode on the device (config 5) ----
    def run_free(self, n_steps, gate_threshold=0.5, prenet_masks=None, seed=0):
        P, B = self.P, self.B
        D = _lib.T2VDecoderInfer()
        _lib.ctypes.memmove(_lib.ctypes.addressof(D.f), _lib.ctypes.addressof(self.S), _lib.ctypes.sizeof(_lib.T2VDecoderSeq))
        D.f.seed = seed
        P1 = torch.empty(2 * B, 256, device=self.dev)
        nfr = torch.full((B,), -1, device=self.dev, dtype=torch.int32)
        # the prenet and the mel / gate projection inside the persistent kernel are FFMA mat-vecs: exact fp32 weights
        Wp1 = P["decoder.prenet.layers.0.linear_layer.weight"]
        Wp2 = P["decoder.prenet.layers.1.linear_layer.weight"]
        D.Wp1, D.Wp2 = Wp1.data_ptr(), Wp2.data_ptr()
        D.Wpg, D.bpg = self.W.get("Wpg_x", self.W["Wpg"]).data_ptr(), self.W["bpg"].data_ptr()
        D.prenet_masks = _lib.ptr(prenet_masks)
        D.O, D.P1 = self.O.data_ptr(), P1.data_ptr()
        D.gate_threshold = gate_threshold
        D.n_frames = nfr.data_ptr()
        L("t2v_decoder_infer_steps", D, self.t, self.t + n_steps)
        self.t += n_steps
        self._keep = (P1, prenet_masks, Wp1, Wp2)
        return nfr

    def outputs(self, n=None):
        """-> mel [B,80,n], gate [B,n,1] (quirk Q7), align [B,n,Ti]"""
        n = self.t if n is None else n
        B = self.B
        O = self.O[:n * B].view(n, B, 84)
        mel = O[:, :, :80].permute(1, 2, 0).contiguous()
        gate = O[:, :, 80:81].permute(1, 0, 2).contiguous()
        return mel, gate, self.buf["align"][:, :n, :].contiguous()
