"""torch.autograd glue: the whole Tacotron2 forward is ONE autograd.Function whose backward is the hand-written
backward pass of t2v.engine, so the reference's train.py (loss.backward(), param.register_hook, clip_grad_norm_,
Adam.step) runs unchanged on top of the CUDA engine."""
import torch

from . import engine
from ._lib import call as L


class Tacotron2Function(torch.autograd.Function):
    """inputs: (cfg, text, in_len, mel_tgt, out_len, *params) in cfg.names order -> the 7 tensor outputs of
    Tacotron2.forward (model.py:545-547; the 8th, emotions, is a pass-through added by the caller)."""

    @staticmethod
    def forward(ctx, cfg, text, in_len, mel_tgt, out_len, *params):
        P = dict(zip(cfg.names, params))
        P.update(cfg.buffers)
        outs, c = engine.forward_train(cfg.ops, P, text, in_len, mel_tgt, out_len, training=cfg.training, rand=cfg.rand,
                                       seed=cfg.seed, mask_padding=cfg.mask_padding, mask_value=cfg.mask_value)
        ctx.cfg, ctx.c, ctx.P = cfg, c, P
        ctx.mark_non_differentiable(outs[3], outs[6])
        ctx.set_materialize_grads(True)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dmel, dpost, dgate, dalign, dmu, dlogvar, dz):
        cfg = ctx.cfg
        grads = engine.backward_train(cfg.ops, ctx.P, ctx.c, dmel.contiguous(), dpost.contiguous(), dgate.contiguous(),
                                      dmu.contiguous(), dlogvar.contiguous())
        ctx.c = None
        out = [None] * 5
        for n in cfg.names:
            out.append(grads.get(n))       # dead parameters (quirk Q6) stay None, like in the reference
        return tuple(out)


class VaeLossFunction(torch.autograd.Function):
    """loss_function.py:27-45 as one reduction kernel forward and one elementwise kernel backward."""

    @staticmethod
    def forward(ctx, mel, post, gate, mu, logvar, mel_tgt, gate_tgt, kl_weight):
        dev = mel.device
        acc = torch.zeros(4, device=dev, dtype=torch.float64)
        out = torch.empty(3, device=dev, dtype=torch.float32)
        args = [t.contiguous() for t in (mel, post, mel_tgt, gate, gate_tgt, mu, logvar)]
        L("t2v_loss_fwd", args[0], args[1], args[2], args[0].numel(), args[3], args[4], args[3].numel(), args[5], args[6],
          args[5].numel(), float(kl_weight), acc, out)
        ctx.save_for_backward(*args)
        ctx.kl_weight = float(kl_weight)
        return out[0], out[1], out[2]

    @staticmethod
    def backward(ctx, g_total, g_recon, g_kl):
        mel, post, tgt, gate, gtgt, mu, logvar = ctx.saved_tensors
        dmel, dpost, dgate = torch.empty_like(mel), torch.empty_like(post), torch.empty_like(gate)
        dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
        g = g_total.contiguous().float()
        L("t2v_loss_bwd", mel, post, tgt, mel.numel(), gate, gtgt, gate.numel(), mu, logvar, mu.numel(), ctx.kl_weight, g,
          dmel, dpost, dgate, dmu, dlv)
        return dmel, dpost, dgate, dmu, dlv, None, None, None
