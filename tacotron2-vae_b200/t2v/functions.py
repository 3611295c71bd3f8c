"""torch.autograd glue: the whole Tacotron2 forward is ONE autograd.Function whose backward is the hand-written
backward pass of t2v.engine, so the reference's train.py (loss.backward(), param.register_hook, clip_grad_norm_,
Adam.step) runs unchanged on top of the CUDA engine."""
import os

import torch

from . import _lib, engine
from ._lib import call as L


_GRAPH_AFTER = int(__import__("os").environ.get("T2V_GRAPH_AFTER", "2"))     # capture a shape on its n-th sighting
_GRAPH_MAX = int(__import__("os").environ.get("T2V_GRAPH_MAX", "3"))         # graphs kept resident (LRU)


_CAPTURE_STREAMS = {}


def _capture_stream(dev):
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        # three tiers: the capture stream (the step's main chain) above the urgent side branches (engine._Branch(urgent=True): -1)
        # above the weight-gradient branches (0); torch clamps to the device's range
        prio = -2 if os.environ.get("T2V_CAPTURE_PRIORITY", "1") != "0" else 0
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=key, priority=prio)
    return _CAPTURE_STREAMS[key]


def flat_copy_runs(names, sizes, ptrs, base, itemsize=4):
    """Which slices of a flat buffer still have to be filled by a copy: `names[i]` owns elements [off_i, off_i + sizes[i]) of the buffer
    at address `base`; a tensor whose data pointer `ptrs[i]` already IS its slice (a producer wrote it in place; falsy = unknown / not
    contiguous) needs nothing.  Returns [(lo, hi, [names...])]: maximal runs of consecutive tensors to concatenate into buffer[lo:hi]."""
    runs, off, cur, lo = [], 0, [], 0
    for n, sz, p in zip(names, sizes, ptrs):
        if p and p == base + itemsize * off:
            if cur:
                runs.append((lo, off, cur))
                cur = []
        else:
            if not cur:
                lo = off
            cur.append(n)
        off += sz
    if cur:
        runs.append((lo, off, cur))
    return runs


class GraphedStep(object):
    """One (shape, mode) instance of the train step captured as two CUDA graphs (forward, backward).

    The engine's forward/backward are fixed launch sequences for a given shape (lengths, seeds and masks live in device
    memory), so capturing them removes the per-launch host cost (~10 us x ~16k launches per step on the B200 hosts) and
    pins every intermediate at a fixed address.  Inputs are copied into static tensors before a replay; the dropout /
    eps seed is a device scalar the kernels dereference (seed argument with bit 63 set, see include/t2v_b200.h)."""

    def __init__(self, cfg, P, text, in_len, mel, out_len):
        dev = text.device
        self.names = list(cfg.names)
        self.s_in = [text.clone(), in_len.clone(), mel.clone(), out_len.clone()]
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        seed_arg = (1 << 63) | self.seed_dev.data_ptr()
        kw = dict(training=cfg.training, rand=None, seed=seed_arg, mask_padding=cfg.mask_padding, mask_value=cfg.mask_value)
        snap = {k: v.clone() for k, v in cfg.buffers.items()}        # the warm-up below must not advance BN statistics
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        self.with_backward = bool(getattr(cfg, "need_grad", True))
        with torch.cuda.stream(side):                                # eager warm-up: lazy inits (smem attributes, TMA entry point)
            outs, c = engine.forward_train(cfg.ops, P, *self.s_in, **kw)
            if self.with_backward:
                engine.backward_train(cfg.ops, P, c, *[torch.zeros_like(outs[i]) for i in (0, 1, 2, 4, 5)])
            del outs, c
        cur.wait_stream(side)
        torch.cuda.synchronize()
        self.g_fwd = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        # captured on a HIGH-priority stream: the graph's main chain (text encoder -> decoder loops -> encoder backward) is the
        # critical path of the step; the side branches (engine._Branch: default priority) fill the SMs it leaves free
        cap = _capture_stream(dev)
        with torch.cuda.graph(self.g_fwd, stream=cap, capture_error_mode="thread_local"):
            self.outs, self.c = engine.forward_train(cfg.ops, P, *self.s_in, **kw)
        self.n_fwd = _lib.launch_count() - n0
        self.direct = False
        if not self.with_backward:           # eval / no_grad forward (validate(), train.py:122-147): nothing to differentiate
            self.c = None
            for k, v in cfg.buffers.items():
                v.copy_(snap[k])
            torch.cuda.synchronize()
            return
        self.s_dout = [torch.zeros_like(self.outs[i]) for i in (0, 1, 2, 4, 5)]
        self.g_bwd = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_bwd, pool=self.g_fwd.pool(), stream=cap, capture_error_mode="thread_local"):
            # the gradients land in ONE flat buffer inside the graph.  When the model owns a persistent flat gradient buffer with
            # the same layout (t2v.optim.FlatGrads: every `.grad` is a view of it, the all-reduce and the fused clip+Adam step
            # run on it) the graph writes straight into that buffer: no clone, no per-parameter copies afterwards; the producers
            # of the large gradients (decoder LSTM matrices, convolution weights: 90 % of the bytes) fill their slices themselves,
            # only the runs of small gradients in between are concatenated into place.
            fg = getattr(cfg, "flat", None)
            dst = {k: v for (k, _), v in zip(fg.named, fg.views)} if fg is not None else None
            grads = engine.backward_train(cfg.ops, P, self.c, *self.s_dout, grad_dst=dst)
            self.live = [n for n in self.names if n in grads]
            self.sizes = [grads[n].numel() for n in self.live]
            self.shapes = [grads[n].shape for n in self.live]
            self.direct = fg is not None and [k for k, _ in fg.named] == self.live and fg.numel == sum(self.sizes)
            self.fg = fg if self.direct else None
            if self.direct:
                self.flat = fg.buffer
                in_place = [grads[n].is_contiguous() and grads[n].data_ptr() for n in self.live]
                for lo, hi, names in flat_copy_runs(self.live, self.sizes, in_place, fg.buffer.data_ptr()):
                    torch.cat([grads[n].reshape(-1) for n in names], out=self.flat[lo:hi])
            else:
                self.flat = torch.cat([grads[n].reshape(-1) for n in self.live])
            del grads
        self.n_bwd = _lib.launch_count() - n0 - self.n_fwd
        for k, v in cfg.buffers.items():
            v.copy_(snap[k])
        torch.cuda.synchronize()

    def forward(self, text, in_len, mel, out_len, seed):
        for s, t in zip(self.s_in, (text, in_len, mel, out_len)):
            s.copy_(t, non_blocking=True)
        self.seed_dev.fill_(int(seed) & 0x7FFFFFFFFFFFFFFF)
        self.g_fwd.replay()
        _lib.add_launch_count(self.n_fwd)
        return [o.clone() for o in self.outs]

    def backward_direct(self, douts):
        """replay the backward graph into the model's flat gradient buffer and make every `.grad` the view of its slice
        (accumulating onto gradients that are already there, like autograd would)"""
        for s, t in zip(self.s_dout, douts):
            s.copy_(t, non_blocking=True)
        fg = self.fg
        prev = None
        if any(p.grad is not None for _, p in fg.named):
            fg.adopt_grads()                     # whatever the caller accumulated so far now sits in the flat buffer
            prev = fg.buffer.clone()
        self.g_bwd.replay()
        _lib.add_launch_count(self.n_bwd)
        if prev is not None:
            fg.buffer.add_(prev)
        for (_, p), v in zip(fg.named, fg.views):
            p.grad = v

    def backward(self, douts):
        for s, t in zip(self.s_dout, douts):
            s.copy_(t, non_blocking=True)
        self.g_bwd.replay()
        _lib.add_launch_count(self.n_bwd)
        flat = self.flat.clone()
        out, off = {}, 0
        for n, sz, shp in zip(self.live, self.sizes, self.shapes):
            out[n] = flat[off:off + sz].view(shp)
            off += sz
        return out


class Tacotron2Function(torch.autograd.Function):
    """inputs: (cfg, text, in_len, mel_tgt, out_len, *params) in cfg.names order -> the 7 tensor outputs of
    Tacotron2.forward (model.py:545-547; the 8th, emotions, is a pass-through added by the caller)."""

    @staticmethod
    def forward(ctx, cfg, text, in_len, mel_tgt, out_len, *params):
        P = dict(zip(cfg.names, params))
        P.update(cfg.buffers)
        ctx.gs = None
        gs = None
        if cfg.graph_cache is not None and cfg.rand is None:
            # A graph is tied to one exact (B, Ti, To): TextMelCollate pads every batch to its own maxima (data_utils.py:82-137) and
            # padding further would change the results (BN statistics and the MSE denominators include padded positions, quirk
            # Q4), so shapes are never bucketed.  A shape is captured only once it REPEATS (fixed-shape training / the bench /
            # bucketed samplers); first sightings run the eager launch sequence, so variable-length data costs nothing extra.
            need_grad = bool(getattr(cfg, "need_grad", True))
            key = (tuple(text.shape), tuple(mel_tgt.shape), cfg.training, need_grad, cfg.ops.precision, cfg.mask_padding,
                   cfg.mask_value, cfg.ops.p_att, cfg.ops.p_dec, params[0].data_ptr(), params[-1].data_ptr())
            gs = cfg.graph_cache.get(key)
            if gs is None:
                seen = cfg.graph_cache.setdefault("_seen", {})
                seen[key] = seen.get(key, 0) + 1
                if len(seen) > 64:
                    seen.pop(next(iter(seen)))
                if seen[key] >= _GRAPH_AFTER:
                    live = [k for k in cfg.graph_cache if k != "_seen"]
                    if len(live) >= _GRAPH_MAX:                      # bounded: every entry pins GBs of workspace
                        cfg.graph_cache.pop(live[0])
                    gs = cfg.graph_cache[key] = GraphedStep(cfg, P, text, in_len, mel_tgt, out_len)
            else:
                cfg.graph_cache[key] = cfg.graph_cache.pop(key)      # LRU order
        if gs is not None:
            outs = gs.forward(text, in_len, mel_tgt, out_len, cfg.seed)
            ctx.gs = gs
            ctx.cfg, ctx.c, ctx.P = cfg, None, None
            ctx.mark_non_differentiable(outs[3], outs[6])
            ctx.set_materialize_grads(True)
            return tuple(outs)
        outs, c = engine.forward_train(cfg.ops, P, text, in_len, mel_tgt, out_len, training=cfg.training, rand=cfg.rand,
                                       seed=cfg.seed, mask_padding=cfg.mask_padding, mask_value=cfg.mask_value)
        ctx.cfg, ctx.c, ctx.P = cfg, c, P
        ctx.mark_non_differentiable(outs[3], outs[6])
        ctx.set_materialize_grads(True)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dmel, dpost, dgate, dalign, dmu, dlogvar, dz):
        cfg = ctx.cfg
        # module-level "after backward" callbacks (distributed.apply_gradient_allreduce): run when the whole backward is done
        for cb in getattr(cfg, "post_backward", ()):
            torch.autograd.Variable._execution_engine.queue_callback(cb)
        if ctx.gs is not None and ctx.gs.direct and not cfg.param_hooks:
            # fast path: gradients go straight into the flat buffer the optimizer / all-reduce work on; autograd sees no
            # parameter gradients (None), so the module-level callbacks stand in for per-parameter hooks
            ctx.gs.backward_direct((dmel, dpost, dgate, dmu, dlogvar))
            return tuple([None] * (5 + len(cfg.names)))
        if ctx.gs is not None:
            grads = ctx.gs.backward((dmel, dpost, dgate, dmu, dlogvar))
        else:
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
        total, recon, kl = out[0], out[1], out[2]
        # the reference only ever backpropagates the total (train.py:215-222); recon / kl are reporting values.  Marking them
        # non-differentiable makes a backward through them fail loudly instead of silently returning zero gradients.
        ctx.mark_non_differentiable(recon, kl)
        return total, recon, kl

    @staticmethod
    def backward(ctx, g_total, g_recon, g_kl):
        mel, post, tgt, gate, gtgt, mu, logvar = ctx.saved_tensors
        dmel, dpost, dgate = torch.empty_like(mel), torch.empty_like(post), torch.empty_like(gate)
        dmu, dlv = torch.empty_like(mu), torch.empty_like(logvar)
        g = g_total.contiguous().float()
        L("t2v_loss_bwd", mel, post, tgt, mel.numel(), gate, gtgt, gate.numel(), mu, logvar, mu.numel(), ctx.kl_weight, g,
          dmel, dpost, dgate, dmu, dlv)
        return dmel, dpost, dgate, dmu, dlv, None, None, None
