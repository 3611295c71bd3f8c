"""Flat gradient buffer + fused clip/Adam step (reference train.py:171-172,226-229 and distributed.py:155-162)."""
import torch

from ._lib import call as L


class FlatGrads(object):
    """One persistent fp32 buffer holding the gradients of every parameter that can receive one (the reference's
    dead parameters, quirk Q6, are excluded so they are never all-reduced or stepped -- torch skips grad=None too)."""

    DEAD = ("speaker_embedding.", "emotion_embedding.", "vae_gst.ref_encoder.convs.0.weight", "vae_gst.ref_encoder.convs.0.bias")

    def __init__(self, module):
        self.named = [(k, p) for k, p in module.named_parameters() if p.requires_grad and not k.startswith(self.DEAD)]
        self.numel = sum(p.numel() for _, p in self.named)
        dev = self.named[0][1].device
        self.buffer = torch.zeros(self.numel, device=dev, dtype=torch.float32)
        self.views = []
        off = 0
        for _, p in self.named:
            self.views.append(self.buffer[off:off + p.numel()].view_as(p))
            off += p.numel()

    def adopt_grads(self):
        """make every .grad a view of the flat buffer (copying once if autograd produced a separate tensor)"""
        for (k, p), v in zip(self.named, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                v.copy_(p.grad)
            p.grad = v

    def scale_(self, a):
        L("t2v_axpby", self.buffer, float(a), self.buffer, 0.0, self.numel)


class FusedAdamClip(object):
    """clip_grad_norm_(max_norm) + torch.optim.Adam(lr, betas, eps, L2 weight_decay) as two kernels over flat buffers
    (t2v_grad_sumsq + t2v_adam_clip_step).  Parameters are re-pointed at one flat buffer so the update is one launch."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-6, max_norm=1.0):
        self.flat = getattr(module, "_t2v_flat_grads", None) or FlatGrads(module)
        module._t2v_flat_grads = self.flat
        n = self.flat.numel
        dev = self.flat.buffer.device
        self.params = torch.empty(n, device=dev, dtype=torch.float32)
        off = 0
        for _, p in self.flat.named:
            v = self.params[off:off + p.numel()].view_as(p)
            v.copy_(p.data)
            p.data = v
            off += p.numel()
        self.m = torch.zeros(n, device=dev)
        self.v = torch.zeros(n, device=dev)
        self.sumsq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.norm = torch.zeros(1, device=dev)
        self.lr, self.betas, self.eps, self.wd, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self._init_groups(module)

    def zero_grad(self, set_to_none=True):
        """like nn.Module.zero_grad(set_to_none=True): the next backward writes the flat buffer instead of accumulating"""
        for _, p in self.flat.named:
            p.grad = None
        if not set_to_none:
            self.flat.buffer.zero_()

    # ---- torch.optim.Optimizer-compatible surface (train.py:92-119, 171-172, 226-229: param_groups lr updates, checkpoints)
    @property
    def param_groups(self):
        return self._groups

    def _init_groups(self, module):
        allp = list(module.parameters())
        self._groups = [dict(params=allp, lr=self.lr, betas=tuple(self.betas), eps=self.eps, weight_decay=self.wd,
                             amsgrad=False, maximize=False)]
        pos = {id(p): i for i, p in enumerate(allp)}
        self._index = [pos[id(p)] for _, p in self.flat.named]      # position of each flat slice in module.parameters()

    def state_dict(self):
        """torch.optim.Adam layout: state[i] = {step, exp_avg, exp_avg_sq} for every parameter that has been stepped
        (the reference's dead parameters, quirk Q6, have no entry -- exactly like torch, which skips grad=None)"""
        state = {}
        if self.step_count > 0:
            off = 0
            for i, (_, p) in zip(self._index, self.flat.named):
                n = p.numel()
                state[i] = dict(step=torch.tensor(float(self.step_count)), exp_avg=self.m[off:off + n].view_as(p).clone(),
                                exp_avg_sq=self.v[off:off + n].view_as(p).clone())
                off += n
        g = dict(self._groups[0])
        g["params"] = list(range(len(self._groups[0]["params"])))
        return dict(state=state, param_groups=[g])

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self._groups[0].update({k: v for k, v in g.items() if k != "params"})
        self.m.zero_()
        self.v.zero_()
        steps = []
        off = 0
        for i, (_, p) in zip(self._index, self.flat.named):
            n = p.numel()
            st = sd["state"].get(i, sd["state"].get(str(i)))
            if st is not None:
                self.m[off:off + n].view_as(p).copy_(st["exp_avg"])
                self.v[off:off + n].view_as(p).copy_(st["exp_avg_sq"])
                steps.append(int(float(st["step"])))
            off += n
        self.step_count = max(steps) if steps else 0

    def step(self, grad_scale=1.0):
        """returns the (device) total gradient norm before clipping, like clip_grad_norm_"""
        self.flat.adopt_grads()
        self.step_count += 1
        g = self._groups[0]
        L("t2v_grad_sumsq", self.flat.buffer, self.flat.numel, float(grad_scale), self.sumsq)
        L("t2v_adam_clip_step", self.params, self.flat.buffer, self.m, self.v, self.flat.numel, self.sumsq, float(grad_scale),
          float(self.max_norm), float(g["lr"]), float(g["betas"][0]), float(g["betas"][1]), float(g["eps"]),
          float(g["weight_decay"]), self.step_count, self.norm)
        return self.norm
