"""Data-parallel gradient exchange (reference distributed.py:126-174) on NCCL over NVLink 5 / NVSwitch.

`apply_gradient_allreduce(module)` keeps the reference's name and contract -- broadcast rank 0's state at start, then
after every backward SUM-all-reduce the gradients and divide by the world size -- with the reference's three extra
passes removed: gradients live in ONE persistent flat fp32 buffer (each `.grad` is a view into it, so there is no
flatten copy and no copy-back), the reduce is a single ncclAllReduce on that buffer, and the 1/world scale is either
applied by one axpby kernel or folded into the fused clip+Adam step (t2v.optim.FusedAdamClip)."""
import torch
import torch.distributed as dist
from torch.autograd import Variable

from t2v import optim as _optim


def apply_gradient_allreduce(module):
    for p in module.state_dict().values():                      # distributed.py:132-135
        if torch.is_tensor(p):
            dist.broadcast(p, 0)
    if getattr(module, "_t2v_allreduce_installed", False):      # train.py calls this twice (86-87 and 177-178)
        return module
    flat = _optim.FlatGrads(module)
    module._t2v_flat_grads = flat
    module.needs_reduction = False

    def allreduce_params():
        if module.needs_reduction:
            module.needs_reduction = False
            flat.adopt_grads()
            dist.all_reduce(flat.buffer)                        # one NCCL SUM over the whole gradient
            flat.scale_(1.0 / dist.get_world_size())

    def hook(*unused):
        Variable._execution_engine.queue_callback(allreduce_params)

    if hasattr(module, "_t2v_post_backward") or type(module).__name__ == "Tacotron2":
        # the B200 engine writes the gradients into the flat buffer itself and queues this callback from its backward
        # (no per-parameter autograd hooks: autograd never sees per-parameter gradients on that path)
        module._t2v_post_backward = list(getattr(module, "_t2v_post_backward", ())) + [allreduce_params]
    else:
        for p in module.parameters():
            if p.requires_grad:
                p.register_hook(hook)

    def set_needs_reduction(self, inp, out):
        self.needs_reduction = True

    module.register_forward_hook(set_needs_reduction)
    module._t2v_allreduce_installed = True
    return module
