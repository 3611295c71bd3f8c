"""N>1 host-side logic of the data-parallel path on CPU (gloo, world size 2): the flat gradient buffer, the
apply_gradient_allreduce contract (distributed.py:126-174: rank-0 broadcast, post-backward SUM/world) and dead-parameter
handling.  The NCCL path itself runs in bench.py --gpus N on the GPU box."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(4, 3)
        self.speaker_embedding = torch.nn.Linear(2, 2)     # "dead" by FlatGrads' naming rule (quirk Q6)

    def forward(self, x):
        return self.a(x).sum()


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tacotron2-vae_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from t2v import optim
    import distributed as t2v_dist
    orig = optim.FlatGrads.scale_
    optim.FlatGrads.scale_ = lambda self, a: self.buffer.mul_(a)        # CPU stand-in for the t2v_axpby kernel
    torch.manual_seed(rank)                                              # different init per rank: broadcast must fix it
    m = _Toy()
    m = t2v_dist.apply_gradient_allreduce(m)
    m = t2v_dist.apply_gradient_allreduce(m)                             # train.py calls it twice (86-87, 177-178)
    w0 = m.a.weight.detach().clone()
    x = torch.full((5, 4), float(rank + 1))
    m(x).backward()
    flat = m._t2v_flat_grads
    out[rank] = dict(w0=w0, g=m.a.weight.grad.clone(), is_view=m.a.weight.grad.data_ptr() == flat.views[0].data_ptr(),
                     n=flat.numel, dead=[k for k, _ in flat.named])
    optim.FlatGrads.scale_ = orig
    dist.destroy_process_group()


def test_gradient_allreduce_gloo_world2():
    mgr = mp.Manager()
    out = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    r0, r1 = out[0], out[1]
    assert torch.equal(r0["w0"], r1["w0"])                               # rank-0 state was broadcast
    expect = (torch.full((3, 4), 5.0 * 1) + torch.full((3, 4), 5.0 * 2)) / 2   # mean over ranks of sum_x
    assert torch.allclose(r0["g"], expect) and torch.allclose(r1["g"], expect)
    assert r0["is_view"] and r1["is_view"]
    assert r0["n"] == 4 * 3 + 3 and all("speaker_embedding" not in k for k in r0["dead"])
