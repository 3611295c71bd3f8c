"""Data-parallel train step on 2 GPUs over NCCL (reference distributed.py:126-174 + train.py:177-229), through the public surface:
apply_gradient_allreduce(model) -> model(x) -> loss.backward() -> gradients are the MEAN over ranks of the per-shard gradients.

BatchNorm uses per-process statistics (no SyncBN, as in the reference), so the check is: the all-reduced gradient of rank r equals
(g(shard 0) + g(shard 1)) / 2 with both shard gradients computed on one GPU by a model without the DP hook.  Skipped with fewer
than 2 devices."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, out):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tacotron2-vae_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    import distributed as t2v_dist
    import model as t2v_model
    from hparams import create_hparams
    from loss_function import Tacotron2Loss_VAE
    from oracle import port
    hp = create_hparams("anneal_function=constant")
    B, Ti, To = 6, 30, 40                                   # config 4's per-rank batch (README recipe: 6 / GPU)
    shards = [port.synthetic_batch(B, Ti, To, seed=10 + r) for r in range(world)]

    def grads_of(m, batch, seed_step):
        m.zero_grad()
        m._step = seed_step                                  # same dropout stream for the same shard in both models
        x, y = m.parse_batch(batch)
        out = m(x)
        loss, _, _, _ = Tacotron2Loss_VAE(hp)(out, y, 0)
        loss.backward()
        torch.cuda.synchronize()
        return {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}, float(loss)

    ref = t2v_model.Tacotron2(hp)
    ref.load_state_dict(port.init_params(1234))
    ref = ref.to(dev).train()
    ref._graph_cache = None
    per_shard = [grads_of(ref, shards[r], 100 + r)[0] for r in range(world)]
    dp = t2v_model.Tacotron2(hp)
    dp.load_state_dict(port.init_params(1234 if rank == 0 else 99))     # different init on rank 1: the broadcast must fix it
    dp = dp.to(dev).train()
    dp._graph_cache = None
    dp = t2v_dist.apply_gradient_allreduce(dp)
    g_dp, loss = grads_of(dp, shards[rank], 100 + rank)
    worst = 0.0
    for k, g in g_dp.items():
        want = sum(ps[k] for ps in per_shard) / world
        err = float((g - want).abs().max() / (want.abs().max() + 1e-20))
        worst = max(worst, err)
    w_equal = float((dp.decoder.attention_rnn.weight_hh - ref.decoder.attention_rnn.weight_hh).abs().max()) == 0.0
    out[rank] = dict(worst=worst, n=len(g_dp), loss=loss, w_equal=w_equal, is_view=all(
        p.grad.data_ptr() == v.data_ptr() for (_, p), v in zip(dp._t2v_flat_grads.named, dp._t2v_flat_grads.views)))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_gradients_are_the_mean_of_the_shard_gradients():
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    for r in range(2):
        print("rank %d: %d gradients, worst max-rel deviation from the shard mean %.2e, loss %.5f" % (r, out[r]["n"], out[r]["worst"], out[r]["loss"]))
        assert out[r]["n"] > 80
        assert out[r]["worst"] <= 2e-4, out[r]          # atomics order / split-K noise only
        assert out[r]["is_view"], "every .grad is a view of the flat all-reduce buffer"
        assert out[r]["w_equal"], "rank 0's parameters were broadcast"
