"""Oracle parity at the BENCHED shapes and the BENCHED precisions.

The CUDA path runs exactly what bench.py times -- the persistent decoder-loop kernels (dec_persist_fwd_kernel / _bwd_kernel /
<INFER>), tensor-core GEMMs, fp16 (default) or tf32 operands -- and is compared with oracle/port.py (the CPU restatement of the
reference, pinned against the real reference by tests/test_oracle_cpu.py and tests/golden) on the same seeded inputs and the same
explicit dropout masks / eps.  Metric: rel-L1 = ||ours - ref||_1 / ||ref||_1; north star = 1e-3 on mel outputs.

Tolerances, stated per assert:
   train forward   mel, mel_post <= 1e-3 (fp16 / tf32 modes);  alignments max-abs <= 1e-3;  gate logits abs <= 2e-3;  loss rel <= 1e-3
   train backward  every live gradient: | ||g|| - ||g_ref|| | <= 1e-2 ||g_ref||;  selected full gradients rel-L2 <= 1e-2
   inference       C1 golden (200 steps) and config 5 (B=16, 128 steps) through the persistent <INFER> kernel: mel <= 1e-3
   bf16 mode       same metrics, achieved error printed, bound 2e-2 (8-bit significand operands in the decoder loops)
"""
import os

import numpy as np
import pytest
import torch

from oracle import port

pytestmark = pytest.mark.gpu


def _model(precision, seed=1234):
    import model as t2v_model
    from hparams import create_hparams
    hp = create_hparams("anneal_function=constant")
    m = t2v_model.Tacotron2(hp)
    m.load_state_dict(port.init_params(seed))
    m = m.cuda()
    m.precision = precision
    return m, hp


def _rand_to(rand, dev):
    r = port.Rand()
    r.enc = [x.to(dev) for x in rand.enc]
    r.prenet = [x.to(dev) for x in rand.prenet]
    r.dec = rand.dec.to(dev)
    r.post = [x.to(dev) for x in rand.post]
    r.eps = rand.eps.to(dev)
    return r


def _l1(a, b):
    return float((a - b).abs().sum() / (b.abs().sum() + 1e-30))


def _oracle_train(B, Ti, To, backward):
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    P = port.init_params(1234)
    batch = port.synthetic_batch(B, Ti, To, seed=0)
    rand = port.Rand.draw(B, Ti, To, seed=1)
    if not backward:
        with torch.no_grad():
            out = port.tacotron2_forward(P, batch[0], batch[1], batch[2], batch[4], True, rand)
            loss, recon, kl = port.vae_loss(out, batch[2], batch[3], 0.001)
        return batch, rand, out, float(loss), float(kl), None
    params = {k: v.clone().requires_grad_(True) for k, v in P.items() if v.dtype.is_floating_point and "running" not in k}
    st = dict(P)
    st.update(params)
    out = port.tacotron2_forward(st, batch[0], batch[1], batch[2], batch[4], True, rand)
    loss, recon, kl = port.vae_loss(out, batch[2], batch[3], 0.001)
    loss.backward()
    grads = {k: (None if v.grad is None else v.grad.detach()) for k, v in params.items()}
    return batch, rand, [o.detach() for o in out], float(loss), float(kl), grads


_ORACLE_CACHE = {}


def _oracle_cached(B, Ti, To, backward):
    key = (B, Ti, To, backward)
    if key not in _ORACLE_CACHE:
        _ORACLE_CACHE.clear()                       # one shape resident at a time (GBs of autograd state at C3)
        _ORACLE_CACHE[key] = _oracle_train(B, Ti, To, backward)
    return _ORACLE_CACHE[key]


OUT_TOL = {"fp16": dict(mel=1e-3, mel_post=1e-3, align=1e-3, gate=2e-3, loss=1e-3, grad=1e-2),
           "tf32": dict(mel=1e-3, mel_post=1e-3, align=1e-3, gate=2e-3, loss=1e-3, grad=1e-2),
           "bf16": dict(mel=2e-2, mel_post=2e-2, align=2e-2, gate=3e-2, loss=1e-2, grad=1e-1)}


@pytest.mark.parametrize("precision", ["fp16", "tf32", "bf16"])
def test_train_step_at_bench_batch_vs_oracle(precision):
    """B=64, Ti=120 (the bench's batch and text length), To=256: forward + loss + backward through the persistent kernels."""
    from loss_function import Tacotron2Loss_VAE
    B, Ti, To = 64, 120, 256
    batch, rand, ref, ref_loss, ref_kl, ref_grads = _oracle_cached(B, Ti, To, True)
    m, hp = _model(precision)
    m.train()
    m._rand = _rand_to(rand, "cuda")
    x, y = m.parse_batch(batch)
    out = m(x)
    loss, recon, kl, klw = Tacotron2Loss_VAE(hp)(out, y, 0)
    loss.backward()
    torch.cuda.synchronize()
    tol = OUT_TOL[precision]
    errs = {}
    for n, o, r in zip(("mel", "mel_post", "gate", "align", "mu", "logvar", "z"), out[:7], ref[:7]):
        o = o.detach().cpu()
        if n == "align":
            errs[n] = float((o - r).abs().max())
        elif n == "gate":
            valid = r < 999.0                                       # padded steps are exactly 1e3 on both sides
            assert bool((o[~valid] == 1e3).all())
            errs[n] = float((o[valid] - r[valid]).abs().max())
        else:
            errs[n] = _l1(o, r)
    errs["loss"] = abs(loss.item() - ref_loss) / abs(ref_loss)
    print("[%s B=%d Ti=%d To=%d] " % (precision, B, Ti, To) + "  ".join("%s %.2e" % kv for kv in errs.items()))
    for n in ("mel", "mel_post", "align", "gate", "loss"):
        assert errs[n] <= tol[n], (precision, n, errs[n])
    for n in ("mu", "logvar", "z"):
        assert errs[n] <= 1e-3, (n, errs[n])
    out_len = batch[4]
    for b in (0, B // 2, B - 1):                                    # padded frames exactly 0 (model.py:515-517)
        assert float(out[0][b, :, int(out_len[b]):].abs().sum()) == 0.0
        assert float(out[1][b, :, int(out_len[b]):].abs().sum()) == 0.0
    # ---- gradients: norm of every parameter, full tensors for one parameter per sub-module (incl. quirk Q10's conv)
    total = float(torch.sqrt(sum((g ** 2).sum() for g in ref_grads.values() if g is not None)))
    params = dict(m.named_parameters())
    worst = ("", 0.0)
    for k, gr in ref_grads.items():
        g = params[k].grad
        if gr is None:
            assert g is None, k
            continue
        assert g is not None, k
        gn, rn = float(g.norm()), float(gr.norm())
        if k.endswith(".0.conv.bias") or (k.startswith("vae_gst.ref_encoder.convs.") and k.endswith(".bias")):
            # a conv bias in front of a training-mode BatchNorm has an exactly-zero true gradient: both sides are rounding noise
            assert gn <= 1e-2 * total and rn <= 1e-2 * total, (k, gn, rn, total)
            continue
        rel = abs(gn - rn) / (rn + 1e-4 * total)
        if rel > worst[1]:
            worst = (k, rel)
        assert rel <= tol["grad"], (k, gn, rn)
    full = ["postnet.convolutions.0.0.conv.weight", "decoder.attention_rnn.weight_hh", "decoder.decoder_rnn.weight_ih",
            "decoder.attention_layer.location_layer.location_conv.conv.weight", "decoder.prenet.layers.0.linear_layer.weight",
            "decoder.linear_projection.linear_layer.weight", "encoder.lstm.weight_hh_l0_reverse", "transcript_embedding.weight",
            "vae_gst.fc1.weight", "encoder.convolutions.0.0.conv.weight"]
    for k in full:
        g, gr = params[k].grad.cpu(), ref_grads[k]
        rel = float((g - gr).norm() / gr.norm())
        print("   grad %-70s rel-L2 %.2e" % (k, rel))
        assert rel <= 3 * tol["grad"], (k, rel)
    print("   worst gradient-norm deviation: %s %.2e" % worst)


@pytest.mark.parametrize("precision", ["fp16"])
def test_forward_at_full_c3_shape_vs_oracle(precision):
    """The exact bench workload (B=64, Ti=120, To=800), forward + loss only (the oracle's autograd state would be >10 GB)."""
    from loss_function import Tacotron2Loss_VAE
    B, Ti, To = 64, 120, 800
    batch, rand, ref, ref_loss, ref_kl, _ = _oracle_cached(B, Ti, To, False)
    m, hp = _model(precision)
    m.train()
    m._rand = _rand_to(rand, "cuda")
    x, y = m.parse_batch(batch)
    with torch.no_grad():
        out = m(x)
        loss, recon, kl, klw = Tacotron2Loss_VAE(hp)(out, y, 0)
    e_mel, e_post = _l1(out[0].cpu(), ref[0]), _l1(out[1].cpu(), ref[1])
    e_align = float((out[3].cpu() - ref[3]).abs().max())
    e_loss = abs(loss.item() - ref_loss) / abs(ref_loss)
    print("[%s C3 full] mel %.2e  mel_post %.2e  align %.2e  loss %.2e" % (precision, e_mel, e_post, e_align, e_loss))
    assert e_mel <= 1e-3 and e_post <= 1e-3, (e_mel, e_post)
    assert e_align <= 2e-3 and e_loss <= 1e-3, (e_align, e_loss)


def _c1_inputs():
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(2, 79, (1, 40), generator=g)
    ids[0, -1] = 1
    refmel = torch.randn(1, 80, 200, generator=g)
    pm = (torch.rand(200, 2, 1, 256, generator=g) >= 0.5).float()
    return ids, refmel, pm


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_inference_c1_golden_through_persistent_kernel(golden_dir, precision):
    """Config 1 (single utterance, 200 steps) against the golden made from the real reference, through the persistent <INFER>
    kernel in the benched precision (the fp32-mode test in test_gpu_parity.py takes the per-step launches)."""
    G = np.load(os.path.join(golden_dir, "inference_c1.npz"))
    m, hp = _model(precision)
    m.eval()
    ids, refmel, pm = _c1_inputs()
    from t2v import _lib
    with torch.no_grad():
        emb = m.transcript_embedding(ids.cuda()).transpose(1, 2)
        enc = m.encoder.inference(emb)
        style, mu, logvar, z = m.vae_gst(refmel.cuda())
        from model import _add_style
        mem = _add_style(enc, style)
        dec = m.decoder
        dec.initialize_decoder_states(mem, mask=None)
        n0 = _lib.launch_count()
        dec._session.run_free(200, 0.5, prenet_masks=pm.cuda().contiguous())
        assert _lib.launch_count() - n0 <= 3, "the free-running loop must be ONE persistent launch (+ memset)"
        mel, gate, align = dec._session.outputs(200)
        post = m.postnet(mel)
    e_mel = _l1(mel.cpu(), torch.from_numpy(G["mel"]))
    e_align = float((align.cpu() - torch.from_numpy(G["align"])).abs().max())
    e_post = _l1((post + mel).cpu(), torch.from_numpy(G["mel_post"]))
    print("[%s C1 persistent infer] mel %.2e  align(max-abs) %.2e  mel_post %.2e" % (precision, e_mel, e_align, e_post))
    assert e_mel <= 1e-3, e_mel
    assert e_align <= 1e-3, e_align
    assert e_post <= 1e-3, e_post
    assert tuple(gate.shape) == (1, 200, 1)


@pytest.mark.parametrize("precision", ["fp16", "tf32"])
def test_inference_config5_vs_oracle(precision):
    """Config 5: B=16, Ti=120 free-running decode.  128 steps against port.decoder_free_running with the same prenet masks
    (mel rel-L1 <= 1e-3, alignments max-abs <= 2e-3), then the full 1000 steps: finite, softmax rows sum to 1, stop flags
    agree with the recorded gate logits."""
    from t2v import engine, infer
    B, Ti, n = 16, 120, 128
    dev = torch.device("cuda")
    P = port.init_params(1234)
    g = torch.Generator().manual_seed(55)
    mem = torch.randn(B, Ti, 512, generator=g)
    pm = (torch.rand(1000, 2, B, 256, generator=g) >= 0.5).float()
    torch.set_num_threads(max(1, min(32, os.cpu_count() or 1)))
    with torch.no_grad():
        rmel, rgate, ralign = port.decoder_free_running(P, mem, n, pm[:n], training=False)
    Pd = {k: v.to(dev) for k, v in P.items()}
    ops = engine.Ops(precision)
    sess = infer.DecoderSession(ops, Pd, mem.to(dev), None, 1000, training=False, seed=5)
    pmd = pm.to(dev).contiguous()
    nfr = sess.run_free(1000, 0.5, prenet_masks=pmd)
    torch.cuda.synchronize()
    mel, gate, align = sess.outputs(1000)
    e_mel = _l1(mel[:, :, :n].cpu(), rmel)
    e_gate = float((gate[:, :n].cpu() - rgate).abs().max())
    e_align = float((align[:, :n].cpu() - ralign).abs().max())
    print("[%s C5 B=16 %d steps] mel %.2e  gate(max-abs) %.2e  align(max-abs) %.2e" % (precision, n, e_mel, e_gate, e_align))
    assert e_mel <= 1e-3, e_mel
    assert e_align <= 2e-3, e_align
    assert e_gate <= 2e-3, e_gate
    assert bool(torch.isfinite(mel).all()) and bool(torch.isfinite(gate).all())
    assert torch.allclose(align.sum(-1), torch.ones(B, 1000, device=dev), atol=1e-4)
    # stop bookkeeping (model.py:453-455): first step whose sigmoid(gate) > 0.5, +1; -1 if never
    hit = torch.sigmoid(gate[:, :, 0]) > 0.5
    first = torch.where(hit.any(1), hit.float().argmax(1) + 1, torch.full((B,), -1, device=dev, dtype=torch.long))
    assert torch.equal(first.int(), nfr.int()), (first, nfr)
