"""Pins oracle/port.py (the CPU restatement) against the real reference's outputs:
 - committed fixtures tests/golden/*.npz (made by oracle/gen_golden.py from /root/reference), and
 - when /root/reference is present (build container), a live run of the reference itself."""
import os

import numpy as np
import pytest
import torch

from oracle import port


def _grad_P(P):
    out = {}
    for k, v in P.items():
        out[k] = v.clone().requires_grad_(True) if (v.dtype.is_floating_point and "running" not in k) else v.clone()
    return out


def _run_port_train(B, Ti, To):
    P = _grad_P(port.init_params(1234))
    text, in_len, mel, gate, out_len, _, _ = port.synthetic_batch(B, Ti, To, seed=0)
    rand = port.Rand.draw(B, Ti, To, seed=1)
    stats = {}
    out = port.tacotron2_forward(P, text, in_len, mel, out_len, True, rand, stats)
    loss, recon, kl = port.vae_loss(out, mel, gate, port.kl_weight("constant", 0))
    loss.backward()
    return P, out, (loss, recon, kl), stats


@pytest.mark.parametrize("tag,B,Ti,To", [("b3", 3, 20, 30), ("b4", 4, 40, 64)])
def test_port_train_step_matches_golden(golden_dir, tag, B, Ti, To):
    G = np.load(os.path.join(golden_dir, "train_step_%s.npz" % tag))
    P, out, (loss, recon, kl), stats = _run_port_train(B, Ti, To)
    for n, o in zip(("mel", "mel_post", "gate", "align", "mu", "logvar", "z"), out):
        ref = torch.from_numpy(G[n])
        assert torch.allclose(o.detach(), ref, rtol=1e-4, atol=2e-5), n
    assert abs(loss.item() - float(G["loss"])) <= 1e-5 * abs(float(G["loss"]))
    assert abs(kl.item() - float(G["kl"])) <= 1e-4 * abs(float(G["kl"]))
    # padded frames exactly 0 / gate exactly 1e3 (model.py:515-517)
    _, _, _, _, out_len, _, _ = port.synthetic_batch(B, Ti, To, seed=0)
    for b in range(B):
        assert float(out[0][b, :, int(out_len[b]):].abs().sum()) == 0.0
        assert bool((out[2][b, int(out_len[b]):] == 1e3).all())
    total = float(np.sqrt(np.sum(np.maximum(G["grad_norms"], 0) ** 2)))
    for k, gn, gp in zip(G["grad_names"], G["grad_norms"], G["grad_probes"]):
        g = P[str(k)].grad
        if gn < 0:                       # dead parameter (quirk Q6): never receives a gradient
            assert g is None or float(g.abs().max()) == 0.0, k
            continue
        # conv biases in front of a training-mode BN have an exactly-zero true gradient: noise only
        tol = 1e-3 * gn + 1e-6 * total
        assert abs(float(g.norm()) - gn) <= tol, (k, float(g.norm()), gn)
        probe = float((g * port.grad_probe(str(k), g.shape)).sum())
        assert abs(probe - gp) <= 2e-3 * gn * np.sqrt(g.numel()) ** 0 + 1e-5 * total + 1e-3 * abs(gp), (k, probe, gp)
    for k in G.files:
        if k.startswith("grad::"):
            g = P[k[6:]].grad
            ref = torch.from_numpy(G[k])
            assert float((g - ref).norm() / ref.norm()) < 1e-4, k
        if k.startswith("buf::"):
            assert torch.allclose(stats[k[5:]].float(), torch.from_numpy(G[k]).float(), rtol=1e-5, atol=1e-6), k


def test_port_inference_c1_matches_golden(golden_dir):
    G = np.load(os.path.join(golden_dir, "inference_c1.npz"))
    P = port.init_params(1234)
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(2, 79, (1, 40), generator=g); ids[0, -1] = 1
    refmel = torch.randn(1, 80, 200, generator=g)
    pm = (torch.rand(200, 2, 1, 256, generator=g) >= 0.5).float()
    with torch.no_grad():
        emb = torch.nn.functional.embedding(ids, P["transcript_embedding.weight"]).transpose(1, 2)
        enc = port.encoder(P, emb, None, False)
        style, mu, _, _ = port.vae_gst(P, refmel, False)
        mel, gate, align = port.decoder_free_running(P, enc + style.unsqueeze(1), 200, pm)
        post = port.postnet(P, mel, False) + mel
    assert torch.allclose(enc, torch.from_numpy(G["enc"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(style, torch.from_numpy(G["style"]), rtol=1e-4, atol=1e-5)
    for name, o in (("mel", mel), ("align", align), ("mel_post", post)):
        ref = torch.from_numpy(G[name])
        assert float((o - ref).abs().sum() / ref.abs().sum()) < 1e-3, name     # free-running: errors feed back
    assert tuple(G["gate"].shape) == (1, 200, 1)                                # quirk Q7


def test_port_stft_matches_golden(golden_dir):
    G = np.load(os.path.join(golden_dir, "stft_mel.npz"))
    g = torch.Generator().manual_seed(3)
    wav = torch.rand(2, 16000, generator=g) * 2 - 1
    wav[1] *= torch.linspace(0, 1, 16000)
    mel = port.mel_spectrogram(wav)
    assert np.allclose(port.mel_filterbank(), G["mel_basis"], rtol=1e-6, atol=1e-8)
    assert torch.allclose(mel, torch.from_numpy(G["mel"]), rtol=1e-4, atol=1e-4)


def test_port_stft_matches_golden_speech(golden_dir):
    """one-second excerpts of the reference's eight recordings (samples/refs/*.wav) through the reference's
    wav -> /32768 -> TacotronSTFT path (data_utils.py:42-59): golden made by oracle/gen_golden.py::stft_speech_fixture"""
    G = np.load(os.path.join(golden_dir, "stft_speech.npz"))
    names = sorted(k[:-4] for k in G.files if k.endswith("_mel"))
    assert len(names) == 8, names                     # all of samples/refs/*.wav
    for name in names:
        wav = torch.from_numpy(G[name + "_wav_i16"].astype(np.float32)) / 32768.0
        mel = port.mel_spectrogram(wav[None])[0]
        ref = torch.from_numpy(G[name + "_mel"])
        assert tuple(mel.shape) == tuple(ref.shape)
        assert torch.allclose(mel, ref, rtol=1e-4, atol=1e-4), name


@pytest.mark.reference
def test_port_matches_live_reference():
    from oracle import ref_shims
    ref = ref_shims.load_reference()
    hp = ref.hparams.create_hparams("anneal_function=constant")
    m = ref.model.Tacotron2(hp)
    m.load_state_dict(port.init_params(99))
    m.train()
    B, Ti, To = 2, 12, 17
    batch = port.synthetic_batch(B, Ti, To, seed=5)
    rand = port.Rand.draw(B, Ti, To, seed=6)
    x, y = m.parse_batch(batch)
    with ref_shims.InjectedRandomness(rand.as_reference_call_list(), rand.eps):
        out = m(x)
    P = port.init_params(99)
    with torch.no_grad():
        o2 = port.tacotron2_forward(P, batch[0], batch[1], batch[2], batch[4], True, rand)
    for a, b in zip(out[:7], o2):
        assert torch.allclose(a, b, rtol=1e-4, atol=2e-5)
    assert set(port.param_shapes()) == set(m.state_dict().keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(port.param_shapes()[k]), k
