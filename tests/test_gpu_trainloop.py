"""The call sequence of the reference's train.py (train.py:150-250: DataLoader(TextMelLoader, TextMelCollate) ->
parse_batch -> model(x) -> criterion -> backward -> clip_grad_norm_ -> Adam.step -> validate() in eval mode ->
save/load checkpoint) on this engine, with wav files generated on the fly (mel extraction by the t2v STFT kernels)."""
import copy
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _make_corpus(tmp_path, n=6):
    from scipy.io.wavfile import write
    rng = np.random.RandomState(0)
    lines = []
    texts = ["감정있는 한국어 목소리 생성", "안녕하세요 반갑습니다", "오늘 날씨가 좋네요", "사과 3개 주세요", "테스트 문장입니다", "음성 합성"]
    for i in range(n):
        secs = 0.25 + 0.1 * (i % 3)
        t = np.arange(int(16000 * secs)) / 16000.0
        wav = 0.4 * np.sin(2 * np.pi * (200 + 50 * i) * t) + 0.05 * rng.randn(t.size)
        path = os.path.join(tmp_path, "u%d.wav" % i)
        write(path, 16000, (np.clip(wav, -1, 1) * 32767).astype(np.int16))
        lines.append("%s|%s|0|%d" % (path, texts[i % len(texts)], i % 4))
    fl = os.path.join(tmp_path, "files.txt")
    open(fl, "w", encoding="utf-8").write("\n".join(lines))
    return fl


def test_reference_train_loop_call_sequence(tmp_path):
    from torch.utils.data import DataLoader
    import model as t2v_model
    from data_utils import TextMelCollate, TextMelLoader
    from hparams import create_hparams
    from logger import Tacotron2Logger
    from loss_function import Tacotron2Loss_VAE
    fl = _make_corpus(str(tmp_path))
    hp = create_hparams("batch_size=3,anneal_function=constant,training_files=%s,validation_files=%s" % (fl, fl))
    torch.manual_seed(hp.seed)
    model = t2v_model.Tacotron2(hp).cuda()
    optimizer = torch.optim.Adam(model.parameters(), lr=hp.learning_rate, weight_decay=hp.weight_decay)
    criterion = Tacotron2Loss_VAE(hp)
    logger = Tacotron2Logger(os.path.join(str(tmp_path), "logs"))
    trainset = TextMelLoader(hp.training_files, hp)
    loader = DataLoader(trainset, num_workers=0, shuffle=False, batch_size=hp.batch_size, drop_last=True,
                        collate_fn=TextMelCollate(hp.n_frames_per_step))
    model.train()
    losses = []
    iteration = 0
    for epoch in range(3):
        for batch in loader:
            model.zero_grad()
            x, y = model.parse_batch(batch)
            y_pred = model(x)
            loss, recon, kl, klw = criterion(y_pred, y, iteration)
            reduced = loss.item()
            loss.backward()
            gn = torch.nn.utils.clip_grad_norm_(model.parameters(), hp.grad_clip_thresh)
            optimizer.step()
            logger.log_training(reduced, gn, hp.learning_rate, 0.0, recon, kl, klw, iteration)
            assert np.isfinite(reduced) and np.isfinite(float(gn))
            losses.append(reduced)
            iteration += 1
    assert losses[-1] < losses[0]                      # it learns
    # dead parameters never receive a gradient (quirk Q6) and Adam leaves them untouched
    assert model.speaker_embedding.linear_layer.weight.grad is None
    # validate(): eval mode forward (running BN statistics, prenet dropout still on)
    model.eval()
    with torch.no_grad():
        x, y = model.parse_batch(next(iter(loader)))
        val = criterion(model(x), y, iteration)[0].item()
    assert np.isfinite(val)
    # checkpoint round trip with the reference's dict layout (train.py:113-119, 100-110)
    ck = os.path.join(str(tmp_path), "checkpoint_0")
    torch.save({"iteration": iteration, "state_dict": model.state_dict(), "optimizer": optimizer.state_dict(),
                "learning_rate": hp.learning_rate}, ck)
    d = torch.load(ck, map_location="cpu")
    model2 = t2v_model.Tacotron2(hp).cuda()
    model2.load_state_dict(d["state_dict"])
    opt2 = torch.optim.Adam(model2.parameters(), lr=hp.learning_rate, weight_decay=hp.weight_decay)
    opt2.load_state_dict(d["optimizer"])
    model2.eval()
    with torch.no_grad():
        emb = model.transcript_embedding(x[0]).transpose(1, 2)
        a = model.encoder.inference(emb)
        b = model2.encoder.inference(model2.transcript_embedding(x[0]).transpose(1, 2))
    assert torch.equal(a, b)


def test_reference_layout_checkpoint_with_dead_param_adam_state(tmp_path):
    """A checkpoint in the reference's layout (train.py:113-119: iteration / state_dict / optimizer / learning_rate) whose Adam state
    has NO entries for the parameters that never receive a gradient (quirk Q6: speaker / emotion embeddings, the outer CoordConv
    weight) -- exactly what torch.optim.Adam saves for them -- loads into (a) torch.optim.Adam as train.py:100-110 does and (b) the fused
    flat-buffer optimizer, and both continue identically."""
    import model as t2v_model
    from hparams import create_hparams
    from loss_function import Tacotron2Loss_VAE
    from oracle import port
    from t2v import optim
    hp = create_hparams("anneal_function=constant")
    batch = port.synthetic_batch(3, 14, 18, seed=2)

    def run_steps(model, opt, n, fused):
        crit = Tacotron2Loss_VAE(hp)
        model.train()
        for it in range(n):
            model._step = 50 + it                       # same dropout stream in every run
            model.zero_grad()
            x, y = model.parse_batch(batch)
            loss = crit(model(x), y, it)[0]
            loss.backward()
            if fused:
                opt.step()
            else:
                torch.nn.utils.clip_grad_norm_(model.parameters(), hp.grad_clip_thresh)
                opt.step()

    # ---- produce the checkpoint with the reference's own optimizer class
    m0 = t2v_model.Tacotron2(hp)
    m0.load_state_dict(port.init_params(1234))
    m0 = m0.cuda()
    m0._graph_cache = None
    o0 = torch.optim.Adam(m0.parameters(), lr=hp.learning_rate, weight_decay=hp.weight_decay)
    run_steps(m0, o0, 2, False)
    ck = os.path.join(str(tmp_path), "checkpoint_2")
    torch.save({"iteration": 2, "state_dict": m0.state_dict(), "optimizer": o0.state_dict(), "learning_rate": hp.learning_rate}, ck)
    d = torch.load(ck, map_location="cpu")
    names = [k for k, _ in m0.named_parameters()]
    dead = [i for i, k in enumerate(names) if k.startswith(t2v_model.Tacotron2._DEAD)]
    assert dead and all(i not in d["optimizer"]["state"] for i in dead)          # torch skipped them: grad was None
    assert len(d["optimizer"]["state"]) == len(names) - len(dead)
    # ---- (a) torch.optim.Adam resume, (b) fused optimizer resume: one more step each, same parameters afterwards
    res = {}
    for fused in (False, True):
        m = t2v_model.Tacotron2(hp).cuda()
        m._graph_cache = None
        m.load_state_dict(d["state_dict"])
        if fused:
            opt = optim.FusedAdamClip(m, lr=hp.learning_rate, weight_decay=hp.weight_decay, max_norm=hp.grad_clip_thresh)
        else:
            opt = torch.optim.Adam(m.parameters(), lr=hp.learning_rate, weight_decay=hp.weight_decay)
        opt.load_state_dict(copy.deepcopy(d["optimizer"]))       # (torch keeps the CPU `step` tensors and bumps them in place)
        for g in opt.param_groups:
            g["lr"] = d["learning_rate"]
        run_steps(m, opt, 1, fused)
        res[fused] = {k: v.detach().clone() for k, v in m.named_parameters()}
        if fused:
            sd = opt.state_dict()
            assert set(sd["state"]) == set(d["optimizer"]["state"]) and int(sd["state"][dead[-1] + 1]["step"]) == 3
    for k in res[False]:
        a, b = res[True][k], res[False][k]
        assert float((a - b).abs().max()) <= 2e-6 + 1e-4 * float(b.abs().max()), k
    for i in dead:
        assert torch.equal(res[True][names[i]], m0.state_dict()[names[i]])        # untouched by both optimizers
