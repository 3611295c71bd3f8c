"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference
(imported from /root/reference via oracle/ref_shims.py) on seeded inputs.  Run here (build
container) only:  python oracle/gen_golden.py.  The fixtures hold reference OUTPUTS; inputs and
weights are regenerated from seeds by oracle/port.py (init_params / synthetic_batch / Rand.draw),
so nothing of the reference's source travels."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import port, ref_shims  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _model(ref, train):
    hp = ref.hparams.create_hparams("anneal_function=constant")
    m = ref.model.Tacotron2(hp)
    m.load_state_dict(port.init_params(1234))
    m.train(train)
    return m, hp


def train_step(ref, tag, B, Ti, To):
    m, hp = _model(ref, True)
    batch = port.synthetic_batch(B, Ti, To, seed=0)
    rand = port.Rand.draw(B, Ti, To, seed=1)
    x, y = m.parse_batch(batch)
    with ref_shims.InjectedRandomness(rand.as_reference_call_list(), rand.eps):
        out = m(x)
    loss, recon, kl, klw = ref.loss_function.Tacotron2Loss_VAE(hp)(out, y, 0)
    loss.backward()
    d = {n: o.detach().numpy() for n, o in zip(("mel", "mel_post", "gate", "align", "mu", "logvar", "z"), out[:7])}
    d.update(loss=np.float32(loss.item()), recon=np.float32(recon.item()), kl=np.float32(kl.item()), klw=np.float32(klw))
    names, norms, probes = [], [], []
    for k, p in m.named_parameters():
        names.append(k)
        if p.grad is None:
            norms.append(-1.0); probes.append(0.0)
        else:
            norms.append(float(p.grad.norm())); probes.append(float((p.grad * port.grad_probe(k, p.shape)).sum()))
    d.update(grad_names=np.array(names), grad_norms=np.array(norms, np.float64), grad_probes=np.array(probes, np.float64))
    for k in ("postnet.convolutions.0.0.conv.weight", "decoder.gate_layer.linear_layer.weight",
              "decoder.attention_layer.location_layer.location_conv.conv.weight", "vae_gst.fc1.weight",
              "decoder.prenet.layers.0.linear_layer.weight"):
        d["grad::" + k] = dict(m.named_parameters())[k].grad.numpy()
    for k, v in m.state_dict().items():
        if "running" in k and (k.startswith("postnet.convolutions.4") or k.startswith("vae_gst.ref_encoder.bns.5")
                               or k.startswith("encoder.convolutions.0")):
            d["buf::" + k] = v.numpy()
    np.savez_compressed(os.path.join(OUT, "train_step_%s.npz" % tag), **d)
    print(tag, "loss", loss.item(), "kl", kl.item())


def inference_c1(ref):
    """Config 1: B=1, 40 ids, ref-mel 80x200, eval mode, 200 fixed decoder steps driven manually
    the way inference.ipynb / synthesizer.py:139-154 do (prenet dropout masks injected)."""
    m, hp = _model(ref, False)
    g = torch.Generator().manual_seed(7)
    ids = torch.randint(2, 79, (1, 40), generator=g); ids[0, -1] = 1
    refmel = torch.randn(1, 80, 200, generator=g)
    n = 200
    pm = (torch.rand(n, 2, 1, 256, generator=g) >= 0.5).float()
    with torch.no_grad():
        emb = m.transcript_embedding(ids).transpose(1, 2)
        enc = m.encoder.inference(emb)
        style, mu, logvar, z = m.vae_gst(refmel)
        mem = enc + style.unsqueeze(1)
        dec = m.decoder
        x = dec.get_go_frame(mem)
        dec.initialize_decoder_states(mem, mask=None)
        mels, gates, aligns = [], [], []
        for t in range(n):
            with ref_shims.InjectedRandomness([pm[t, 0], pm[t, 1]], None):
                p = dec.prenet(x)
            mel, gate, w = dec.decode(p)
            mels.append(mel); gates.append(gate); aligns.append(w)
            x = mel
        mel_o, gate_o, align_o = dec.parse_decoder_outputs(mels, gates, aligns)
        post = m.postnet(mel_o) + mel_o
    np.savez_compressed(os.path.join(OUT, "inference_c1.npz"), enc=enc.numpy(), style=style.numpy(), mu=mu.numpy(),
                        mel=mel_o.numpy(), gate=gate_o.numpy(), align=align_o.numpy(), mel_post=post.numpy())
    print("c1 mel", mel_o.shape, float(mel_o.abs().mean()))


def stft_fixture(ref):
    st = ref.layers.TacotronSTFT(1024, 256, 1024, 80, 16000, 0.0, 8000.0)
    g = torch.Generator().manual_seed(3)
    wav = torch.rand(2, 16000, generator=g) * 2 - 1
    wav[1] *= torch.linspace(0, 1, 16000)
    mel = st.mel_spectrogram(wav)
    np.savez_compressed(os.path.join(OUT, "stft_mel.npz"), mel=mel.numpy(), mel_basis=st.mel_basis.numpy())
    print("stft", mel.shape)


SPEECH_FILES = ("ref_neu", "recorded_hap", "ref_ang", "ref_hap", "ref_sad", "recorded_ang", "recorded_neu", "recorded_sad")


def stft_speech_fixture(ref):
    """one-second excerpts of the reference's eight recordings (samples/refs/*.wav) through the real
    load_wav_to_torch -> / max_wav_value -> TacotronSTFT.mel_spectrogram path (data_utils.py:42-59)"""
    from scipy.io import wavfile
    st = ref.layers.TacotronSTFT(1024, 256, 1024, 80, 16000, 0.0, 8000.0)
    out = {}
    for name in SPEECH_FILES:
        sr, data = wavfile.read(os.path.join(ref_shims.REFERENCE_ROOT, "samples", "refs", name + ".wav"))
        assert data.dtype == np.int16, data.dtype
        lo = min(len(data) // 3, max(0, len(data) - sr))
        x = data[lo:lo + sr].copy()
        wav = torch.from_numpy(x.astype(np.float32)) / 32768.0
        out[name + "_wav_i16"] = x
        out[name + "_sr"] = np.int32(sr)
        out[name + "_mel"] = st.mel_spectrogram(wav[None])[0].numpy()
        print("speech", name, sr, x.shape, out[name + "_mel"].shape)
    np.savez_compressed(os.path.join(OUT, "stft_speech.npz"), **out)


def text_fixture():
    # README.md:16-24 known-answer vector for "감정있는 한국어 목소리 생성" (korean_cleaners)
    ids = [2, 21, 57, 14, 25, 62, 13, 41, 61, 4, 39, 45, 79, 20, 21, 45, 2, 34, 42, 13, 25, 79, 8, 29, 42, 11, 29, 7, 41,
           79, 11, 22, 62, 11, 25, 62, 1]
    np.savez_compressed(os.path.join(OUT, "text_ids.npz"), ids=np.array(ids, np.int64))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    ref = ref_shims.load_reference()
    torch.set_num_threads(8)
    train_step(ref, "b3", 3, 20, 30)
    train_step(ref, "b4", 4, 40, 64)
    inference_c1(ref)
    stft_fixture(ref)
    stft_speech_fixture(ref)
    text_fixture()
