"""STFT / mel front-end (reference layers.py:75-92, stft.py:77-105) and the vocoder hand-off surface (synthesizer.py:112-168)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _stft():
    from layers import TacotronSTFT
    return TacotronSTFT(1024, 256, 1024, 80, 16000, 0.0, 8000.0).cuda()


def test_fused_stft_mel_matches_golden_and_gemm_path(golden_dir):
    """the fused FFT kernel against (1) the golden made by the real reference (conv1d with the Fourier basis) and (2) the exact-fp32
    GEMM path of this repo, on the golden's waveforms and on white noise U(-1,1) of 160 000 samples x 4 (SURVEY 8d STFT input)"""
    from t2v import frontend
    G = np.load(os.path.join(golden_dir, "stft_mel.npz"))
    st = _stft()
    g = torch.Generator().manual_seed(3)
    wav = torch.rand(2, 16000, generator=g) * 2 - 1
    wav[1] *= torch.linspace(0, 1, 16000)
    mel = st.mel_spectrogram(wav.cuda())
    assert tuple(mel.shape) == tuple(G["mel"].shape)
    assert torch.allclose(mel.cpu(), torch.from_numpy(G["mel"]), rtol=1e-3, atol=1e-3)
    noise = (torch.rand(4, 160000, generator=g) * 2 - 1).cuda()
    a = st.mel_spectrogram(noise)
    b = frontend.mel_spectrogram_gemm(noise, st.stft_fn, st.mel_basis)
    assert tuple(a.shape) == (4, 80, 160000 // 256 + 1)
    err = float((a - b).abs().max())
    print("fused FFT vs fp32 DFT-GEMM path, max-abs on log-mel: %.2e" % err)
    assert err <= 1e-3
    # odd lengths / a single frame pair boundary
    for S in (513, 1000, 4097, 25601):
        w = (torch.rand(1, S, generator=g) * 2 - 1).cuda()
        assert float((st.mel_spectrogram(w) - frontend.mel_spectrogram_gemm(w, st.stft_fn, st.mel_basis)).abs().max()) <= 1e-3, S


def test_fused_stft_mel_matches_golden_speech(golden_dir):
    """the fused FFT kernel on real speech: one-second excerpts of two of the reference's recordings (samples/refs/*.wav) against
    the mel the real reference computed from them (oracle/gen_golden.py::stft_speech_fixture); quiet frames exercise the clip.
    (The fixture holds all eight recordings; the CPU oracle is pinned on all of them in tests/test_oracle_cpu.py.)"""
    G = np.load(os.path.join(golden_dir, "stft_speech.npz"))
    st = _stft()
    for name in ("ref_neu", "recorded_hap"):
        wav = torch.from_numpy(G[name + "_wav_i16"].astype(np.float32)) / 32768.0
        mel = st.mel_spectrogram(wav[None].cuda())[0].cpu()
        ref = torch.from_numpy(G[name + "_mel"])
        assert tuple(mel.shape) == tuple(ref.shape)
        # speech spans 80 dB across mel bands: both implementations are fp32 transforms whose error scales with the LOUDEST bin of a
        # frame, so the quiet bands are compared in the linear domain (relative to the frame maximum) and the log values where the
        # band is within 40 dB of it
        lin, rlin = mel.exp(), ref.exp()
        fmax = rlin.max(dim=0, keepdim=True).values
        err_lin = float(((lin - rlin).abs() / fmax).max())
        loud = rlin >= 1e-2 * fmax
        err_log = float((mel - ref).abs()[loud].max())
        print("speech %s: linear mel error / frame max %.2e, log-mel max-abs on bands within 40 dB of the frame max %.2e" % (name, err_lin, err_log))
        assert err_lin <= 5e-5 and err_log <= 2e-3, (name, err_lin, err_log)


def test_batched_mel_equals_per_utterance_mel():
    """data_utils.batch_mel_spectrogram (one fused launch for a ragged batch + tail fix-up) == mel_spectrogram of every utterance alone"""
    from data_utils import batch_mel_spectrogram
    st = _stft()
    g = torch.Generator().manual_seed(5)
    wavs = [torch.rand(n, generator=g) * 2 - 1 for n in (16000, 9000, 12345, 700, 16000, 2048)]
    got = batch_mel_spectrogram(st, wavs)
    for w, m in zip(wavs, got):
        ref = st.mel_spectrogram(w.view(1, -1).cuda())[0]
        assert tuple(m.shape) == tuple(ref.shape)
        assert float((m - ref).abs().max()) <= 1e-5, w.numel()


def test_synthesizer_handoff_contract():
    """Synthesizer.synthesize (emotion-ratio path) returns the PRE-postnet mel as fp32 contiguous [1, 80, T] on the GPU -- the tensor
    waveglow.infer consumes (glow.py:251-256: ConvTranspose1d(80, 80, 1024, stride 256) over [B, 80, T]); quirk Q8."""
    from synthesizer import Synthesizer, handoff
    from oracle import port

    class FakeVocoder(object):                   # the first op of WaveGlow.infer: shape / dtype / device contract only
        def __init__(self):
            self.upsample = torch.nn.ConvTranspose1d(80, 80, 1024, stride=256).cuda()
            self.seen = None

        def infer(self, spect, sigma=1.0):
            assert spect.dtype == torch.float32 and spect.is_contiguous() and spect.is_cuda
            self.seen = tuple(spect.shape)
            up = self.upsample(spect)
            return up[:, 0, :-(1024 - 256)]

    syn = Synthesizer()
    syn.hparams.max_decoder_steps = 40
    voc = FakeVocoder()
    syn.load(state_dict=port.init_params(1234), vocoder=voc)
    g = torch.Generator().manual_seed(0)
    syn.centroids = {n: torch.randn(32, generator=g).numpy() for n in ("neu", "sad", "ang", "hap")}
    out = syn.synthesize("감정있는 한국어 목소리 생성", condition_on_ref=False, ratios=(0.5, 0.2, 0.2, 0.1))
    mel = out["mel_outputs"]
    assert mel.dim() == 3 and mel.shape[0] == 1 and mel.shape[1] == 80 and 1 <= mel.shape[2] <= 40
    assert mel.dtype == torch.float32 and mel.is_contiguous() and mel.is_cuda
    assert voc.seen == tuple(mel.shape)
    assert out["audio"].shape[1] == mel.shape[2] * 256
    assert tuple(out["mel_outputs_postnet"].shape) == tuple(mel.shape)
    assert out["gate_outputs"].shape[-1] == 1                     # quirk Q7: [B, N, 1] at inference
    with pytest.raises(ValueError):
        handoff(mel.cpu())
