"""Text -> mel synthesis surface of the reference's synthesizer.py (Synthesizer.load / load_mel / synthesize, synthesizer.py:46-168)
on the B200 engine, up to the vocoder hand-off.

What the reference does per request (synthesizer.py:112-163): text_to_sequence -> transcript_embedding -> encoder.inference ->
style from a reference wav (vae_gst) or from a mix of per-emotion latent centroids (vae_gst.fc3) -> style broadcast-add ->
step-wise prenet / decode loop until the gate fires -> postnet -> `waveglow.infer(mel_outputs, sigma=0.666)` on the PRE-postnet
mel (quirk Q8).  Here the decode loop is the persistent free-running kernel (Tacotron2.inference / Decoder.inference); WaveGlow itself
is out of scope: `vocoder` is any object with `.infer(mel[B,80,T] fp32 contiguous CUDA, sigma=...)`
(waveglow/glow.py:251-292), and `handoff()` checks that contract.
"""
import os

import numpy as np
import torch

from hparams import create_hparams
from layers import TacotronSTFT
from model import Tacotron2
from text import text_to_sequence
from utils import load_wav_to_torch

EMOTIONS = ("neu", "sad", "ang", "hap")          # class ids 0..3 of the filelists (synthesizer.py:107-110)


def handoff(mel_outputs):
    """The tensor WaveGlow.infer consumes (glow.py:251-256: ConvTranspose1d(n_mel, n_mel, 1024, stride=256) on [B, n_mel, T]):
    pre-postnet mel, fp32, contiguous, on the GPU."""
    if mel_outputs.dim() != 3 or mel_outputs.size(1) != 80:
        raise ValueError("vocoder hand-off expects [B, 80, T], got %s" % (tuple(mel_outputs.shape),))
    if not mel_outputs.is_cuda:
        raise ValueError("vocoder hand-off expects a CUDA tensor")
    return mel_outputs.detach().float().contiguous()


class Synthesizer(object):
    def __init__(self, hparams=None):
        self.hparams = hparams or create_hparams()
        self.hparams.sampling_rate = 16000              # synthesizer.py:50-51
        self.hparams.max_decoder_steps = 600
        hp = self.hparams
        self.stft = TacotronSTFT(hp.filter_length, hp.hop_length, hp.win_length, hp.n_mel_channels, hp.sampling_rate,
                                 hp.mel_fmin, hp.mel_fmax)
        self.model = None
        self.vocoder = None
        self.centroids = {}

    # ---- synthesizer.py:59-69
    def load_mel(self, path):
        audio, sampling_rate = load_wav_to_torch(path)
        if sampling_rate != self.hparams.sampling_rate:
            raise ValueError("{} SR doesn't match target {} SR".format(sampling_rate, self.stft.sampling_rate))
        audio_norm = (audio / self.hparams.max_wav_value).unsqueeze(0)
        return self.stft.cuda().mel_spectrogram(audio_norm.cuda())

    # ---- synthesizer.py:75-110
    def load(self, checkpoint_path=None, vocoder=None, state_dict=None, emotion_filelist=None):
        """checkpoint_path: a reference-format checkpoint ({'state_dict': ...}, train.py:113-119); vocoder: object with .infer();
        emotion_filelist: `wav|text|speaker|emotion` lines used to compute (and cache next to the checkpoint) the per-emotion means
        of the latent z."""
        self.model = Tacotron2(self.hparams).cuda()
        if state_dict is None and checkpoint_path is not None:
            state_dict = torch.load(checkpoint_path, map_location="cpu")["state_dict"]
        if state_dict is not None:
            self.model.load_state_dict(state_dict)
        self.model.eval()
        self.vocoder = vocoder
        if emotion_filelist:
            self.compute_centroids(emotion_filelist, cache_next_to=checkpoint_path)
        return self

    @torch.no_grad()
    def compute_centroids(self, filelist, cache_next_to=None):
        npz = None
        if cache_next_to:
            tag = os.path.basename(filelist).rsplit("_", 1)[-1].split(".")[0]
            npz = os.path.join(os.path.dirname(cache_next_to), os.path.basename(cache_next_to) + "_" + tag + ".npz")
        if npz and os.path.exists(npz):
            d = np.load(npz)
            zs, emotions = d["zs"], d["emotions"]
        else:
            zs, emotions = [], []
            with open(filelist, encoding="utf-8") as f:
                for line in f:
                    audio_path, _, _, emotion = line.strip().split("|")
                    _, _, _, z = self.model.vae_gst(self.load_mel(audio_path))
                    zs.append(z.cpu())
                    emotions.append(int(emotion))
            zs, emotions = torch.cat(zs, 0).numpy(), np.array(emotions)
            if npz:
                np.savez(npz, zs=zs, emotions=emotions)
        for i, name in enumerate(EMOTIONS):
            if (emotions == i).any():
                self.centroids[name] = zs[emotions == i].mean(0)
        return self.centroids

    # ---- synthesizer.py:112-168
    @torch.no_grad()
    def synthesize(self, text, path=None, condition_on_ref=False, ref_audio=None, ratios=(1.0, 0.0, 0.0, 0.0), sigma=0.666):
        """-> dict(mel_outputs [1,80,T] (the vocoder hand-off), mel_outputs_postnet, gate_outputs, alignments, audio or None)"""
        if self.model is None:
            raise RuntimeError("call load() first")
        seq = np.array(text_to_sequence(text, ["korean_cleaners"]))[None, :]
        seq = torch.from_numpy(seq).cuda().long()
        if condition_on_ref:
            mel, mel_post, gate, align = self.model.inference(seq, ref_mel=self.load_mel(ref_audio))
        else:
            order = ("neu", "sad", "hap", "ang")       # the reference mixes ratios in this order (synthesizer.py:128-129)
            z = sum(float(r) * torch.from_numpy(np.asarray(self.centroids[n], dtype=np.float32)) for r, n in zip(ratios, order))
            mel, mel_post, gate, align = self.model.inference(seq, z=z.view(1, -1).cuda())
        out = dict(mel_outputs=handoff(mel), mel_outputs_postnet=mel_post, gate_outputs=gate, alignments=align, audio=None)
        if self.vocoder is not None:
            audio = self.vocoder.infer(out["mel_outputs"], sigma=sigma)
            out["audio"] = audio
            if path:
                from scipy.io.wavfile import write
                write(path, self.hparams.sampling_rate, audio[0].detach().float().cpu().numpy())
        return out
