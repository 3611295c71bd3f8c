"""Dataset + collate with the reference's contract (reference data_utils.py:11-137).

`TextMelLoader` yields (text ids IntTensor, mel [80,T], speaker one-hot, emotion one-hot); the mel comes from the
t2v STFT kernels when a GPU is present (the reference computes it on the CPU inside the DataLoader worker,
data_utils.py:42-52).  `TextMelCollate` produces exactly the 7-tuple `Tacotron2.parse_batch` consumes: sorted by text
length (descending), zero-padded, gate target 1 from the last frame on."""
import random

import numpy as np
import torch
import torch.utils.data

import layers
from text import text_to_sequence
from utils import load_filepaths_and_text, load_wav_to_torch


class TextMelLoader(torch.utils.data.Dataset):
    """defer_mel=True: __getitem__ returns the normalised waveform instead of the mel (pure CPU work, safe in DataLoader worker
    processes); TextMelCollate(..., stft=loader.stft) then computes the mels of the whole batch in one fused GPU launch."""

    def __init__(self, audiopaths_and_text, hparams, defer_mel=False):
        self.defer_mel = defer_mel
        self.audiopaths_and_text = load_filepaths_and_text(audiopaths_and_text)
        self.text_cleaners = hparams.text_cleaners
        self.max_wav_value = hparams.max_wav_value
        self.sampling_rate = hparams.sampling_rate
        self.load_mel_from_disk = hparams.load_mel_from_disk
        self.n_speakers, self.n_emotions = hparams.n_speakers, hparams.n_emotions
        self.stft = layers.TacotronSTFT(hparams.filter_length, hparams.hop_length, hparams.win_length,
                                        hparams.n_mel_channels, hparams.sampling_rate, hparams.mel_fmin, hparams.mel_fmax)
        if torch.cuda.is_available():
            self.stft = self.stft.cuda()
        random.seed(1234)
        random.shuffle(self.audiopaths_and_text)

    def get_mel(self, filename):
        if self.load_mel_from_disk:
            mel = torch.from_numpy(np.load(filename))
            assert mel.size(0) == self.stft.n_mel_channels, "Mel dimension mismatch"
            return mel
        audio, sr = load_wav_to_torch(filename)
        if sr != self.stft.sampling_rate:
            raise ValueError("{} SR doesn't match target {} SR".format(sr, self.stft.sampling_rate))
        if self.defer_mel:
            return audio / self.max_wav_value                  # 1-D waveform; the collate function batches the STFT
        wav = (audio / self.max_wav_value).unsqueeze(0)
        dev = self.stft.mel_basis.device
        return self.stft.mel_spectrogram(wav.to(dev)).squeeze(0).cpu()

    def get_text(self, text):
        return torch.IntTensor(text_to_sequence(text, self.text_cleaners))

    @staticmethod
    def _one_hot(index, n):
        v = torch.zeros(n)
        v[int(index)] = 1
        return v

    def get_mel_text_pair(self, item):
        path, text, speaker, emotion = item[0], item[1], item[2], item[3]
        return (self.get_text(text), self.get_mel(path), self._one_hot(speaker, self.n_speakers),
                self._one_hot(emotion, self.n_emotions))

    def __getitem__(self, index):
        return self.get_mel_text_pair(self.audiopaths_and_text[index])

    def __len__(self):
        return len(self.audiopaths_and_text)


def batch_mel_spectrogram(stft, wavs):
    """Batched wav -> mel on the GPU: `wavs` = list of 1-D float waveforms in [-1, 1] of different lengths.  They are
    zero-padded into one [B, S_max] tensor and go through ONE fused STFT/mel launch; every utterance then keeps its own
    len // hop + 1 frames.  Frames that touch the padding differ from the per-utterance reflect padding, so the last
    ceil(n_fft / 2 / hop) = 2 frames of shorter utterances are recomputed from a per-utterance tail call."""
    dev = stft.mel_basis.device
    lens = [int(w.numel()) for w in wavs]
    S = max(lens)
    batch = torch.zeros(len(wavs), S, device=dev)
    for i, w in enumerate(wavs):
        batch[i, :lens[i]] = w.to(dev, non_blocking=True)
    mel = stft.mel_spectrogram(batch)
    hop, n_fft = stft.hop_length, stft.filter_length
    out = []
    for i, n in enumerate(lens):
        nF = n // hop + 1
        m = mel[i, :, :nF].clone()
        if n < S:       # the frames whose window crosses the end of this utterance need ITS reflection, not the batch padding
            k = min(nF, (n_fft // 2 + hop - 1) // hop + 1)
            tail0 = max(0, (nF - k) * hop - n_fft // 2)                 # first sample the last k frames can see
            if tail0 == 0:
                m = stft.mel_spectrogram(batch[i:i + 1, :n])[0]
            else:
                # re-run the tail with enough left context that its first frames are interior frames; keep only the last k
                ctx = (tail0 // hop) * hop
                tail = stft.mel_spectrogram(batch[i:i + 1, ctx:n])[0]
                m[:, nF - k:] = tail[:, tail.shape[1] - k:]
        out.append(m)
    return out


class TextMelCollate(object):
    def __init__(self, n_frames_per_step, stft=None):
        self.n_frames_per_step = n_frames_per_step
        self.stft = stft             # with TextMelLoader(defer_mel=True): batch items carry waveforms, mels are made here

    def __call__(self, batch):
        if self.stft is not None and batch[0][1].dim() == 1:
            mels = batch_mel_spectrogram(self.stft, [x[1] for x in batch])
            batch = [(x[0], m.cpu(), x[2], x[3]) for x, m in zip(batch, mels)]
        in_len, order = torch.sort(torch.LongTensor([len(x[0]) for x in batch]), dim=0, descending=True)
        B = len(batch)
        text = torch.zeros(B, int(in_len[0]), dtype=torch.long)
        speakers = torch.zeros(B, len(batch[0][2]), dtype=torch.long)
        emotions = torch.zeros(B, len(batch[0][3]), dtype=torch.long)
        n_mel = batch[0][1].size(0)
        To = max(x[1].size(1) for x in batch)
        if To % self.n_frames_per_step:
            To += self.n_frames_per_step - To % self.n_frames_per_step
        mel = torch.zeros(B, n_mel, To)
        gate = torch.zeros(B, To)
        out_len = torch.zeros(B, dtype=torch.long)
        for i, j in enumerate(order.tolist()):
            t, m, s, e = batch[j]
            text[i, :t.size(0)] = t
            speakers[i] = s
            emotions[i] = e
            mel[i, :, :m.size(1)] = m
            gate[i, m.size(1) - 1:] = 1
            out_len[i] = m.size(1)
        return text, in_len, mel, gate, out_len, speakers, emotions
