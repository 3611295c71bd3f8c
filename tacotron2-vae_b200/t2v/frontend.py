"""STFT + mel front-end on the t2v kernels (reference stft.py:77-105, layers.py:75-92, audio_processing.py:77-83).

The windowed DFT is a GEMM over overlapping hop-strided rows of the reflect-padded waveform (no framing copy); the
magnitude, mel projection and log-compression are small kernels around a second GEMM."""
import math

import numpy as np
import torch

from . import engine
from ._lib import call as L


def slaney_mel_filterbank(sr, n_fft, n_mels=80, fmin=0.0, fmax=None):
    """The filterbank librosa 0.6.0 `filters.mel` builds by default (Slaney mel scale, area-normalised triangles),
    which is what reference layers.py:62-63 asks for."""
    fmax = sr / 2.0 if fmax is None else fmax
    lin_step, knee_hz = 200.0 / 3.0, 1000.0
    knee_mel, log_step = knee_hz / lin_step, math.log(6.4) / 27.0
    to_mel = lambda f: knee_mel + math.log(f / knee_hz) / log_step if f >= knee_hz else f / lin_step
    mels = np.linspace(to_mel(fmin), to_mel(fmax), n_mels + 2)
    edges = np.where(mels >= knee_mel, knee_hz * np.exp(log_step * (mels - knee_mel)), lin_step * mels)
    freqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    fb = np.zeros((n_mels, freqs.size), dtype=np.float64)
    for m in range(n_mels):
        lo, mid, hi = edges[m], edges[m + 1], edges[m + 2]
        rise = (freqs - lo) / (mid - lo)
        fall = (hi - freqs) / (hi - mid)
        fb[m] = np.clip(np.minimum(rise, fall), 0.0, None) * (2.0 / (hi - lo))
    return fb.astype(np.float32)


class STFT(torch.nn.Module):
    """Holds the hann-windowed Fourier basis [2*(n_fft/2+1), n_fft] (reference stft.py:53-75)."""

    def __init__(self, filter_length=800, hop_length=200, win_length=800, window="hann"):
        super().__init__()
        assert window == "hann" and win_length == filter_length
        self.filter_length, self.hop_length, self.win_length, self.window = filter_length, hop_length, win_length, window
        n, nb = filter_length, filter_length // 2 + 1
        ang = 2.0 * np.pi * np.outer(np.arange(nb), np.arange(n)) / n
        win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)          # scipy get_window('hann', fftbins=True)
        basis = np.concatenate([np.cos(ang) * win, -np.sin(ang) * win], 0).astype(np.float32)
        self.register_buffer("forward_basis", torch.from_numpy(basis).unsqueeze(1))     # [2nb,1,n] like the reference

    def transform(self, wav):
        """-> (magnitude [B,nb,frames], None).  Phase is only needed by Griffin-Lim (out of scope)."""
        mag_rows, rows_pb, nF = _stft_magnitude(wav, self)
        B = wav.shape[0]
        nb = self.filter_length // 2 + 1
        out = mag_rows.view(B, rows_pb, -1)[:, :nF, :nb].transpose(1, 2).contiguous()
        return out, None


def _stft_magnitude(wav, stft):
    if not wav.is_cuda:
        raise RuntimeError("the t2v STFT runs on CUDA tensors only")
    wav = wav.contiguous().float()
    dev = wav.device
    B, S = wav.shape
    n, hop = stft.filter_length, stft.hop_length
    nb = n // 2 + 1
    pad = n // 2
    nF = S // hop + 1
    rows_pb = (S + 2 * pad + hop - 1) // hop + (n + hop - 1) // hop
    Lp = rows_pb * hop
    WP = torch.empty(B, Lp, device=dev)
    L("t2v_reflect_pad", wav, WP, B, S, pad, Lp)
    Rtot = B * rows_pb
    M = Rtot - (n + hop - 1) // hop
    ld_ft = (2 * nb + 3) // 4 * 4
    FT = torch.empty(Rtot, ld_ft, device=dev)
    basis = stft.forward_basis.view(2 * nb, n)
    # frames are overlapping rows (stride hop) of the padded waveform: exact fp32 GEMM (|X| of quiet bins cancels)
    engine.Ops.gemm(WP, hop, 1, basis, n, 1, FT, ld_ft, M, 2 * nb, n)
    ld_mag = (nb + 3) // 4 * 4
    MAG = torch.zeros(Rtot, ld_mag, device=dev)
    L("t2v_stft_mag", FT, ld_ft, MAG, ld_mag, M, nb)
    return MAG, rows_pb, nF


def _fused_tables(stft, mel_basis):
    """constant tables of the fused kernel (hann window, FFT twiddles, non-zero band of every mel filter), cached on the STFT object"""
    dev = mel_basis.device
    key = (str(dev), mel_basis.data_ptr())
    tabs = getattr(stft, "_fused", None)
    if tabs is None or tabs[0] != key:
        n = stft.filter_length
        win = 0.5 - 0.5 * np.cos(2.0 * np.pi * np.arange(n) / n)
        ang = -2.0 * np.pi * np.arange(n) / n
        tw = np.stack([np.cos(ang), np.sin(ang)], 1)
        nz = (mel_basis.detach().cpu().numpy() != 0)
        lo = np.where(nz.any(1), nz.argmax(1), 0)
        hi = np.where(nz.any(1), nz.shape[1] - nz[:, ::-1].argmax(1), 0)
        tabs = (key, torch.from_numpy(win.astype(np.float32)).to(dev), torch.from_numpy(tw.astype(np.float32)).contiguous().to(dev),
                torch.from_numpy(lo.astype(np.int32)).to(dev), torch.from_numpy(hi.astype(np.int32)).to(dev))
        stft._fused = tabs
    return tabs[1:]


def mel_spectrogram(wav, stft, mel_basis):
    """wav [B,S] -> log-mel [B,n_mel,S//hop+1].  The reference recipe (n_fft 1024, hop 256) runs as ONE fused kernel
    (stft_fused.cu: FFT in shared memory, no intermediate in HBM); other geometries take the GEMM path below."""
    if not wav.is_cuda:
        raise RuntimeError("the t2v STFT runs on CUDA tensors only")
    if stft.filter_length == 1024 and stft.hop_length == 256 and mel_basis.shape[1] == 513 and mel_basis.shape[0] <= 128 \
            and wav.shape[1] > 512:
        wav = wav.contiguous().float()
        B, S = wav.shape
        nF = S // 256 + 1
        win, tw, lo, hi = _fused_tables(stft, mel_basis)
        out = torch.empty(B, mel_basis.shape[0], nF, device=wav.device)
        L("t2v_stft_mel_fused", wav, B, S, win, tw, mel_basis.contiguous(), lo, hi, out, mel_basis.shape[0], nF, 1e-5)
        return out
    return mel_spectrogram_gemm(wav, stft, mel_basis)


def mel_spectrogram_gemm(wav, stft, mel_basis):
    MAG, rows_pb, nF = _stft_magnitude(wav, stft)
    dev = wav.device
    B = wav.shape[0]
    n_mel, nb = mel_basis.shape
    Rtot = B * rows_pb
    MEL = torch.empty(Rtot, n_mel, device=dev)
    engine.Ops.gemm(MAG, MAG.shape[1], 1, mel_basis.contiguous(), nb, 1, MEL, n_mel, Rtot, n_mel, nb)
    out = torch.empty(B, n_mel, nF, device=dev)
    L("t2v_mel_log", MEL, n_mel, out, B, n_mel, nF, rows_pb, 1e-5)
    return out
