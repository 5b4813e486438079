"""CPU ORACLE for the decoder glue.  TEST INFRASTRUCTURE ONLY (imported by tests/ only).

Restates reference src/decoder/voicebox/util/model_util.py:24-38,50-54 and vocoder/meldataset.py:55-78 with the SAME torch
calls the reference makes (repeat_interleave + mode; reflect pad + torch.stft + filterbank matmul + log).  The reference's
own functions cannot be imported here (librosa / matplotlib are not installed), so:
  * the STFT / log / mode part is pinned by construction (identical torch ops);
  * the mel filterbank (librosa.filters.mel defaults: Slaney scale + Slaney normalisation) is restated below in float64 numpy
    and cross-checked in tests/test_glue_oracle.py against torchaudio.functional.melscale_fbanks(norm="slaney",
    mel_scale="slaney"), an independent implementation of the same published formula.  Parity with librosa itself is unpinned.
"""
from __future__ import annotations

import math

import numpy as np
import torch

mel_mean, mel_std = -5.5419, 2.1575                       # model_util.py:18-19


def process_unit(unit: torch.Tensor, sampling_rate: int = 22050, hop_size: int = 256):
    """model_util.py:50-54 verbatim semantics."""
    unit = torch.repeat_interleave(unit, sampling_rate // 50)
    new_length = len(unit) // hop_size * hop_size
    unit = unit[:new_length].reshape(-1, hop_size).mode(1)[0]
    return unit.unsqueeze(0), new_length


def slaney_mel_filterbank(sr=22050, n_fft=1024, n_mels=80, fmin=0.0, fmax=8000.0) -> np.ndarray:
    """librosa.filters.mel(sr=, n_fft=, n_mels=, fmin=, fmax=) with default htk=False, norm='slaney' (meldataset.py:63)."""
    f_sp, min_log_hz = 200.0 / 3, 1000.0
    min_log_mel, logstep = min_log_hz / f_sp, math.log(6.4) / 27.0

    def hz_to_mel(f):
        return min_log_mel + math.log(f / min_log_hz) / logstep if f >= min_log_hz else f / f_sp

    def mel_to_hz(m):
        return min_log_hz * math.exp(logstep * (m - min_log_mel)) if m >= min_log_mel else f_sp * m

    mels = np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2)
    mel_f = np.array([mel_to_hz(m) for m in mels])
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    w = np.zeros((n_mels, n_fft // 2 + 1))
    for i in range(n_mels):
        lower = (fftfreqs - mel_f[i]) / (mel_f[i + 1] - mel_f[i])
        upper = (mel_f[i + 2] - fftfreqs) / (mel_f[i + 2] - mel_f[i + 1])
        w[i] = np.maximum(0.0, np.minimum(lower, upper)) * (2.0 / (mel_f[i + 2] - mel_f[i]))
    return w.astype(np.float32)


def mel_spectrogram(y: torch.Tensor, n_fft=1024, num_mels=80, sampling_rate=22050, hop_size=256, win_size=1024, fmin=0, fmax=8000):
    """meldataset.py:55-78; y fp32 [1, n] in [-1, 1] -> [1, num_mels, frames] (log-mel, not mean/std-normalised)."""
    basis = torch.from_numpy(slaney_mel_filterbank(sampling_rate, n_fft, num_mels, fmin, fmax))
    window = torch.hann_window(win_size)
    pad = int((n_fft - hop_size) / 2)
    y = torch.nn.functional.pad(y.unsqueeze(1), (pad, pad), mode="reflect").squeeze(1)
    spec = torch.stft(y, n_fft, hop_length=hop_size, win_length=win_size, window=window, center=False, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True)
    spec = torch.sqrt(torch.real(spec * spec.conj() + 1e-9))
    spec = torch.matmul(basis, spec)
    return torch.log(torch.clamp(spec, min=1e-5))


def get_mel(audio: torch.Tensor, length=None, **kw):
    """model_util.py:24-38 for float audio already at the target rate."""
    if length:
        audio = audio[:length]
    return mel_spectrogram(audio.float().clamp(-1, 1).unsqueeze(0), **kw)
