"""Decoder glue around the sampler: the host-side mirror of reference
src/decoder/voicebox/util/model_util.py — `process_unit` (:50-54), `get_mel` (:24-38), and the mel-producing part of
`reconstruct_speech` (:72-103; the vocoder call at :104 goes to usdm_b200.vocoder).  The arithmetic runs in libtvb.so
(usdm_b200/csrc/glue.cuh); this module only computes constants (the Slaney mel filterbank) and moves pointers.
"""
from __future__ import annotations

import ctypes as C
import math
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import engine as E

mel_mean = -5.5419          # model_util.py:18
mel_std = 2.1575            # model_util.py:19

# BigVGAN 22 kHz / 80 band config the reference reads from `vocoder.h` (model_util.py:73,81)
DEFAULT_HPS = SimpleNamespace(sampling_rate=22050, n_fft=1024, hop_size=256, win_size=1024, num_mels=80, fmin=0, fmax=8000)


def _hz_to_mel(f):
    """Slaney mel scale (librosa.hz_to_mel, htk=False): linear below 1 kHz, log above."""
    f = np.asarray(f, dtype=np.float64)
    f_sp = 200.0 / 3
    mels = f / f_sp
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-10) / min_log_hz) / logstep, mels)


def _mel_to_hz(m):
    m = np.asarray(m, dtype=np.float64)
    f_sp = 200.0 / 3
    min_log_hz, min_log_mel, logstep = 1000.0, 1000.0 / f_sp, math.log(6.4) / 27.0
    return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)


def mel_filterbank(sr: int, n_fft: int, n_mels: int, fmin: float, fmax: float) -> np.ndarray:
    """`librosa.filters.mel(sr, n_fft, n_mels, fmin, fmax)` with its defaults (Slaney scale, Slaney area normalisation), which is
    what meldataset.py:63 calls: triangular filters on the rFFT bin frequencies, float32 [n_mels, n_fft // 2 + 1]."""
    fftfreqs = np.linspace(0.0, sr / 2.0, n_fft // 2 + 1)
    mel_f = _mel_to_hz(np.linspace(_hz_to_mel(fmin), _hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fftfreqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    w = np.maximum(0.0, np.minimum(lower, upper))
    w *= (2.0 / (mel_f[2:n_mels + 2] - mel_f[:n_mels]))[:, None]
    return w.astype(np.float32)


_basis_cache = {}


def _basis(hps, device):
    key = (hps.sampling_rate, hps.n_fft, hps.num_mels, hps.fmin, hps.fmax, str(device))
    if key not in _basis_cache:
        _basis_cache[key] = torch.from_numpy(mel_filterbank(hps.sampling_rate, hps.n_fft, hps.num_mels, hps.fmin, hps.fmax)).to(device).contiguous()
    return _basis_cache[key]


def _dev_index(device) -> int:
    device = torch.device(device)
    if device.type != "cuda":
        raise RuntimeError("usdm_b200.glue runs on a CUDA (sm_100a) device only; there is no CPU path")
    return device.index if device.index is not None else torch.cuda.current_device()


def _stream(idx):
    return C.c_void_p(torch.cuda.current_stream(idx).cuda_stream)


def process_unit(unit: torch.Tensor, hps=DEFAULT_HPS, device="cuda"):
    """model_util.py:50-54: 50 Hz units -> one unit per mel frame; returns (LongTensor [1, T] on device, new_length in samples)."""
    idx = _dev_index(device)
    lib = E.load_library()
    u = torch.as_tensor(unit).reshape(-1).to(device=f"cuda:{idx}", dtype=torch.int64).contiguous()
    rep = hps.sampling_rate // 50
    n_out = u.numel() * rep // hps.hop_size
    out = torch.empty(n_out, dtype=torch.int64, device=u.device)
    E.check(lib.tvb_process_unit(idx, C.c_void_p(u.data_ptr()), u.numel(), rep, hps.hop_size, C.c_void_p(out.data_ptr()), n_out, _stream(idx)))
    return out.unsqueeze(0), n_out * hps.hop_size


def mel_spectrogram(audio: torch.Tensor, hps=DEFAULT_HPS, device="cuda", normalize: bool = True) -> torch.Tensor:
    """meldataset.py:55-78 on a mono float waveform in [-1, 1] ([n] or [1, n]); returns [1, num_mels, frames].
    normalize=True additionally applies (mel - mel_mean) / mel_std (model_util.py:82)."""
    if hps.win_size != hps.n_fft:
        raise ValueError("win_size must equal n_fft (as in the reference's vocoder config)")
    idx = _dev_index(device)
    lib = E.load_library()
    y = torch.as_tensor(audio).reshape(-1).to(device=f"cuda:{idx}", dtype=torch.float32).contiguous()
    n = y.numel()
    n_frames = (n + (hps.n_fft - hps.hop_size) - hps.n_fft) // hps.hop_size + 1
    out = torch.empty((hps.num_mels, n_frames), dtype=torch.float32, device=y.device)
    basis = _basis(hps, y.device)
    mean, std = (mel_mean, mel_std) if normalize else (0.0, 1.0)
    E.check(lib.tvb_mel_spectrogram(idx, C.c_void_p(y.data_ptr()), n, hps.n_fft, hps.hop_size, hps.num_mels, C.c_void_p(basis.data_ptr()),
                                    mean, std, C.c_void_p(out.data_ptr()), n_frames, _stream(idx)))
    return out.unsqueeze(0)


def load_wav(path: str):
    """Stand-in for train_util.load_wav_to_torch (:36-38; torchaudio.load is unusable without torchcodec).  torchaudio.load
    returns float32 in [-1, 1] whatever the file's sample format (normalize=True), so PCM data is scaled here by its own full
    scale: int16 / 2**15, int32 (and 24-bit, which scipy widens to int32 left-justified) / 2**31, uint8 (offset binary)
    (x - 128) / 2**7; float files pass through.  Returns (float32 tensor [n] of the first channel, sample rate)."""
    from scipy.io import wavfile
    sr, data = wavfile.read(path)
    if data.ndim > 1:
        data = data[:, 0]
    if data.dtype == np.int16:
        data = data.astype(np.float32) / 32768.0
    elif data.dtype == np.int32:
        data = (data.astype(np.float64) / 2147483648.0).astype(np.float32)
    elif data.dtype == np.uint8:
        data = (data.astype(np.float32) - 128.0) / 128.0
    elif data.dtype.kind == "f":
        data = data.astype(np.float32)
    else:
        raise ValueError(f"{path}: unsupported wav sample format {data.dtype}")
    return torch.from_numpy(np.ascontiguousarray(data)), sr


def get_mel(filepath_or_audio, length: Optional[int] = None, hps=DEFAULT_HPS, device="cuda", normalize: bool = False) -> torch.Tensor:
    """model_util.py:24-38 (resampling is not provided: the file must already be at hps.sampling_rate).  A file always takes the
    reference's float branch (`audio.clamp(-1, 1)`, :33): its loader hands out floats.  The dither branch of :33
    (`(audio + rand) / 32768`) is only reachable with an integer tensor, which only a caller passing raw int16 PCM produces."""
    if isinstance(filepath_or_audio, str):
        audio, sr = load_wav(filepath_or_audio)
        if sr != hps.sampling_rate:
            raise ValueError(f"{filepath_or_audio}: {sr} Hz, expected {hps.sampling_rate} Hz")
    else:
        audio = torch.as_tensor(filepath_or_audio)
    if length:
        audio = audio[:length]
    if audio.dtype.is_floating_point:
        audio = audio.float().clamp(-1, 1)
    else:
        audio = (audio.float() + torch.rand(audio.shape)) / 32768.0        # model_util.py:33 (dither for raw integer PCM)
    return mel_spectrogram(audio, hps, device, normalize=normalize)


def _hps_of(vocoder, hps):
    """The reference reads every audio parameter from `vocoder.h` (model_util.py:73-81).  BigVGAN's own config carries
    sampling_rate / hop_size / num_mels (and n_fft / win_size / fmin / fmax when it was loaded from the published config.json);
    whatever it does not name falls back to `hps`."""
    h = getattr(vocoder, "h", None)
    if h is None:
        return hps
    get = h.get if hasattr(h, "get") else (lambda k, d=None: getattr(h, k, d))
    pick = lambda k: getattr(hps, k) if get(k, None) is None else get(k)
    return SimpleNamespace(**{k: pick(k) for k in ("sampling_rate", "n_fft", "hop_size", "win_size", "num_mels", "fmin", "fmax")})


@torch.inference_mode()
def reconstruct_mel(agent_unit, voicebox, reference_audio=None, reference_unit=None, n_timesteps: int = 50, hps=DEFAULT_HPS,
                    device="cuda", solver: str = "heun", gradient_scale: float = 1.0) -> torch.Tensor:
    """The mel-producing part of reconstruct_speech (model_util.py:72-103).  `reference_unit` are the 50 Hz units of the
    reference audio (the reference obtains them from its XLS-R unit extractor, :78-80, which is outside this path).
    Returns the de-normalised mel [1, num_mels, T_agent] that the reference hands to its vocoder (:103-104)."""
    agent, _ = process_unit(agent_unit, hps, device)
    if reference_audio is not None:
        ref_unit, new_length = process_unit(reference_unit, hps, device)
        ref_mel = get_mel(reference_audio, new_length, hps, device, normalize=True)
        P, T = ref_unit.shape[-1], ref_unit.shape[-1] + agent.shape[-1]
        dummy_y = torch.zeros(1, hps.num_mels, T, device=agent.device)
        dummy_y[:, :, :P] = ref_mel[:, :, :P]
        unit = torch.cat([ref_unit, agent], dim=-1)
        y = voicebox.generate(unit, dummy_y, torch.tensor([T]), n_timesteps=n_timesteps, solver=solver, gradient_scale=gradient_scale,
                              speech_prompt=True, prompt_lengths=torch.tensor([P]))
        y = y[:, :, P:]
    else:
        T = agent.shape[-1]
        dummy_y = torch.zeros(1, hps.num_mels, T, device=agent.device)
        y = voicebox.generate(agent, dummy_y, torch.tensor([T]), n_timesteps=n_timesteps, solver=solver, gradient_scale=gradient_scale,
                              speech_prompt=False)
    return y * mel_std + mel_mean


@torch.inference_mode()
def reconstruct_speech(agent_unit, voicebox, vocoder, reference_audio=None, reference_unit=None, n_timesteps: int = 50, hps=DEFAULT_HPS,
                       device="cuda", solver: str = "heun", gradient_scale: float = 1.0):
    """reconstruct_speech (model_util.py:72-105) end to end on the B200: units -> mel (usdm_b200.Voicebox) -> waveform
    (usdm_b200.vocoder.BigVGAN), returned like the reference as a clamped float32 numpy array [T_agent * hop].  Like the
    reference (`process_unit(agent_unit, vocoder.h, device)`, :73) the audio parameters come from `vocoder.h`."""
    hps = _hps_of(vocoder, hps)
    y_dec = reconstruct_mel(agent_unit, voicebox, reference_audio, reference_unit, n_timesteps, hps, device, solver, gradient_scale)
    return vocoder.forward(y_dec).cpu().squeeze().clamp(-1, 1).numpy()          # model_util.py:104

