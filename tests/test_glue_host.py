"""CPU: host-side behaviour of the decoder glue mirror (usdm_b200/glue.py, usdm_b200/vocoder.py) that needs no device."""
import numpy as np
import pytest
import torch
from scipy.io import wavfile

from util import ROOT  # noqa: F401  (puts the repo root on sys.path)

from usdm_b200 import glue
from usdm_b200.vocoder import BIGVGAN_22K_80, AttrDict


@pytest.mark.parametrize("dtype,scale,offset", [(np.int16, 32768.0, 0.0), (np.int32, 2147483648.0, 0.0), (np.uint8, 128.0, 128.0), (np.float32, 1.0, 0.0)])
def test_load_wav_returns_unit_scale_floats_like_torchaudio(tmp_path, dtype, scale, offset):
    """train_util.load_wav_to_torch (:36-38) = torchaudio.load: float32 in [-1, 1] for every sample format, so get_mel's
    clamp branch (model_util.py:33) is the one a file takes and the prompt mel is deterministic."""
    rng = np.random.default_rng(3)
    x = rng.uniform(-0.9, 0.9, 4000)
    if dtype == np.float32:
        pcm = x.astype(np.float32)
    else:
        info = np.iinfo(dtype)
        pcm = np.clip(np.round(x * scale + offset), info.min, info.max).astype(dtype)
    path = str(tmp_path / "a.wav")
    wavfile.write(path, 22050, pcm)
    audio, sr = glue.load_wav(path)
    assert sr == 22050 and audio.dtype == torch.float32 and audio.shape == (4000,)
    want = (pcm.astype(np.float64) - offset) / scale
    assert np.array_equal(audio.numpy(), want.astype(np.float32))
    assert float(audio.abs().max()) <= 1.0
    again, _ = glue.load_wav(path)
    assert torch.equal(audio, again)


def test_load_wav_first_channel_of_stereo(tmp_path):
    pcm = (np.arange(200, dtype=np.int16).reshape(100, 2) * 37) % 2000
    path = str(tmp_path / "s.wav")
    wavfile.write(path, 22050, pcm.astype(np.int16))
    audio, _ = glue.load_wav(path)
    assert np.array_equal(audio.numpy(), pcm[:, 0].astype(np.float32) / 32768.0)


def test_vocoder_h_reads_as_attributes_like_the_reference_attrdict():
    """reference callers: process_unit(agent_unit, vocoder.h, device), vocoder.h.num_mels (model_util.py:73,84)."""
    h = AttrDict(BIGVGAN_22K_80)
    assert h.num_mels == 80 and h["hop_size"] == 256 and h.sampling_rate == 22050
    assert (h.n_fft, h.win_size, h.fmin, h.fmax) == (1024, 1024, 0, 8000)
    h.extra = 3
    assert h["extra"] == 3
    with pytest.raises(AttributeError):
        h.missing
    from usdm_b200 import BigVGAN
    voc = BigVGAN()
    assert voc.h.num_mels == 80 and voc.h["upsample_rates"] == [4, 4, 2, 2, 2, 2]
    hps = glue._hps_of(voc, glue.DEFAULT_HPS)
    assert (hps.sampling_rate, hps.hop_size, hps.n_fft, hps.num_mels, hps.fmax) == (22050, 256, 1024, 80, 8000)
    small = BigVGAN(dict(BIGVGAN_22K_80, sampling_rate=16000, hop_size=256, n_fft=None))
    assert glue._hps_of(small, glue.DEFAULT_HPS).sampling_rate == 16000


def test_get_mel_rejects_cpu_and_wrong_rate(tmp_path):
    path = str(tmp_path / "r.wav")
    wavfile.write(path, 16000, np.zeros(2000, np.int16))
    with pytest.raises(ValueError):
        glue.get_mel(path)
