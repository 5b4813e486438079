"""CPU: the decoder-glue oracle (oracle/glue_oracle.py) and the product's host-side constants (usdm_b200/glue.py)."""
import os

import numpy as np
import pytest
import torch

from util import ROOT  # noqa: F401
from oracle import glue_oracle as G


def test_slaney_filterbank_matches_torchaudio_and_product():
    ta = pytest.importorskip("torchaudio")
    fb = ta.functional.melscale_fbanks(513, 0.0, 8000.0, 80, 22050, norm="slaney", mel_scale="slaney").t().numpy()
    mine = G.slaney_mel_filterbank()
    assert mine.shape == (80, 513)
    assert np.abs(mine - fb).max() < 1e-5 * np.abs(fb).max()        # float32 round-off of two independent evaluations
    from usdm_b200 import glue
    prod = glue.mel_filterbank(22050, 1024, 80, 0, 8000)
    assert np.abs(prod - mine).max() < 1e-8
    # every filter is a non-negative triangle with unit area in Hz (Slaney normalisation): sum * bin width ~ 1 once a band
    # spans enough FFT bins for the sampled sum to resolve it (the lowest bands are only ~3 bins wide)
    area = mine.sum(1) * (22050 / 1024)
    assert mine.min() >= 0 and (mine.max(1) > 0).all() and np.allclose(area[40:], 1.0, atol=0.05)


def test_process_unit_semantics():
    g = torch.Generator().manual_seed(0)
    u = torch.randint(0, 10000, (250,), generator=g)             # 5 s of 50 Hz units
    out, new_len = G.process_unit(u)
    assert out.shape == (1, 430) and new_len == 430 * 256        # 5 s -> 430 frames
    # frame 220 is the first exact 128/128 tie between two units (441 * 128 = 256 * 220 + 128): torch.mode picks the smaller
    a, b = int(u[127]), int(u[128])
    assert int(out[0, 220]) == min(a, b)


def test_mel_of_reference_prompt_shape_if_sample_present():
    path = "/root/reference/samples/3_0_d1581_ref.wav"
    if not os.path.exists(path):
        pytest.skip("reference samples not mounted")
    from scipy.io import wavfile
    sr, d = wavfile.read(path)
    assert sr == 22050
    n = len(d) // 256 * 256
    mel = G.get_mel(torch.from_numpy(d), n)
    assert mel.shape == (1, 80, 272)                              # P = 272
    assert torch.isfinite(mel).all() and float(mel.min()) >= np.log(1e-5) - 1e-6
