"""GPU: decoder glue kernels (usdm_b200/csrc/glue.cuh through usdm_b200/glue.py) against the torch oracle."""
import numpy as np
import pytest
import torch

from util import ROOT, model, rel_l2  # noqa: F401
from oracle import glue_oracle as G

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 7, 250, 1000, 4321])
def test_process_unit_bit_exact(n):
    from usdm_b200 import glue
    g = torch.Generator().manual_seed(n)
    u = torch.randint(0, 10000, (n,), generator=g)
    if n >= 250:
        u[100:140] = u[100]                                       # long runs and immediate repeats
        u[141] = u[139]
    want, wl = G.process_unit(u)
    got, gl = glue.process_unit(u)
    assert gl == wl and got.shape == want.shape and got.dtype == torch.int64
    assert torch.equal(got.cpu(), want)


def test_process_unit_general_ratio_multiple_units_per_window():
    """rep < hop: several units per window, the same value may occur in separate runs; ties -> smallest."""
    from usdm_b200 import engine
    import ctypes as C
    lib = engine.load_library()
    g = torch.Generator().manual_seed(3)
    u = torch.randint(0, 4, (997,), generator=g)
    rep, hop = 7, 50
    rep_seq = torch.repeat_interleave(u, rep)
    n_out = len(rep_seq) // hop
    want = rep_seq[:n_out * hop].reshape(-1, hop).mode(1)[0]
    ud = u.cuda()
    out = torch.empty(n_out, dtype=torch.int64, device="cuda")
    engine.check(lib.tvb_process_unit(0, C.c_void_p(ud.data_ptr()), len(u), rep, hop, C.c_void_p(out.data_ptr()), n_out,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    assert torch.equal(out.cpu(), want)


@pytest.mark.parametrize("n", [256 * 3, 256 * 272, 69796 // 256 * 256, 22050 * 10 // 256 * 256])
def test_mel_spectrogram_vs_oracle(n):
    from usdm_b200 import glue
    g = torch.Generator().manual_seed(n)
    t = torch.arange(n) / 22050.0
    y = 0.3 * torch.sin(2 * np.pi * 220 * t) + 0.2 * torch.sin(2 * np.pi * 3100 * t + 1.0) + 0.05 * torch.randn(n, generator=g)
    y = y.clamp(-1, 1)
    want = G.mel_spectrogram(y.unsqueeze(0))
    got = glue.mel_spectrogram(y, normalize=False).cpu()
    assert got.shape == want.shape == (1, 80, n // 256)
    err = (got - want).abs().max().item()
    print(f"log-mel max-abs err vs torch.stft oracle: {err:.2e}")
    assert err < 2e-4                                              # fp32 FFT round-off under the log
    gotn = glue.mel_spectrogram(y, normalize=True).cpu()
    assert (gotn - (want - G.mel_mean) / G.mel_std).abs().max().item() < 2e-4


def test_reconstruct_mel_runs_end_to_end_and_matches_manual_pipeline():
    """glue.reconstruct_mel == process_unit + get_mel + generate + de-normalise assembled by hand (model_util.py:72-103)."""
    from usdm_b200 import glue
    m = model(0, False, "bf16")
    g = torch.Generator().manual_seed(1)
    agent = torch.randint(0, 10000, (40,), generator=g)
    ref_unit = torch.randint(0, 10000, (25,), generator=g)
    n = 22050
    ref_audio = (0.1 * torch.randn(n, generator=g)).clamp(-1, 1)
    torch.manual_seed(5)
    out = glue.reconstruct_mel(agent, m, ref_audio, ref_unit, n_timesteps=4, solver="euler")
    ru, new_len = G.process_unit(ref_unit)
    au, _ = G.process_unit(agent)
    rm = (G.get_mel(ref_audio, new_len) - G.mel_mean) / G.mel_std
    P, T = ru.shape[-1], ru.shape[-1] + au.shape[-1]
    cond = torch.zeros(1, 80, T)
    cond[:, :, :P] = rm
    torch.manual_seed(5)
    want = m.generate(torch.cat([ru, au], -1).cuda(), cond.cuda(), torch.tensor([T]), 4, solver="euler", gradient_scale=1.0,
                      speech_prompt=True, prompt_lengths=torch.tensor([P]))[:, :, P:] * G.mel_std + G.mel_mean
    torch.cuda.synchronize()
    assert out.shape == (1, 80, au.shape[-1])
    e = rel_l2(out, want)
    print("reconstruct_mel vs manual pipeline rel-L2", e)
    assert e < 1e-3


def test_reconstruct_speech_units_to_waveform_matches_oracle_chain():
    """glue.reconstruct_speech (model_util.py:72-105): units -> mel -> wav on the GPU (both models on their fp32 validation paths)
    against the oracle chain fed the SAME mel-stage noise: oracle Voicebox Euler solve + oracle BigVGAN forward."""
    import numpy as np
    from usdm_b200 import glue, BigVGAN
    from oracle import vocoder_oracle as V
    from util import weights, CFG, O
    m = model(0, False, "fp32")
    vsd = V.init_vocoder_weights(11, V.SMALL_CFG)
    voc = BigVGAN(V.SMALL_CFG)
    voc.load_state_dict(vsd)
    voc = voc.cuda().eval().set_precision("fp32")
    g = torch.Generator().manual_seed(2)
    agent = torch.randint(0, 10000, (30,), generator=g)
    ref_unit = torch.randint(0, 10000, (20,), generator=g)
    ref_audio = (0.1 * torch.randn(22050, generator=g)).clamp(-1, 1)
    torch.manual_seed(7)
    audio = glue.reconstruct_speech(agent, m, voc, ref_audio, ref_unit, n_timesteps=2, solver="euler")
    au, _ = G.process_unit(agent)
    assert isinstance(audio, np.ndarray) and audio.dtype == np.float32 and audio.shape == (au.shape[-1] * 256,)
    assert np.abs(audio).max() <= 1.0
    # the mel the GPU path produced, then the oracle vocoder on it: isolates the mel -> wav hand-off (de-normalisation, layout, clamp)
    torch.manual_seed(7)
    mel = glue.reconstruct_mel(agent, m, ref_audio, ref_unit, n_timesteps=2, solver="euler")
    want = V.bigvgan_forward(vsd, mel.cpu(), V.SMALL_CFG).squeeze().clamp(-1, 1).numpy()
    err = float(np.linalg.norm(audio - want) / np.linalg.norm(want))
    print("reconstruct_speech vs oracle vocoder on the same mel: rel-L2", err)
    assert err < 1e-5
    m.set_precision("bf16")
