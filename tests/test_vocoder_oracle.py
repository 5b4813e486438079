"""CPU: the vocoder oracle against the goldens made from the REAL reference BigVGAN (oracle/make_golden_vocoder.py),
against the live reference when /root/reference exists, and the host mirror's state-dict layout."""
import math
import os

import pytest
import torch

from util import golden, rel_l2, ROOT
from oracle import vocoder_oracle as V


@pytest.mark.parametrize("name", ["vocoder_small_b2.pt", "vocoder_small_wn.pt", "vocoder_full_t6.pt"])
def test_oracle_matches_reference_golden(name):
    g = golden(name)
    sd = V.init_vocoder_weights(g["seed"], g["cfg"], weight_norm=g["weight_norm"])
    assert math.isclose(V.state_dict_checksum(sd), g["checksum"], rel_tol=1e-9), "seeded weights drifted from the fixture's"
    wav = V.bigvgan_forward(sd, g["mel"], g["cfg"])
    assert wav.shape == g["wav"].shape == (g["mel"].shape[0], 1, g["mel"].shape[2] * 256)
    assert rel_l2(wav, g["wav"]) < 1e-5
    assert float(g["wav"].abs().max()) < 0.999, "fixture must not sit in tanh saturation"


@pytest.mark.skipif(not os.path.isdir("/root/reference/src/decoder"), reason="reference tree not present")
def test_oracle_matches_live_reference():
    from oracle.make_golden_vocoder import build_reference
    torch.manual_seed(3)
    for wn in (False, True):
        sd = V.init_vocoder_weights(21, V.SMALL_CFG, weight_norm=wn)
        ref = build_reference(V.SMALL_CFG, sd, wn)
        mel = torch.randn(2, 80, 7)
        with torch.no_grad():
            want = ref(mel)
        assert rel_l2(V.bigvgan_forward(sd, mel, V.SMALL_CFG), want) < 1e-5
        # the reference's registered filters equal the restated kaiser-sinc
        assert torch.allclose(ref.resblocks[0].activations[0].downsample.lowpass.filter.flatten(), V.kaiser_sinc_filter1d(0.25, 0.3, 12), atol=1e-7)


def test_polyphase_form_of_activation1d():
    """The per-phase formulas the CUDA kernel implements (vocoder.cuh) equal the reference's conv_transpose / conv form."""
    f = V.kaiser_sinc_filter1d(0.25, 0.3, 12)
    for T in (1, 2, 5, 13):
        x = torch.randn(T, dtype=torch.float64)
        cl = lambda i: x[min(max(i, 0), T - 1)]
        u = torch.zeros(2 * T, dtype=torch.float64)
        for q in range(T):
            u[2 * q] = 2 * sum(f[2 * m + 1].double() * cl(q + 2 - m) for m in range(6))
            u[2 * q + 1] = 2 * sum(f[2 * m].double() * cl(q + 3 - m) for m in range(6))
        want_u = V.upsample2(x.view(1, 1, T).float(), f)[0, 0]
        assert torch.allclose(u.float(), want_u, atol=1e-5)
        y = torch.stack([sum(f[j].double() * u[min(max(2 * t + j - 5, 0), 2 * T - 1)] for j in range(12)) for t in range(T)])
        want_y = V.downsample2(u.view(1, 1, -1).float(), f)[0, 0]
        assert torch.allclose(y.float(), want_y, atol=1e-5)


def test_transposed_conv_as_three_tap_conv():
    """ConvTranspose1d(k = 2u, stride u, pad u/2) == 3-tap conv with u*C_out outputs per input step (vocoder.cuh k_voc_pack_convT)."""
    for u, k in ((4, 8), (2, 4)):
        cin, cout, T = 5, 3, 7
        W, b, x = torch.randn(cin, cout, k), torch.randn(cout), torch.randn(1, cin, T)
        want = torch.nn.functional.conv_transpose1d(x, W, b, stride=u, padding=(k - u) // 2)
        pad = (k - u) // 2
        Wc = torch.zeros(u * cout, 3, cin)
        for r in range(u):
            for s in range(3):
                j = r + pad + (1 - s) * u
                if 0 <= j < k:
                    Wc[r * cout:(r + 1) * cout, s] = W[:, :, j].t()
        xp = torch.nn.functional.pad(x, (1, 1))[0].t()
        out = torch.stack([sum(Wc[:, s] @ xp[t + s] for s in range(3)) for t in range(T)]).view(T * u, cout) + b
        assert torch.allclose(out.t(), want[0], atol=1e-5)


def test_host_mirror_state_dict_layout_and_errors():
    from usdm_b200.vocoder import BigVGAN, BIGVGAN_22K_80, kaiser_sinc_filter
    assert BIGVGAN_22K_80 == V.BIGVGAN_22K_80
    m = BigVGAN(V.SMALL_CFG)
    sd = V.init_vocoder_weights(5, V.SMALL_CFG)
    assert set(m.state_dict()) == set(sd)
    assert all(m.state_dict()[k].shape == v.shape for k, v in sd.items())
    m.load_state_dict(sd)
    m.load_state_dict(V.init_vocoder_weights(5, V.SMALL_CFG, weight_norm=True))       # weight_g / weight_v layout loads too
    assert len(m._extra_state) == 2 * sum(1 for k in sd if k.endswith(".weight"))
    assert torch.allclose(kaiser_sinc_filter().flatten(), V.kaiser_sinc_filter1d(0.25, 0.3, 12), atol=0)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 80, 4))                                                       # CPU tensor: no fallback
    with pytest.raises(NotImplementedError):
        BigVGAN(dict(V.SMALL_CFG, resblock="2"))
    full = BigVGAN()
    assert abs(sum(p.numel() for p in full.parameters()) / 1e6 - 112.4) < 1.0          # vocoder/README.md:73: 112M


def test_save_and_from_pretrained_round_trip(tmp_path):
    from usdm_b200.vocoder import BigVGAN
    m = BigVGAN(V.SMALL_CFG)
    m.load_state_dict(V.init_vocoder_weights(6, V.SMALL_CFG))
    m._save_pretrained(tmp_path)
    m2 = BigVGAN._from_pretrained(model_id=str(tmp_path))
    assert m2.h["upsample_initial_channel"] == V.SMALL_CFG["upsample_initial_channel"]
    assert all(torch.equal(v, m2.state_dict()[k]) for k, v in m.state_dict().items())
