"""Generate tests/golden/vocoder_*.pt from the REAL reference BigVGAN (run in the authoring container only).

    python -m oracle.make_golden_vocoder

TEST INFRASTRUCTURE ONLY.  Imports /root/reference/src/decoder/voicebox/vocoder/models.py unmodified.  Its sibling
utils.py / meldataset.py import matplotlib and librosa (plotting / dataset helpers, never called on the forward path)
which are not installed here, so inert stub modules are registered first; nothing else is shimmed.  Weights are NOT stored: oracle.vocoder_oracle
.init_vocoder_weights(seed, cfg, weight_norm) re-creates them on any box (fingerprinted by `checksum`).
"""
from __future__ import annotations

import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import vocoder_oracle as V      # noqa: E402
from oracle.ref_harness import REF_DECODER  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def load_reference_bigvgan():
    class _Stub(types.ModuleType):                       # attribute access yields an inert callable: import-time only
        def __getattr__(self, k):
            if k.startswith("__"):
                raise AttributeError(k)
            return lambda *a, **kw: None
    for name in ("matplotlib", "matplotlib.pylab", "matplotlib.pyplot", "librosa", "librosa.util", "librosa.filters"):
        if name not in sys.modules:
            sys.modules[name] = _Stub(name)
    if REF_DECODER not in sys.path:
        sys.path.insert(0, REF_DECODER)
    from voicebox.vocoder.models import BigVGAN          # noqa: E402
    from voicebox.vocoder.env import AttrDict            # noqa: E402
    return BigVGAN, AttrDict


def build_reference(cfg: dict, sd: dict, weight_norm: bool):
    BigVGAN, AttrDict = load_reference_bigvgan()
    model = BigVGAN(AttrDict(cfg)).eval()
    if not weight_norm:
        model.remove_weight_norm()                        # model_util.py:68
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected, unexpected
    assert all(".filter" in k for k in missing), missing
    return model


def main():
    torch.set_grad_enabled(False)
    os.makedirs(GOLD, exist_ok=True)
    cases = [("vocoder_small_b2", V.SMALL_CFG, 11, 2, 9, False), ("vocoder_small_wn", V.SMALL_CFG, 12, 1, 5, True),
             ("vocoder_full_t6", V.BIGVGAN_22K_80, 13, 1, 6, False)]
    for name, cfg, seed, B, T, wn in cases:
        sd = V.init_vocoder_weights(seed, cfg, weight_norm=wn)
        if wn:   # torch >= 2.1 parametrizations use different key names; the old-style hook API (models.py:11) keeps weight_g/_v
            pass
        model = build_reference(cfg, sd, wn)
        g = torch.Generator().manual_seed(seed + 100)
        mel = torch.randn(B, cfg["num_mels"], T, generator=g) * 1.2
        wav = model(mel)
        mine = V.bigvgan_forward(sd, mel, cfg)
        err = float((wav - mine).norm() / wav.norm())
        print(f"{name}: wav {tuple(wav.shape)} rms {float(wav.pow(2).mean().sqrt()):.4f}  oracle vs reference rel-L2 {err:.3e}")
        assert err < 1e-5
        # also pin the reference's own registered filter buffers against the restated kaiser-sinc
        ref_filt = model.activation_post.upsample.filter.flatten()
        assert torch.allclose(ref_filt, V.kaiser_sinc_filter1d(0.25, 0.3, 12), atol=1e-7)
        torch.save(dict(cfg=cfg, seed=seed, weight_norm=wn, mel=mel, wav=wav, checksum=V.state_dict_checksum(sd),
                        torch_version=torch.__version__), os.path.join(GOLD, name + ".pt"))


if __name__ == "__main__":
    main()
