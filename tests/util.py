"""Shared helpers for the test-suite (the oracle is imported here and ONLY used as the checker)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import voicebox_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
CFG = dict(O.DEFAULT_MODEL_CFG)

_sd_cache = {}
_model_cache = {}


def golden(name):
    return torch.load(os.path.join(GOLD, name), weights_only=False)


def weights(seed=0, perturb=False):
    key = (seed, perturb)
    if key not in _sd_cache:
        _sd_cache[key] = O.init_reference_weights(seed, CFG, perturb=perturb)
    return _sd_cache[key]


def model(seed=0, perturb=False, precision="bf16"):
    """usdm_b200.Voicebox carrying the same random-init weights as the oracle state dict."""
    import usdm_b200
    key = (seed, perturb)
    if key not in _model_cache:
        m = usdm_b200.Voicebox(**CFG)
        m.load_state_dict(weights(seed, perturb))
        _model_cache[key] = m.eval()
    m = _model_cache[key]
    m.set_precision(precision)
    m.set_option("attn_tc", 1)
    m.set_option("conv_tc", 1)
    m.set_option("gemm_2cta", 1)
    m.set_option("gemm_mc", 0)
    m.set_option("gemm_bn", 1)
    m.set_option("resid_l2", 0)
    m.set_option("attn_v", 10)
    m.set_option("attn_win", 1)
    m.set_option("attn_poly", 4)
    m.set_option("ln_fuse", 1)
    m.set_option("ln1_fold", 1)
    m.set_option("zigzag", 1)
    m.set_option("graph", 1)
    m.set_option("pdl", 1)
    m.set_option("frame_tc", 1)
    m.set_option("out_tc", 1)
    m.set_option("conv_halo", 1)
    return m


def rel_l2(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a, b):
    return float((a.detach().double().cpu() - b.detach().double().cpu()).abs().max())


def oracle_on(device):
    """Oracle state dict + estimator closure on `device` (fp32, TF32 off)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return device
