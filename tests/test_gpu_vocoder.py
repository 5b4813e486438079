"""GPU: the CUDA vocoder path (include/tvb_vocoder.h through usdm_b200.vocoder.BigVGAN) against the oracle and the goldens
made from the real reference BigVGAN.  fp32 validation path: <= 1e-5 rel-L2 of the reference wav; bf16 tcgen05 path: <= 3e-2
(bf16 operands, fp32 accumulation / residual stream; no reference-side tolerance exists for the vocoder)."""
import pytest
import torch

from util import golden, rel_l2
from oracle import vocoder_oracle as V

pytestmark = pytest.mark.gpu
TOL_FP32, TOL_BF16 = 1e-5, 3e-2
_cache = {}


def vocoder(cfg_name, seed, wn=False, precision="bf16"):
    from usdm_b200.vocoder import BigVGAN
    key = (cfg_name, seed, wn)
    if key not in _cache:
        cfg = dict(small=V.SMALL_CFG, full=V.BIGVGAN_22K_80)[cfg_name]
        m = BigVGAN(cfg)
        m.load_state_dict(V.init_vocoder_weights(seed, cfg, weight_norm=wn))
        _cache[key] = m.cuda().eval()
    return _cache[key].set_precision(precision)


@pytest.mark.parametrize("name,cfg", [("vocoder_small_b2.pt", "small"), ("vocoder_small_wn.pt", "small"), ("vocoder_full_t6.pt", "full")])
@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_matches_reference_golden(name, cfg, precision):
    g = golden(name)
    m = vocoder(cfg, g["seed"], g["weight_norm"], precision)
    wav = m(g["mel"].cuda())
    assert wav.shape == g["wav"].shape
    err = rel_l2(wav, g["wav"])
    print(f"{name} {precision}: rel-L2 vs reference {err:.3e}")
    assert err < (TOL_FP32 if precision == "fp32" else TOL_BF16)
    assert m.kernel_launches() > 0


@pytest.mark.parametrize("option", ["halo", "act_vec2"])
@pytest.mark.parametrize("name,cfg", [("vocoder_small_b2.pt", "small"), ("vocoder_full_t6.pt", "full")])
def test_ablation_kernels_match_reference_golden(name, cfg, option):
    """The kernels behind the tvb_voc_set_option switches (one activation fetch per tap instead of the halo tile; the scalar
    activation kernel instead of the two-channel one) stay parity-green: they are the A/B record of DESIGN.md §9."""
    g = golden(name)
    m = vocoder(cfg, g["seed"], g["weight_norm"], "bf16")
    try:
        m.set_option(option, 0)
        err = rel_l2(m(g["mel"].cuda()), g["wav"])
    finally:
        m.set_option(option, 1)
    print(f"{name} bf16 {option}=0: rel-L2 vs reference {err:.3e}")
    assert err < TOL_BF16


@pytest.mark.parametrize("T", [1, 2, 5, 33, 64, 65, 200])
def test_activation_kernel_edges(T):
    """Activation1d alone (replicate edges on both FIRs) against the oracle, down to one-sample signals."""
    m = vocoder("small", 11, False, "fp32")
    m(torch.zeros(2, 80, 16, device="cuda"))                   # sizes the workspace
    sd = V.init_vocoder_weights(11, V.SMALL_CFG)
    g = torch.Generator().manual_seed(T)
    for name, C in (("resblocks.0.activations.0", 96), ("resblocks.16.activations.3", 3), ("activation_post", 3)):
        x = torch.randn(2, C, T, generator=g) * 2
        want = V.activation1d(sd, name, x, V.SMALL_CFG)
        got = m.debug_activation(name, x.cuda())
        assert rel_l2(got, want) < 2e-6, (name, T)


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("bf16", TOL_BF16)])
@pytest.mark.parametrize("B,T", [(1, 1), (3, 37), (2, 130)])
def test_ragged_shapes_vs_oracle(B, T, precision, tol):
    """Lengths that are not multiples of the 128-step tile / 64-step activation chunk, several time tiles, tiny T."""
    sd = V.init_vocoder_weights(11, V.SMALL_CFG)
    g = torch.Generator().manual_seed(B * 1000 + T)
    mel = torch.randn(B, 80, T, generator=g) * 1.2
    want = V.bigvgan_forward(sd, mel, V.SMALL_CFG)
    got = vocoder("small", 11, False, precision)(mel.cuda())
    assert rel_l2(got, want) < tol


@pytest.mark.parametrize("precision", ["fp32", "bf16"])
def test_batch_rows_are_independent_and_deterministic(precision):
    """Size-independent property: utterance b of a batch equals the same utterance run alone, bit for bit (no cross-talk
    through the time-axis zero padding of the 3-D tensor map), and a repeated call is bit-identical."""
    m = vocoder("small", 11, False, precision)
    g = torch.Generator().manual_seed(5)
    mel = (torch.randn(3, 80, 150, generator=g) * 1.2).cuda()
    full = m(mel)
    assert torch.equal(full, m(mel))
    for b in range(3):
        assert torch.equal(full[b:b + 1], m(mel[b:b + 1]))
    assert torch.isfinite(full).all() and float(full.abs().max()) <= 1.0


def test_full_config_bf16_vs_fp32_path_at_bench_shape():
    """BASELINE-sized model (112 M parameters) on 2 x 120 frames: the tcgen05 path against the fp32 CUDA path (itself pinned to
    the reference above); also exercises every N tile shape (256, 192, 96, 48, 32-padded) and the 3-D map at real sizes."""
    g = torch.Generator().manual_seed(9)
    mel = (torch.randn(2, 80, 120, generator=g) * 1.2).cuda()
    want = vocoder("full", 13, False, "fp32")(mel)
    got = vocoder("full", 13, False, "bf16")(mel)
    err = rel_l2(got, want)
    print(f"full config bf16 vs fp32 path: {err:.3e}")
    assert err < TOL_BF16


def test_errors():
    from usdm_b200.vocoder import BigVGAN
    m = vocoder("small", 11)
    with pytest.raises(ValueError):
        m(torch.zeros(1, 79, 8, device="cuda"))
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 80, 8))
    bad = BigVGAN(dict(V.SMALL_CFG, upsample_kernel_sizes=[8, 8, 4, 4, 4, 6]))
    with pytest.raises(ValueError):
        bad.cuda()(torch.zeros(1, 80, 8, device="cuda"))


def test_vocode_sharded_ragged_batch_equals_per_utterance_reference_calls():
    """parallel.vocode_sharded (world 1 here; the world-2 logic is tests/test_parallel_gloo.py): a ragged batch is vocoded
    utterance by utterance at true length -- the reference's one-utterance call (model_util.py:104) -- not at the padded length."""
    from usdm_b200 import parallel
    sd = V.init_vocoder_weights(11, V.SMALL_CFG)
    m = vocoder("small", 11, False, "fp32")
    g = torch.Generator().manual_seed(3)
    lens = torch.tensor([20, 7, 20, 13])
    mel = torch.randn(4, 80, 20, generator=g) * 1.2
    for u, t in enumerate(lens.tolist()):
        mel[u, :, t:] = 0
    wav = parallel.vocode_sharded(m, mel.cuda(), lens, hop=256).cpu()
    assert wav.shape == (4, 1, 20 * 256)
    for u, t in enumerate(lens.tolist()):
        want = V.bigvgan_forward(sd, mel[u:u + 1, :, :t], V.SMALL_CFG)
        assert rel_l2(wav[u:u + 1, :, :t * 256], want) < TOL_FP32
        assert float(wav[u, :, t * 256:].abs().max()) == 0.0 if t < 20 else True


def test_full_model_at_maximum_length_properties():
    """The 112 M-parameter model at BASELINE's longest utterance (20 s = 1723 frames, 441 k samples, 13.8 k time tiles in the last
    stage): the tcgen05 path against the fp32 CUDA path, row independence and determinism -- the size-independent checks that
    replace an oracle run the CPU cannot finish in seconds."""
    g = torch.Generator().manual_seed(21)
    mel = (torch.randn(2, 80, 1723, generator=g) * 1.2).cuda()
    m = vocoder("full", 13, False, "bf16")
    got = m(mel)
    assert got.shape == (2, 1, 1723 * 256) and torch.isfinite(got).all()
    assert torch.equal(got, m(mel))
    assert torch.equal(got[1:2], m(mel[1:2]))
    want = vocoder("full", 13, False, "fp32")(mel)
    err = rel_l2(got, want)
    print(f"full model, 2 x 1723 frames: bf16 vs fp32 path {err:.3e}")
    assert err < TOL_BF16
