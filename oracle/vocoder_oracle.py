"""CPU/torch-fp32 ORACLE for the BigVGAN v1 vocoder forward.  TEST INFRASTRUCTURE ONLY.

Functional restatement over a state-dict with the reference's key names of
  /root/reference/src/decoder/voicebox/vocoder/models.py:132-222            (BigVGAN.__init__ / forward)
  /root/reference/src/decoder/voicebox/vocoder/models.py:27-85              (AMPBlock1)
  /root/reference/src/decoder/voicebox/vocoder/activations.py:52-119        (SnakeBeta; Snake :8-50)
  /root/reference/src/decoder/voicebox/vocoder/alias_free_torch/act.py, resample.py, filter.py  (Activation1d)
It may be imported ONLY by tests/, __graft_entry__.smoke() and tools/bench_vocoder.py's cpu baseline leg, and only as
the checker.  The product (usdm_b200/) never imports it.

Hyper-parameters: the reference downloads `nvidia/bigvgan_22khz_80band` (model_util.py:64-67); its config.json is not in
the reference tree.  BIGVGAN_22K_80 below restates the published configs/bigvgan_22khz_80band.json of NVIDIA/BigVGAN v1
(112 M parameters, matches vocoder/README.md:73's table row).

Parity pinning: the reference ships no vocoder tests.  oracle/make_golden_vocoder.py executes the REAL reference BigVGAN
(unmodified models.py; matplotlib, which utils.py imports only for plotting, is stubbed) on seeded weights / mels and
writes tests/golden/vocoder_*.pt; tests/test_vocoder_oracle.py compares this oracle against them and, when
/root/reference is present, against the live reference modules.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

BIGVGAN_22K_80 = dict(
    num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4], upsample_initial_channel=1536,
    resblock="1", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    activation="snakebeta", snake_logscale=True, sampling_rate=22050, hop_size=256,
    # audio front-end entries of the same published config.json (read through `vocoder.h` by get_mel, model_util.py:35-36)
    n_fft=1024, win_size=1024, fmin=0, fmax=8000,
)
SMALL_CFG = dict(BIGVGAN_22K_80, upsample_initial_channel=192)       # same topology, 96..3 channels: CPU-sized parity cases


def kaiser_sinc_filter1d(cutoff: float, half_width: float, kernel_size: int) -> torch.Tensor:
    """alias_free_torch/filter.py:24-50 (the [kernel_size] filter; the reference keeps it as [1, 1, k])."""
    half_size = kernel_size // 2
    delta_f = 4 * half_width
    A = 2.285 * (half_size - 1) * math.pi * delta_f + 7.95
    if A > 50.0:
        beta = 0.1102 * (A - 8.7)
    elif A >= 21.0:
        beta = 0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0)
    else:
        beta = 0.0
    window = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    if kernel_size % 2 == 0:
        time = torch.arange(-half_size, half_size) + 0.5
    else:
        time = torch.arange(kernel_size) - half_size
    filt = 2 * cutoff * window * torch.sinc(2 * cutoff * time)
    return filt / filt.sum()


def resblock_count(h: dict) -> int:
    return len(h["upsample_rates"]) * len(h["resblock_kernel_sizes"])


def init_vocoder_weights(seed: int = 0, h: Optional[dict] = None, weight_norm: bool = False) -> Dict[str, torch.Tensor]:
    """Seeded random state-dict with the reference's key names (fp32, CPU).

    NOT the reference constructor's init (std 0.01 would make every activation vanish): fan-in scaled normals so that all
    stages carry O(1) signals, and random log-scale alpha / beta.  weight_norm=True emits `weight_g` / `weight_v` pairs
    (the layout before `remove_weight_norm()`, models.py:224-233) with g != ||v||.
    """
    h = h or BIGVGAN_22K_80
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, torch.Tensor] = {}

    def conv(name, cout, cin, k, transposed=False, gain=1.0):
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        fan = cin * k
        w = torch.randn(shape, generator=g) * (gain / math.sqrt(fan))
        if weight_norm:
            sd[name + ".weight_v"] = w
            nrm = w.flatten(1).norm(dim=1).view(-1, 1, 1)
            sd[name + ".weight_g"] = nrm * (0.75 + 0.5 * torch.rand(nrm.shape, generator=g))
        else:
            sd[name + ".weight"] = w
        sd[name + ".bias"] = 0.1 * torch.randn(cout, generator=g)

    def act(name, ch):
        sd[name + ".act.alpha"] = 0.5 * torch.randn(ch, generator=g)
        sd[name + ".act.beta"] = 0.5 * torch.randn(ch, generator=g)
        filt = kaiser_sinc_filter1d(0.25, 0.3, 12).view(1, 1, 12)
        sd[name + ".upsample.filter"] = filt.clone()
        sd[name + ".downsample.lowpass.filter"] = filt.clone()

    c0 = h["upsample_initial_channel"]
    conv("conv_pre", c0, h["num_mels"], 7)
    ch = c0
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        conv(f"ups.{i}.0", c0 // 2 ** (i + 1), c0 // 2 ** i, k, transposed=True, gain=math.sqrt(k / u))
    for i in range(len(h["upsample_rates"])):
        ch = c0 // 2 ** (i + 1)
        for j, k in enumerate(h["resblock_kernel_sizes"]):
            n = i * len(h["resblock_kernel_sizes"]) + j
            for m in range(3):
                conv(f"resblocks.{n}.convs1.{m}", ch, ch, k)
                conv(f"resblocks.{n}.convs2.{m}", ch, ch, k, gain=0.5)
            for m in range(6):
                act(f"resblocks.{n}.activations.{m}", ch)
    act("activation_post", ch)
    conv("conv_post", 1, ch, 7, gain=0.2)      # keeps tanh out of saturation so wav errors are not masked
    return sd


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    return float(sum(v.double().abs().sum() * (1 + (i % 7)) for i, (k, v) in enumerate(sorted(sd.items()))))


def _weight(sd, name):
    """Plain `weight`, or g * v / ||v|| over all dims but 0 (torch.nn.utils.weight_norm default dim=0)."""
    if name + ".weight" in sd:
        return sd[name + ".weight"].float()
    v, g = sd[name + ".weight_v"].float(), sd[name + ".weight_g"].float()
    return v * (g / v.flatten(1).norm(dim=1).view(-1, 1, 1))


def snake(x, alpha, beta, logscale: bool):
    """activations.py:107-119 (SnakeBeta); Snake (:40-50) is the beta == alpha case."""
    a, b = alpha.view(1, -1, 1), beta.view(1, -1, 1)
    if logscale:
        a, b = torch.exp(a), torch.exp(b)
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a) ** 2


def upsample2(x, filt):
    """resample.py:10-30 with ratio 2, kernel 12: replicate pad 5, transposed depthwise conv (x ratio), crop 15 each side."""
    C = x.shape[1]
    k, ratio = filt.numel(), 2
    pad = k // ratio - 1
    pad_left = pad * ratio + (k - ratio) // 2
    pad_right = pad * ratio + (k - ratio + 1) // 2
    x = F.pad(x, (pad, pad), mode="replicate")
    x = ratio * F.conv_transpose1d(x, filt.view(1, 1, k).expand(C, -1, -1), stride=ratio, groups=C)
    return x[..., pad_left:-pad_right]


def downsample2(x, filt):
    """resample.py:33-48 -> filter.py:53-94: replicate pad (k/2 - 1, k/2), depthwise conv stride 2."""
    C = x.shape[1]
    k = filt.numel()
    x = F.pad(x, (k // 2 - 1, k // 2), mode="replicate")
    return F.conv1d(x, filt.view(1, 1, k).expand(C, -1, -1), stride=2, groups=C)


def activation1d(sd, name, x, h):
    """alias_free_torch/act.py:21-27: up x2 -> snake -> down x2."""
    beta_key = name + ".act.beta"
    alpha = sd[name + ".act.alpha"].float()
    beta = sd[beta_key].float() if beta_key in sd else alpha
    x = upsample2(x, sd[name + ".upsample.filter"].float())
    x = snake(x, alpha, beta, h["snake_logscale"])
    return downsample2(x, sd[name + ".downsample.lowpass.filter"].float())


def amp_block1(sd, n, x, k, dilations, h):
    """models.py:70-79."""
    for m, d in enumerate(dilations):
        xt = activation1d(sd, f"resblocks.{n}.activations.{2 * m}", x, h)
        xt = F.conv1d(xt, _weight(sd, f"resblocks.{n}.convs1.{m}"), sd[f"resblocks.{n}.convs1.{m}.bias"].float(),
                      dilation=d, padding=(k * d - d) // 2)
        xt = activation1d(sd, f"resblocks.{n}.activations.{2 * m + 1}", xt, h)
        xt = F.conv1d(xt, _weight(sd, f"resblocks.{n}.convs2.{m}"), sd[f"resblocks.{n}.convs2.{m}.bias"].float(),
                      dilation=1, padding=(k - 1) // 2)
        x = xt + x
    return x


@torch.no_grad()
def bigvgan_forward(sd: Dict[str, torch.Tensor], mel: torch.Tensor, h: Optional[dict] = None, taps: Optional[dict] = None):
    """models.py:192-222: mel fp32 [B, num_mels, T] -> wav [B, 1, T * prod(upsample_rates)].

    `taps`, if given, receives intermediate tensors (conv_pre output, each stage's output) for layer-wise parity tests.
    """
    h = h or BIGVGAN_22K_80
    assert h["resblock"] == "1", "the shipped 22 kHz checkpoint uses AMPBlock1 (models.py:155)"
    nk = len(h["resblock_kernel_sizes"])
    x = F.conv1d(mel.float(), _weight(sd, "conv_pre"), sd["conv_pre.bias"].float(), padding=3)
    if taps is not None:
        taps["conv_pre"] = x
    for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
        x = F.conv_transpose1d(x, _weight(sd, f"ups.{i}.0"), sd[f"ups.{i}.0.bias"].float(), stride=u, padding=(k - u) // 2)
        if taps is not None:
            taps[f"ups.{i}"] = x
        xs = None
        for j in range(nk):
            r = amp_block1(sd, i * nk + j, x, h["resblock_kernel_sizes"][j], h["resblock_dilation_sizes"][j], h)
            xs = r if xs is None else xs + r
        x = xs / nk
        if taps is not None:
            taps[f"stage.{i}"] = x
    x = activation1d(sd, "activation_post", x, h)
    x = F.conv1d(x, _weight(sd, "conv_post"), sd["conv_post.bias"].float(), padding=3)
    return torch.tanh(x)
