"""Host-side mirror of the reference vocoder `BigVGAN` (src/decoder/voicebox/vocoder/models.py:132-327) over libtvb.so.

Same constructor (`BigVGAN(h, use_cuda_kernel=False)` with the hyper-parameter dict `h`), same state-dict key names
(conv_pre / ups.{i}.0 / resblocks.{n}.convs1|convs2.{m} / resblocks.{n}.activations.{m}.act.alpha|beta and the registered
`upsample.filter` / `downsample.lowpass.filter` buffers / activation_post / conv_post), same `forward(mel) -> wav`,
`remove_weight_norm()` and `from_pretrained` layout (config.json + bigvgan_generator.pt).  `forward` runs
include/tvb_vocoder.h's tvb_voc_forward (tcgen05 implicit-GEMM convolutions + fused anti-aliased snake kernels) on CUDA
tensors only; there is no CPU or PyTorch fallback.  Both weight layouts load: plain `weight` (after
`remove_weight_norm()`, model_util.py:68) and `weight_g` / `weight_v` pairs, which the library folds once.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os
from typing import Optional

import torch
from torch import nn

from . import engine as E

try:
    from huggingface_hub import PyTorchModelHubMixin
except Exception:  # pragma: no cover
    class PyTorchModelHubMixin:  # type: ignore
        pass

# the published hyper-parameters of nvidia/bigvgan_22khz_80band (the checkpoint model_util.py:64-67 downloads)
BIGVGAN_22K_80 = dict(
    num_mels=80, upsample_rates=[4, 4, 2, 2, 2, 2], upsample_kernel_sizes=[8, 8, 4, 4, 4, 4], upsample_initial_channel=1536,
    resblock="1", resblock_kernel_sizes=[3, 7, 11], resblock_dilation_sizes=[[1, 3, 5], [1, 3, 5], [1, 3, 5]],
    activation="snakebeta", snake_logscale=True, sampling_rate=22050, hop_size=256,
    # the audio front-end entries of the same config.json, which get_mel reads through `vocoder.h` (model_util.py:35-36)
    n_fft=1024, win_size=1024, fmin=0, fmax=8000,
)


class AttrDict(dict):
    """The reference's `vocoder.h` (vocoder/env.py:8-11): a dict whose keys also read and write as attributes."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k) from None

    def __setattr__(self, k, v):
        self[k] = v


def kaiser_sinc_filter(cutoff: float = 0.25, half_width: float = 0.3, kernel_size: int = 12) -> torch.Tensor:
    """The [1, 1, k] low-pass buffer `UpSample1d` / `LowPassFilter1d` register (alias_free_torch/filter.py:24-50)."""
    half = kernel_size // 2
    A = 2.285 * (half - 1) * math.pi * (4 * half_width) + 7.95
    beta = 0.1102 * (A - 8.7) if A > 50.0 else (0.5842 * (A - 21) ** 0.4 + 0.07886 * (A - 21.0) if A >= 21.0 else 0.0)
    win = torch.kaiser_window(kernel_size, beta=beta, periodic=False)
    t = (torch.arange(-half, half) + 0.5) if kernel_size % 2 == 0 else (torch.arange(kernel_size) - half).float()
    f = 2 * cutoff * win * torch.sinc(2 * cutoff * t)
    return (f / f.sum()).view(1, 1, kernel_size)


class _Conv(nn.Module):
    def __init__(self, cout, cin, k, transposed=False):
        super().__init__()
        self.weight = nn.Parameter(torch.randn((cin, cout, k) if transposed else (cout, cin, k)) * 0.01)   # utils.py:39-42
        self.bias = nn.Parameter(torch.zeros(cout))


class _Snake(nn.Module):
    def __init__(self, ch, beta: bool, logscale: bool):
        super().__init__()
        init = torch.zeros(ch) if logscale else torch.ones(ch)                                              # activations.py:84-89
        self.alpha = nn.Parameter(init.clone())
        if beta:
            self.beta = nn.Parameter(init.clone())


class _Filter(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("filter", kaiser_sinc_filter())


class _Down(nn.Module):
    def __init__(self):
        super().__init__()
        self.lowpass = _Filter()


class _Activation1d(nn.Module):
    def __init__(self, ch, beta, logscale):
        super().__init__()
        self.act = _Snake(ch, beta, logscale)
        self.upsample = _Filter()
        self.downsample = _Down()


class _AMPBlock1(nn.Module):
    def __init__(self, ch, k, beta, logscale):
        super().__init__()
        self.convs1 = nn.ModuleList([_Conv(ch, ch, k) for _ in range(3)])
        self.convs2 = nn.ModuleList([_Conv(ch, ch, k) for _ in range(3)])
        self.activations = nn.ModuleList([_Activation1d(ch, beta, logscale) for _ in range(6)])


class BigVGAN(nn.Module, PyTorchModelHubMixin):
    """Drop-in for voicebox.vocoder.models.BigVGAN (inference only; parameters are storage for the CUDA library)."""

    def __init__(self, h: Optional[dict] = None, use_cuda_kernel: bool = False):
        super().__init__()
        h = dict(h or BIGVGAN_22K_80)
        self.h = AttrDict(h)                 # reference callers use attribute access (`vocoder.h.num_mels`, model_util.py:73-84)
        if str(h.get("resblock", "1")) != "1":
            raise NotImplementedError("only resblock '1' (AMPBlock1) is implemented: it is what bigvgan_22khz_80band ships")
        if h["activation"] not in ("snake", "snakebeta"):
            raise NotImplementedError("activation incorrectly specified. check the config file and look for 'activation'.")   # models.py:67
        beta, logscale = h["activation"] == "snakebeta", bool(h.get("snake_logscale", False))
        c0 = h["upsample_initial_channel"]
        self.num_kernels, self.num_upsamples = len(h["resblock_kernel_sizes"]), len(h["upsample_rates"])
        self.conv_pre = _Conv(c0, h["num_mels"], 7)
        self.ups = nn.ModuleList([nn.ModuleList([_Conv(c0 // 2 ** (i + 1), c0 // 2 ** i, k, transposed=True)])
                                  for i, k in enumerate(h["upsample_kernel_sizes"])])
        self.resblocks = nn.ModuleList()
        ch = c0
        for i in range(self.num_upsamples):
            ch = c0 // 2 ** (i + 1)
            for k in h["resblock_kernel_sizes"]:
                self.resblocks.append(_AMPBlock1(ch, k, beta, logscale))
        self.activation_post = _Activation1d(ch, beta, logscale)
        self.conv_post = _Conv(1, ch, 7)
        self.hop = int(math.prod(h["upsample_rates"]))
        self._precision = E.PRECISION_BF16
        self._handle: Optional[C.c_void_p] = None
        self._handle_device: Optional[int] = None
        self._weights_version = 0
        self._uploaded_version = -1
        self._options: dict = {}
        self._extra_state: dict = {}          # weight_g / weight_v tensors kept verbatim when a weight-normed checkpoint is loaded

    # ------------------------------------------------------------------ reference API
    def remove_weight_norm(self):
        """models.py:224-233.  The library folds g * v / ||v|| at upload; nothing to strip on the host mirror."""
        return None

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        """Accepts both layouts; `weight_g` / `weight_v` pairs are handed to the library untouched (folded on the device)."""
        sd, extra = {}, {}
        for k, v in state_dict.items():
            (extra if k.endswith(("weight_g", "weight_v")) else sd)[k] = v
        for k in list(extra):
            if k.endswith("weight_v"):          # placeholder so that nn.Module's strict check passes; the library uses g and v
                sd[k[:-2]] = extra[k]
        out = super().load_state_dict(sd, strict=strict, **kw)
        self._extra_state = {k: v.detach().clone() for k, v in extra.items()}
        self._weights_version += 1
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._weights_version += 1
        return out

    def set_precision(self, precision: str) -> "BigVGAN":
        self._precision = {"bf16": E.PRECISION_BF16, "fp32": E.PRECISION_FP32_VALIDATE}[precision]
        if self._handle is not None:
            E.check(E.load_library().tvb_voc_set_precision(self._handle, self._precision))
        return self

    def set_option(self, name: str, value: int) -> "BigVGAN":
        self._options[name] = int(value)
        if self._handle is not None:
            E.check(E.load_library().tvb_voc_set_option(self._handle, name.encode(), int(value)))
        return self

    def _config(self) -> E.TvbVocConfig:
        h, c = self.h, E.TvbVocConfig()
        c.num_mels, c.upsample_initial_channel = h["num_mels"], h["upsample_initial_channel"]
        c.num_upsamples, c.num_kernels = self.num_upsamples, self.num_kernels
        for i, (u, k) in enumerate(zip(h["upsample_rates"], h["upsample_kernel_sizes"])):
            c.upsample_rates[i], c.upsample_kernel_sizes[i] = u, k
        for j, (k, d) in enumerate(zip(h["resblock_kernel_sizes"], h["resblock_dilation_sizes"])):
            c.resblock_kernel_sizes[j] = k
            for m in range(3):
                c.resblock_dilation_sizes[j][m] = d[m]
        c.snake_logscale, c.snakebeta = int(bool(h.get("snake_logscale", False))), int(h["activation"] == "snakebeta")
        return c

    def _release(self):
        if getattr(self, "_handle", None) is not None:
            try:
                E.load_library().tvb_voc_destroy(self._handle)
            except Exception:
                pass
            object.__setattr__(self, "_handle", None)      # plain attribute write: nn.Module.__setattr__ can fail at interpreter shutdown

    def __del__(self):
        self._release()

    def _engine(self, device: torch.device):
        lib = E.load_library()
        if not torch.cuda.is_available():
            raise RuntimeError("usdm_b200 needs a CUDA (sm_100a) device; there is no CPU path")
        idx = device.index if device.index is not None else torch.cuda.current_device()
        if self._handle is not None and (self._handle_device != idx or self._uploaded_version != self._weights_version):
            self._release()
        if self._handle is None:
            hd = C.c_void_p()
            cfg = self._config()
            E.check(lib.tvb_voc_create(C.byref(cfg), idx, C.byref(hd)))
            self._handle, self._handle_device = hd, idx
            try:
                sd = dict(self.state_dict())
                for k, v in self._extra_state.items():            # weight-normed checkpoint: ship g and v, not the placeholder
                    sd.pop(k[:-2], None)
                    sd[k] = v
                for key, ten in sd.items():
                    t = ten.detach().to(torch.float32).contiguous()
                    shape = (C.c_int64 * t.dim())(*t.shape)
                    E.check(lib.tvb_voc_load_weight(hd, key.encode(), C.c_void_p(t.data_ptr()), shape, t.dim()))
                E.check(lib.tvb_voc_finalize_weights(hd))
                E.check(lib.tvb_voc_set_precision(hd, self._precision))
                for k, v in self._options.items():
                    E.check(lib.tvb_voc_set_option(hd, k.encode(), v))
            except Exception:
                self._release()
                raise
            self._uploaded_version = self._weights_version
        return lib, self._handle

    @torch.no_grad()
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """models.py:192-222: mel fp32 [B, num_mels, T] -> wav [B, 1, T * hop]."""
        if not x.is_cuda:
            raise RuntimeError("usdm_b200.BigVGAN runs on CUDA tensors only (no CPU path); move the mel to the B200")
        if x.dim() != 3 or x.shape[1] != self.h["num_mels"]:
            raise ValueError(f"expected a [B, {self.h['num_mels']}, T] mel, got {tuple(x.shape)}")
        lib, hd = self._engine(x.device)
        mel = x.to(torch.float32).contiguous()
        B, _, T = mel.shape
        wav = torch.empty(B, 1, T * self.hop, device=x.device, dtype=torch.float32)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        E.check(lib.tvb_voc_forward(hd, C.c_void_p(mel.data_ptr()), B, T, C.c_void_p(wav.data_ptr()), st))
        return wav

    # ------------------------------------------------------------------ measurement / test hooks
    def debug_activation(self, name: str, x: torch.Tensor) -> torch.Tensor:
        lib, hd = self._engine(x.device)
        x = x.to(torch.float32).contiguous()
        out = torch.empty_like(x)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        E.check(lib.tvb_voc_debug_activation(hd, name.encode(), C.c_void_p(x.data_ptr()), x.shape[0], x.shape[2], C.c_void_p(out.data_ptr()), st))
        return out

    def set_profile(self, on: bool) -> None:
        if self._handle is not None:
            E.check(E.load_library().tvb_voc_set_profile(self._handle, int(on)))

    def read_profile(self) -> dict:
        ms = (C.c_float * 19)()
        E.check(E.load_library().tvb_voc_profile_read(self._handle, ms, 19))
        n = self.num_upsamples
        return dict(conv=ms[0], activation=ms[1], other=ms[2], conv_by_stage=[ms[3 + i] for i in range(n)],
                    activation_by_stage=[ms[11 + i] for i in range(n)])

    def kernel_launches(self) -> int:
        return 0 if self._handle is None else int(E.load_library().tvb_voc_kernel_launches(self._handle))

    def device_bytes(self) -> int:
        return 0 if self._handle is None else int(E.load_library().tvb_voc_device_bytes(self._handle))

    # ------------------------------------------------------------------ checkpoint layout (models.py:236-327)
    def _save_pretrained(self, save_directory) -> None:
        torch.save({"generator": self.state_dict()}, os.path.join(str(save_directory), "bigvgan_generator.pt"))
        with open(os.path.join(str(save_directory), "config.json"), "w") as f:
            json.dump(dict(self.h), f, indent=4)

    @classmethod
    def _from_pretrained(cls, *, model_id, revision=None, cache_dir=None, force_download=False, proxies=None, resume_download=None,
                         local_files_only=False, token=None, map_location="cpu", strict=False, use_cuda_kernel=False, **model_kwargs):
        if os.path.isdir(model_id):
            cfg_file, model_file = os.path.join(model_id, "config.json"), os.path.join(model_id, "bigvgan_generator.pt")
        else:
            from huggingface_hub import hf_hub_download
            kw = dict(repo_id=model_id, revision=revision, cache_dir=cache_dir, force_download=force_download, proxies=proxies,
                      token=token, local_files_only=local_files_only)
            cfg_file = hf_hub_download(filename="config.json", **kw)
            model_file = hf_hub_download(filename="bigvgan_generator.pt", **kw)
        with open(cfg_file) as f:
            h = json.load(f)
        model = cls(h, use_cuda_kernel=use_cuda_kernel)
        ckpt = torch.load(model_file, map_location=map_location)
        model.load_state_dict(ckpt["generator"], strict=False)
        return model
