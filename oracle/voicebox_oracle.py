"""CPU/torch-fp32 ORACLE for the token-Voicebox mel generator.  TEST INFRASTRUCTURE ONLY.

This file is a from-the-spec restatement of the reference hot path
  /root/reference/src/decoder/voicebox/model/networks.py:302-374   (Transformer.forward)
  /root/reference/src/decoder/voicebox/model/voicebox.py:51-150    (CFM.sample / solve_euler / solve_heun / generate)
written as plain functional torch fp32 over a state-dict with the reference's key names.

It may be imported ONLY by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, and only as the checker (or as the timed CPU baseline).  The product (usdm_b200/) never imports it.

Parity pinning: the reference ships no tests / golden vectors.  This oracle is pinned
against outputs of the reference itself executed in the authoring container (oracle/make_golden.py writes
tests/golden/*.pt from the real reference; tests/test_oracle_golden.py compares).  When /root/reference
is present the test-suite also checks the oracle against the live reference modules.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

DEFAULT_MODEL_CFG = dict(  # src/decoder/voicebox/configs/YOUR_DATA_NAME/config.json:18-32 (+ data block :6,:11)
    n_feats=80, n_tokens=10000, embedding_dim=1280, hidden_size=1024, intermediate_size=4096,
    num_attention_heads=16, num_hidden_layers=24, convpos_width=31, convpos_groups=16, convpos_depth=2,
    attention_dropout=0.0, activation_dropout=0.1, hidden_dropout=0.0, solver="euler", sigma_min=1e-4,
)


# ----------------------------------------------------------------------------------------------
# Random-init weights that reproduce the reference constructor's RNG consumption order
# (voicebox.py:153-176 -> networks.py:270-300).  Used so that the GPU box, which has no copy of the
# reference, can rebuild bit-identical "random-init" weights from a seed.
# ----------------------------------------------------------------------------------------------
def init_reference_weights(seed: int = 0, cfg: Optional[dict] = None, perturb: bool = False) -> Dict[str, torch.Tensor]:
    """State-dict (reference key names, fp32, CPU) equal to `torch.manual_seed(seed); Voicebox(**cfg)`.

    perturb=True additionally randomises LayerNorm affine parameters and the weight-norm gains so that
    parity tests exercise them (torch's defaults are 1/0 and g=||v||).  The perturbation is drawn
    from a separate generator and is applied identically to the real reference by oracle/make_golden.py.
    """
    import torch.nn as nn
    cfg = dict(DEFAULT_MODEL_CFG if cfg is None else cfg)
    H, E, Fd, M = cfg["hidden_size"], cfg["embedding_dim"], cfg["intermediate_size"], cfg["n_feats"]
    n_layers, G, Wd = cfg["num_hidden_layers"], cfg["convpos_groups"], cfg["convpos_width"]
    sd: Dict[str, torch.Tensor] = {}
    torch.manual_seed(seed)

    def put(prefix, mod):
        for k, v in mod.state_dict().items():
            sd[f"{prefix}.{k}"] = v.detach().clone()

    # networks.py:291-300, in order
    put("estimator.embed", nn.Embedding(cfg["n_tokens"] + 1, E))
    put("estimator.proj_in", nn.Conv1d(2 * M + E, H, kernel_size=1))
    for i in range(cfg["convpos_depth"]):
        conv = nn.Conv1d(H, H, kernel_size=Wd, padding=Wd // 2, groups=G)   # networks.py:70-76
        v = conv.weight.detach().clone()
        g = v.norm(2, dim=(0, 1), keepdim=True)                              # weight_norm(dim=2): one gain per tap
        sd[f"estimator.pos_conv_embeds.{i}.conv.bias"] = conv.bias.detach().clone()
        sd[f"estimator.pos_conv_embeds.{i}.conv.parametrizations.weight.original0"] = g
        sd[f"estimator.pos_conv_embeds.{i}.conv.parametrizations.weight.original1"] = v
    put("estimator.layer_norm", nn.LayerNorm(H))
    for l in range(n_layers):                                                # networks.py:236-248, :119-147, :214-223
        p = f"estimator.layers.{l}"
        put(p + ".attention.k_proj", nn.Linear(H, H))
        put(p + ".attention.v_proj", nn.Linear(H, H))
        put(p + ".attention.q_proj", nn.Linear(H, H))
        put(p + ".attention.out_proj", nn.Linear(H, H))
        put(p + ".layer_norm", nn.LayerNorm(H))
        put(p + ".feed_forward.intermediate_dense", nn.Linear(H, Fd))
        put(p + ".feed_forward.output_dense", nn.Linear(Fd, H))
        put(p + ".final_layer_norm", nn.LayerNorm(H))
    for m in range(n_layers // 2):
        put(f"estimator.skip_connections_layers.{m}", nn.Linear(2 * H, H))
    put("estimator.proj_out", nn.Conv1d(H, M, kernel_size=1))

    if perturb:
        gen = torch.Generator().manual_seed(seed + 7919)
        for k in sorted(sd):
            if "layer_norm.weight" in k:
                sd[k] = 1.0 + 0.2 * torch.randn(sd[k].shape, generator=gen)
            elif "layer_norm.bias" in k:
                sd[k] = 0.1 * torch.randn(sd[k].shape, generator=gen)
            elif k.endswith("original0"):
                sd[k] = sd[k] * (1.0 + 0.3 * torch.rand(sd[k].shape, generator=gen))
    return sd


def state_dict_checksum(sd: Dict[str, torch.Tensor]) -> float:
    """Cheap order-independent fingerprint used to check that two boxes rebuilt the same weights."""
    acc = 0.0
    for k in sorted(sd):
        acc += float(sd[k].double().abs().sum())
    return acc


# ----------------------------------------------------------------------------------------------
# One estimator call (networks.py:302-374)
# ----------------------------------------------------------------------------------------------
def alibi_slopes(n_heads: int) -> List[float]:
    """networks.py:99-115 (xformers get_slopes)."""
    def pow2(n):
        start = 2 ** (-(2 ** -(math.log2(n) - 3)))
        return [start * start ** i for i in range(n)]
    if math.log2(n_heads).is_integer():
        return pow2(n_heads)
    c = 2 ** math.floor(math.log2(n_heads))
    return pow2(c) + alibi_slopes(2 * c)[0::2][: n_heads - c]


def time_token(t: torch.Tensor, dim: int, scale: float = 1000.0) -> torch.Tensor:
    """networks.py:19-28. t: [B] -> [B, dim] fp32 (sin half then cos half)."""
    half = dim // 2
    f = torch.exp(torch.arange(half, device=t.device).float() * -(math.log(10000) / (half - 1)))
    a = scale * t.reshape(-1, 1).float() * f.unsqueeze(0)
    return torch.cat((a.sin(), a.cos()), dim=-1)


def effective_conv_weight(g: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    """weight_norm(dim=2): W = g * v / ||v||_(0,1)   (networks.py:78-82)."""
    return v * (g / v.norm(2, dim=(0, 1), keepdim=True))


def _ln(x, sd, prefix, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[prefix + ".weight"], sd[prefix + ".bias"], eps)


def estimator_forward(sd: Dict[str, torch.Tensor], cfg: dict, x: torch.Tensor, y: torch.Tensor,
                      cond: torch.Tensor, t: torch.Tensor, lengths: torch.Tensor,
                      taps: Optional[dict] = None) -> torch.Tensor:
    """Velocity estimate. x int64 [B,T]; y, cond fp32 [B,M,T]; t [B] (or [B,1,1]); lengths int64 [B] -> [B,M,T].

    `taps`, when given, is filled with intermediate activations (for kernel-level debugging in tests).
    """
    H, E, nh = cfg["hidden_size"], cfg["embedding_dim"], cfg["num_attention_heads"]
    nl, G, Wd = cfg["num_hidden_layers"], cfg["convpos_groups"], cfg["convpos_width"]
    dh = H // nh
    B, T = x.shape
    S = T + 1
    dev = y.device
    p = "estimator."

    # 1. embed * sqrt(E) ; y ; cond  -> proj_in (1x1 conv == per-frame linear)        (:305-307)
    u = torch.cat([sd[p + "embed.weight"][x] * math.sqrt(E), y.transpose(1, 2), cond.transpose(1, 2)], dim=-1)
    h = F.linear(u, sd[p + "proj_in.weight"][:, :, 0], sd[p + "proj_in.bias"])        # [B,T,H]
    # 2. time token at position 0                                                        (:312-313)
    h = torch.cat([time_token(t.reshape(B), H).unsqueeze(1), h], dim=1)                # [B,S,H]
    # 3. masks; zero padded rows                                                         (:314-315,:333)
    L = lengths.to(dev) + 1
    pos = torch.arange(S, device=dev)
    valid = pos.unsqueeze(0) < L.unsqueeze(1)                                          # [B,S] bool
    m = valid.float().unsqueeze(-1)
    h = h * m
    if taps is not None: taps["h_in"] = h.clone()
    # 4. positional convs + residual                                                     (:343-346,:87-95)
    r = h
    q = h
    for i in range(cfg["convpos_depth"]):
        cp = f"{p}pos_conv_embeds.{i}.conv."
        w = effective_conv_weight(sd[cp + "parametrizations.weight.original0"], sd[cp + "parametrizations.weight.original1"])
        c = F.conv1d(q.transpose(1, 2), w, sd[cp + "bias"], padding=Wd // 2, groups=G)
        if Wd % 2 == 0:
            c = c[:, :, :-1]
        q = F.gelu(c).transpose(1, 2) * m
    h = q + r
    # 5. LN0, mask                                                                       (:348-352)
    h = _ln(h, sd, p + "layer_norm") * m
    if taps is not None: taps["h0"] = h.clone()
    # 6. additive attention term: symmetric ALiBi, key column 0 exempt, masked keys -> finfo.min   (:319-341)
    slopes = torch.tensor(alibi_slopes(nh), device=dev, dtype=torch.float32)
    dist = (pos.unsqueeze(0) - pos.unsqueeze(1)).abs().float()                         # [S,S]
    dist[:, 0] = 0.0
    bias = -slopes.view(1, nh, 1, 1) * dist.view(1, 1, S, S)                           # [1,nh,S,S]
    keyvalid = valid.view(B, 1, 1, S)
    bias = torch.where(keyvalid, bias, torch.full_like(bias[:, :1, :1, :1], torch.finfo(torch.float32).min))

    # 7. layers with LIFO skip stack                                                     (:352-368, :250-266, :149-210)
    skips = [h]
    for l in range(nl):
        lp = f"{p}layers.{l}."
        if l >= nl // 2:
            s = skips.pop()
            sp = f"{p}skip_connections_layers.{l - nl // 2}."
            h = F.linear(torch.cat([h, s], dim=-1), sd[sp + "weight"], sd[sp + "bias"])
        h = h * m
        qh = (F.linear(h, sd[lp + "attention.q_proj.weight"], sd[lp + "attention.q_proj.bias"]) * dh ** -0.5)
        kh = F.linear(h, sd[lp + "attention.k_proj.weight"], sd[lp + "attention.k_proj.bias"])
        vh = F.linear(h, sd[lp + "attention.v_proj.weight"], sd[lp + "attention.v_proj.bias"])
        qh = qh.view(B, S, nh, dh).transpose(1, 2)
        kh = kh.view(B, S, nh, dh).transpose(1, 2)
        vh = vh.view(B, S, nh, dh).transpose(1, 2)
        sc = torch.matmul(qh, kh.transpose(-1, -2)) + bias
        pr = torch.softmax(sc, dim=-1)
        ctx = torch.matmul(pr, vh).transpose(1, 2).reshape(B, S, H)
        a = F.linear(ctx, sd[lp + "attention.out_proj.weight"], sd[lp + "attention.out_proj.bias"])
        h = _ln((h + a) * m, sd, lp + "layer_norm")
        f = F.linear(F.gelu(F.linear(h, sd[lp + "feed_forward.intermediate_dense.weight"], sd[lp + "feed_forward.intermediate_dense.bias"])),
                     sd[lp + "feed_forward.output_dense.weight"], sd[lp + "feed_forward.output_dense.bias"])
        h = _ln((h + f) * m, sd, lp + "final_layer_norm") * m
        if taps is not None: taps[f"layer{l}"] = h.clone()
        if l < nl // 2 - 1:
            skips.append(h)
    assert not skips
    # 8. proj_out, mask, drop the time token                                             (:369-374)
    v = F.linear(h * m, sd[p + "proj_out.weight"][:, :, 0], sd[p + "proj_out.bias"]) * m
    return v[:, 1:, :].transpose(1, 2).contiguous()


# ----------------------------------------------------------------------------------------------
# CFG + solvers (voicebox.py:51-150), batched (the reference's own loop only accepts B == 1)
# ----------------------------------------------------------------------------------------------
def sample_cfg(sd, cfg, x, z, cond, cond_lengths, t, gradient_scale: float, speech_prompt: bool,
               estimator=None) -> torch.Tensor:
    """voicebox.py:51-72.  t: scalar tensor or [B]. Returns the guided velocity [B,M,T]."""
    est = estimator or (lambda *a: estimator_forward(sd, cfg, *a))
    B = z.shape[0]
    t = t.reshape(-1).expand(B) if t.numel() == 1 else t.reshape(B)
    if not speech_prompt:
        cond = torch.zeros_like(cond)                                                  # :56-58 (prompt_length == 0)
    if gradient_scale > 0:
        xx = torch.cat([torch.full_like(x, cfg["n_tokens"]), x], dim=0)                # :61  [uncond ; cond]
        zz = torch.cat([z, z], dim=0)
        cc = torch.cat([torch.zeros_like(cond), cond], dim=0)                          # :63
        tt = torch.cat([t, t], dim=0)
        ll = torch.cat([cond_lengths, cond_lengths], dim=0)
        v = est(xx, zz, cc, tt, ll)
        vu, vc = torch.chunk(v, 2, dim=0)
        return vc + gradient_scale * (vc - vu)                                         # :69-71
    return est(x, z, cond, t, cond_lengths)


def _renoise(z, cond, t, eps, sigma_min, P):
    """voicebox.py:89-92: overwrite the first P frames with the flow-matching interpolant of fresh noise."""
    z = z.clone()
    z[:, :, :P] = ((1 - (1 - sigma_min) * t) * eps + t * cond)[:, :, :P]
    return z


def solve_euler(sd, cfg, x, z, cond, cond_lengths, t_span, gradient_scale, speech_prompt, prompt_len: int,
                noises: Optional[Sequence[torch.Tensor]] = None, trace: Optional[list] = None, estimator=None):
    """voicebox.py:74-99 with an explicit per-step noise list (noises[k] ~ randn_like(cond), step k)."""
    t = t_span[0]
    dt = t_span[1] - t_span[0]
    n = len(t_span) - 1
    for k in range(1, n + 1):
        v = sample_cfg(sd, cfg, x, z, cond, cond_lengths, t, gradient_scale, speech_prompt, estimator)
        z_in = z
        z = z + dt * v
        t = t + dt
        if speech_prompt:
            z = _renoise(z, cond, t, noises[k - 1], cfg["sigma_min"], prompt_len)
        if trace is not None:
            trace.append(dict(z_in=z_in, v=v, z_out=z, t_next=t.clone(), dt=dt.clone()))
        if k < n:
            dt = t_span[k + 1] - t
    return z


def solve_heun(sd, cfg, x, z, cond, cond_lengths, t_span, gradient_scale, speech_prompt, prompt_len: int,
               noises: Optional[Sequence[torch.Tensor]] = None, trace: Optional[list] = None, estimator=None):
    """voicebox.py:101-138. Noise consumption order: predictor noise then corrector noise per step."""
    t = t_span[0]
    dt = t_span[1] - t_span[0]
    n = len(t_span) - 1
    ni = 0
    for k in range(1, n + 1):
        v = sample_cfg(sd, cfg, x, z, cond, cond_lengths, t, gradient_scale, speech_prompt, estimator)
        z_hat = z + dt * v
        t = t + dt
        if speech_prompt:
            z_hat = _renoise(z_hat, cond, t, noises[ni], cfg["sigma_min"], prompt_len); ni += 1
        if k < n:
            v_hat = sample_cfg(sd, cfg, x, z_hat, cond, cond_lengths, t, gradient_scale, speech_prompt, estimator)
            z_hat = z + dt * (v + v_hat) / 2
            if speech_prompt:
                z_hat = _renoise(z_hat, cond, t, noises[ni], cfg["sigma_min"], prompt_len); ni += 1
        z = z_hat
        if trace is not None:
            trace.append(dict(z_out=z, t_next=t.clone()))
        if k < n:
            dt = t_span[k + 1] - t
    return z


def n_noise_draws(solver: str, n_timesteps: int) -> int:
    """How many randn_like(cond) draws the reference makes after the initial z when speech_prompt=True."""
    if solver == "euler":
        return n_timesteps
    n = (n_timesteps + 1) // 2
    return 2 * n - 1


def generate(sd, cfg, x, cond, cond_lengths, n_timesteps, solver="euler", gradient_scale=0.0,
             speech_prompt=False, prompt_lengths=None, generator: Optional[torch.Generator] = None,
             trace: Optional[list] = None):
    """voicebox.py:140-150, batched; draws noise in the reference's order (z first, then per (sub)step)."""
    def randn():
        return torch.randn(cond.shape, generator=generator, dtype=cond.dtype, device=cond.device)
    z = randn()
    n = (n_timesteps + 1) // 2 if solver == "heun" else n_timesteps
    t_span = torch.linspace(0, 1, n + 1, device=cond.device).to(z)
    P = int(prompt_lengths[0]) if (speech_prompt and prompt_lengths is not None) else 0
    noises = [randn() for _ in range(n_noise_draws(solver, n_timesteps))] if speech_prompt else None
    fn = {"euler": solve_euler, "heun": solve_heun}[solver]
    return fn(sd, cfg, x, z, cond, cond_lengths, t_span, gradient_scale, speech_prompt, P, noises, trace)


# FLOP model of SURVEY.md §8(d): algorithmic FLOPs of one estimator call over valid tokens only.
def cfm_forward_loss(sd, cfg, x, mask, x1, lengths, t=None, z=None, estimator=None):
    """Training-loss forward, reference voicebox.py:28-49 (CFM.forward), no autograd.  t [B,1,1] ~ U(0,1) and z ~ N(0,1)
    are drawn with the reference's call order (torch.rand then randn_like) unless given."""
    if t is None:
        t = torch.rand([x1.shape[0], 1, 1], device=x1.device, dtype=x1.dtype)
    if z is None:
        z = torch.randn_like(x1)
    sigma_min = cfg["sigma_min"]
    loss_mask = (torch.arange(0, x.shape[-1], device=lengths.device).unsqueeze(0) < lengths.unsqueeze(1)).float().unsqueeze(1)
    y = (1 - (1 - sigma_min) * t) * z + t * x1
    cond = x1 * mask
    u = x1 - (1 - sigma_min) * z
    est = estimator or (lambda x_, y_, c_, t_, l_: estimator_forward(sd, cfg, x_, y_, c_, t_.reshape(-1), l_))
    ue = est(x, y, cond, t, lengths)
    loss_mask = loss_mask * (1 - mask)
    ue = ue * loss_mask
    u = u * loss_mask
    return F.mse_loss(ue, u, reduction="sum") / (torch.sum(loss_mask) * u.shape[1])


def estimator_flops(token_counts: Sequence[int], cfg: Optional[dict] = None) -> float:
    cfg = cfg or DEFAULT_MODEL_CFG
    H, Fd, E, M = cfg["hidden_size"], cfg["intermediate_size"], cfg["embedding_dim"], cfg["n_feats"]
    nl = cfg["num_hidden_layers"]
    dense = 2 * ((2 * M + E) * H                                       # proj_in
                 + cfg["convpos_depth"] * H * (H // cfg["convpos_groups"]) * cfg["convpos_width"]
                 + nl * (4 * H * H + 2 * H * Fd)                       # q,k,v,o + ffn
                 + (nl // 2) * 2 * H * H                               # skip merges
                 + H * M)                                              # proj_out
    attn = nl * 4 * H                                                  # per token per key
    return float(sum(S * dense + attn * S * S for S in token_counts))
