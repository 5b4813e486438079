"""Multi-GPU host logic: utterances are independent (no cross-sequence op anywhere in the estimator), so a batch is
sharded over ranks with NO data-path collective; the only collective is one all-gather of the output mels
(SURVEY.md §8(e)).  Works over any torch.distributed backend (NCCL on B200s, gloo in the CPU tests)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist

_DENSE = 665.55e6     # FLOP / token
_ATTN = 98304.0       # FLOP / token^2


def utterance_cost(frames: int) -> float:
    s = frames + 1
    return s * _DENSE + _ATTN * s * s


def shard_utterances(lengths: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time greedy balance of the per-utterance cost over ranks. Deterministic on every rank."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0.0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += utterance_cost(int(lengths[i]))
    return shards


def micro_batches(indices: Sequence[int], lengths: Sequence[int], max_tokens: int) -> List[List[int]]:
    """Split one rank's utterances (already length-sorted, longest first) into batches of at most max_tokens CFG tokens."""
    out: List[List[int]] = []
    cur: List[int] = []
    tok = 0
    for i in indices:
        t = 2 * (int(lengths[i]) + 1)
        if cur and tok + t > max_tokens:
            out.append(cur)
            cur, tok = [], 0
        cur.append(i)
        tok += t
    if cur:
        out.append(cur)
    return out


def generate_sharded(generate_fn: Callable, x: torch.Tensor, cond: torch.Tensor, lengths: torch.Tensor, *,
                     max_tokens: int = 80000, group: Optional[dist.ProcessGroup] = None, **gen_kwargs) -> torch.Tensor:
    """Every rank calls this with the SAME full inputs (x [N,T], cond [N,M,T], lengths [N]); each rank synthesises its
    shard with `generate_fn` (e.g. Voicebox.generate) and one all-gather returns the full [N,M,T] mel tensor on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N, M, T = cond.shape
    lens = [int(v) for v in lengths.tolist()]
    shards = shard_utterances(lens, world)
    mine = shards[rank]
    n_max = max(len(s) for s in shards)
    local = torch.zeros((n_max, M, T), dtype=torch.float32, device=cond.device)
    pos = 0
    for mb in micro_batches(mine, lens, max_tokens):
        idx = torch.tensor(mb, dtype=torch.long, device=cond.device)
        tmax = max(max(lens[i] for i in mb), 1)
        lens_mb = torch.tensor([lens[i] for i in mb], dtype=torch.int64)          # host tensor: lengths may live on either device
        out = generate_fn(x[idx][:, :tmax].contiguous(), cond[idx][:, :, :tmax].contiguous(), lens_mb, **gen_kwargs)
        local[pos:pos + len(mb), :, :tmax] = out
        pos += len(mb)
    if world == 1:
        gathered = local.unsqueeze(0)
    else:
        flat = torch.empty((world * n_max, M, T), dtype=torch.float32, device=cond.device)   # dim-0 concat: valid for NCCL and gloo
        dist.all_gather_into_tensor(flat, local, group=group)
        gathered = flat.view(world, n_max, M, T)
    full = torch.zeros((N, M, T), dtype=torch.float32, device=cond.device)
    for r, s in enumerate(shards):
        if s:
            full[torch.tensor(s, dtype=torch.long, device=cond.device)] = gathered[r, :len(s)]
    return full


def vocode_sharded(vocode_fn: Callable, mel: torch.Tensor, lengths: torch.Tensor, *, hop: int = 256, max_frames: int = 16000,
                   group: Optional[dist.ProcessGroup] = None) -> torch.Tensor:
    """Vocoder counterpart of generate_sharded: every rank calls this with the SAME mel [N, M, T] / lengths [N]; each rank runs its
    shard through `vocode_fn` (e.g. BigVGAN.forward: [n, M, t] -> [n, 1, t * hop]) and one all-gather returns the waveforms
    [N, 1, T * hop] (zeros past lengths * hop) on every rank.

    The reference vocoder has no length argument (models.py:192): padded frames would be synthesised as content and leak into the
    last frames of a shorter utterance through the convolutions' receptive field.  Utterances are therefore batched only with
    others of exactly the same length, each at its true length -- what the reference's one-utterance call computes."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    N, M, T = mel.shape
    lens = [int(v) for v in lengths.tolist()]
    shards = shard_utterances(lens, world)
    n_max = max(len(s) for s in shards)
    local = torch.zeros((n_max, 1, T * hop), dtype=torch.float32, device=mel.device)
    slot = {u: i for i, u in enumerate(shards[rank])}
    by_len: dict = {}
    for u in shards[rank]:
        by_len.setdefault(lens[u], []).append(u)
    for t, group_idx in sorted(by_len.items()):
        if t <= 0:
            continue
        per_call = max(1, max_frames // t)
        for s in range(0, len(group_idx), per_call):
            mb = group_idx[s:s + per_call]
            idx = torch.tensor(mb, dtype=torch.long, device=mel.device)
            wav = vocode_fn(mel[idx][:, :, :t].contiguous())
            for j, u in enumerate(mb):
                local[slot[u], :, :t * hop] = wav[j]
    if world == 1:
        gathered = local.unsqueeze(0)
    else:
        flat = torch.empty((world * n_max, 1, T * hop), dtype=torch.float32, device=mel.device)
        dist.all_gather_into_tensor(flat, local, group=group)
        gathered = flat.view(world, n_max, 1, T * hop)
    full = torch.zeros((N, 1, T * hop), dtype=torch.float32, device=mel.device)
    for r, s in enumerate(shards):
        if s:
            full[torch.tensor(s, dtype=torch.long, device=mel.device)] = gathered[r, :len(s)]
    return full
