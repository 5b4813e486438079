"""Algorithmic FLOP model of one estimator evaluation: counted over VALID tokens only,
multiply-add = 2, per kernel family, so that measured kernel times can be turned into roofline fractions."""
from __future__ import annotations

from typing import Dict, Sequence


def per_token_flops(cfg: dict) -> Dict[str, float]:
    H, F, E, M = cfg["hidden_size"], cfg["intermediate_size"], cfg["embedding_dim"], cfg["n_feats"]
    nl = cfg["num_hidden_layers"]
    return {
        "gemm": 2.0 * (nl * (4 * H * H + 2 * H * F) + (nl // 2) * 2 * H * H),       # q,k,v,o + FFN + skip merges
        "posconv": 2.0 * cfg["convpos_depth"] * H * (H // cfg["convpos_groups"]) * cfg["convpos_width"],
        "frame_proj": 2.0 * (2 * M + E) * H,                                        # proj_in (reference op count)
        "output": 2.0 * H * M,                                                      # proj_out
    }


def estimator_flops(token_counts: Sequence[int], cfg: dict) -> Dict[str, float]:
    """FLOPs of one estimator call over sequences with `token_counts` tokens (frames + 1 each), by kernel family."""
    pt = per_token_flops(cfg)
    tokens = float(sum(token_counts))
    out = {k: v * tokens for k, v in pt.items()}
    out["attention"] = float(cfg["num_hidden_layers"] * 4 * cfg["hidden_size"]) * float(sum(s * s for s in token_counts))
    out["total"] = sum(out.values())
    return out
