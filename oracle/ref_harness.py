"""Harness around the REAL reference (only usable where /root/reference exists, i.e. the authoring container).

TEST INFRASTRUCTURE ONLY.  Nothing here is imported by the product, by `-m gpu` tests, smoke() or bench.py.
It imports `voicebox.model.voicebox.Voicebox` from /root/reference/src/decoder unmodified and adds the two
shims SURVEY.md §8(c) lists: a no-op `.cuda()` on CPU (networks.py:319 hard-codes it) and a batched driver,
because the reference's own `sample()` only accepts B == 1 (voicebox.py:52).
"""
from __future__ import annotations

import os
import sys

import torch

REF_ROOT = os.environ.get("TVB_REFERENCE_ROOT", "/root/reference")
REF_DECODER = os.path.join(REF_ROOT, "src", "decoder")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_DECODER, "voicebox", "model", "voicebox.py"))


def load_reference_class():
    if not reference_available():
        raise RuntimeError("reference tree not present")
    if REF_DECODER not in sys.path:
        sys.path.insert(0, REF_DECODER)
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self      # networks.py:319 shim (CPU only)
    from voicebox.model.voicebox import Voicebox             # noqa: E402
    return Voicebox


def build_reference(cfg: dict, state_dict=None, seed: int = 0):
    """Reference model in eval mode; either seeded default init or a given state dict."""
    Voicebox = load_reference_class()
    kw = {k: v for k, v in cfg.items() if k not in ("n_feats", "n_tokens")}
    torch.manual_seed(seed)
    model = Voicebox(cfg["n_feats"], cfg["n_tokens"], **kw).eval()
    if state_dict is not None:
        model.load_state_dict(state_dict)
    return model


def reference_estimator(model):
    """Callable with the oracle's estimator signature that runs the reference Transformer.forward."""
    def est(x, y, cond, t, lengths):
        with torch.no_grad():
            return model.estimator(x, y, cond, t.reshape(-1, 1, 1), lengths)
    return est
