"""Generate tests/golden/*.pt from the REAL reference (run in the authoring container only).

    python -m oracle.make_golden [--big]

TEST INFRASTRUCTURE ONLY. The fixtures pin oracle/voicebox_oracle.py (and through it the CUDA path) to outputs
of the unmodified reference at /root/reference, which cannot travel to the GPU box.  Weights are NOT stored:
they are re-created on any box by oracle.voicebox_oracle.init_reference_weights(seed, perturb) (asserted
bit-identical to the reference constructor here, and fingerprinted by `weights_checksum`).
"""
from __future__ import annotations

import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as R          # noqa: E402
from oracle import voicebox_oracle as O      # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _inputs(seed, B, T, lengths, prompt_len):
    g = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 10000, (B, T), generator=g)
    y = torch.randn(B, 80, T, generator=g)
    cond = torch.randn(B, 80, T, generator=g)
    cond[:, :, prompt_len:] = 0
    L = torch.tensor(lengths, dtype=torch.int64)
    return x, y, cond, L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--big", action="store_true", help="also write the config-1 sized fixture (minutes of CPU)")
    args = ap.parse_args()
    os.makedirs(GOLD, exist_ok=True)
    cfg = dict(O.DEFAULT_MODEL_CFG)
    torch.set_grad_enabled(False)

    # --- weights: default init (seed 0) must equal the reference constructor bit for bit
    model = R.build_reference(cfg, seed=0)
    sd0 = O.init_reference_weights(0, cfg)
    rsd = model.state_dict()
    assert set(rsd) == set(sd0) and all(torch.equal(rsd[k], sd0[k]) for k in sd0), "init order drifted"
    sdp = O.init_reference_weights(0, cfg, perturb=True)
    meta = dict(cfg=cfg, checksum_default=O.state_dict_checksum(sd0), checksum_perturbed=O.state_dict_checksum(sdp),
                torch_version=torch.__version__)

    # --- 1. estimator, ragged batch, perturbed weights
    model.load_state_dict(sdp)
    est = R.reference_estimator(model)
    x, y, cond, L = _inputs(11, 3, 50, [50, 29, 1], 50)
    t = torch.tensor([0.3, 0.7, 0.11])
    v = est(x, y, cond, t, L)
    torch.save(dict(meta=meta, weights=dict(seed=0, perturb=True), x=x, y=y, cond=cond, t=t, lengths=L, v=v),
               os.path.join(GOLD, "estimator_ragged.pt"))

    # --- 2. reference generate() (B=1 only), both solvers, with/without prompt, global-RNG seeded
    x, _, cond, L = _inputs(12, 1, 40, [40], 16)
    outs = {}
    for solver in ("euler", "heun"):
        for sp in (True, False):
            for gs in (1.0, 0.0):
                torch.manual_seed(5)
                outs[(solver, sp, gs)] = model.generate(x, cond, L, 5, solver=solver, gradient_scale=gs, speech_prompt=sp,
                                                        prompt_lengths=torch.tensor([16]))
    torch.save(dict(meta=meta, weights=dict(seed=0, perturb=True), x=x, cond=cond, lengths=L, n_timesteps=5, seed=5,
                    prompt_len=16, outs=outs), os.path.join(GOLD, "generate_b1.pt"))

    # --- 3. batched ragged Euler trace through the reference estimator (batched driver = oracle loop)
    x, _, cond, L = _inputs(13, 2, 48, [48, 33], 12)
    g = torch.Generator().manual_seed(21)
    z0 = torch.randn(cond.shape, generator=g)
    noises = [torch.randn(cond.shape, generator=g) for _ in range(4)]
    t_span = torch.linspace(0, 1, 5)
    trace = []
    zT = O.solve_euler(sdp, cfg, x, z0, cond, L, t_span, 0.7, True, 12, noises, trace, estimator=est)
    torch.save(dict(meta=meta, weights=dict(seed=0, perturb=True), x=x, cond=cond, lengths=L, z0=z0, noises=noises,
                    t_span=t_span, gradient_scale=0.7, prompt_len=12, final=zT,
                    v=[s["v"] for s in trace], z_out=[s["z_out"] for s in trace]),
               os.path.join(GOLD, "euler_trace_b2.pt"))

    # --- 4. config 1 (BASELINE.json configs[0]): B=1, 430 target + 272 prompt frames, 32 Euler steps, CFG 1.0
    if args.big:
        model.load_state_dict(sd0)
        P, Tt = 272, 430
        T = P + Tt
        torch.manual_seed(1234)
        x = torch.randint(0, 10000, (1, T))
        cond = torch.zeros(1, 80, T)
        cond[:, :, :P] = torch.randn(1, 80, P)        # synthetic N(0,1) prompt mel (wav front-end is row f-1, later)
        L = torch.tensor([T])
        torch.manual_seed(1)
        z0 = torch.randn_like(cond)
        noises = [torch.randn_like(cond) for _ in range(32)]
        trace = []
        zT = O.solve_euler(sd0, cfg, x, z0, cond, L, torch.linspace(0, 1, 33), 1.0, True, P, noises, trace, estimator=est)
        torch.manual_seed(1)
        ref = model.generate(x, cond, L, 32, solver="euler", gradient_scale=1.0, speech_prompt=True,
                             prompt_lengths=torch.tensor([P]))
        assert torch.equal(ref, zT), "batched driver must be bit-identical to generate() at B=1"
        keep = (0, 15, 31)
        torch.save(dict(meta=meta, weights=dict(seed=0, perturb=False), x=x, cond=cond, lengths=L, prompt_len=P,
                        noise_seed=1, final=zT, steps=keep,
                        z_in=[trace[k]["z_in"] for k in keep], v=[trace[k]["v"] for k in keep]),
                   os.path.join(GOLD, "config1_euler32.pt"))
    print("golden written to", GOLD)


if __name__ == "__main__":
    main()
