"""GPU: the fp32 validation path (TVB_PRECISION_FP32_VALIDATE) against the golden outputs of the real reference
and against the oracle; tolerance 1e-5 relative on the velocity (BASELINE.json north_star)."""
import pytest
import torch

from util import CFG, O, golden, max_abs, model, rel_l2, weights

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.mark.timeout(900)
def test_estimator_ragged_vs_reference_golden():
    g = golden("estimator_ragged.pt")
    m = model(0, True, "fp32")
    v = m.estimator(g["x"].cuda(), g["y"].cuda(), g["cond"].cuda(), g["t"], g["lengths"])
    torch.cuda.synchronize()
    err = rel_l2(v, g["v"])
    print("fp32 estimator rel-L2 vs reference:", err)
    assert err < TOL
    for b, L in enumerate(g["lengths"].tolist()):
        assert float(v[b, :, L:].abs().sum()) == 0.0


@pytest.mark.timeout(900)
def test_euler_trace_vs_reference_golden():
    g = golden("euler_trace_b2.pt")
    m = model(0, True, "fp32")
    out, vel = m.solve_euler(g["x"].cuda(), g["z0"].cuda(), g["cond"].cuda(), g["lengths"], g["t_span"], g["gradient_scale"], True,
                             torch.tensor([g["prompt_len"]]), return_velocities=True, noises=g["noises"])
    torch.cuda.synchronize()
    # free-running: errors compound slightly over the 4 steps
    for k in range(len(g["v"])):
        e = rel_l2(vel[k], g["v"][k])
        print(f"step {k} velocity rel-L2 {e:.3e}")
        assert e < 5 * TOL
    assert rel_l2(out, g["final"]) < 5 * TOL


@pytest.mark.timeout(900)
@pytest.mark.parametrize("solver", ["euler", "heun"])
@pytest.mark.parametrize("speech_prompt", [True, False])
@pytest.mark.parametrize("gs", [1.0, 0.0])
def test_generate_b1_vs_reference_generate(solver, speech_prompt, gs):
    """Same semantics as reference generate(): CFG on/off, prompt on/off, both solvers (noise supplied explicitly so the
    CPU-generator stream of the golden run is reproduced)."""
    g = golden("generate_b1.pt")
    m = model(0, True, "fp32")
    gen = torch.Generator().manual_seed(g["seed"])
    cond = g["cond"]
    z = torch.randn(cond.shape, generator=gen)
    n = g["n_timesteps"]
    n = (n + 1) // 2 if solver == "heun" else n
    t_span = torch.linspace(0, 1, n + 1)
    noises = [torch.randn(cond.shape, generator=gen) for _ in range(O.n_noise_draws(solver, g["n_timesteps"]))] if speech_prompt else None
    fn = m.solve_euler if solver == "euler" else m.solve_heun
    out = fn(g["x"].cuda(), z.cuda(), cond.cuda(), g["lengths"], t_span, gs, speech_prompt, torch.tensor([g["prompt_len"]]), noises=noises)
    torch.cuda.synchronize()
    ref = g["outs"][(solver, speech_prompt, gs)]
    e = rel_l2(out, ref)
    print(solver, speech_prompt, gs, "final rel-L2", e, "max-abs", max_abs(out, ref))
    assert e < 5 * TOL


@pytest.mark.timeout(900)
def test_config1_teacher_forced_velocity():
    g = golden("config1_euler32.pt")
    m = model(0, False, "fp32")
    t_span = torch.linspace(0, 1, 33)
    for idx, k in enumerate(g["steps"]):
        t = t_span[0]; dt = t_span[1] - t_span[0]
        for s in range(1, k + 1):
            t = t + dt; dt = t_span[s + 1] - t
        v, _ = m.sample(g["x"].cuda(), g["z_in"][idx].cuda(), g["cond"].cuda(), g["lengths"], t, 1.0, True)
        torch.cuda.synchronize()
        e = rel_l2(v, g["v"][idx])
        print(f"config1 step {k}: fp32 velocity rel-L2 {e:.3e}")
        assert e < TOL
