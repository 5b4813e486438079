"""CPU: the oracle restatement reproduces the REAL reference's outputs stored in tests/golden (made by
oracle/make_golden.py in the authoring container), and — where /root/reference is mounted — the live reference."""
import pytest
import torch

from util import CFG, O, golden, max_abs, rel_l2, weights

TOL = 2e-5   # fp32 vs fp32, different op order (measured ~1.4e-6)


def _check_meta(g):
    w = g["weights"]
    sd = weights(w["seed"], w["perturb"])
    want = g["meta"]["checksum_perturbed" if w["perturb"] else "checksum_default"]
    got = O.state_dict_checksum(sd)
    assert abs(got - want) <= 1e-9 * abs(want), "random-init weights differ from the ones the golden was made with"
    return sd


def test_estimator_ragged_matches_reference_output():
    g = golden("estimator_ragged.pt")
    sd = _check_meta(g)
    with torch.no_grad():
        v = O.estimator_forward(sd, CFG, g["x"], g["y"], g["cond"], g["t"], g["lengths"])
    assert rel_l2(v, g["v"]) < TOL
    for b, L in enumerate(g["lengths"].tolist()):        # padded frames exactly zero (networks.py:372)
        assert float(v[b, :, L:].abs().sum()) == 0.0


@pytest.mark.parametrize("solver", ["euler", "heun"])
@pytest.mark.parametrize("speech_prompt", [True, False])
@pytest.mark.parametrize("gs", [1.0, 0.0])
def test_generate_b1_matches_reference_generate(solver, speech_prompt, gs):
    g = golden("generate_b1.pt")
    sd = _check_meta(g)
    gen = torch.Generator().manual_seed(g["seed"])       # same stream as torch.manual_seed(seed) + randn_like
    with torch.no_grad():
        out = O.generate(sd, CFG, g["x"], g["cond"], g["lengths"], g["n_timesteps"], solver=solver, gradient_scale=gs,
                         speech_prompt=speech_prompt, prompt_lengths=torch.tensor([g["prompt_len"]]), generator=gen)
    ref = g["outs"][(solver, speech_prompt, gs)]
    assert out.shape == ref.shape
    assert rel_l2(out, ref) < TOL, max_abs(out, ref)


def test_batched_euler_trace_matches_reference_estimator_loop():
    g = golden("euler_trace_b2.pt")
    sd = _check_meta(g)
    trace = []
    with torch.no_grad():
        zT = O.solve_euler(sd, CFG, g["x"], g["z0"], g["cond"], g["lengths"], g["t_span"], g["gradient_scale"], True,
                           g["prompt_len"], g["noises"], trace)
    for k in range(len(trace)):
        assert rel_l2(trace[k]["v"], g["v"][k]) < TOL
        assert rel_l2(trace[k]["z_out"], g["z_out"][k]) < TOL
    assert rel_l2(zT, g["final"]) < TOL


def test_config1_teacher_forced_steps():
    """BASELINE configs[0] (B=1, 430 target + 272 prompt frames): per-step velocity, teacher-forced on the reference z."""
    g = golden("config1_euler32.pt")
    sd = _check_meta(g)
    t_span = torch.linspace(0, 1, 33)
    for idx, k in enumerate(g["steps"][:2]):
        # reference t at step k: accumulated fp32 adds of dt (voicebox.py:87,96)
        t = t_span[0]
        dt = t_span[1] - t_span[0]
        for s in range(1, k + 1):
            t = t + dt
            dt = t_span[s + 1] - t
        with torch.no_grad():
            v = O.sample_cfg(sd, CFG, g["x"], g["z_in"][idx], g["cond"], g["lengths"], t, 1.0, True)
        assert rel_l2(v, g["v"][idx]) < TOL


def test_flop_model_matches_survey():
    assert abs(O.estimator_flops([1]) - (665.55e6 + 98304)) < 0.01e6
    assert abs(O.estimator_flops([862] * 32) / 1e12 - 20.70) < 0.02


def test_live_reference_if_mounted():
    from oracle import ref_harness as R
    if not R.reference_available():
        pytest.skip("reference tree not mounted on this box")
    model = R.build_reference(CFG, seed=0)
    sd = weights(0, False)
    rsd = model.state_dict()
    assert set(rsd) == set(sd) and all(torch.equal(rsd[k], sd[k]) for k in sd)
    g = golden("estimator_ragged.pt")
    model.load_state_dict(weights(0, True))
    v = R.reference_estimator(model)(g["x"], g["y"], g["cond"], g["t"], g["lengths"])
    assert torch.equal(v, g["v"]) or rel_l2(v, g["v"]) < 1e-6
