"""CPU: the host-side mirror carries the reference's parameter names / shapes / init order and API surface."""
import inspect

import torch

from util import CFG, O, weights


def test_state_dict_keys_shapes_and_seeded_init_match_reference_layout():
    import usdm_b200
    torch.manual_seed(0)
    m = usdm_b200.Voicebox(**CFG)
    sd = weights(0, False)              # bit-identical to the reference constructor (asserted in make_golden / live test)
    msd = m.state_dict()
    assert list(msd.keys()) == list(sd.keys()) or set(msd) == set(sd)
    for k in sd:
        assert msd[k].shape == sd[k].shape, k
        assert torch.equal(msd[k], sd[k]), k
    assert m.nparams == sum(v.numel() for v in sd.values())
    assert abs(m.nparams / 1e6 - 345.91) < 0.01          # SURVEY §0 [probed]


def test_api_surface_matches_reference_signatures():
    import usdm_b200
    V = usdm_b200.Voicebox
    assert list(inspect.signature(V.__init__).parameters)[1:] == [
        "n_feats", "n_tokens", "embedding_dim", "hidden_size", "intermediate_size", "num_attention_heads",
        "num_hidden_layers", "convpos_width", "convpos_groups", "convpos_depth", "attention_dropout",
        "activation_dropout", "hidden_dropout", "solver", "sigma_min"]                      # voicebox.py:154-156
    assert list(inspect.signature(V.generate).parameters)[1:] == [
        "x", "cond", "cond_lengths", "n_timesteps", "solver", "gradient_scale", "speech_prompt", "prompt_lengths"]  # :141
    g = inspect.signature(V.generate).parameters
    assert g["solver"].default == "euler" and g["gradient_scale"].default == 0.0 and g["speech_prompt"].default is False
    assert list(inspect.signature(V.sample).parameters)[1:] == ["x", "z", "cond", "cond_lengths", "t", "gradient_scale", "speech_prompt"]
    for name in ("solve_euler", "solve_heun"):
        assert list(inspect.signature(getattr(V, name)).parameters)[1:9] == [
            "x", "z", "cond", "cond_lengths", "t_span", "gradient_scale", "speech_prompt", "prompt_lengths"]
    assert hasattr(V, "from_pretrained")                                                       # PyTorchModelHubMixin


def test_load_state_dict_marks_weights_dirty():
    import usdm_b200
    small = dict(CFG, num_hidden_layers=2)
    m = usdm_b200.Voicebox(**small)
    v0 = m._weights_version
    m.load_state_dict(m.state_dict())
    assert m._weights_version > v0


def test_save_pretrained_from_pretrained_round_trip(tmp_path):
    """the HF PyTorchModelHubMixin layout (config.json + model.safetensors) the reference publishes/loads
    (voicebox.py:10,153; model_util.py:59-62) round-trips through the mirror with the reference's key names, offline."""
    import usdm_b200
    small = dict(CFG, num_hidden_layers=2)
    torch.manual_seed(3)
    m = usdm_b200.Voicebox(**small)
    m.save_pretrained(str(tmp_path))
    import os
    assert os.path.exists(tmp_path / "model.safetensors")
    m2 = usdm_b200.Voicebox.from_pretrained(str(tmp_path), **small)
    sd, sd2 = m.state_dict(), m2.state_dict()
    assert list(sd) == list(sd2)
    for k in sd:
        assert torch.equal(sd[k], sd2[k]), k
    assert any(k.endswith("parametrizations.weight.original1") for k in sd)      # weight-norm keys as in the reference checkpoint
    assert m2._weights_version > 0                                             # the engine will re-pack on first use
