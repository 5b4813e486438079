"""GPU: sampler-level behaviour through the public API (generate / solve_* / sample), bf16 path, oracle as checker."""
import pytest
import torch

from util import CFG, O, golden, max_abs, model, rel_l2, weights

pytestmark = pytest.mark.gpu


def _inputs(seed, B, T, lengths, P):
    gen = torch.Generator().manual_seed(seed)
    x = torch.randint(0, 10000, (B, T), generator=gen)
    cond = torch.randn(B, 80, T, generator=gen)
    cond[:, :, P:] = 0
    return x, cond, torch.tensor(lengths)


@pytest.mark.timeout(900)
def test_generate_matches_oracle_noise_stream_and_shapes():
    """generate() draws z and the per-step prompt noise from torch's CUDA generator in the reference's order."""
    torch.backends.cuda.matmul.allow_tf32 = False
    x, cond, L = _inputs(5, 2, 96, [96, 57], 24)
    m = model(0, True, "bf16")
    torch.manual_seed(77)
    out = m.generate(x.cuda(), cond.cuda(), L, 4, solver="euler", gradient_scale=0.7, speech_prompt=True, prompt_lengths=torch.tensor([24]))
    torch.cuda.synchronize()
    assert out.shape == cond.shape and out.dtype == torch.float32
    sd = {k: v.cuda() for k, v in weights(0, True).items()}
    torch.manual_seed(77)
    cc = cond.cuda()
    z0 = torch.randn_like(cc)
    noises = [torch.randn_like(cc) for _ in range(4)]
    with torch.no_grad():
        ref = O.solve_euler(sd, CFG, x.cuda(), z0, cc, L.cuda(), torch.linspace(0, 1, 5, device="cuda"), 0.7, True, 24, noises)
    e = rel_l2(out, ref)
    print("generate vs oracle (same CUDA noise stream): rel-L2", e, "max-abs", max_abs(out, ref))
    assert e < 2e-2
    assert m.generate(x.cuda(), cond.cuda(), L, 4, solver="rk4") is None      # reference: unknown solver -> None


@pytest.mark.timeout(900)
def test_ragged_batch_equals_per_utterance_and_padding_is_ignored():
    """ragged == trimmed per-utterance (reference: 7.7e-7) and padded-region garbage has no effect (reference: 0)."""
    x, cond, L = _inputs(6, 3, 200, [200, 131, 64], 200)
    gen = torch.Generator().manual_seed(9)
    y = torch.randn(3, 80, 200, generator=gen)
    t = torch.tensor([0.4, 0.4, 0.4])
    m = model(0, True, "bf16")
    v = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L).clone()
    # garbage in the padded region must not change anything, bit for bit
    x2, y2, c2 = x.clone(), y.clone(), cond.clone()
    for b, l in enumerate(L.tolist()):
        x2[b, l:] = 1234; y2[b, :, l:] = 1e3; c2[b, :, l:] = -1e3
    v2 = m.estimator(x2.cuda(), y2.cuda(), c2.cuda(), t, L)
    torch.cuda.synchronize()
    assert torch.equal(v, v2)
    for b, l in enumerate(L.tolist()):
        vb = m.estimator(x[b:b + 1, :l].cuda(), y[b:b + 1, :, :l].contiguous().cuda(), cond[b:b + 1, :, :l].contiguous().cuda(), t[b:b + 1], torch.tensor([l]))
        torch.cuda.synchronize()
        e = rel_l2(vb, v[b:b + 1, :, :l])
        print(f"utt {b}: ragged vs alone rel-L2 {e:.3e}")
        assert e < 1e-6, "each sequence is computed independently of its batch neighbours"


@pytest.mark.timeout(900)
def test_heun_and_cfg_off_and_prompt_off_vs_oracle():
    torch.backends.cuda.matmul.allow_tf32 = False
    x, cond, L = _inputs(7, 2, 80, [80, 41], 20)
    sd = {k: v.cuda() for k, v in weights(0, True).items()}
    m = model(0, True, "bf16")
    for solver, gs, sp in [("heun", 1.0, True), ("heun", 0.0, False), ("euler", 0.0, True), ("euler", 1.5, False)]:
        gen = torch.Generator().manual_seed(31)
        z0 = torch.randn(cond.shape, generator=gen)
        nt = 5
        n = (nt + 1) // 2 if solver == "heun" else nt
        noises = [torch.randn(cond.shape, generator=gen) for _ in range(O.n_noise_draws(solver, nt))]
        t_span = torch.linspace(0, 1, n + 1)
        fn = m.solve_euler if solver == "euler" else m.solve_heun
        out = fn(x.cuda(), z0.cuda(), cond.cuda(), L, t_span, gs, sp, torch.tensor([20]), noises=noises if sp else None)
        ofn = O.solve_euler if solver == "euler" else O.solve_heun
        with torch.no_grad():
            ref = ofn(sd, CFG, x.cuda(), z0.cuda(), cond.cuda(), L.cuda(), t_span.cuda(), gs, sp, 20, [n_.cuda() for n_ in noises])
        torch.cuda.synchronize()
        e = rel_l2(out, ref)
        print(solver, gs, sp, "rel-L2", e)
        assert e < 2e-2
        # padded frames beyond the prompt: v == 0 there, so z0 is carried through unchanged, as in the reference
        assert torch.equal(out[1, :, 41:].cpu(), z0[1, :, 41:]) and rel_l2(ref[1, :, 41:], z0[1, :, 41:]) < 1e-6


@pytest.mark.timeout(900)
def test_edge_cases_zero_length_and_single_frame():
    x, cond, L = _inputs(8, 3, 32, [0, 1, 32], 32)
    gen = torch.Generator().manual_seed(2)
    y = torch.randn(3, 80, 32, generator=gen)
    m = model(0, True, "bf16")
    v = m.estimator(x.cuda(), y.cuda(), cond.cuda(), torch.tensor([0.2, 0.2, 0.2]), L)
    torch.cuda.synchronize()
    assert torch.isfinite(v).all()
    assert float(v[0].abs().sum()) == 0.0 and float(v[1, :, 1:].abs().sum()) == 0.0
    with pytest.raises(ValueError):
        m.estimator(x.cuda(), y.cuda(), cond.cuda(), torch.tensor([0.2, 0.2, 0.2]), torch.tensor([0, 1, 33]))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_cfm_forward_loss_value_vs_oracle(precision, tol):
    """Voicebox.forward(x, mask, x1, lengths) = the reference's flow-matching loss (voicebox.py:28-49) with the same
    RNG call order; per-sequence t goes through tvb_estimator_forward."""
    torch.backends.cuda.matmul.allow_tf32 = False
    gen = torch.Generator().manual_seed(13)
    B, T = 3, 150
    x = torch.randint(0, 10000, (B, T), generator=gen).cuda()
    x1 = torch.randn(B, 80, T, generator=gen).cuda()
    L = torch.tensor([150, 90, 33]).cuda()
    mask = torch.zeros(B, 80, T).cuda()
    mask[0, :, 20:60] = 1                      # span-masked context frames (UnitMelBatchCollate, data_util.py:94-147)
    mask[1, :, :30] = 1
    sd = {k: v.cuda() for k, v in weights(0, True).items()}
    m = model(0, True, precision)
    torch.manual_seed(99)
    got = m(x, mask, x1, L)
    torch.manual_seed(99)
    with torch.no_grad():
        want = O.cfm_forward_loss(sd, CFG, x, mask, x1, L)
    torch.cuda.synchronize()
    rel = abs(float(got) - float(want)) / abs(float(want))
    print(f"CFM.forward loss [{precision}]: {float(got):.6f} vs oracle {float(want):.6f} (rel {rel:.2e})")
    assert got.dim() == 0 and torch.isfinite(got)
    assert rel < tol


@pytest.mark.timeout(900)
@pytest.mark.parametrize("solver", ["euler", "heun"])
def test_graph_replay_of_small_batches_is_bit_identical_to_direct_launches(solver):
    """tvb_solve replays the estimator pass of small batches as a CUDA graph (option "graph"); same kernels, same order ->
    the mel must be bit-identical to direct launches, also when the same handle is reused with another batch in between."""
    x, cond, L = _inputs(21, 2, 120, [120, 77], 30)
    m = model(0, True, "bf16")
    gen = torch.Generator().manual_seed(4)
    z0 = torch.randn(cond.shape, generator=gen)
    nt = 6
    n = (nt + 1) // 2 if solver == "heun" else nt
    noises = [torch.randn(cond.shape, generator=gen) for _ in range(O.n_noise_draws(solver, nt))]
    t_span = torch.linspace(0, 1, n + 1)
    fn = m.solve_euler if solver == "euler" else m.solve_heun
    run = lambda: fn(x.cuda(), z0.cuda(), cond.cuda(), L, t_span, 1.0, True, torch.tensor([30]), noises=noises).clone()
    m.set_option("graph", 0)
    direct = run()
    m.set_option("graph", 1)
    g1 = run()                                   # captures during this solve
    x2, cond2, L2 = _inputs(22, 1, 64, [64], 16)
    m.generate(x2.cuda(), cond2.cuda(), L2, 2, gradient_scale=1.0, speech_prompt=True, prompt_lengths=torch.tensor([16]))   # another plan
    g2 = run()                                   # re-plans, re-captures
    g3 = run()                                   # same plan as the previous call: graphs reused
    torch.cuda.synchronize()
    assert torch.equal(direct, g1) and torch.equal(direct, g2) and torch.equal(direct, g3)


@pytest.mark.timeout(900)
def test_generate_sharded_single_rank_on_gpu_inputs():
    """parallel.generate_sharded with device-resident inputs (world 1: no process group), micro-batched by token budget."""
    from usdm_b200 import parallel
    x, cond, L = _inputs(23, 5, 90, [90, 40, 66, 12, 90], 20)
    m = model(0, True, "bf16")
    def gen_fn(xx, cc, ll, **kw):          # deterministic stand-in for generate(): fixed initial state, 2 Euler steps, no prompt / CFG
        z = torch.full(cc.shape, 0.1, device=cc.device)
        return m.solve_euler(xx, z, cc, ll, torch.linspace(0, 1, 3), 0.0, False, None)
    full = parallel.generate_sharded(gen_fn, x.cuda(), cond.cuda(), L, max_tokens=300)
    torch.cuda.synchronize()
    assert full.shape == (5, 80, 90) and torch.isfinite(full).all()
    for b, l in enumerate(L.tolist()):
        alone = gen_fn(x[b:b + 1, :l].cuda(), cond[b:b + 1, :, :l].contiguous().cuda(), torch.tensor([l]))
        assert rel_l2(full[b:b + 1, :, :l], alone) < 1e-6


@pytest.mark.timeout(900)
def test_unit_ids_outside_the_table_are_reported_not_dereferenced():
    """The reference's nn.Embedding asserts device-side on ids outside [0, n_tokens] (networks.py:305).  Here the lookup stays in
    range (no illegal address), the call's output is flagged invalid, and the flag surfaces without any device synchronisation
    inside the library: through check_async() once the stream has been synchronised, or at the next sampling call."""
    x, cond, L = _inputs(12, 2, 48, [48, 30], 16)
    m = model(0, True, "bf16")
    y = torch.randn(2, 80, 48, generator=torch.Generator().manual_seed(1))
    t = torch.tensor([0.3, 0.3])
    good = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)
    torch.cuda.synchronize()
    m.check_async()                                            # nothing pending
    for bad_id in (10001, -1, 2 ** 40):
        xb = x.clone()
        xb[1, 7] = bad_id
        out = m.estimator(xb.cuda(), y.cuda(), cond.cuda(), t, L)
        torch.cuda.synchronize()                               # would raise on an illegal address
        assert torch.isfinite(out).all()
        with pytest.raises(ValueError, match="unit ids outside"):
            m.check_async()
        m.check_async()                                        # reported once, then cleared
        m.estimator(xb.cuda(), y.cuda(), cond.cuda(), t, L)
        torch.cuda.synchronize()
        with pytest.raises(ValueError, match="unit ids outside"):      # ... or by the next call on the handle
            m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)
    xu = x.clone()
    xu[0, :] = 10000                                           # the unconditional id itself is a valid table row
    m.estimator(xu.cuda(), y.cuda(), cond.cuda(), t, L)
    torch.cuda.synchronize()
    m.check_async()
    again = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)
    assert torch.equal(good, again)


@pytest.mark.timeout(900)
def test_sequences_beyond_the_supported_length_are_refused():
    """Key positions reach the attention kernel as two bf16 terms that are exact below 2^15: longer inputs are an error, not a wrong bias."""
    m = model(0, True, "bf16")
    T = 32768
    x = torch.zeros(1, T, dtype=torch.int64, device="cuda")
    y = torch.zeros(1, 80, T, device="cuda")
    with pytest.raises(ValueError, match="not supported"):
        m.estimator(x, y, y, torch.tensor([0.5]), torch.tensor([T]))
    v = m.estimator(x[:, :40], y[:, :, :40].contiguous(), y[:, :, :40].contiguous(), torch.tensor([0.5]), torch.tensor([40]))   # the handle is still usable
    torch.cuda.synchronize()
    assert torch.isfinite(v).all()


@pytest.mark.timeout(900)
def test_large_logits_switch_the_model_to_the_running_maximum_kernel():
    """The default attention kernel (v10) supports logits inside a weight-dependent bound and the library reports a violation
    asynchronously.  The host mirror checks the FIRST call after a weight upload synchronously and, when the bound is exceeded,
    switches the model to the running-maximum kernel (attn_v = 8) and repeats the call - so a checkpoint with large logits works
    out of the box, and the result is the reference's."""
    import usdm_b200
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.clone() for k, v in weights(0, True).items()}
    for name in ("q_proj", "k_proj"):
        sd[f"estimator.layers.3.attention.{name}.weight"] *= 12.0
        sd[f"estimator.layers.3.attention.{name}.bias"] *= 12.0
    m = usdm_b200.Voicebox(**CFG)
    m.load_state_dict(sd)
    m.eval()
    x, cond, L = _inputs(21, 2, 160, [160, 97], 32)
    y = torch.randn(2, 80, 160, generator=torch.Generator().manual_seed(3))
    t = torch.tensor([0.6, 0.6])
    with pytest.warns(RuntimeWarning, match="running-maximum"):
        v = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)
    assert m._options["attn_v"] == 8
    m.check_async()                                            # nothing pending: the repeated call was clean
    sdc = {k: w.cuda() for k, w in sd.items()}
    with torch.no_grad():
        ref = O.estimator_forward(sdc, CFG, x.cuda(), y.cuda(), cond.cuda(), t.cuda(), L.cuda())
    e = rel_l2(v, ref)
    print("large-logit weights, v8 after the automatic switch: rel-L2 vs oracle", e)
    assert e < 3e-2
    v2 = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)   # later calls: no synchronisation, same kernel, same bits
    assert torch.equal(v, v2)


@pytest.mark.timeout(900)
def test_large_row_offsets_switch_the_model_to_the_separate_layernorm():
    """The LayerNorm folded into the FFN-in GEMM multiplies the bf16-rounded RAW rows and removes mean * colsum afterwards: a row whose
    mean is many standard deviations loses precision there (the reference rounds after normalising).  The library flags such rows
    (|mean| * rstd > 8); the host mirror's first-call check then switches the model to ln1_fold = 0 and repeats the call."""
    import usdm_b200
    torch.backends.cuda.matmul.allow_tf32 = False
    sd = {k: v.clone() for k, v in weights(0, True).items()}
    sd["estimator.layers.2.final_layer_norm.bias"] += 40.0          # every hidden channel of layer 3's input shifted by 40
    m = usdm_b200.Voicebox(**CFG)
    m.load_state_dict(sd)
    m.eval()
    x, cond, L = _inputs(22, 2, 160, [160, 97], 32)
    y = torch.randn(2, 80, 160, generator=torch.Generator().manual_seed(4))
    t = torch.tensor([0.6, 0.6])
    with pytest.warns(RuntimeWarning, match="ln1_fold = 0"):
        v = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)
    assert m._options["ln1_fold"] == 0      # (the offset also inflates |q|, |k|: the attention kernel may have been switched as well)
    m.check_async()
    sdc = {k: w.cuda() for k, w in sd.items()}
    with torch.no_grad():
        ref = O.estimator_forward(sdc, CFG, x.cuda(), y.cuda(), cond.cuda(), t.cuda(), L.cuda())
    e = rel_l2(v, ref)
    print("row-offset weights, ln1_fold = 0 after the automatic switch: rel-L2 vs oracle", e)
    assert e < 2e-2
    m.set_option("ln1_fold", 1)                                      # what the switch avoided
    vf = m.estimator(x.cuda(), y.cuda(), cond.cuda(), t, L)
    torch.cuda.synchronize()
    print("  the same with the folded LayerNorm forced on: rel-L2 vs oracle", rel_l2(vf, ref))
    with pytest.raises(ValueError, match="ln1_fold = 0"):
        m.check_async()


@pytest.mark.timeout(900)
def test_second_device_in_the_same_process():
    """Kernel attributes (the > 48 KB dynamic shared-memory opt-in) are per device: a handle created on another GPU of the same
    process must set them up again.  Needs two devices (gpurun --gpus 2); skipped otherwise."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    import usdm_b200
    x, cond, L = _inputs(13, 2, 64, [64, 40], 16)
    y = torch.randn(2, 80, 64, generator=torch.Generator().manual_seed(2))
    t = torch.tensor([0.4, 0.4])
    m0 = model(0, True, "bf16")
    v0 = m0.estimator(x.cuda(0), y.cuda(0), cond.cuda(0), t, L)
    m1 = usdm_b200.Voicebox(**CFG)
    m1.load_state_dict(weights(0, True))
    with torch.cuda.device(1):
        v1 = m1.estimator(x.cuda(1), y.cuda(1), cond.cuda(1), t, L)
        z = m1.generate(x.cuda(1), cond.cuda(1), L, 2, solver="heun", gradient_scale=1.0, speech_prompt=True, prompt_lengths=torch.tensor([16]))
        torch.cuda.synchronize(1)
    torch.cuda.synchronize(0)
    assert torch.isfinite(z).all()
    assert torch.equal(v0.cpu(), v1.cpu())                     # same bits on both devices
    from oracle import vocoder_oracle as V
    voc = usdm_b200.BigVGAN(V.SMALL_CFG)
    voc.load_state_dict(V.init_vocoder_weights(11, V.SMALL_CFG))
    mel = torch.randn(1, 80, 24, generator=torch.Generator().manual_seed(3))
    w0 = voc.cuda(0)(mel.cuda(0)).cpu()
    with torch.cuda.device(1):
        w1 = voc.cuda(1)(mel.cuda(1)).cpu()
    assert torch.equal(w0, w1)
