"""GPU: the tcgen05 attention kernels alone (usdm_b200/csrc/attn4_tc.cuh, attn_tc.cuh) through tvb_debug_attention, against
fp32 torch math on the SAME bf16 q|k|v rows: softmax_j(q_i.k_j - slope_h |i-j| [j != 0]) with keys j >= len masked
(reference networks.py:319-341, :174-193).  The only rounding that differs from the fp32 math is P -> bf16 before PV."""
import ctypes as C
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

H, NH = 1024, 16
LOG2E = 1.4426950408889634


def _slopes():
    start = 2 ** (-2 ** -(math.log2(NH) - 3))
    return torch.tensor([start * start ** i for i in range(NH)], dtype=torch.float32)


def _reference(qkv, starts, lens, slopes):
    out = torch.zeros(qkv.shape[0], H, device=qkv.device)
    for st, L in zip(starts, lens):
        blk = qkv[st:st + L].float()
        q = blk[:, :H].view(L, NH, 64).transpose(0, 1)          # carries 0.125 * log2(e)
        k = blk[:, H:2 * H].view(L, NH, 64).transpose(0, 1)
        v = blk[:, 2 * H:].view(L, NH, 64).transpose(0, 1)
        s = torch.matmul(q, k.transpose(1, 2)) / LOG2E
        idx = torch.arange(L, device=qkv.device)
        dist = (idx[:, None] - idx[None, :]).abs().float()
        bias = -slopes.to(qkv.device)[:, None, None] * dist[None]
        bias[:, :, 0] = 0.0
        p = torch.softmax(s + bias, dim=-1)
        out[st:st + L] = torch.matmul(p, v).transpose(0, 1).reshape(L, H)
    return out


def _run(lens, version, poly, seed=0, scale=1.0, spikes=(), return_ctx=False):
    from usdm_b200 import engine
    lib = engine.load_library()
    g = torch.Generator(device="cuda").manual_seed(seed)
    starts, o = [], 0
    for L in lens:
        starts.append(o)
        o += L + 16
    rows = (o + 255) // 256 * 256
    qkv = torch.randn(rows, 3 * H, device="cuda", generator=g) * scale
    qkv[:, :H] *= 0.125 * LOG2E
    for seq, j, height in spikes:       # one key whose score towers `height` (log2 units) above the rest for every query of the sequence
        st, L = starts[seq], lens[seq]
        qkv[st:st + L, :H] += 1.0 / 8.0                               # + e per head, |e| = 1 over the 64 head dims
        qkv[st + j, H:2 * H] += height / 8.0                          # + height * e  ->  q.k gains ~height
    qkv = qkv.to(torch.bfloat16)
    ctx = torch.zeros(rows, H, device="cuda", dtype=torch.bfloat16)
    slopes = _slopes().cuda()
    s_arr = (C.c_int32 * len(lens))(*starts)
    l_arr = (C.c_int32 * len(lens))(*lens)
    rc = lib.tvb_debug_attention(0, C.c_void_p(qkv.data_ptr()), C.c_void_p(ctx.data_ptr()), s_arr, l_arr, len(lens), rows, H, NH,
                                 C.c_void_p(slopes.data_ptr()), version, poly, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    engine.check(rc)
    torch.cuda.synchronize()
    ref = _reference(qkv, starts, lens, slopes)
    valid = torch.zeros(rows, dtype=torch.bool, device="cuda")
    for st, L in zip(starts, lens):
        valid[st:st + L] = True
    assert float(ctx[~valid].float().abs().sum()) == 0.0          # rows outside every sequence are never written
    a, b = ctx[valid].float(), ref[valid]
    if return_ctx:
        return float((a - b).norm() / b.norm()), ctx[valid]
    return float((a - b).norm() / b.norm()), float((a - b).abs().max())


@pytest.mark.timeout(600)
@pytest.mark.parametrize("version,poly", [(10, 0), (10, 2), (10, 3), (10, 4), (100, 4), (9, 0), (9, 2), (9, 3), (9, 4), (90, 3), (8, 0), (8, 2), (8, 3), (8, 4), (7, 0), (7, 2), (7, 3), (7, 4), (6, 0), (6, 2), (6, 3), (6, 4), (5, 0), (5, 2), (5, 3), (5, 4), (4, 0), (4, 2), (4, 3), (4, 4), (3, 0), (2, 0)])
def test_attention_kernel_vs_fp32_math(version, poly):
    # lengths cover: 1 key, < 1 block, exact block multiples, odd tails, one / two query halves, the configs[1] length
    lens = [1134, 1, 127, 128, 129, 256, 257, 384, 700, 389, 2, 513]
    rel, mx = _run(lens, version, poly)
    print(f"attention v{version} poly={poly}: rel-L2 {rel:.3e}, max-abs {mx:.3e}")
    assert rel < 4e-3


@pytest.mark.timeout(600)
@pytest.mark.parametrize("version,poly", [(4, 0), (4, 3), (5, 0), (5, 3), (6, 0), (6, 3), (7, 0), (7, 3), (8, 0), (8, 3)])
def test_attention_peaky_scores_exercise_rescale(version, poly):
    """Large-magnitude q.k (std ~ 40 in the log2 domain) so that block maxima jump by more than 2^16 and the lazy O rescale runs."""
    rel, mx = _run([1134, 640, 300], version, poly, seed=1, scale=3.0)
    print(f"attention v{version} poly={poly}, peaky: rel-L2 {rel:.3e}, max-abs {mx:.3e}")
    assert rel < 6e-3


@pytest.mark.timeout(600)
@pytest.mark.parametrize("poly", [0, 2, 3, 4])
def test_attention_v8_single_key_spike_beyond_the_polynomial_range(poly):
    """The fast path's overflow guard: one key 200 log2 units above everything else, sitting in a LATER key block at a column the
    FMA-pipe polynomial handles (columns 4, 5, 10, 11, ... of each 32-column chunk for poly = 3) and, in other sequences, at a MUFU
    column and in the first block.  2^200 is beyond both the polynomial's exponent range and fp32; the guard must send the block
    to the careful path, after which the softmax is one-hot on the spike."""
    lens = [700, 640, 513, 300]
    spikes = [(0, 2 * 128 + 4, 200.0), (1, 3 * 128 + 37, 200.0), (2, 1 * 128 + 1, 200.0), (3, 7, 200.0)]
    rel, mx = _run(lens, 8, poly, seed=5, scale=0.3, spikes=spikes)
    print(f"attention v8 poly={poly}, single-key spikes: rel-L2 {rel:.3e}, max-abs {mx:.3e}")
    assert rel < 6e-3


@pytest.mark.timeout(600)
@pytest.mark.parametrize("version", [4, 5, 6, 7, 8, 9, 90, 10, 100])
def test_attention_long_form(version):
    """configs[2]-like length (30 s + prompt): 23 key blocks, 12 query tiles per head."""
    rel, mx = _run([2856, 2584], version, 0, seed=2)
    print(f"attention v{version} long-form: rel-L2 {rel:.3e}, max-abs {mx:.3e}")
    assert rel < 4e-3


@pytest.mark.timeout(600)
@pytest.mark.parametrize("version", [6, 8, 9, 10])
def test_attention_persistent_many_items_per_cta(version):
    """More work items than SMs and very uneven ones (1-key, one-half, two-half tiles interleaved), so that every CTA of the
    persistent kernel crosses several item boundaries with different half counts and heads."""
    lens = [1134, 3, 700, 129, 1134, 64, 389, 257, 900, 1, 513, 1134, 200, 1100, 77, 640] * 2
    rel, mx = _run(lens, version, 3, seed=4)
    print(f"attention v{version} persistent, 32 ragged sequences: rel-L2 {rel:.3e}, max-abs {mx:.3e}")
    assert rel < 4e-3


@pytest.mark.timeout(600)
@pytest.mark.parametrize("poly", [0, 3])
@pytest.mark.parametrize("scale", [0.3, 1.0, 2.0])
@pytest.mark.parametrize("v", [9, 10])
def test_attention_v9_fixed_reference_and_window(v, poly, scale):
    """attn9_tc.cuh: (a) the fixed softmax reference s(i, key 0) + shift: scale 2.0 gives a logit bound 2 max|q| max|k| ~ 160 in
    the log2 domain, i.e. a non-zero shift, scale 0.3 a bound of a few units; (b) the ALiBi window (version 9) against the same kernel
    with every key block processed (version 90): skipped keys are below 2^-40 of their row sum, so the two may differ by an odd
    last-place rounding at most."""
    lens = [1134, 700, 2856, 129, 513, 300]
    rel_w, ctx_w = _run(lens, v, poly, seed=7, scale=scale, return_ctx=True)
    rel_f, ctx_f = _run(lens, 10 * v, poly, seed=7, scale=scale, return_ctx=True)
    diff = (ctx_w.float() - ctx_f.float()).abs()
    changed = int((diff > 0).sum())
    print(f"attention v{v} poly={poly} scale={scale}: rel-L2 window {rel_w:.3e} / full {rel_f:.3e}; elements that differ between the two: {changed} of {diff.numel()}, max {float(diff.max()):.3e}")
    assert rel_w < 6e-3 and rel_f < 6e-3
    assert changed <= diff.numel() * 1e-4
    assert float((diff / ctx_f.float().abs().clamp_min(1e-3)).max()) < 1e-2          # one bf16 ulp of the output


@pytest.mark.timeout(600)
@pytest.mark.parametrize("v", [9, 10])
def test_attention_v9_key0_spike_and_far_key_spike(v):
    """Scores far from the reference in both directions: a spike ON the time token (every other exponent becomes very negative) and
    a spike on a far key inside the logit bound (exponent close to the top of the range)."""
    lens = [700, 640]
    spikes = [(0, 0, 60.0), (1, 3 * 128 + 37, 60.0)]
    rel, mx = _run(lens, v, 3, seed=5, scale=0.3, spikes=spikes)
    print(f"attention v{v} spikes: rel-L2 {rel:.3e}, max-abs {mx:.3e}")
    assert rel < 6e-3
