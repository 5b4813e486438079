"""GPU: the tcgen05/TMA bf16 GEMM (usdm_b200/csrc/gemm_tc.cuh) against torch on the same bf16 operands."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


def _run(M, N, K, epi, seed=0):
    from usdm_b200 import engine
    lib = engine.load_library()
    g = torch.Generator(device="cuda").manual_seed(seed)
    a = (torch.randn(M, K, device="cuda", generator=g)).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    of = torch.full((M, N), float("nan"), device="cuda")
    ob = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = lib.tvb_debug_gemm_bf16(0, p(a), p(w), p(bias), p(resid), p(of), p(ob), M, N, K, epi,
                                 C.c_void_p(torch.cuda.current_stream().cuda_stream))
    engine.check(rc)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t() + bias
    return ref, resid, of, ob


@pytest.mark.timeout(300)
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (256, 512, 128), (512, 256, 512), (1024, 1024, 1024),
                                   (768, 3072, 1024), (256, 1024, 4096), (148 * 128 * 2 + 256, 1024, 1024)])
def test_gemm_pair_f32_epilogue(M, N, K):
    """CTA-pair (cta_group::2, M=256) kernel."""
    ref, _, of, _ = _run(M, N, K, 4 | 0x100)
    err = (of - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.timeout(300)
@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (256, 512, 128), (128, 256, 512), (1024, 1024, 1024),
                                   (384, 3072, 1024), (256, 1024, 4096), (148 * 128 * 2 + 128, 1024, 1024)])
def test_gemm_f32_epilogue(M, N, K):
    ref, _, of, _ = _run(M, N, K, 4)
    err = (of - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.timeout(300)
@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (256, 384, 128), (512, 256, 512), (1536, 1024, 1024), (1536, 4096, 1024),
                                   (256, 1024, 4096), (74 * 256 * 2 + 256, 128, 256)])
def test_gemm_pair_bn128_f32_epilogue(M, N, K):
    """CTA-pair kernel with 128-column tiles (the small-batch tiling: cta_group::2, M=256, N=128)."""
    ref, _, of, _ = _run(M, N, K, 4 | 0x400)
    err = (of - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.timeout(300)
@pytest.mark.parametrize("pair", [0, 0x100, 0x400])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_epilogues(epi, pair):
    M, N, K = 512, 1024, 1024
    ref, resid, of, ob = _run(M, N, K, epi | pair, seed=epi + 1)
    if epi == 0:
        want = ref.to(torch.bfloat16).float(); got = ob.float()
    elif epi == 1:
        want = torch.nn.functional.gelu(ref).to(torch.bfloat16).float(); got = ob.float()
    elif epi == 2:
        want = ref + resid; got = of
    else:
        want = ref; got = of
        assert (ob.float() - ref.to(torch.bfloat16).float()).abs().max().item() < 0.05
    err = (got - want).abs().max().item()
    assert err < 0.04, f"epilogue {epi}: max abs err {err}"      # one bf16 ulp at |x| in [4, 8) is 0.03125


@pytest.mark.timeout(300)
@pytest.mark.parametrize("pair", [0x100, 0x400])
@pytest.mark.parametrize("epi", [2, 3, 4])
@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (512, 1024, 1024), (1536, 1024, 2048)])
def test_gemm_direct_fp32_epilogues(M, N, K, epi, pair):
    """The unstaged epilogue (registers <-> global memory through tcgen05.ld.16x256b fragments, flag 0x800) gives the staged epilogue's
    bits: residual add (2), fp32 + bf16 copy (3), plain fp32 (4), 256- and 128-column tiles."""
    ref, resid, of_s, ob_s = _run(M, N, K, epi | pair, seed=epi + 21)
    _, _, of_d, ob_d = _run(M, N, K, epi | pair | 0x800, seed=epi + 21)
    assert torch.equal(of_s, of_d), f"fp32 outputs differ: max {(of_s - of_d).abs().max().item()}"
    if epi == 3:
        assert torch.equal(ob_s.float(), ob_d.float())
    want = ref + resid if epi == 2 else ref
    assert (of_d - want).abs().max().item() < 0.04


@pytest.mark.timeout(300)
@pytest.mark.parametrize("M,N,K", [(512, 256, 64), (512, 512, 128), (1024, 256, 512), (1024, 1024, 1024),
                                   (1536, 3072, 1024), (512, 1024, 4096), (148 * 128 * 2 + 512, 1024, 1024)])
def test_gemm_cluster4_multicast_f32_epilogue(M, N, K):
    """Cluster of two CTA pairs sharing the W tile by TMA multicast (gemm4_tc_kernel)."""
    ref, _, of, _ = _run(M, N, K, 4 | 0x200)
    err = (of - ref).abs().max().item()
    assert err < 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"


@pytest.mark.timeout(300)
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_cluster4_epilogues(epi):
    M, N, K = 1024, 1024, 1024
    ref, resid, of, ob = _run(M, N, K, epi | 0x200, seed=epi + 11)
    if epi == 0:
        want = ref.to(torch.bfloat16).float(); got = ob.float()
    elif epi == 1:
        want = torch.nn.functional.gelu(ref).to(torch.bfloat16).float(); got = ob.float()
    elif epi == 2:
        want = ref + resid; got = of
    else:
        want = ref; got = of
        assert (ob.float() - ref.to(torch.bfloat16).float()).abs().max().item() < 0.05
    err = (got - want).abs().max().item()
    assert err < 0.04, f"epilogue {epi}: max abs err {err}"      # one bf16 ulp at |x| in [4, 8) is 0.03125
