// Shared device helpers for the token-Voicebox sm_100a kernels: mbarrier / TMA / tcgen05 PTX wrappers.
// Everything here is inline PTX for sm_100a; there is no fallback path for other architectures.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>

namespace tvb {

// ------------------------------------------------------------------------------------------
// Host side: per-DEVICE one-time set-up.  cudaFuncSetAttribute (the > 48 KB dynamic shared-memory opt-in) and occupancy answers
// belong to the device that was current when they were made; one process may hold handles on several GPUs (tvb_create takes a
// device index), so the "already done" state is one bit / one value per device ordinal, updated atomically (the set-up itself is
// idempotent, two threads racing on the same device only repeat it).
// ------------------------------------------------------------------------------------------
constexpr int kMaxDevices = 64;
inline int current_device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return d & (kMaxDevices - 1);
}
struct PerDeviceOnce {
    std::atomic<uint64_t> bits{0};
    bool done() const { return (bits.load(std::memory_order_acquire) >> current_device_slot()) & 1ull; }
    void mark() { bits.fetch_or(1ull << current_device_slot(), std::memory_order_release); }
};
struct PerDeviceInt {                                  // 0 = not computed yet on that device
    std::atomic<int> v[kMaxDevices];
    PerDeviceInt() { for (auto& x : v) x.store(0, std::memory_order_relaxed); }
    int get() const { return v[current_device_slot()].load(std::memory_order_acquire); }
    void set(int x) { v[current_device_slot()].store(x, std::memory_order_release); }
};

// ------------------------------------------------------------------------------------------
// small utilities
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float gelu_erf(float x) {          // F.gelu default (networks.py:45)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// ------------------------------------------------------------------------------------------
// programmatic dependent launch: a kernel launched with the programmatic-stream-serialization attribute may start (and run its
// set-up: barrier init, TMEM allocation, descriptor prefetch) while the previous kernel of the stream drains; pdl_wait() blocks
// until that kernel has completed and flushed.  pdl_launch_dependents() lets the NEXT kernel do the same with this one.
// Both are no-ops for kernels launched the ordinary way.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {     // generic-proxy smem writes -> visible to async proxy (UMMA/TMA)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: a waiting thread is parked by the hardware until the phase completes (or the hint
// expires) instead of re-issuing the poll every few cycles - fewer issue slots and less power spent by the many warps that
// wait on barriers in these warp-specialised kernels (the chip runs at its power cap, so idle polling costs clock).
// TVB_MBAR_SPIN (tools/attn_lab.cu builds only): 1 = try_wait without the hint, 2 = test_wait polling.  Measured on the attention kernels
// (profiles/r02/attn_lab_wait*.txt): 225.8 / 230.5 us against 228.5 us with the hint -- the wake-up is not what the per-block chain waits for.
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if defined(TVB_MBAR_SPIN) && TVB_MBAR_SPIN == 2
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#elif defined(TVB_MBAR_SPIN)
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
#endif
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}

// ------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor), 2-D tiled loads with mbarrier completion
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c_inner,
                                            int32_t c_outer) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c_inner), "r"(c_outer)
        : "memory");
}

// Pull a 2-D box into the L2 only (no shared-memory destination, no completion to wait for): a later TMA load of the same box then
// sees L2 latency instead of HBM latency under load.
__device__ __forceinline__ void tma_prefetch_l2_2d(const CUtensorMap* map, int32_t c_inner, int32_t c_outer) {
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];" ::"l"(reinterpret_cast<uint64_t>(map)), "r"(c_inner),
                 "r"(c_outer)
                 : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05: TMEM allocation, MMA issue, commit, TMEM loads
// ------------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result) {   // executed by ONE full warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t tmem_base) {    // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by one thread.
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem desc]: A is 128 lanes x (K/2) 32-bit columns of packed bf16 pairs.
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrive when all previously issued tcgen05.mma of this thread have completed (implies fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}


// ------------------------------------------------------------------------------------------
// CTA-pair (cta_group::2) variants: two CTAs of one cluster on adjacent SMs share one M=256 MMA.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// mbarrier arrive on the barrier at the same smem offset in CTA `cta` of this cluster.
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
    asm volatile(
        "{\n\t.reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}\n" ::"r"(smem_u32(bar)),
        "r"(cta)
        : "memory");
}
// TMA load into THIS CTA's smem whose completion bytes are credited to the pair LEADER's (even CTA's) mbarrier.
__device__ __forceinline__ void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c_inner,
                                                 int32_t c_outer) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;   // clear the peer bit -> leader CTA's barrier
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar), "r"(c_inner), "r"(c_outer)
        : "memory");
}
// Same, multicast: the box lands at the same smem offset in every CTA of `cta_mask`, and in each of them the bytes are credited
// to the barrier (same offset) of that CTA's pair leader.
__device__ __forceinline__ void tma_load_2d_2cta_mc(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c_inner,
                                                    int32_t c_outer, uint16_t cta_mask) {
    const uint32_t mbar = smem_u32(bar) & 0xFEFFFFFFu;
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(mbar), "r"(c_inner), "r"(c_outer), "h"(cta_mask)
        : "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc_2cta(uint32_t* smem_result) {   // one full warp in EACH CTA of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "n"(kCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc_2cta(uint32_t tmem_base) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma_ss_2cta(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// arrive (once all prior MMAs retire) on the barrier at this smem offset in every CTA of `cta_mask`.
__device__ __forceinline__ void umma_commit_2cta(uint64_t* bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                     smem_u32(bar)),
                 "h"(cta_mask)
                 : "memory");
}


// TMA store of a 2-D box from this CTA's smem to global (bulk async-group completion).
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* smem_src, int32_t c_inner, int32_t c_outer) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(smem_u32(smem_src)), "r"(c_inner), "r"(c_outer)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() {   // at most N committed groups still reading their smem source
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// erf-GELU (F.gelu default) with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7): ~14 FP ops + 2 MUFU instead
// of the ~30-instruction erff; used in the GEMM epilogue where the result is rounded to bf16 anyway.
__device__ __forceinline__ float gelu_erf_fast(float x) {
    const float z = x * 0.70710678118654752440f;
    const float az = fabsf(z);
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, az, 1.0f)));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float y = fmaf(-p * t, e, 1.0f);          // erf(|z|)
    const float h = 0.5f * x;
    return fmaf(h, copysignf(y, z), h);
}

// The same function on two elements with packed fp32x2 arithmetic (FMUL2 / FFMA2: one issue slot per pair; every rounding is the scalar
// version's, so the results are bit-identical): 12 packed FP + 4 MUFU + 4 logic instructions per pair instead of 24 + 4 + 4.  The GELU
// epilogue of the FFN-in GEMM is issue-bound (ncu: ~6000 issue slots per 8192-cycle tile on two warps per scheduler).
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t f2_mul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ uint64_t f2_dup(float x) { return f2_pack(x, x); }
__device__ __forceinline__ void gelu_erf_fast2(float& x0, float& x1) {
    const uint64_t x2 = f2_pack(x0, x1);
    const uint64_t z2 = f2_mul(x2, f2_dup(0.70710678118654752440f));
    const uint64_t az2 = z2 & 0x7fffffff7fffffffull;
    float d0, d1, a0, a1;
    f2_unpack(f2_fma(f2_dup(0.3275911f), az2, f2_dup(1.0f)), d0, d1);
    f2_unpack(f2_mul(f2_mul(z2, z2), f2_dup(-1.4426950408889634f)), a0, a1);
    float t0, t1, e0, e1;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t0) : "f"(d0));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t1) : "f"(d1));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
    const uint64_t t2 = f2_pack(t0, t1), e2 = f2_pack(e0, e1);
    // -p by the negated coefficients: fma(-a, t, -b) == -fma(a, t, b) exactly
    uint64_t np = f2_fma(f2_dup(-1.061405429f), t2, f2_dup(1.453152027f));
    np = f2_fma(np, t2, f2_dup(-1.421413741f));
    np = f2_fma(np, t2, f2_dup(0.284496736f));
    np = f2_fma(np, t2, f2_dup(-0.254829592f));
    const uint64_t y2 = f2_fma(f2_mul(np, t2), e2, f2_dup(1.0f));          // erf(|z|) = 1 - p t e
    const uint64_t h2 = f2_mul(f2_dup(0.5f), x2);
    const uint64_t s2 = (y2 & 0x7fffffff7fffffffull) | (z2 & 0x8000000080000000ull);   // copysign(y, z)
    f2_unpack(f2_fma(h2, s2, h2), x0, x1);
}

// Instruction descriptor, kind::f16: bf16 x bf16 -> fp32 (layout: cute/arch/mma_sm100_desc.hpp InstrDescriptor).
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t M, uint32_t N, bool a_mn_major = false, bool b_mn_major = false) {
    return (1u << 4)                         // c_format = F32
           | (1u << 7)                       // a_format = BF16
           | (1u << 10)                      // b_format = BF16
           | ((a_mn_major ? 1u : 0u) << 15)  // a_major
           | ((b_mn_major ? 1u : 0u) << 16)  // b_major
           | ((N >> 3) << 17)                // n_dim
           | ((M >> 4) << 24);               // m_dim
}
// Shared-memory matrix descriptor, 128-byte swizzle (bf16: 64 elements per swizzled row), version 1 (Blackwell).
//  K-major operand : rows are MN, 128 B of K per row; SBO = 1024 B between 8-row groups; LBO field = 1 (unused).
//  MN-major operand: rows are K, 128 B (64 elements) of MN per row; SBO = 1024 B between 8-K groups; LBO = stride
//                    between 64-element MN blocks (unused when MN extent is 64).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= 1ull << 46;   // descriptor version
    d |= 2ull << 61;   // SWIZZLE_128B
    return d;
}

// TMEM -> registers: this warp's 32 lanes x 32 consecutive fp32 columns (lane i of the warp reads TMEM lane base+i).
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// Same wait, but the 32 destination registers of the load it completes are tied to it as in/out operands, so that no use of
// them can be scheduled above the wait when other work is interleaved between the load and the wait.
__device__ __forceinline__ void tmem_ld_wait_dep(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]),
                   "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]),
                   "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :
                 : "memory");
}
// 16 TMEM lanes x 32 columns in the MMA C-fragment layout (tools/tmem_layout.cu): thread t holds, per 8-column block i, r[4i], r[4i+1] =
// row (t / 4), columns 8i + 2 (t % 4) + {0, 1} and r[4i+2], r[4i+3] = row (t / 4) + 8, the same columns.  taddr's lane field is the first of
// the 16 lanes (inside the warp's own 32-lane quadrant).
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait_dep16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

}  // namespace tvb
