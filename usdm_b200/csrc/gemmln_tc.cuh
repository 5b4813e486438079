// Residual GEMM with the following LayerNorm fused into its epilogue:
//     x = A W^T + bias + resid            (fp32, never written)
//     y = LayerNorm(x; gamma, beta)       -> out_f32 [M, 1024] (the post-LN residual stream) and out_bf16 (next GEMM operand)
// for the two N = hidden = 1024 projections of an encoder layer (attention out + LN1, FFN out + LN2; reference
// networks.py:208 + :256-259 and :230 + :260-263).  Replaces gemm2_tc_kernel<EPI_RESID_*> + k_layernorm, i.e. removes one
// 151 MB fp32 read (plus the write-back stall it caused, profiles/r01/ncu_ln_summary.txt) and one launch per LayerNorm.
//
// A LayerNorm row spans the whole N = 1024, four 256-column tiles: a CLUSTER OF 8 CTAs = 4 CTA pairs computes the four tiles of one
// 256-row block with tcgen05.mma.cta_group::2 exactly like gemm2_tc_kernel (each pair: TMA producer warp, MMA warp in the
// leader, 8 epilogue warps, 2 TMEM accumulator stages), and the epilogues exchange per-row partial (sum, sum of squares)
// through distributed shared memory:
//   pass 1  thread == row: x for its 128 columns into registers (TMEM load + bias + residual tile staged by TMA), partial sums;
//           the accumulator stage is released here already;
//   xchg    every thread stores its partial into the statistics buffer of the 4 CTAs that hold the same rows (st.shared::cluster)
//           and one lane per warp arrives (release.cluster) on their barrier; 4 CTAs x 8 warps arrivals complete a phase;
//   pass 2  mean / rstd from the 8 partials, normalise from registers, fp32 tile -> swizzled staging -> TMA store, bf16 -> global.
// On a 148-SM B200 16 such clusters are resident (128 SMs); the chip is power-capped in these GEMMs, so the
// idle SMs cost clock headroom rather than throughput.
#pragma once
#include "gemm_tc.cuh"

namespace tvb {

constexpr int GL_STAGES = 3;
constexpr int GL_EPI_BUFS = 3;
constexpr int GL_SMEM_EPI = GL_STAGES * G2_STAGE_BYTES;                                   //  96 KB operand ring
constexpr int GL_SMEM_STAT = GL_SMEM_EPI + G2_EPI_WARPS * GL_EPI_BUFS * G2_EPI_TILE;      // + 96 KB staging
constexpr int GL_STAT_BYTES = 2 * 4 * 2 * 128 * 8;                                        // [parity][source pair][column half][row] float2
constexpr int GL_SMEM_BAR = GL_SMEM_STAT + GL_STAT_BYTES;
constexpr int GL_SMEM_BYTES = GL_SMEM_BAR + 1024 + 512;
constexpr int GL_N = 1024;
constexpr int GL_THREADS = 384;    // warpgroup 0: TMA / MMA / TMEM-alloc warps (+1 idle), warpgroups 1-2: the 8 epilogue warps (setmaxnreg 232)

struct GemmLnParams {
    int M, K;                 // M % 256 == 0, K % 64 == 0; N is GL_N
    const float* bias;        // [N]
    const float* ln_g;        // [N] LayerNorm applied to the output
    const float* ln_b;
    float eps;
    __nv_bfloat16* out_bf16;  // [M, N]
};
struct GemmLnMaps {
    CUtensorMap a, b;         // A [M, K] box 128 x 64; W [N, K] box 128 x 64 (bf16, 128-B swizzle)
    CUtensorMap out, resid;   // fp32 [M, N] box 32 x 32
};

__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
    return r;
}
__device__ __forceinline__ void st_cluster_f32x2(uint32_t raddr, float a, float b) {
    asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(raddr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_arrive_release_cluster(uint32_t raddr_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_acquire_cluster(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u)
            : "memory");
    }
}

__global__ void __cluster_dims__(8, 1, 1) __launch_bounds__(GL_THREADS, 1)
gemmln_tc_kernel(const __grid_constant__ GemmLnMaps tm, const GemmLnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GL_SMEM_BAR);
    uint64_t* empty_bar = full_bar + GL_STAGES;
    uint64_t* tfull_bar = empty_bar + GL_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* stat_bar = tempty_bar + 2;                          // [2] by tile parity, 4 CTAs x 8 warps arrivals
    uint64_t* rbar = stat_bar + 2;                                // [8 warps][GL_EPI_BUFS] residual-tile arrival
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + GL_EPI_BUFS * G2_EPI_WARPS);
    float2* stat = reinterpret_cast<float2*>(smem + GL_SMEM_STAT);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t cta = crank & 1;                               // position inside the pair, 0 = leader
    const int pair_id = static_cast<int>(crank >> 1);             // 0..3 = which 256 of the 1024 output columns
    const int num_tiles = p.M / 256;                              // 256-row blocks, one per cluster step
    const int num_kb = p.K / GEMM_BK;
    const int cluster_id = blockIdx.x >> 3, num_clusters = gridDim.x >> 3;
    const uint16_t pair_mask = static_cast<uint16_t>(0x3u << (2 * pair_id));

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tm.a); tma_prefetch_desc(&tm.b); tma_prefetch_desc(&tm.out); tma_prefetch_desc(&tm.resid);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < GL_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * G2_EPI_WARPS); mbar_init(&stat_bar[s], 4 * G2_EPI_WARPS);
        }
        for (int s = 0; s < GL_EPI_BUFS * G2_EPI_WARPS; ++s) mbar_init(&rbar[s], 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc_2cta<512>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp_idx == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = tile * 256 + static_cast<int>(cta) * 128;
                const int n0 = pair_id * 256 + static_cast<int>(cta) * 128;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * G2_STAGE_BYTES;
                    if (cta == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);   // bytes of BOTH CTAs of the pair
                    tma_load_2d_2cta(sa, &tm.a, &full_bar[stage], kb * GEMM_BK, m0);
                    tma_load_2d_2cta(sa + G2_A_BYTES, &tm.b, &full_bar[stage], kb * GEMM_BK, n0);
                    if (++stage == GL_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (cta == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(256, 256);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int as = it & 1;
                mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + G2_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) umma_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2cta(&empty_bar[stage], pair_mask);
                    if (++stage == GL_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&tfull_bar[as], pair_mask);
            }
        }
    }
    } else {
        // ------------------------------------------ epilogue ------------------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
        const int ew = warp_idx - 4;
        const int quad = warp_idx & 3;
        const int chalf = ew >> 2;                                // which 128 of this CTA's 256 accumulator columns
        uint8_t* stg = smem + GL_SMEM_EPI + ew * GL_EPI_BUFS * G2_EPI_TILE;
        uint64_t* rb = rbar + ew * GL_EPI_BUFS;
        const int r7 = lane & 7;
        const int rloc = quad * 32 + lane;                        // row inside this CTA's 128 rows
        const int col0 = pair_id * 256 + chalf * 128;             // first output column of this thread
        uint32_t rph0 = 0, rph12 = 0;                             // residual-barrier parities: buffer 0 is used twice per tile
        // statistics slots this thread writes (same offset in every destination CTA) and the barrier addresses
        const uint32_t stat_local = smem_u32(stat);
        const uint32_t statbar_local = smem_u32(stat_bar);
        auto issue_resid = [&](int tile, int c, int buf) {        // one elected lane
            mbar_arrive_expect_tx(&rb[buf], G2_EPI_TILE);
            tma_load_2d(stg + buf * G2_EPI_TILE, &tm.resid, &rb[buf], col0 + c * 32, tile * 256 + static_cast<int>(cta) * 128 + quad * 32);
        };
        if (lane == 0 && cluster_id < num_tiles)
            for (int c = 0; c < 3; ++c) issue_resid(cluster_id, c, c);
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int as = it & 1;
            const int row0 = tile * 256 + static_cast<int>(cta) * 128 + quad * 32;
            mbar_wait(&tfull_bar[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 256 + chalf * 128;
            float v[128];
            float s1 = 0.f, s2 = 0.f;
            // ---- pass 1: x = acc + bias + resid into registers, partial row statistics
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int buf = (c == 3) ? 0 : c;
                uint32_t r[32];
                tmem_ld_32x32(t_acc + c * 32, r);
                tmem_ld_wait();
                mbar_wait(&rb[buf], buf == 0 ? rph0 : rph12);    // residual tile landed
                if (buf == 0) rph0 ^= 1;
                const uint8_t* rowp = stg + buf * G2_EPI_TILE + lane * 128;
                const float4* b4 = reinterpret_cast<const float4*>(p.bias + col0 + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 rr = *reinterpret_cast<const float4*>(rowp + ((j ^ r7) << 4));
                    const float4 b = __ldg(b4 + j);
                    const float x0 = __uint_as_float(r[4 * j]) + b.x + rr.x, x1 = __uint_as_float(r[4 * j + 1]) + b.y + rr.y;
                    const float x2 = __uint_as_float(r[4 * j + 2]) + b.z + rr.z, x3 = __uint_as_float(r[4 * j + 3]) + b.w + rr.w;
                    v[c * 32 + 4 * j] = x0; v[c * 32 + 4 * j + 1] = x1; v[c * 32 + 4 * j + 2] = x2; v[c * 32 + 4 * j + 3] = x3;
                    s1 += (x0 + x1) + (x2 + x3);
                    s2 = fmaf(x0, x0, fmaf(x1, x1, fmaf(x2, x2, fmaf(x3, x3, s2))));
                }
                if (c == 0) {                                     // buffer 0 is free again: fetch the fourth residual chunk into it
                    __syncwarp();
                    if (lane == 0) issue_resid(tile, 3, 0);
                }
            }
            rph12 ^= 1;
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], static_cast<uint32_t>(2 * pair_id));   // accumulator drained
            // ---- exchange: my partial to the 4 CTAs holding these rows (same position in their pair), then wait for all 8 partials
            {
                const uint32_t slot = stat_local + static_cast<uint32_t>((((as * 4 + pair_id) * 2 + chalf) * 128 + rloc) * 8);
                const uint32_t sbar = statbar_local + static_cast<uint32_t>(as * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) st_cluster_f32x2(mapa_u32(slot, static_cast<uint32_t>(2 * q) + cta), s1, s2);
                // ONE arrival per warp and destination (1024 per-thread remote arrivals per tile serialised on the barrier and cost
                // ~13 us per tile): the warp barrier orders the lanes' stores before lane 0's release.cluster arrive (cumulative)
                __syncwarp();
                if (lane == 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) mbar_arrive_release_cluster(mapa_u32(sbar, static_cast<uint32_t>(2 * q) + cta));
                }
            }
            mbar_wait_acquire_cluster(&stat_bar[as], (it >> 1) & 1);
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const float2 ps = stat[(as * 8 + q) * 128 + rloc];
                t1 += ps.x; t2 += ps.y;
            }
            const float mean = t1 * (1.0f / GL_N);
            const float rstd = rsqrtf(fmaxf(t2 * (1.0f / GL_N) - mean * mean, 0.f) + p.eps);
            // ---- pass 2: normalise from registers; fp32 tile through swizzled staging + TMA store, bf16 copy straight to global
            __nv_bfloat16* orow = p.out_bf16 + static_cast<size_t>(row0 + lane) * GL_N + col0;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const int buf = (c == 3) ? 0 : c;
                if (c == 3) {                                     // buffer 0 holds the store of chunk 0: it must have been read out
                    if (lane == 0) tma_store_wait_read<2>();
                    __syncwarp();
                }
                uint8_t* rowp = stg + buf * G2_EPI_TILE + lane * 128;
                const float4* g4 = reinterpret_cast<const float4*>(p.ln_g + col0 + c * 32);
                const float4* be4 = reinterpret_cast<const float4*>(p.ln_b + col0 + c * 32);
                uint32_t pk[16];
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 g = __ldg(g4 + j), be = __ldg(be4 + j);
                    const float y0 = fmaf((v[c * 32 + 4 * j] - mean) * rstd, g.x, be.x);
                    const float y1 = fmaf((v[c * 32 + 4 * j + 1] - mean) * rstd, g.y, be.y);
                    const float y2 = fmaf((v[c * 32 + 4 * j + 2] - mean) * rstd, g.z, be.z);
                    const float y3 = fmaf((v[c * 32 + 4 * j + 3] - mean) * rstd, g.w, be.w);
                    *reinterpret_cast<float4*>(rowp + ((j ^ r7) << 4)) = make_float4(y0, y1, y2, y3);
                    pk[2 * j] = pack_bf16x2(y0, y1);
                    pk[2 * j + 1] = pack_bf16x2(y2, y3);
                }
                uint4* o4 = reinterpret_cast<uint4*>(orow + c * 32);
#pragma unroll
                for (int j = 0; j < 4; ++j) o4[j] = make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]);
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tm.out, stg + buf * G2_EPI_TILE, col0 + c * 32, row0);
                    tma_store_commit();
                }
            }
            // all four stores must have read their staging tiles before the next tile's residual lands in them
            if (lane == 0) {
                tma_store_wait_read<0>();
                const int nt = tile + num_clusters;
                if (nt < num_tiles)
                    for (int c = 0; c < 3; ++c) issue_resid(nt, c, c);
            }
            __syncwarp();
        }
    }

    tc_fence_before();
    cluster_sync_all();                                           // nobody exits while a peer may still signal it or write its statistics
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

}  // namespace tvb
