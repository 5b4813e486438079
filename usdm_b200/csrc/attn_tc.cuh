// Fused masked self-attention on tcgen05 with the reference's additive term computed on the fly.
//
// Replaces reference networks.py:319-341 (materialised [B,16,S,S] ALiBi + mask tensors) and :174-193
// (bmm QK^T -> + bias + mask -> fp32 softmax -> bmm PV).  Per (sequence, head, 128-query tile):
//     S  = Q K^T                      tcgen05.mma  M=128 N=128 K=64, Q/K tiles by TMA (SWIZZLE_128B, K-major)
//     t  = S*log2(e) - slope_h*log2(e)*|i-j| [j != 0] ;  keys j >= len masked          (fp32, registers)
//     P  = exp2(t - running_max)      online softmax, fp32 row max / row sum
//     O += P V                        tcgen05.mma  M=128 N=64  K=128, P staged in smem as bf16 (K-major, swizzled by hand),
//                                     V tile used in place as an MN-major operand
// q is pre-scaled by head_dim^-0.5 (folded into W_q / b_q at weight-finalise time, exact: 2^-3).
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 softmax + epilogue (thread == query row == TMEM lane).
// TMEM: S at columns [0,128), O at [128,192) -> 256 allocated columns, two CTAs resident per SM so that one CTA's
// softmax overlaps the other's MMAs.
#pragma once
#include "common.cuh"

namespace tvb {

constexpr int ATT_BQ = 128;      // queries per CTA
constexpr int ATT_BK = 128;      // keys per block
constexpr int ATT_D = 64;        // head dim
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;                    // 16 KB: Q, K, V tiles and each half of P
constexpr int ATT_SMEM_Q = 0;
constexpr int ATT_SMEM_KV = ATT_TILE_BYTES;                     // 2 stages x (K, V)
constexpr int ATT_SMEM_P = ATT_SMEM_KV + 4 * ATT_TILE_BYTES;    // 2 x 16 KB
constexpr int ATT_SMEM_BAR = ATT_SMEM_P + 2 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_BAR + 128;
constexpr int ATT_THREADS = 192;
constexpr float LOG2E = 1.4426950408889634f;

__global__ void __launch_bounds__(ATT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, const float* __restrict__ slopes,
               const int4* __restrict__ work) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;    // [2]
    uint64_t* kv_empty = bars + 3;   // [2]
    uint64_t* s_full = bars + 5;
    uint64_t* p_ready = bars + 6;
    uint64_t* o_full = bars + 7;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int4 w = work[blockIdx.x];
    const int q0 = w.y, start = w.z, len = w.w;
    const int head = blockIdx.y;
    const int nkb = (len + ATT_BK - 1) / ATT_BK;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        mbar_init(&kv_full[0], 1); mbar_init(&kv_full[1], 1);
        mbar_init(&kv_empty[0], 1); mbar_init(&kv_empty[1], 1);
        mbar_init(s_full, 1);
        mbar_init(p_ready, 128);
        mbar_init(o_full, 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_S = tmem_base;
    const uint32_t tmem_O = tmem_base + 128;

    if (warp_idx == 0) {
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, ATT_TILE_BYTES);
            tma_load_2d(smem + ATT_SMEM_Q, &tmQKV, q_full, head * ATT_D, start + q0);
            for (int j = 0; j < nkb; ++j) {
                const int slot = j & 1;
                mbar_wait(&kv_empty[slot], ((j >> 1) & 1) ^ 1);
                uint8_t* sk = smem + ATT_SMEM_KV + slot * 2 * ATT_TILE_BYTES;
                mbar_arrive_expect_tx(&kv_full[slot], 2 * ATT_TILE_BYTES);
                tma_load_2d(sk, &tmQKV, &kv_full[slot], H + head * ATT_D, start + j * ATT_BK);
                tma_load_2d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[slot], 2 * H + head * ATT_D, start + j * ATT_BK);
            }
        }
    } else if (warp_idx == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                  // A, B K-major
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);      // B (= V) MN-major
            const uint32_t sQ = smem_u32(smem + ATT_SMEM_Q);
            const uint32_t sP = smem_u32(smem + ATT_SMEM_P);
            auto issue_qk = [&](int j) {
                const int slot = j & 1;
                mbar_wait(&kv_full[slot], (j >> 1) & 1);
                tc_fence_after();
                const uint32_t sK = smem_u32(smem + ATT_SMEM_KV + slot * 2 * ATT_TILE_BYTES);
                const uint64_t da = make_smem_desc_sw128(sQ, 16, 1024);
                const uint64_t db = make_smem_desc_sw128(sK, 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_S, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                umma_commit(s_full);
            };
            mbar_wait(q_full, 0);
            issue_qk(0);
            for (int j = 0; j < nkb; ++j) {
                const int slot = j & 1;
                mbar_wait(p_ready, j & 1);
                tc_fence_after();
                const uint32_t sV = smem_u32(smem + ATT_SMEM_KV + slot * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
#pragma unroll
                for (int k = 0; k < ATT_BK / 16; ++k) {
                    // A: P half (k / 4), 32 B per K step inside the swizzle atom.  B: V rows [16k, 16k+16) -> 2048 B per K step.
                    const uint64_t da = make_smem_desc_sw128(sP + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sV + k * 2048, 8192, 1024);
                    umma_ss(tmem_O, da, db, idesc_pv, (j | k) != 0);
                }
                umma_commit(&kv_empty[slot]);
                if (j + 1 < nkb) issue_qk(j + 1);
                else umma_commit(o_full);
            }
        }
    } else {
        // ------------------------------ softmax + epilogue: one thread per query row ------------------------------
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane;                 // row inside the tile == TMEM lane
        const int i = q0 + r;                           // query position in the sequence
        const float fi = static_cast<float>(i);
        const float slope2 = slopes[head] * LOG2E;
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        uint8_t* sP = smem + ATT_SMEM_P;
        float m_run = -INFINITY, l_run = 0.f;
        for (int j = 0; j < nkb; ++j) {
            const int k0 = j * ATT_BK;
            const bool tail = (k0 + ATT_BK > len);
            mbar_wait(s_full, j & 1);
            tc_fence_after();
            // pass 1: row max of t = S*log2e + bias
            float mx = -INFINITY;
#pragma unroll 1
            for (int c = 0; c < ATT_BK / 32; ++c) {
                uint32_t s[32];
                tmem_ld_32x32(tmem_S + lane_off + c * 32, s);
                tmem_ld_wait();
#pragma unroll
                for (int jj = 0; jj < 32; ++jj) {
                    const int key = k0 + c * 32 + jj;
                    float bias = -slope2 * fabsf(fi - static_cast<float>(key));
                    if (key == 0) bias = 0.f;
                    float t = fmaf(__uint_as_float(s[jj]), LOG2E, bias);
                    if (tail && key >= len) t = -INFINITY;
                    mx = fmaxf(mx, t);
                }
            }
            const float m_new = fmaxf(m_run, mx);
            if (j > 0 && __any_sync(0xffffffffu, m_new > m_run)) {
                const float alpha = exp2f(m_run - m_new);
                l_run *= alpha;
#pragma unroll 1
                for (int c = 0; c < ATT_D / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld_32x16(tmem_O + lane_off + c * 16, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int jj = 0; jj < 16; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
                    tmem_st_32x16(tmem_O + lane_off + c * 16, o);
                }
                tmem_st_wait();
            }
            m_run = m_new;
            // pass 2: P = exp2(t - m), row sum, bf16 P tile into smem (K-major, 128-B swizzle: chunk ^= row & 7)
            float lsum = 0.f;
#pragma unroll 1
            for (int c = 0; c < ATT_BK / 32; ++c) {
                uint32_t s[32];
                tmem_ld_32x32(tmem_S + lane_off + c * 32, s);
                tmem_ld_wait();
                uint32_t pk[16];
#pragma unroll
                for (int jj = 0; jj < 32; jj += 2) {
                    float pv[2];
#pragma unroll
                    for (int u = 0; u < 2; ++u) {
                        const int key = k0 + c * 32 + jj + u;
                        float bias = -slope2 * fabsf(fi - static_cast<float>(key));
                        if (key == 0) bias = 0.f;
                        const float t = fmaf(__uint_as_float(s[jj + u]), LOG2E, bias);
                        float p = exp2f(t - m_new);
                        if (tail && key >= len) p = 0.f;
                        lsum += p;
                        pv[u] = p;
                    }
                    pk[jj >> 1] = pack_bf16x2(pv[0], pv[1]);
                }
                // 32 keys = 4 chunks of 16 B; key block c -> half (c >> 1), chunk base (c & 1) * 4
                uint8_t* rowp = sP + (c >> 1) * ATT_TILE_BYTES + r * 128;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int chunk = ((c & 1) * 4 + q) ^ (r & 7);
                    *reinterpret_cast<uint4*>(rowp + chunk * 16) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
                }
            }
            l_run += lsum;
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(p_ready);
        }
        // epilogue: O / l -> ctx (bf16), only rows that belong to the sequence
        mbar_wait(o_full, 0);
        tc_fence_after();
        const float inv = 1.0f / l_run;
        __nv_bfloat16* dst = ctx + static_cast<size_t>(start + i) * H + head * ATT_D;
#pragma unroll 1
        for (int c = 0; c < ATT_D / 16; ++c) {
            uint32_t o[16];
            tmem_ld_32x16(tmem_O + lane_off + c * 16, o);
            tmem_ld_wait();
            if (i < len) {
                uint4 a, b;
                a.x = pack_bf16x2(__uint_as_float(o[0]) * inv, __uint_as_float(o[1]) * inv);
                a.y = pack_bf16x2(__uint_as_float(o[2]) * inv, __uint_as_float(o[3]) * inv);
                a.z = pack_bf16x2(__uint_as_float(o[4]) * inv, __uint_as_float(o[5]) * inv);
                a.w = pack_bf16x2(__uint_as_float(o[6]) * inv, __uint_as_float(o[7]) * inv);
                b.x = pack_bf16x2(__uint_as_float(o[8]) * inv, __uint_as_float(o[9]) * inv);
                b.y = pack_bf16x2(__uint_as_float(o[10]) * inv, __uint_as_float(o[11]) * inv);
                b.z = pack_bf16x2(__uint_as_float(o[12]) * inv, __uint_as_float(o[13]) * inv);
                b.w = pack_bf16x2(__uint_as_float(o[14]) * inv, __uint_as_float(o[15]) * inv);
                reinterpret_cast<uint4*>(dst + c * 16)[0] = a;
                reinterpret_cast<uint4*>(dst + c * 16)[1] = b;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

inline cudaError_t launch_attn_tc_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                      const int4* work, int nwork, cudaStream_t st) {
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr = true;
    }
    dim3 grid(nwork, nheads);
    attn_tc_kernel<<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tmQKV, ctx, H, slopes, work);
    return cudaGetLastError();
}

}  // namespace tvb
