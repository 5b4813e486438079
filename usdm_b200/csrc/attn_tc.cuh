// Fused masked self-attention on tcgen05 with the reference's additive term computed on the fly.
//
// Replaces reference networks.py:319-341 (materialised [B,16,S,S] ALiBi + mask tensors) and :174-193
// (bmm QK^T -> + bias + mask -> fp32 softmax -> bmm PV).  One CTA = one (sequence, head, 256-query tile):
//     S  = Q K^T                      tcgen05.mma  M=128 N=128 K=64 per 128-query half, Q/K tiles by TMA (SWIZZLE_128B)
//     t  = S*log2(e) - slope_h*log2(e)*|i-j| [j != 0] ;  keys j >= len masked                 (fp32, registers)
//     P  = exp2(t - running_max)      single pass over S held in registers, fp32 row max / row sum (online softmax)
//     O += P V                        tcgen05.mma  M=128 N=64  K=128, P staged in smem as bf16 (K-major, swizzled by hand),
//                                     V tile used in place as an MN-major operand
// q is pre-scaled by head_dim^-0.5 (folded into W_q / b_q at weight-finalise time, exact: 2^-3).
// Warp roles (384 threads = 3 warpgroups): warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocation (warpgroup 0
// gives most of its registers back with setmaxnreg), warps 4..7 softmax of query half A, warps 8..11 softmax of query
// half B (thread == query row == TMEM lane; setmaxnreg raises them to 216 registers so a 128-column score row lives in
// registers without spilling).  The MMA warp ping-pongs between the halves so that the tensor pipe works on one half
// while the other half's softmax runs; K/V tiles are shared by both halves.
// PTMEM = true keeps P in tensor memory (bf16 pairs written with tcgen05.st over the dead S columns, consumed as the
// TMEM A operand of the PV MMA); PTMEM = false stages P through shared memory.
// TMEM (512 columns): S_A [0,128) S_B [128,256) O_A [256,320) O_B [320,384).
//
// Off-diagonal key blocks have a known sign of (i - j), so the bias is affine in j and is evaluated with packed
// fp32x2 FMAs (FFMA2/FADD2); only the one diagonal block per half needs |i - j|.
#pragma once
#include "common.cuh"

namespace tvb {

constexpr int ATT_BQ = 256;      // queries per CTA (two halves of 128)
constexpr int ATT_BK = 128;      // keys per block
constexpr int ATT_D = 64;        // head dim
constexpr int ATT_KV_STAGES = 3;
constexpr int ATT_TILE_BYTES = 128 * 64 * 2;                                  // 16 KB: Q half, K, V tiles and each half of a P tile
constexpr int ATT_SMEM_Q = 0;                                                 // 2 x 16 KB
constexpr int ATT_SMEM_KV = 2 * ATT_TILE_BYTES;                               // stages x (K, V)
constexpr int ATT_SMEM_P = ATT_SMEM_KV + ATT_KV_STAGES * 2 * ATT_TILE_BYTES;  // 2 halves x 32 KB
constexpr int ATT_SMEM_BAR = ATT_SMEM_P + 4 * ATT_TILE_BYTES;
constexpr int ATT_SMEM_BYTES = ATT_SMEM_BAR + 256;
constexpr int ATT_THREADS = 384;   // warpgroup 0: TMA / MMA / TMEM-alloc warps (+1 idle), warpgroups 1-2: softmax of query halves A / B
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint64_t pack2f(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack2f(uint64_t v, float& lo, float& hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t fadd2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
__device__ __forceinline__ float ex2_approx(float x) {      // 2^x, MUFU.EX2; inputs are <= 0 here, tiny results flush to 0
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

template <bool PTMEM>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, const float* __restrict__ slopes,
               const int4* __restrict__ work, const float qk_scale /* log2(e), or 1 when it is already folded into W_q */) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + ATT_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;                      // [3]
    uint64_t* kv_empty = bars + 1 + ATT_KV_STAGES;     // [3]
    uint64_t* s_full = bars + 1 + 2 * ATT_KV_STAGES;   // [2] per half
    uint64_t* p_ready = s_full + 2;                    // [2]
    uint64_t* o_full = p_ready + 2;                    // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int4 w = work[blockIdx.x];
    const int q0 = w.y, start = w.z, len = w.w;
    const int head = blockIdx.y;
    const int nkb = (len + ATT_BK - 1) / ATT_BK;
    const bool two = (q0 + 128 < len);                 // second query half has rows

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < ATT_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int t = 0; t < 2; ++t) { mbar_init(&s_full[t], 1); mbar_init(&p_ready[t], 128); mbar_init(&o_full[t], 1); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp_idx == 0) {
        // ------------------------------------------ TMA producer ------------------------------------------
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, (two ? 2 : 1) * ATT_TILE_BYTES);
            tma_load_2d(smem + ATT_SMEM_Q, &tmQKV, q_full, head * ATT_D, start + q0);
            if (two) tma_load_2d(smem + ATT_SMEM_Q + ATT_TILE_BYTES, &tmQKV, q_full, head * ATT_D, start + q0 + 128);
            int slot = 0; uint32_t phase = 0;
            for (int j = 0; j < nkb; ++j) {
                mbar_wait(&kv_empty[slot], phase ^ 1);
                uint8_t* sk = smem + ATT_SMEM_KV + slot * 2 * ATT_TILE_BYTES;
                mbar_arrive_expect_tx(&kv_full[slot], 2 * ATT_TILE_BYTES);
                tma_load_2d(sk, &tmQKV, &kv_full[slot], H + head * ATT_D, start + j * ATT_BK);
                tma_load_2d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[slot], 2 * H + head * ATT_D, start + j * ATT_BK);
                if (++slot == ATT_KV_STAGES) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------ MMA issuer ------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                  // A, B K-major
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);      // B (= V) MN-major
            const uint32_t sQ = smem_u32(smem + ATT_SMEM_Q);
            const uint32_t sP = smem_u32(smem + ATT_SMEM_P);
            auto issue_qk = [&](int half, int slot) {
                const uint32_t sK = smem_u32(smem + ATT_SMEM_KV + slot * 2 * ATT_TILE_BYTES);
                const uint64_t da = make_smem_desc_sw128(sQ + half * ATT_TILE_BYTES, 16, 1024);
                const uint64_t db = make_smem_desc_sw128(sK, 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_ss(tmem_base + half * 128, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                umma_commit(&s_full[half]);
            };
            auto issue_pv = [&](int half, int slot, int j) {
                const uint32_t sV = smem_u32(smem + ATT_SMEM_KV + slot * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
                const uint32_t sPh = sP + half * 2 * ATT_TILE_BYTES;
#pragma unroll
                for (int k = 0; k < ATT_BK / 16; ++k) {
                    // B: V rows [16k, 16k+16) -> 2048 B per K step.
                    const uint64_t db = make_smem_desc_sw128(sV + k * 2048, 8192, 1024);
                    if constexpr (PTMEM) {
                        // A: P in TMEM over the first 64 columns of this half's S region, 8 columns (16 bf16) per K step
                        umma_ts(tmem_base + 256 + half * 64, tmem_base + half * 128 + k * 8, db, idesc_pv, (j | k) != 0);
                    } else {
                        // A: P half-tile (k / 4) in smem, 32 B per K step inside the swizzle atom
                        const uint64_t da = make_smem_desc_sw128(sPh + (k >> 2) * ATT_TILE_BYTES + (k & 3) * 32, 16, 1024);
                        umma_ss(tmem_base + 256 + half * 64, da, db, idesc_pv, (j | k) != 0);
                    }
                }
            };
            mbar_wait(q_full, 0);
            mbar_wait(&kv_full[0], 0);
            tc_fence_after();
            issue_qk(0, 0);
            if (two) issue_qk(1, 0);
            int slot = 0; uint32_t phase = 0;
            for (int j = 0; j < nkb; ++j) {
                int nslot = slot + 1; uint32_t nphase = phase;
                if (nslot == ATT_KV_STAGES) { nslot = 0; nphase ^= 1; }
                const bool more = (j + 1 < nkb);
                // ---- half A
                mbar_wait(&p_ready[0], j & 1);
                tc_fence_after();
                issue_pv(0, slot, j);
                if (!two) umma_commit(&kv_empty[slot]);
                if (more) {
                    mbar_wait(&kv_full[nslot], nphase);
                    tc_fence_after();
                    issue_qk(0, nslot);
                } else {
                    umma_commit(&o_full[0]);
                }
                // ---- half B
                if (two) {
                    mbar_wait(&p_ready[1], j & 1);
                    tc_fence_after();
                    issue_pv(1, slot, j);
                    umma_commit(&kv_empty[slot]);
                    if (more) issue_qk(1, nslot);
                    else umma_commit(&o_full[1]);
                }
                slot = nslot; phase = nphase;
            }
        }
    }
    } else {
        // ------------------------------ softmax + epilogue: one thread per query row ------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int half = (warp_idx - 4) >> 2;
        if (half == 0 || two) {
            const int quad = warp_idx & 3;
            const int r = quad * 32 + lane;                 // row inside the half == TMEM lane
            const int qt = q0 + half * 128;                 // first query of this half
            const int i = qt + r;                           // query position in the sequence
            const float fi = static_cast<float>(i);
            const float slope2 = slopes[head] * LOG2E;
            const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
            const uint32_t tmem_S = tmem_base + half * 128 + lane_off;
            const uint32_t tmem_O = tmem_base + 256 + half * 64 + lane_off;
            uint8_t* sP = smem + ATT_SMEM_P + half * 2 * ATT_TILE_BYTES + r * 128;
            const uint64_t L2 = pack2f(qk_scale, qk_scale);
            float m_run = -INFINITY, l_run = 0.f;
            for (int j = 0; j < nkb; ++j) {
                const int k0 = j * ATT_BK;
                mbar_wait(&s_full[half], j & 1);
                tc_fence_after();
                uint32_t s[128];
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_S + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
                tmem_ld_wait();
                const float s0raw = __uint_as_float(s[0]);
                // ---- t = S*log2e + bias, in place
                if (k0 != qt) {
                    // whole block on one side of the diagonal: |i-j| = sg*(i-j)  ->  bias = -sg*slope2*i + sg*slope2*j
                    const float a = (k0 < qt) ? slope2 : -slope2;
                    const float cb = a * (static_cast<float>(k0) - fi);
                    const uint64_t inc = pack2f(2.f * a, 2.f * a);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float cc = fmaf(a, static_cast<float>(c * 32), cb);     // restart per 32 keys: bounded rounding drift
                        uint64_t ct = pack2f(cc, cc + a);
#pragma unroll
                        for (int e = c * 32; e < c * 32 + 32; e += 2) {
                            const uint64_t t2 = ffma2(pack2f(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), L2, ct);
                            ct = fadd2(ct, inc);
                            float lo, hi;
                            unpack2f(t2, lo, hi);
                            s[e] = __float_as_uint(lo);
                            s[e + 1] = __float_as_uint(hi);
                        }
                    }
                } else {
                    const float d0 = fi - static_cast<float>(k0);
#pragma unroll
                    for (int e = 0; e < 128; ++e) {
                        const float bias = -slope2 * fabsf(d0 - static_cast<float>(e));
                        s[e] = __float_as_uint(fmaf(__uint_as_float(s[e]), qk_scale, bias));
                    }
                }
                if (k0 == 0) s[0] = __float_as_uint(s0raw * qk_scale);       // key 0 (time token) carries no distance penalty
                if (k0 + ATT_BK > len) {
#pragma unroll
                    for (int e = 0; e < 128; ++e)
                        if (k0 + e >= len) s[e] = __float_as_uint(-INFINITY);
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int e = 0; e < 128; e += 4) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[e]), __uint_as_float(s[e + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])));
                }
                const float m_new = fmaxf(m_run, fmaxf(mx0, mx1));
                if (j > 0 && __any_sync(0xffffffffu, m_new > m_run)) {
                    const float alpha = ex2_approx(m_run - m_new);
                    l_run *= alpha;
#pragma unroll
                    for (int c = 0; c < ATT_D / 16; ++c) {
                        uint32_t o[16];
                        tmem_ld_32x16(tmem_O + c * 16, o);
                        tmem_ld_wait();
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
                        tmem_st_32x16(tmem_O + c * 16, o);
                    }
                    tmem_st_wait();
                }
                m_run = m_new;
                // ---- P = exp2(t - m), row sum, bf16 P tile into smem (K-major, 128-B swizzle: chunk ^= row & 7)
                const uint64_t nm = pack2f(-m_new, -m_new);
                uint64_t acc2 = pack2f(0.f, 0.f);
#pragma unroll
                for (int g = 0; g < 16; ++g) {          // 16 groups of 8 keys = one 16-byte chunk each
                    uint32_t pk[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int e = g * 8 + u * 2;
                        float lo, hi;
                        unpack2f(fadd2(pack2f(__uint_as_float(s[e]), __uint_as_float(s[e + 1])), nm), lo, hi);
                        lo = ex2_approx(lo);
                        hi = ex2_approx(hi);
                        acc2 = fadd2(acc2, pack2f(lo, hi));
                        pk[u] = pack_bf16x2(lo, hi);
                    }
                    if constexpr (PTMEM) {
                        // bf16 pairs (keys 2c, 2c+1) -> 32-bit column c of this row's P, over the dead S columns
                        s[g * 4 + 0] = pk[0]; s[g * 4 + 1] = pk[1]; s[g * 4 + 2] = pk[2]; s[g * 4 + 3] = pk[3];
                    } else {
                        // keys [8g, 8g+8): half-tile (g >> 3), chunk (g & 7)
                        const int chunk = (g & 7) ^ (r & 7);
                        *reinterpret_cast<uint4*>(sP + (g >> 3) * ATT_TILE_BYTES + chunk * 16) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
                    }
                }
                if constexpr (PTMEM) {
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_S + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                    tmem_st_wait();
                }
                float a0, a1;
                unpack2f(acc2, a0, a1);
                l_run += a0 + a1;
                if constexpr (!PTMEM) fence_proxy_async_smem();
                tc_fence_before();
                mbar_arrive(&p_ready[half]);
            }
            // epilogue: O / l -> ctx (bf16), only rows that belong to the sequence
            mbar_wait(&o_full[half], 0);
            tc_fence_after();
            const float inv = 1.0f / l_run;
            __nv_bfloat16* dst = ctx + static_cast<size_t>(start + i) * H + head * ATT_D;
#pragma unroll
            for (int c = 0; c < ATT_D / 16; ++c) {
                uint32_t o[16];
                tmem_ld_32x16(tmem_O + c * 16, o);
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
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

inline cudaError_t launch_attn_tc_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                      const int4* work, int nwork, bool p_in_tmem, float qk_scale, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(attn_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    dim3 grid(nwork, nheads);
    if (p_in_tmem) attn_tc_kernel<true><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tmQKV, ctx, H, slopes, work, qk_scale);
    else           attn_tc_kernel<false><<<grid, ATT_THREADS, ATT_SMEM_BYTES, st>>>(tmQKV, ctx, H, slopes, work, qk_scale);
    return cudaGetLastError();
}

}  // namespace tvb
