// Fused masked self-attention, version 5: attn4_tc.cuh with TWO threads per query row.
//
// Why (ncu of v4, profiles/r01/ncu_attn4_*): MUFU.EX2 issues once per 8 cycles per SM sub-partition and v4 kept that pipe only
// 45 % busy, the issue slots 33 % busy; with 8 softmax warps (2 per sub-partition, thread == whole 128-key row) each warp's
// load -> max -> exp2 -> convert -> store chain ran back to back with nothing to overlap it.  Here a row is split between two
// threads (keys [0,64) and [64,128) of every block), i.e. 16 softmax warps = 4 per sub-partition, each with a 64-register score
// row.  Everything else is v4: TMA-fed QK^T / PV on tcgen05, three rotating S buffers in TMEM, the ALiBi slope term added by
// an extra K=16 MMA step, lazy (2^16) reference-maximum update, P written back to TMEM as bf16 and consumed as the TMEM A operand.
//   * the two threads of a row agree on the block maximum through shared memory (one named barrier per half and block),
//     so the softmax is an exact two-pass one per block and never has to be redone;
//   * P of keys [0,64) goes to columns [0,32) of the S buffer, P of keys [64,128) to columns [64,96): each thread only
//     overwrites score columns it has loaded itself; the PV MMA addresses its TMEM A operand per 16-key step accordingly;
//   * O (64 columns) is rescaled by the part-0 thread of the row and written out half by each thread.
// Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocation, 3 idle, 4..19 softmax: half = (w-4)/8, part = ((w-4)/4)&1,
// TMEM lane quadrant = w & 3.  640 threads, 1 CTA per SM; setmaxnreg gives the softmax threads 104 registers.
#pragma once
#include "attn4_tc.cuh"

namespace tvb {

constexpr int A5_THREADS = 640;
constexpr int A5_SMEM_XCH = A4_SMEM_BAR + 256;                 // fp32 [2 halves][2 parities][2 parts][128 rows]
constexpr int A5_SMEM_BYTES = A5_SMEM_XCH + 2 * 2 * 2 * 128 * 4;

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

template <int POLY>
__global__ void __launch_bounds__(A5_THREADS, 1)
attn5_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
             const float* __restrict__ slopes, const int4* __restrict__ work) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A4_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;                       // [4]
    uint64_t* kv_empty = kv_full + A4_KV_STAGES;        // [4]
    uint64_t* s_full = kv_empty + A4_KV_STAGES;         // [3] per S buffer
    uint64_t* p_ready = s_full + 3;                     // [3] per S buffer, 256 arrivals
    uint64_t* pv_done = p_ready + 3;                    // [2] per half
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);
    float* xch = reinterpret_cast<float*>(smem + A5_SMEM_XCH);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int4 w = work[blockIdx.x / nheads];
    const int head = blockIdx.x % nheads;
    const int q0 = w.y, start = w.z, len = w.w;
    const int nkb = (len + ATT_BK - 1) / ATT_BK;
    const bool two = (q0 + 128 < len);                  // second query half has rows
    const int nh = two ? 2 : 1;
    const int qb0 = q0 / ATT_BK;                        // key block that is diagonal for half A (half B: qb0 + 1)
    const float slope2 = slopes[head] * LOG2E;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp_idx == 1 && lane == 0) {
        mbar_init(q_full, 1);
        for (int s = 0; s < A4_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 3; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 256); }
        for (int t = 0; t < 2; ++t) mbar_init(&pv_done[t], 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
    if (warp_idx >= 4 && warp_idx < 8) {
        // bias operand tile, see attn4_tc.cuh: columns [0,16) {jrel x3}, [16,32) {+sigma hi/mid/lo}, [32,48) {-sigma hi/mid/lo}
        const int r = threadIdx.x - 128;
        const __nv_bfloat16 hi = __float2bfloat16_rn(slope2);
        const float rem1 = slope2 - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(rem1);
        const __nv_bfloat16 lo = __float2bfloat16_rn(rem1 - __bfloat162float(mid));
        const __nv_bfloat16 jr = __float2bfloat16_rn(static_cast<float>(r));
        auto bits = [](__nv_bfloat16 v) { return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v)); };
        uint8_t* rowp = smem + A4_SMEM_EXT + r * 128;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        const uint4 ck = make_uint4(bits(jr) | (bits(jr) << 16), bits(jr), 0, 0);
        const uint4 cp = make_uint4(bits(hi) | (bits(mid) << 16), bits(lo), 0, 0);
        const uint4 cn = make_uint4(cp.x ^ 0x80008000u, cp.y ^ 0x8000u, 0, 0);
        const int x7 = r & 7;
        *reinterpret_cast<uint4*>(rowp + ((0 ^ x7) << 4)) = ck;
        *reinterpret_cast<uint4*>(rowp + ((1 ^ x7) << 4)) = zero;
        *reinterpret_cast<uint4*>(rowp + ((2 ^ x7) << 4)) = cp;
        *reinterpret_cast<uint4*>(rowp + ((3 ^ x7) << 4)) = zero;
        *reinterpret_cast<uint4*>(rowp + ((4 ^ x7) << 4)) = cn;
        *reinterpret_cast<uint4*>(rowp + ((5 ^ x7) << 4)) = zero;
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
    if (warp_idx == 0) {
        // ------------------------------------------ TMA producer ------------------------------------------
        if (lane == 0) {
            mbar_arrive_expect_tx(q_full, nh * ATT_TILE_BYTES);
            tma_load_2d(smem + A4_SMEM_Q, &tmQKV, q_full, head * ATT_D, start + q0);
            if (two) tma_load_2d(smem + A4_SMEM_Q + ATT_TILE_BYTES, &tmQKV, q_full, head * ATT_D, start + q0 + 128);
            int slot = 0; uint32_t phase = 0;
            for (int j = 0; j < nkb; ++j) {
                mbar_wait(&kv_empty[slot], phase ^ 1);
                uint8_t* sk = smem + A4_SMEM_KV + slot * 2 * ATT_TILE_BYTES;
                mbar_arrive_expect_tx(&kv_full[slot], 2 * ATT_TILE_BYTES);
                tma_load_2d(sk, &tmQKV, &kv_full[slot], H + head * ATT_D, start + j * ATT_BK);
                tma_load_2d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[slot], 2 * H + head * ATT_D, start + j * ATT_BK);
                if (++slot == A4_KV_STAGES) { slot = 0; phase ^= 1; }
            }
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------ MMA issuer ------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                  // A, B K-major
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);      // B (= V) MN-major
            const uint32_t sQ = smem_u32(smem + A4_SMEM_Q);
            const uint64_t dext = make_smem_desc_sw128(smem_u32(smem + A4_SMEM_EXT), 16, 1024);
            const int total = nkb * nh;
            // step n: half n % nh, key block n / nh, S buffer n % 3
            auto issue_qk = [&](int n) {
                const int half = two ? (n & 1) : 0;
                const int j = two ? (n >> 1) : n;
                const int slot = j % A4_KV_STAGES;
                if (half == 0) {                                                          // first user of this K/V stage
                    mbar_wait(&kv_full[slot], (j / A4_KV_STAGES) & 1);
                    tc_fence_after();
                }
                const uint32_t d = tmem_base + (n % 3) * 128;
                const uint64_t da = make_smem_desc_sw128(sQ + half * ATT_TILE_BYTES, 16, 1024);
                const uint64_t db = make_smem_desc_sw128(smem_u32(smem + A4_SMEM_KV + slot * 2 * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                const int jd = qb0 + half;
                if (j != jd) umma_ss(d, dext + (j < jd ? 2 : 4), dext, idesc_qk, 1);       // +sigma*jrel left of the diagonal, -sigma*jrel right
                umma_commit(&s_full[n % 3]);
            };
            auto issue_pv = [&](int n) {
                const int half = two ? (n & 1) : 0;
                const int j = two ? (n >> 1) : n;
                const int slot = j % A4_KV_STAGES;
                mbar_wait(&p_ready[n % 3], (n / 3) & 1);
                tc_fence_after();
                const uint32_t sV = smem_u32(smem + A4_SMEM_KV + slot * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
                const uint32_t tP = tmem_base + (n % 3) * 128;
                const uint32_t tO = tmem_base + 384 + half * 64;
#pragma unroll
                for (int k = 0; k < ATT_BK / 16; ++k) {
                    const uint64_t db = make_smem_desc_sw128(sV + k * 2048, 8192, 1024);   // V rows [16k, 16k+16)
                    // P of keys [0,64) sits in columns [0,32), of keys [64,128) in columns [64,96); 8 columns (16 bf16) per step
                    umma_ts(tO, tP + (k < 4 ? k * 8 : 64 + (k - 4) * 8), db, idesc_pv, (j | k) != 0);
                }
                umma_commit(&pv_done[half]);
                if (half == nh - 1) umma_commit(&kv_empty[slot]);                         // every MMA reading this stage has been issued
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            for (int n = 0; n < 3 && n < total; ++n) issue_qk(n);
            for (int n = 0; n < total; ++n) {
                issue_pv(n);
                if (n + 3 < total) issue_qk(n + 3);
            }
        }
    }
    } else {
        // ------------------------------ softmax + epilogue: two threads per query row ------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
        const int sw = warp_idx - 4;
        const int half = sw >> 3;
        if (half == 0 || two) {
            const int part = (sw >> 2) & 1;                 // 0: keys [0,64) of each block, 1: keys [64,128)
            const int quad = warp_idx & 3;
            const int r = quad * 32 + lane;                 // row inside the half == TMEM lane
            const int qt = q0 + half * 128;                 // first query of this half
            const int jd = qb0 + half;                      // diagonal key block
            const int i = qt + r;                           // query position in the sequence
            const int c0 = part * 64;                       // first key column of this thread inside a block
            const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
            const uint32_t tmem_O = tmem_base + 384 + half * 64 + part * 32 + lane_off;
            const float a0d = slope2 * static_cast<float>(c0 - r);   // diagonal block: -sigma |r - c| = -|sigma c - sigma r|
            float* xmine = xch + half * 512 + part * 128 + r;        // + parity * 256
            float* xother = xch + half * 512 + (part ^ 1) * 128 + r;
            const int bar_id = 1 + half;
            float m_ref = 0.f, l_run = 0.f;
            for (int j = 0; j < nkb; ++j) {
                const int n = two ? 2 * j + half : j;
                const int k0 = j * ATT_BK;
                const int buf = n % 3;
                const uint32_t tmem_S = tmem_base + buf * 128 + c0 + lane_off;
                const bool diag = (j == jd);
                // additive term of this block: per-row constant (off-diagonal; the j-dependent part came out of the MMA) or none
                const float c_row = diag ? 0.f : -slope2 * fabsf(static_cast<float>(i - k0));
                mbar_wait(&s_full[buf], (n / 3) & 1);
                tc_fence_after();
                uint32_t s[64];
                tmem_ld_32x32(tmem_S, *reinterpret_cast<uint32_t(*)[32]>(&s[0]));
                tmem_ld_32x32(tmem_S + 32, *reinterpret_cast<uint32_t(*)[32]>(&s[32]));
                tmem_ld_wait();
                if (diag) {
                    const float s0raw = __uint_as_float(s[0]);
                    float a0v = a0d;
                    asm volatile("" : "+f"(a0v));              // keep the bias values from being hoisted out of the block loop
#pragma unroll
                    for (int e = 0; e < 64; ++e) {
                        const float a = fmaf(slope2, static_cast<float>(e), a0v);
                        s[e] = __float_as_uint(__uint_as_float(s[e]) - fabsf(a));
                    }
                    if (k0 == 0 && part == 0) s[0] = __float_as_uint(s0raw);      // key 0 (time token): no distance penalty
                } else if (k0 == 0 && part == 0) {
                    s[0] = __float_as_uint(__uint_as_float(s[0]) - c_row);
                }
                if (k0 + ATT_BK > len) {
#pragma unroll
                    for (int e = 0; e < 64; ++e)
                        if (k0 + c0 + e >= len) s[e] = __float_as_uint(-INFINITY);
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int e = 0; e < 64; e += 4) {
                    mx0 = fmaxf(mx0, fmaxf(__uint_as_float(s[e]), __uint_as_float(s[e + 1])));
                    mx1 = fmaxf(mx1, fmaxf(__uint_as_float(s[e + 2]), __uint_as_float(s[e + 3])));
                }
                const float m_loc = fmaxf(mx0, mx1) + c_row;             // my half of the row, true (biased) domain
                xmine[(j & 1) * 256] = m_loc;
                named_bar_sync(bar_id, 256);
                const float m_blk = fmaxf(m_loc, xother[(j & 1) * 256]);
                bool pv_seen = false;
                if (j == 0) {
                    m_ref = m_blk;
                } else {
                    const bool grow = m_blk > m_ref + A4_RESCALE_LOG2;   // identical in both threads of the row
                    if (__any_sync(0xffffffffu, grow)) {
                        const float m_new = grow ? m_blk : m_ref;
                        const float alpha = ex2_approx(m_ref - m_new);    // 1 for rows whose reference stays
                        m_ref = m_new;
                        l_run *= alpha;
                        mbar_wait(&pv_done[half], (j - 1) & 1);          // O is only stable once the previous PV has retired
                        pv_seen = true;
                        tc_fence_after();
#pragma unroll
                        for (int c = 0; c < 2; ++c) {                    // my 32 of the 64 O columns
                            uint32_t o[16];
                            tmem_ld_32x16(tmem_O + c * 16, o);
                            tmem_ld_wait();
#pragma unroll
                            for (int jj = 0; jj < 16; ++jj) o[jj] = __float_as_uint(__uint_as_float(o[jj]) * alpha);
                            tmem_st_32x16(tmem_O + c * 16, o);
                        }
                        tmem_st_wait();
                    }
                }
                // ---- P = exp2(s + c_row - m_ref), partial row sum, bf16 pairs over my own first 32 score columns
                const float off = c_row - m_ref;
                const uint64_t off2 = pack2f(off, off);
                uint64_t acc2 = pack2f(0.f, 0.f);
#pragma unroll
                for (int u = 0; u < 32; ++u) {
                    const uint64_t x2 = fadd2(pack2f(__uint_as_float(s[2 * u]), __uint_as_float(s[2 * u + 1])), off2);
                    uint64_t p2;
                    if (POLY > 0 && (u % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
                        p2 = exp2_poly2(x2);
                    } else {
                        float lo, hi;
                        unpack2f(x2, lo, hi);
                        p2 = pack2f(ex2_approx(lo), ex2_approx(hi));
                    }
                    acc2 = fadd2(acc2, p2);
                    float lo, hi;
                    unpack2f(p2, lo, hi);
                    s[u] = pack_bf16x2(lo, hi);
                }
                float a0, a1;
                unpack2f(acc2, a0, a1);
                l_run += a0 + a1;
                // every thread observes every phase of pv_done exactly once and BEFORE it lets PV(j) start (attn4_tc.cuh)
                if (j > 0 && !pv_seen) mbar_wait(&pv_done[half], (j - 1) & 1);
                tmem_st_32x16(tmem_S, *reinterpret_cast<uint32_t(*)[16]>(&s[0]));
                tmem_st_32x16(tmem_S + 16, *reinterpret_cast<uint32_t(*)[16]>(&s[16]));
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_ready[buf]);
            }
            // epilogue: row sum of both parts, O / l -> ctx (bf16), each thread 32 of the 64 columns, only rows of the sequence
            xmine[(nkb & 1) * 256] = l_run;
            named_bar_sync(bar_id, 256);
            const float inv = 1.0f / (l_run + xother[(nkb & 1) * 256]);
            mbar_wait(&pv_done[half], (nkb - 1) & 1);
            tc_fence_after();
            __nv_bfloat16* dst = ctx + static_cast<size_t>(start + i) * H + head * ATT_D + part * 32;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
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

template <int POLY>
inline cudaError_t launch_attn5_t(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  int nwork, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn5_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, A5_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    attn5_kernel<POLY><<<nwork * nheads, A5_THREADS, A5_SMEM_BYTES, st>>>(tmQKV, ctx, H, nheads, slopes, work);
    return cudaGetLastError();
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e).
inline cudaError_t launch_attn5_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                    const int4* work, int nwork, int poly, cudaStream_t st) {
    switch (poly) {
        case 0: return launch_attn5_t<0>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 2: return launch_attn5_t<2>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 3: return launch_attn5_t<3>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 4: return launch_attn5_t<4>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
