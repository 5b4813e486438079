// Fused masked self-attention, version 8: persistent CTAs (attn6_tc.cuh) + split converged issue warps (attn7_tc.cuh).
//
// Evidence (profiles/r01/attn4_trace.txt and the exp2-free timing ablation of v7): a non-persistent CTA spends ~10.9 k of its
// ~40.5 k cycles before the first and after the last softmax block (27 %; 39 % once the exponentials are taken out), and v6's
// persistence was eaten by its single divergent issue thread walking two item cursors.  Here:
//   warp 0  TMA producer (one elected lane): Q double-buffered per item, K/V ring running across items;
//   warp 1  PV issuer, warp 2 QK^T issuer: both converged, all counters incremental (no div / mod), elect.sync around the
//           tcgen05 instructions only; S-buffer hand-back through s_free[buffer] as in v7;
//   warp 3  rewrites the head-dependent bias operand tile one item ahead (double-buffered);
//   warps 4-11 softmax, one thread per query row, exactly v4's block body; O is handed back (o_free) as soon as it is read.
// Item parameters are prefetched one item ahead by every role.  Work items should arrive sorted by cost (build_plan does).
#pragma once
#include "attn6_tc.cuh"

namespace tvb {

// Two MUFU.EX2 whose issue is tied (through an always-true predicate the compiler cannot prove) to `pace`, a value produced by the
// tail of an EARLIER element pair.  Without it ptxas emits the 32 MUFUs of a chunk back to back (8 issue cycles each on this warp)
// and only then the adds / converts; with it the program order alternates MUFU pairs with the tails of the pairs before, so the FMA
// and ALU pipes work inside the MUFU issue gaps (measured pipe costs: tools/pipe_bench.cu).
__device__ __forceinline__ void ex2_pair_paced(float& lo, float& hi, float pace) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.ge.f32 p, %2, 0f00000000;\n\t"
        "@p ex2.approx.ftz.f32 %0, %0;\n\t"
        "@p ex2.approx.ftz.f32 %1, %1;\n\t}\n"
        : "+f"(lo), "+f"(hi)
        : "f"(pace));
}

template <int POLY, bool PACE>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn8_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
             const float* __restrict__ slopes, const int4* __restrict__ work, int n_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A6_SMEM_BAR);
    uint64_t* q_full = bars + 0;                        // [2] per Q buffer
    uint64_t* q_empty = q_full + 2;                     // [2]
    uint64_t* ext_full = q_empty + 2;                   // [2] per bias-tile buffer
    uint64_t* ext_empty = ext_full + 2;                 // [2]
    uint64_t* kv_full = ext_empty + 2;                  // [4]
    uint64_t* kv_empty = kv_full + A4_KV_STAGES;        // [4]
    uint64_t* s_full = kv_empty + A4_KV_STAGES;         // [3] per S buffer
    uint64_t* p_ready = s_full + 3;                     // [3] per S buffer
    uint64_t* pv_done = p_ready + 3;                    // [2 halves][2]: PV number c of a half arrives on [half][c & 1] (phase c >> 1), so a
                                                        //   thread may wait for PV(c) without having observed every earlier phase (see below)
    uint64_t* o_free = pv_done + 4;                     // [2] per half: O has been read out by the 128 softmax threads
    uint64_t* s_free = o_free + 2;                      // [3] per S buffer: the PV MMA that read P from it has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 3);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&ext_full[s], 1); mbar_init(&ext_empty[s], 1);
            mbar_init(&pv_done[2 * s], 1); mbar_init(&pv_done[2 * s + 1], 1); mbar_init(&o_free[s], 128);
        }
        for (int s = 0; s < A4_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 3; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 128); mbar_init(&s_free[s], 1); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                      // set-up above overlapped the tail of the QKV GEMM

    if (warp_idx < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp_idx == 0) {
        // ------------------------------------------ TMA producer ------------------------------------------
        if (elect_one()) {
            int slot = 0; uint32_t phase = 0;
            int itc = 0;
            for (int t = blockIdx.x; t < n_items; t += G, ++itc) {
                const A6Item w = a6_load_item(work, t, nheads);
                const int qb = itc & 1;
                mbar_wait(&q_empty[qb], ((itc >> 1) & 1) ^ 1);
                uint8_t* sq = smem + A6_SMEM_Q + qb * 2 * ATT_TILE_BYTES;
                mbar_arrive_expect_tx(&q_full[qb], w.nh * ATT_TILE_BYTES);
                tma_load_2d(sq, &tmQKV, &q_full[qb], w.head * ATT_D, w.start + w.q0);
                if (w.two) tma_load_2d(sq + ATT_TILE_BYTES, &tmQKV, &q_full[qb], w.head * ATT_D, w.start + w.q0 + 128);
                for (int j = 0; j < w.nkb; ++j) {
                    mbar_wait(&kv_empty[slot], phase ^ 1);
                    uint8_t* sk = smem + A6_SMEM_KV + slot * 2 * ATT_TILE_BYTES;
                    mbar_arrive_expect_tx(&kv_full[slot], 2 * ATT_TILE_BYTES);
                    tma_load_2d(sk, &tmQKV, &kv_full[slot], H + w.head * ATT_D, w.start + j * ATT_BK);
                    tma_load_2d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[slot], 2 * H + w.head * ATT_D, w.start + j * ATT_BK);
                    if (++slot == A4_KV_STAGES) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 3) {
        // ------------------------------------------ bias-tile writer ------------------------------------------
        // K-major, 128-B swizzle, one 128 x 64 bf16 tile per buffer: columns [0,16) key side {jrel x3, 0...}, [16,32) query side
        // {+sigma hi, mid, lo, 0...}, [32,48) {-sigma hi, mid, lo, 0...}; 16-B chunk c of row r lives at chunk position c ^ (r & 7).
        int itc = 0;
        for (int t = blockIdx.x; t < n_items; t += G, ++itc) {
            const int eb = itc & 1;
            const float slope2 = slopes[t % nheads] * LOG2E;
            mbar_wait(&ext_empty[eb], ((itc >> 1) & 1) ^ 1);
            const __nv_bfloat16 hi = __float2bfloat16_rn(slope2);
            const float rem1 = slope2 - __bfloat162float(hi);
            const __nv_bfloat16 mid = __float2bfloat16_rn(rem1);
            const __nv_bfloat16 lo = __float2bfloat16_rn(rem1 - __bfloat162float(mid));
            auto bits = [](__nv_bfloat16 v) { return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v)); };
            const uint4 zero = make_uint4(0, 0, 0, 0);
            const uint4 cp = make_uint4(bits(hi) | (bits(mid) << 16), bits(lo), 0, 0);
            const uint4 cn = make_uint4(cp.x ^ 0x80008000u, cp.y ^ 0x8000u, 0, 0);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = lane + 32 * rr;
                const __nv_bfloat16 jr = __float2bfloat16_rn(static_cast<float>(r));
                const uint4 ck = make_uint4(bits(jr) | (bits(jr) << 16), bits(jr), 0, 0);
                uint8_t* rowp = smem + A6_SMEM_EXT + eb * ATT_TILE_BYTES + r * 128;
                const int x7 = r & 7;
                *reinterpret_cast<uint4*>(rowp + ((0 ^ x7) << 4)) = ck;
                *reinterpret_cast<uint4*>(rowp + ((1 ^ x7) << 4)) = zero;
                *reinterpret_cast<uint4*>(rowp + ((2 ^ x7) << 4)) = cp;
                *reinterpret_cast<uint4*>(rowp + ((3 ^ x7) << 4)) = zero;
                *reinterpret_cast<uint4*>(rowp + ((4 ^ x7) << 4)) = cn;
                *reinterpret_cast<uint4*>(rowp + ((5 ^ x7) << 4)) = zero;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ext_full[eb]);
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------ PV issuer (converged warp) ------------------------------------------
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);          // B (= V) MN-major
        const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_KV + ATT_TILE_BYTES), 8192, 1024);
        int buf = 0, slot = 0, cntA = 0, cntB = 0;          // cntX: items in which half X has taken part so far
        int pvA = 0, pvB = 0;                               // PV MMAs issued so far per half
        uint32_t bpar = 0;
        A6Item w = a6_load_item(work, blockIdx.x, nheads);
        for (int t = blockIdx.x; t < n_items; t += G) {
            A6Item wn = w;
            if (t + G < n_items) wn = a6_load_item(work, t + G, nheads);                 // prefetch the next item's parameters
            for (int j = 0; j < w.nkb; ++j) {
                for (int half = 0; half < w.nh; ++half) {
                    mbar_wait(&p_ready[buf], bpar);
                    if (j == 0) {                                                         // this PV overwrites O: the previous item's
                        const int cnt = half ? cntB : cntA;                               // epilogue must have read it out
                        if (cnt > 0) mbar_wait(&o_free[half], (cnt - 1) & 1);
                        if (half) ++cntB; else ++cntA;
                    }
                    tc_fence_after();
                    const uint32_t tP = tmem_base + buf * 128;
                    const uint32_t tO = tmem_base + 384 + half * 64;
                    const uint64_t dv = dv0 + static_cast<uint64_t>(slot * (2 * ATT_TILE_BYTES >> 4));
                    const int pvc = half ? pvB : pvA;
                    if (half) ++pvB; else ++pvA;
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < ATT_BK / 16; ++k)                             // V rows [16k, 16k+16); P: 8 columns per K step
                            umma_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc_pv, (j | k) != 0);
                        umma_commit(&pv_done[2 * half + (pvc & 1)]);
                        umma_commit(&s_free[buf]);
                        if (half == w.nh - 1) umma_commit(&kv_empty[slot]);               // every MMA reading this stage has been issued
                    }
                    __syncwarp();
                    if (++buf == 3) { buf = 0; bpar ^= 1; }
                }
                if (++slot == A4_KV_STAGES) slot = 0;
            }
            w = wn;
        }
    } else if (warp_idx == 2) {
        // ------------------------------------------ QK issuer (converged warp) ------------------------------------------
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                      // A, B K-major
        const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_Q), 16, 1024);
        const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_KV), 16, 1024);
        const uint64_t de0 = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_EXT), 16, 1024);
        int buf = 0, slot = 0, itc = 0;
        uint32_t fpar = 1, kvpar = 0;                       // s_free parity: the first lap passes on the fresh barriers
        A6Item w = a6_load_item(work, blockIdx.x, nheads);
        for (int t = blockIdx.x; t < n_items; t += G, ++itc) {
            A6Item wn = w;
            if (t + G < n_items) wn = a6_load_item(work, t + G, nheads);
            const int qb = itc & 1;
            mbar_wait(&q_full[qb], (itc >> 1) & 1);
            mbar_wait(&ext_full[qb], (itc >> 1) & 1);
            const uint64_t dext = de0 + static_cast<uint64_t>(qb * (ATT_TILE_BYTES >> 4));
            for (int j = 0; j < w.nkb; ++j) {
                mbar_wait(&kv_full[slot], kvpar);
                for (int half = 0; half < w.nh; ++half) {
                    mbar_wait(&s_free[buf], fpar);          // the PV that consumed this buffer's previous P has retired
                    tc_fence_after();
                    const uint32_t d = tmem_base + buf * 128;
                    const uint64_t da = dq0 + static_cast<uint64_t>((qb * 2 + half) * (ATT_TILE_BYTES >> 4));
                    const uint64_t db = dk0 + static_cast<uint64_t>(slot * (2 * ATT_TILE_BYTES >> 4));
                    const int jd = w.qb0 + half;
                    const bool last = (j == w.nkb - 1) && (half == w.nh - 1);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                        if (j != jd) umma_ss(d, dext + (j < jd ? 2 : 4), dext, idesc_qk, 1);   // +sigma*jrel left of the diagonal, -sigma*jrel right
                        umma_commit(&s_full[buf]);
                        if (last) {                                                       // last QK of the item: Q and bias buffers are free
                            umma_commit(&q_empty[qb]);
                            umma_commit(&ext_empty[qb]);
                        }
                    }
                    __syncwarp();
                    if (++buf == 3) { buf = 0; fpar ^= 1; }
                }
                if (++slot == A4_KV_STAGES) { slot = 0; kvpar ^= 1; }
            }
            w = wn;
        }
    }
    } else {
        // ------------------------------ softmax + epilogue: one thread per query row ------------------------------
        asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
        const int half = (warp_idx - 4) >> 2;
        const int quad = warp_idx & 3;
        const int r = quad * 32 + lane;                     // row inside the half == TMEM lane
        const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
        const uint32_t tmem_O = tmem_base + 384 + half * 64 + lane_off;
        int gbase = 0;                                      // global step counter at the start of the item
        int pvbase = 0;                                     // PV MMAs of this half issued before the item
        A6Item wnext = a6_load_item(work, blockIdx.x, nheads);
        float slope_next = slopes[wnext.head];
        for (int t = blockIdx.x; t < n_items; t += G) {
            const A6Item w = wnext;
            const float slope2 = slope_next * LOG2E;
            if (t + G < n_items) {                          // prefetch the next item's parameters: consumed one item later
                wnext = a6_load_item(work, t + G, nheads);
                slope_next = slopes[wnext.head];
            }
            const int nkb = w.nkb, len = w.len;
            if (half == 0 || w.two) {
            const int qt = w.q0 + half * 128;               // first query of this half
            const int jd = w.qb0 + half;                    // diagonal key block
            const int i = qt + r;                           // query position in the sequence
            const float a0d = -slope2 * static_cast<float>(r);   // diagonal block: -sigma |r - e| = min(A, -A), A = sigma e - sigma r
            float m_ref = 0.f, l_run = 0.f;
            // PV(c) of this half retired?  Two alternating barriers per half make this wait unambiguous without observing every
            // phase: the scores of the current block exist, hence (s_free chain, in-order PV issue) PV(c - 2) of this half has
            // retired, hence barrier [c & 1] is in the phase of PV(c) or past it.
            auto wait_pv = [&](int c) { mbar_wait(&pv_done[2 * half + (c & 1)], (c >> 1) & 1); };
            // s[0,64): bf16 pairs of P (output); fast path: s[64,96) / s[96,128) double-buffer 32-column chunks of S.  The first chunk
            // of block j + 1 is requested before P of block j is stored (the scores are there already), so its latency and the
            // barrier round trip overlap the store / hand-over of block j.
            uint32_t s[128];
            uint32_t(&ca)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
            uint32_t(&cb)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
            for (int j = 0; j < nkb; ++j) {
                const int ng = gbase + (w.two ? 2 * j + half : j);
                    const int k0 = j * ATT_BK;
                    const int buf = ng % 3;
                    const uint32_t tmem_S = tmem_base + buf * 128 + lane_off;
                    const bool diag = (j == jd);
                    const bool partial = (k0 + ATT_BK > len);
                    // additive term of this block: per-row constant (off-diagonal; the j-dependent part came out of the MMA) or none
                    const float c_row = diag ? 0.f : -slope2 * fabsf(static_cast<float>(i - k0));
                    if (j == 0) {                            // later blocks: waited for and first chunk requested at the end of block j - 1
                        mbar_wait(&s_full[buf], (ng / 3) & 1);
                        tc_fence_after();
                    }
                    float lsum = 0.f;
                    bool need_full = (j == 0);
                    if (j > 0) {
                        // ---- fast path: exponentiate against the standing reference maximum while the next chunk streams out of
                        // TMEM (the TMEM read port and MUFU are both near saturation: neither may wait for the other); the block
                        // maximum is only checked afterwards.
                        const float off = c_row - m_ref;
                        const uint64_t off2 = pack2f(off, off);
                        uint64_t acc2 = pack2f(0.f, 0.f), accp2 = pack2f(0.f, 0.f);
                        float pace_a = 0.f, pace_b = 0.f;
#ifdef ATT8_MAX_GUARD
                        float mxa = -INFINITY, mxb = -INFINITY;
#else
                        float mxp = -INFINITY;               // largest argument handed to the polynomial (see the guard below)
#endif
                        auto proc = [&](uint32_t(&cur)[32], const int c) {
                            if (diag) {
                                const float s0raw = __uint_as_float(cur[0]);
                                float a0v = a0d;
                                asm volatile("" : "+f"(a0v));                   // keep the 128 bias values from being hoisted out of the block loop
#pragma unroll
                                for (int e = 0; e < 32; ++e) {
                                    const float a = fmaf(slope2, static_cast<float>(c * 32 + e), a0v);
                                    cur[e] = __float_as_uint(__uint_as_float(cur[e]) + fminf(a, -a));
                                }
                                if (c == 0 && k0 == 0) cur[0] = __float_as_uint(s0raw);     // key 0 (time token): no distance penalty
                            } else if (c == 0 && k0 == 0) {
                                cur[0] = __float_as_uint(__uint_as_float(cur[0]) - c_row);
                            }
                            if (partial) {
#pragma unroll
                                for (int e = 0; e < 32; ++e)
                                    if (k0 + c * 32 + e >= len) cur[e] = __float_as_uint(-INFINITY);
                            }
#pragma unroll
                            for (int u = 0; u < 16; ++u) {
                                const float x0 = __uint_as_float(cur[2 * u]), x1 = __uint_as_float(cur[2 * u + 1]);
#ifdef ATT8_MAX_GUARD
                                if (u & 1) mxb = fmaxf(mxb, fmaxf(x0, x1));
                                else mxa = fmaxf(mxa, fmaxf(x0, x1));
#endif
                                const uint64_t x2 = fadd2(pack2f(x0, x1), off2);
                                uint64_t p2;
                                if (POLY > 0 && (u % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
#ifndef ATT8_MAX_GUARD
                                    {
                                        float g0, g1;
                                        unpack2f(x2, g0, g1);
                                        mxp = fmaxf(mxp, fmaxf(g0, g1));
                                    }
#endif
                                    p2 = exp2_poly2(x2);
                                    accp2 = fadd2(accp2, p2);               // polynomial pairs: own accumulator, off the pacing chain
                                } else {
                                    float lo, hi;
                                    unpack2f(x2, lo, hi);
                                    if constexpr (PACE) {
                                        ex2_pair_paced(lo, hi, pace_a);
                                    } else {
                                        lo = ex2_approx(lo);
                                        hi = ex2_approx(hi);
                                    }
                                    p2 = pack2f(lo, hi);
                                    acc2 = fadd2(acc2, p2);
                                    if constexpr (PACE) {
                                        float al, ah;
                                        unpack2f(acc2, al, ah);
                                        pace_a = pace_b;                    // the MUFU pair after the next one waits for this sum
                                        pace_b = al;
                                    }
                                }
                                float lo, hi;
                                unpack2f(p2, lo, hi);
                                s[c * 16 + u] = pack_bf16x2(lo, hi);
                            }
                        };
                        tmem_ld_wait_dep(ca);                // chunk 0 was requested at the end of the previous block
                        tmem_ld_32x32(tmem_S + 32, cb);
                        proc(ca, 0);
                        tmem_ld_wait_dep(cb);
                        tmem_ld_32x32(tmem_S + 64, ca);
                        proc(cb, 1);
                        tmem_ld_wait_dep(ca);
                        tmem_ld_32x32(tmem_S + 96, cb);
                        proc(ca, 2);
                        tmem_ld_wait_dep(cb);
                        proc(cb, 3);
                        float a0, a1;
                        unpack2f(fadd2(acc2, accp2), a0, a1);
                        lsum = a0 + a1;
                        // P must stay <= 2^16: if any row of the warp outgrew its reference, redo the block the careful way (the scores are
                        // still in TMEM: P is stored only after this check).
#ifdef ATT8_MAX_GUARD
                        const float m_blk = fmaxf(mxa, mxb) + c_row;          // block maximum in the true (biased) domain
                        need_full = __any_sync(0xffffffffu, m_blk > m_ref + A4_RESCALE_LOG2);
#else
                        // The row sum of the block is the guard: every p is >= 0, so sum < 2^16 implies max p < 2^16, and an overflowed
                        // MUFU exponential (inf / NaN) fails the comparison too.  The FMA-pipe polynomial is only valid up to 2^126 and
                        // would wrap silently beyond, so the pairs it handles (one in POLY) keep a running maximum of their arguments.
                        // This drops two thirds of the fast path's FMNMX3 (one per pair of scores before); a false alarm only costs
                        // the careful path, which then keeps the reference.
                        need_full = __any_sync(0xffffffffu, !(lsum < 65536.0f) || mxp > A4_RESCALE_LOG2);
#endif
                    }
                    bool pv_seen = false;
                    if (need_full) {
                        // ---- full path (first block, or a reference maximum has to move): whole row in registers, maximum first
#pragma unroll
                        for (int c = 0; c < 4; ++c) tmem_ld_32x32(tmem_S + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
                        tmem_ld_wait();
                        if (diag) {
                            const float s0raw = __uint_as_float(s[0]);
                            float a0v = a0d;
                            asm volatile("" : "+f"(a0v));
#pragma unroll
                            for (int e = 0; e < 128; ++e) {
                                const float a = fmaf(slope2, static_cast<float>(e), a0v);
                                s[e] = __float_as_uint(__uint_as_float(s[e]) + fminf(a, -a));
                            }
                            if (k0 == 0) s[0] = __float_as_uint(s0raw);
                        } else if (k0 == 0) {
                            s[0] = __float_as_uint(__uint_as_float(s[0]) - c_row);
                        }
                        if (partial) {
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
                        const float m_blk = fmaxf(mx0, mx1) + c_row;
                        if (j == 0) {
                            m_ref = m_blk;
                        } else {
                            const float m_new = (m_blk > m_ref + A4_RESCALE_LOG2) ? m_blk : m_ref;
                            const float alpha = ex2_approx(m_ref - m_new);    // 1 for rows whose reference stays
                            m_ref = m_new;
                            l_run *= alpha;
                            wait_pv(pvbase + j - 1);                                  // O is only stable once the previous PV has retired
                            pv_seen = true;
                            tc_fence_after();
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
                        const float off = c_row - m_ref;
                        const uint64_t off2 = pack2f(off, off);
                        uint64_t acc2 = pack2f(0.f, 0.f);
#pragma unroll
                        for (int u = 0; u < 64; ++u) {
                            const uint64_t x2 = fadd2(pack2f(__uint_as_float(s[2 * u]), __uint_as_float(s[2 * u + 1])), off2);
                            float lo, hi;
                            unpack2f(x2, lo, hi);
                            lo = ex2_approx(lo);
                            hi = ex2_approx(hi);
                            acc2 = fadd2(acc2, pack2f(lo, hi));
                            s[u] = pack_bf16x2(lo, hi);
                        }
                        float a0, a1;
                        unpack2f(acc2, a0, a1);
                        lsum = a0 + a1;
                    }
                    l_run += lsum;
                    if (j + 1 < nkb) {                       // scores of the next block: wait, then request their first 32 columns
                        const int ng1 = ng + w.nh;
                        mbar_wait(&s_full[ng1 % 3], (ng1 / 3) & 1);
                        tc_fence_after();
                        tmem_ld_32x32(tmem_base + (ng1 % 3) * 128 + lane_off, ca);
                    }
                    // ---- bf16 pairs of P into the first 64 columns of this S buffer (A operand of the PV MMA)
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_S + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                    tmem_st_wait();
                    tc_fence_before();
                    mbar_arrive(&p_ready[buf]);

            }
            // An item whose second half has no rows is skipped by the half-B threads.  They must not run ahead to the s_full wait
            // of their next step while this item's QK MMAs are still outstanding (parity waits are only unambiguous within one
            // phase): they wait on a named barrier that the half-A threads reach after observing the item's last scores.
            if (!w.two) asm volatile("bar.arrive 1, 256;" ::: "memory");
            // epilogue: read O, hand the accumulator back, then O / l -> ctx (bf16), only rows that belong to the sequence
            if (nkb >= 2) wait_pv(pvbase + nkb - 2);       // (one-half items: makes the next wait unambiguous, see wait_pv)
            wait_pv(pvbase + nkb - 1);
            tc_fence_after();
            uint32_t o[64];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld_32x32(tmem_O + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&o[c * 32]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_free[half]);
            const float inv = 1.0f / l_run;
            if (i < len) {
                __nv_bfloat16* dst = ctx + static_cast<size_t>(w.start + i) * H + w.head * ATT_D;
#pragma unroll
                for (int c = 0; c < 8; ++c)
                    reinterpret_cast<uint4*>(dst)[c] =
                        make_uint4(pack_bf16x2(__uint_as_float(o[8 * c]) * inv, __uint_as_float(o[8 * c + 1]) * inv),
                                   pack_bf16x2(__uint_as_float(o[8 * c + 2]) * inv, __uint_as_float(o[8 * c + 3]) * inv),
                                   pack_bf16x2(__uint_as_float(o[8 * c + 4]) * inv, __uint_as_float(o[8 * c + 5]) * inv),
                                   pack_bf16x2(__uint_as_float(o[8 * c + 6]) * inv, __uint_as_float(o[8 * c + 7]) * inv));
            }
            pvbase += nkb;
            } else {
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            gbase += nkb * w.nh;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

template <int POLY, bool PACE>
inline cudaError_t launch_attn8_t(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  int nwork, int num_sms, cudaStream_t st, bool pdl) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn8_kernel<POLY, PACE>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int n_items = nwork * nheads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ATT_THREADS); cfg.dynamicSmemBytes = A6_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, attn8_kernel<POLY, PACE>, tmQKV, ctx, H, nheads, slopes, work, n_items);
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e).
inline cudaError_t launch_attn8_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                    const int4* work, int nwork, int poly, int num_sms, cudaStream_t st, bool pace = true, bool pdl = false) {
    if (pace) switch (poly) {
        case 0: return launch_attn8_t<0, true>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        case 2: return launch_attn8_t<2, true>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        case 3: return launch_attn8_t<3, true>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        case 4: return launch_attn8_t<4, true>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        default: return cudaErrorInvalidValue;
    }
    switch (poly) {
        case 0: return launch_attn8_t<0, false>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        case 2: return launch_attn8_t<2, false>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        case 3: return launch_attn8_t<3, false>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        case 4: return launch_attn8_t<4, false>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st, pdl);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
