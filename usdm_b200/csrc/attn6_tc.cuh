// Fused masked self-attention, version 6: attn4_tc.cuh made PERSISTENT.
//
// Why (clock64 trace of v4, profiles/r01/attn4_trace.txt): one v4 CTA lives ~40.5 k cycles for a 9-block work item, of which
// ~6.4 k pass before the first scores exist (TMEM allocation, barrier set-up, then ~4.3 k cycles waiting for the first Q / K
// tiles from a cold start) and ~4.5 k after the last softmax (final PV, normalise, store), with one CTA per SM and
// nothing to overlap them; 2560 equal items on 148 SMs add a 4 % tail.  Here one CTA per SM walks a static round-robin list
// of (sequence, head, 256-query tile) items; every role keeps running across item boundaries:
//   * the TMA producer streams Q (double-buffered) and the K/V ring of the next item while the current one is computed;
//   * the MMA thread issues the first QK^T steps of the next item before the last PV steps of the current one (the three
//     rotating S buffers and all barriers are indexed by counters that run across items);
//   * warp 3 rewrites the head-dependent bias operand tile (double-buffered) one item ahead;
//   * the softmax threads hand O back (o_free) as soon as they have read it, and then normalise / store while the tensor
//     core already accumulates the next item.
// Everything else (bias by the tensor core, lazy reference maximum, P in TMEM, barrier protocol) is attn4_tc.cuh.
// TMEM (512 columns): S buffers [0,128) [128,256) [256,384), O_A [384,448), O_B [448,512).
#pragma once
#include "attn4_tc.cuh"

namespace tvb {

constexpr int A6_SMEM_Q = 0;                                                   // 2 items x 2 halves x 16 KB
constexpr int A6_SMEM_KV = 4 * ATT_TILE_BYTES;                                 // stages x (K, V)
constexpr int A6_SMEM_EXT = A6_SMEM_KV + A4_KV_STAGES * 2 * ATT_TILE_BYTES;    // 2 x 16 KB bias operand tiles
constexpr int A6_SMEM_BAR = A6_SMEM_EXT + 2 * ATT_TILE_BYTES;
constexpr int A6_SMEM_BYTES = A6_SMEM_BAR + 256;

struct A6Item {           // one work item as every role sees it
    int q0, start, len, nkb, nh, qb0, head;
    bool two;
};
__device__ __forceinline__ A6Item a6_load_item(const int4* __restrict__ work, int t, int nheads) {
    const int4 w = work[t / nheads];
    A6Item it;
    it.head = t % nheads;
    it.q0 = w.y; it.start = w.z; it.len = w.w;
    it.nkb = (w.w + ATT_BK - 1) / ATT_BK;
    it.two = (w.y + 128 < w.w);
    it.nh = it.two ? 2 : 1;
    it.qb0 = w.y / ATT_BK;
    return it;
}

template <int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn6_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
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
    uint64_t* pv_done = p_ready + 3;                    // [2] per half
    uint64_t* o_free = pv_done + 2;                     // [2] per half: O has been read out by the 128 softmax threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&ext_full[s], 1); mbar_init(&ext_empty[s], 1);
            mbar_init(&pv_done[s], 1); mbar_init(&o_free[s], 128);
        }
        for (int s = 0; s < A4_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int s = 0; s < 3; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 128); }
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
        // ------------------------------------------ MMA issuer ------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                  // A, B K-major
            constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);      // B (= V) MN-major
            // Two cursors walk the same flattened sequence of steps (item, key block, half): QK runs three steps ahead of PV.
            struct Cursor {
                int t, itc, n, total, kvbase;     // item id, local item count, step inside the item, steps of the item, first K/V block
                A6Item w;
            };
            auto open_item = [&](Cursor& c) {
                if (c.t < n_items) { c.w = a6_load_item(work, c.t, nheads); c.total = c.w.nkb * c.w.nh; c.n = 0; }
            };
            auto advance = [&](Cursor& c) {
                if (++c.n == c.total) { c.kvbase += c.w.nkb; c.t += G; ++c.itc; open_item(c); }
            };
            Cursor cq{static_cast<int>(blockIdx.x), 0, 0, 0, 0, {}}, cp{static_cast<int>(blockIdx.x), 0, 0, 0, 0, {}};
            open_item(cq);
            open_item(cp);
            int gq = 0, gp = 0;                      // global step counters (S buffer = g % 3)
            int ofree_cnt[2] = {0, 0};               // items in which each half has taken part so far
            auto issue_qk = [&]() {
                const A6Item& w = cq.w;
                const int half = w.two ? (cq.n & 1) : 0;
                const int j = w.two ? (cq.n >> 1) : cq.n;
                const int qb = cq.itc & 1;
                if (cq.n == 0) {                                                          // first step of an item: its Q and bias tile
                    mbar_wait(&q_full[qb], (cq.itc >> 1) & 1);
                    mbar_wait(&ext_full[qb], (cq.itc >> 1) & 1);
                    tc_fence_after();
                }
                const int kb = cq.kvbase + j;
                const int slot = kb % A4_KV_STAGES;
                if (half == 0) {                                                          // first user of this K/V stage
                    mbar_wait(&kv_full[slot], (kb / A4_KV_STAGES) & 1);
                    tc_fence_after();
                }
                const uint32_t d = tmem_base + (gq % 3) * 128;
                const uint64_t da = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_Q + (qb * 2 + half) * ATT_TILE_BYTES), 16, 1024);
                const uint64_t db = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_KV + slot * 2 * ATT_TILE_BYTES), 16, 1024);
                const uint64_t dext = make_smem_desc_sw128(smem_u32(smem + A6_SMEM_EXT + qb * ATT_TILE_BYTES), 16, 1024);
#pragma unroll
                for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                const int jd = w.qb0 + half;
                if (j != jd) umma_ss(d, dext + (j < jd ? 2 : 4), dext, idesc_qk, 1);       // +sigma*jrel left of the diagonal, -sigma*jrel right
                umma_commit(&s_full[gq % 3]);
                if (cq.n == cq.total - 1) {                                               // last QK of the item: Q and bias buffers are free
                    umma_commit(&q_empty[qb]);
                    umma_commit(&ext_empty[qb]);
                }
                ++gq;
                advance(cq);
            };
            auto issue_pv = [&]() {
                const A6Item& w = cp.w;
                const int half = w.two ? (cp.n & 1) : 0;
                const int j = w.two ? (cp.n >> 1) : cp.n;
                const int kb = cp.kvbase + j;
                const int slot = kb % A4_KV_STAGES;
                mbar_wait(&p_ready[gp % 3], (gp / 3) & 1);
                if (j == 0) {                                                             // this PV overwrites O: the previous item's
                    if (ofree_cnt[half] > 0) mbar_wait(&o_free[half], (ofree_cnt[half] - 1) & 1);   // epilogue must have read it
                    ++ofree_cnt[half];
                }
                tc_fence_after();
                const uint32_t sV = smem_u32(smem + A6_SMEM_KV + slot * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
                const uint32_t tP = tmem_base + (gp % 3) * 128;
                const uint32_t tO = tmem_base + 384 + half * 64;
#pragma unroll
                for (int k = 0; k < ATT_BK / 16; ++k) {
                    const uint64_t db = make_smem_desc_sw128(sV + k * 2048, 8192, 1024);   // V rows [16k, 16k+16)
                    umma_ts(tO, tP + k * 8, db, idesc_pv, (j | k) != 0);                   // P: 8 columns (16 bf16) per K step
                }
                umma_commit(&pv_done[half]);
                if (half == w.nh - 1) umma_commit(&kv_empty[slot]);                       // every MMA reading this stage has been issued
                ++gp;
                advance(cp);
            };
            for (int n = 0; n < 3 && cq.t < n_items; ++n) issue_qk();
            while (cp.t < n_items) {
                issue_pv();
                if (cq.t < n_items) issue_qk();
            }
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
        for (int t = blockIdx.x; t < n_items; t += G) {
            const A6Item w = a6_load_item(work, t, nheads);
            const int nkb = w.nkb, len = w.len;
            if (half == 0 || w.two) {
            const float slope2 = slopes[w.head] * LOG2E;
            const int qt = w.q0 + half * 128;               // first query of this half
            const int jd = w.qb0 + half;                    // diagonal key block
            const int i = qt + r;                           // query position in the sequence
            const float a0d = -slope2 * static_cast<float>(r);   // diagonal block: -sigma |r - e| = min(A, -A), A = sigma e - sigma r
            float m_ref = 0.f, l_run = 0.f;
            for (int j = 0; j < nkb; ++j) {
                const int ng = gbase + (w.two ? 2 * j + half : j);
                    const int k0 = j * ATT_BK;
                    const int buf = ng % 3;
                    const uint32_t tmem_S = tmem_base + buf * 128 + lane_off;
                    const bool diag = (j == jd);
                    const bool partial = (k0 + ATT_BK > len);
                    // additive term of this block: per-row constant (off-diagonal; the j-dependent part came out of the MMA) or none
                    const float c_row = diag ? 0.f : -slope2 * fabsf(static_cast<float>(i - k0));
                    mbar_wait(&s_full[buf], (ng / 3) & 1);
                    tc_fence_after();
                    // s[0,64): bf16 pairs of P (output); fast path: s[64,96) / s[96,128) double-buffer 32-column chunks of S
                    uint32_t s[128];
                    uint32_t(&ca)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
                    uint32_t(&cb)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
                    float lsum = 0.f;
                    bool need_full = (j == 0);
                    if (j > 0) {
                        // ---- fast path: exponentiate against the standing reference maximum while the next chunk streams out of
                        // TMEM (the TMEM read port and MUFU are both near saturation: neither may wait for the other); the block
                        // maximum is only checked afterwards.
                        const float off = c_row - m_ref;
                        const uint64_t off2 = pack2f(off, off);
                        uint64_t acc2 = pack2f(0.f, 0.f);
                        float mxa = -INFINITY, mxb = -INFINITY;
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
                                if (u & 1) mxb = fmaxf(mxb, fmaxf(x0, x1));
                                else mxa = fmaxf(mxa, fmaxf(x0, x1));
                                const uint64_t x2 = fadd2(pack2f(x0, x1), off2);
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
                                s[c * 16 + u] = pack_bf16x2(lo, hi);
                            }
                        };
                        tmem_ld_32x32(tmem_S, ca);
                        tmem_ld_wait_dep(ca);
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
                        unpack2f(acc2, a0, a1);
                        lsum = a0 + a1;
                        const float m_blk = fmaxf(mxa, mxb) + c_row;          // block maximum in the true (biased) domain
                        // P must stay <= 2^16: if any row of the warp outgrew its reference, redo the block the careful way
                        need_full = __any_sync(0xffffffffu, m_blk > m_ref + A4_RESCALE_LOG2);
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
                            mbar_wait(&pv_done[half], (pvbase + j - 1) & 1);          // O is only stable once the previous PV has retired
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
                    // Every thread observes every phase of pv_done exactly once and BEFORE it lets PV(j) start (parity waits are only
                    // unambiguous within one phase).  PV(j-1) was issued a whole block ago, so this does not stall.
                    if (j > 0 && !pv_seen) mbar_wait(&pv_done[half], (pvbase + j - 1) & 1);
                    l_run += lsum;
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
            mbar_wait(&pv_done[half], (pvbase + nkb - 1) & 1);
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

template <int POLY>
inline cudaError_t launch_attn6_t(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  int nwork, int num_sms, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn6_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, A6_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int n_items = nwork * nheads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    attn6_kernel<POLY><<<grid, ATT_THREADS, A6_SMEM_BYTES, st>>>(tmQKV, ctx, H, nheads, slopes, work, n_items);
    return cudaGetLastError();
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e).
inline cudaError_t launch_attn6_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                    const int4* work, int nwork, int poly, int num_sms, cudaStream_t st) {
    switch (poly) {
        case 0: return launch_attn6_t<0>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st);
        case 2: return launch_attn6_t<2>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st);
        case 3: return launch_attn6_t<3>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st);
        case 4: return launch_attn6_t<4>(tmQKV, ctx, H, nheads, slopes, work, nwork, num_sms, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
