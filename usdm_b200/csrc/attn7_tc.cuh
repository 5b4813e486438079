// Fused masked self-attention, version 7: attn4_tc.cuh with the MMA issue split over two converged warps.
//
// The clock64 trace of v4 (profiles/r01/attn4_trace.txt) shows its steady state (3.25 k cycles per key block for both halves)
// is set by the single MMA-issuing thread: it needs ~1.6 k cycles per step (8 PV + 5 QK tcgen05.mma, three commits, two
// barrier waits, descriptor arithmetic done in a divergent single-thread region, hence outside the uniform datapath) and
// receives every P tile ~1 k cycles after the softmax threads finished it.  Here
//   * warp 1 issues the PV MMAs, warp 2 (after allocating TMEM) the QK^T MMAs: the two instruction streams run in parallel;
//   * both warps execute their loops converged (all 32 lanes; descriptors are base + counter * constant in uniform registers)
//     and only the tcgen05 instructions sit under elect.sync;
//   * the S-buffer hand-back that v4 got from in-order issue of one thread (QK of step n+3 after PV of step n) is now an explicit
//     barrier s_free[buffer], committed after the PV that read P from the buffer.
// Softmax threads, TMEM map and everything else are v4's.
#pragma once
#include "attn4_tc.cuh"

namespace tvb {

// POLY: 0 = all exponentials on MUFU; n > 0 = one element pair in every n goes to the FMA-pipe polynomial.
template <int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn7_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
             const float* __restrict__ slopes, const int4* __restrict__ work, long long* __restrict__ trace) {
    // trace (test hook, normally nullptr): clock64() stamps of the first 16 CTAs, 256 slots each: [0] entry, [1] set-up done,
    // half A [16+2j] scores of block j visible / [17+2j] P of block j handed over, [80] output stored; half B the same at +80;
    // MMA thread [176] Q landed, [177+n] P of step n received.
    long long* tr = (trace != nullptr && blockIdx.x < 16) ? trace + blockIdx.x * 256 : nullptr;
    long long* trf = tr ? trace + 16 * 256 + blockIdx.x * 256 : nullptr;   // fine stamps of the softmax row 0: [half * 128 + 4 j + k]
    if (tr && threadIdx.x == 128) tr[0] = clock64();
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A4_SMEM_BAR);
    uint64_t* q_full = bars + 0;
    uint64_t* kv_full = bars + 1;                       // [4]
    uint64_t* kv_empty = kv_full + A4_KV_STAGES;        // [4]
    uint64_t* s_full = kv_empty + A4_KV_STAGES;         // [3] per S buffer
    uint64_t* p_ready = s_full + 3;                     // [3] per S buffer (a warp may run one block ahead of its half: per-half
                                                        //     counting would mix the arrivals of two blocks)
    uint64_t* pv_done = p_ready + 3;                    // [2] per half
    uint64_t* s_free = pv_done + 2;                     // [3] per S buffer: the PV MMA that read P from it has retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 3);

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
        for (int s = 0; s < 3; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 128); mbar_init(&s_free[s], 1); }
        for (int t = 0; t < 2; ++t) mbar_init(&pv_done[t], 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
    if (warp_idx >= 4 && warp_idx < 8) {
        // Bias operand tile (K-major, 128-B swizzle, one 128 x 64 bf16 tile): columns [0,16) key side {jrel, jrel, jrel, 0...},
        // [16,32) query side {+sigma hi, mid, lo, 0...}, [32,48) query side {-sigma hi, mid, lo, 0...}; 16-B chunk c of row r
        // lives at chunk position c ^ (r & 7).
        const int r = threadIdx.x - 128;
        const __nv_bfloat16 hi = __float2bfloat16_rn(slope2);
        const float rem1 = slope2 - __bfloat162float(hi);
        const __nv_bfloat16 mid = __float2bfloat16_rn(rem1);
        const __nv_bfloat16 lo = __float2bfloat16_rn(rem1 - __bfloat162float(mid));
        const __nv_bfloat16 jr = __float2bfloat16_rn(static_cast<float>(r));
        auto bits = [](__nv_bfloat16 v) { return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v)); };
        uint8_t* rowp = smem + A4_SMEM_EXT + r * 128;
        const uint4 zero = make_uint4(0, 0, 0, 0);
        const uint32_t sgn = 0x80008000u;
        const uint4 ck = make_uint4(bits(jr) | (bits(jr) << 16), bits(jr), 0, 0);
        const uint4 cp = make_uint4(bits(hi) | (bits(mid) << 16), bits(lo), 0, 0);
        const uint4 cn = make_uint4(cp.x ^ sgn, cp.y ^ 0x8000u, 0, 0);
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
    if (tr && threadIdx.x == 128) tr[1] = clock64();

    if (warp_idx < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp_idx == 0) {
        // ------------------------------------------ TMA producer ------------------------------------------
        if (elect_one()) {
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
        // ------------------------------------------ PV issuer ------------------------------------------
        // The whole warp runs the loop converged (addresses and descriptors stay in uniform registers); one elected lane
        // issues.  Measured on v4 (profiles/r01/attn4_trace.txt): ONE thread issuing both GEMMs of a step needed ~1.6 k cycles
        // per step and received every P ~1 k cycles late - the issue thread, not MUFU, set the steady state.
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);          // B (= V) MN-major
        const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem + A4_SMEM_KV + ATT_TILE_BYTES), 8192, 1024);
        int buf = 0, slot = 0;
        uint32_t bpar = 0;
        for (int j = 0; j < nkb; ++j) {
            for (int half = 0; half < nh; ++half) {
                mbar_wait(&p_ready[buf], bpar);
                tc_fence_after();
                if (tr && lane == 0 && j * nh + half < 70) tr[177 + j * nh + half] = clock64();
                const uint32_t tP = tmem_base + buf * 128;
                const uint32_t tO = tmem_base + 384 + half * 64;
                const uint64_t dv = dv0 + static_cast<uint64_t>(slot * (2 * ATT_TILE_BYTES >> 4));
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < ATT_BK / 16; ++k)                                 // V rows [16k, 16k+16); P: 8 columns per K step
                        umma_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc_pv, (j | k) != 0);
                    umma_commit(&pv_done[half]);
                    umma_commit(&s_free[buf]);
                    if (half == nh - 1) umma_commit(&kv_empty[slot]);                     // every MMA reading this stage has been issued
                }
                __syncwarp();
                if (++buf == 3) { buf = 0; bpar ^= 1; }
            }
            if (++slot == A4_KV_STAGES) slot = 0;
        }
    } else if (warp_idx == 2) {
        // ------------------------------------------ QK issuer ------------------------------------------
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                      // A, B K-major
        const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem + A4_SMEM_Q), 16, 1024);
        const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem + A4_SMEM_KV), 16, 1024);
        const uint64_t dext = make_smem_desc_sw128(smem_u32(smem + A4_SMEM_EXT), 16, 1024);
        mbar_wait(q_full, 0);
        int buf = 0, slot = 0, n = 0;
        uint32_t fpar = 1, kvpar = 0;                       // s_free parity: first lap passes on the fresh barriers
        for (int j = 0; j < nkb; ++j) {
            mbar_wait(&kv_full[slot], kvpar);
            for (int half = 0; half < nh; ++half, ++n) {
                mbar_wait(&s_free[buf], fpar);              // the PV that consumed this buffer's previous P has retired
                tc_fence_after();
                const uint32_t d = tmem_base + buf * 128;
                const uint64_t da = dq0 + static_cast<uint64_t>(half * (ATT_TILE_BYTES >> 4));
                const uint64_t db = dk0 + static_cast<uint64_t>(slot * (2 * ATT_TILE_BYTES >> 4));
                const int jd = qb0 + half;
                if (elect_one()) {
#pragma unroll
                    for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                    if (j != jd) umma_ss(d, dext + (j < jd ? 2 : 4), dext, idesc_qk, 1);   // +sigma*jrel left of the diagonal, -sigma*jrel right
                    umma_commit(&s_full[buf]);
                }
                __syncwarp();
                if (++buf == 3) { buf = 0; fpar ^= 1; }
            }
            if (++slot == A4_KV_STAGES) { slot = 0; kvpar ^= 1; }
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
            const int jd = qb0 + half;                      // diagonal key block
            const int i = qt + r;                           // query position in the sequence
            const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
            const uint32_t tmem_O = tmem_base + 384 + half * 64 + lane_off;
            const float a0d = -slope2 * static_cast<float>(r);   // diagonal block: -sigma |r - e| = min(A, -A), A = sigma e - sigma r
            float m_ref = 0.f, l_run = 0.f;
            for (int j = 0; j < nkb; ++j) {
                const int n = two ? 2 * j + half : j;
                const int k0 = j * ATT_BK;
                const int buf = n % 3;
                const uint32_t tmem_S = tmem_base + buf * 128 + lane_off;
                const bool diag = (j == jd);
                const bool partial = (k0 + ATT_BK > len);
                // additive term of this block: per-row constant (off-diagonal; the j-dependent part came out of the MMA) or none
                const float c_row = diag ? 0.f : -slope2 * fabsf(static_cast<float>(i - k0));
                mbar_wait(&s_full[buf], (n / 3) & 1);
                tc_fence_after();
                if (tr && r == 0 && j < 32) tr[16 + half * 80 + 2 * j] = clock64();
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
                            if (POLY >= 100 && (POLY & 1)) {
                                p2 = x2;                                                 // timing ablation: no exponential at all
                            } else if (POLY > 0 && POLY < 100 && (u % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
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
                    if (POLY >= 100 && (POLY & 2)) {                                  // timing ablation: scores never leave TMEM
#pragma unroll
                        for (int e = 0; e < 32; ++e) { ca[e] = __float_as_uint(-1.0f - e); cb[e] = __float_as_uint(-2.0f - e); }
                        asm volatile("" : "+r"(ca[0]), "+r"(cb[0]));
                        proc(ca, 0); proc(cb, 1); proc(ca, 2); proc(cb, 3);
                    } else {
                    tmem_ld_32x32(tmem_S, ca);
                    tmem_ld_wait_dep(ca);
                    if (trf && r == 0 && j < 32) trf[half * 128 + 4 * j + 0] = clock64();     // first 32 score columns in registers
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
                    }
                    if (trf && r == 0 && j < 32) trf[half * 128 + 4 * j + 1] = clock64();     // exponentials / sums / bf16 packing done
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
                        mbar_wait(&pv_done[half], (j - 1) & 1);          // O is only stable once the previous PV has retired
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
                if (j > 0 && !pv_seen) mbar_wait(&pv_done[half], (j - 1) & 1);
                if (trf && r == 0 && j < 32) trf[half * 128 + 4 * j + 2] = clock64();         // previous PV observed
                l_run += lsum;
                // ---- bf16 pairs of P into the first 64 columns of this S buffer (A operand of the PV MMA)
                if (!(POLY >= 100 && (POLY & 4))) {                                   // (timing ablation 104: P never stored)
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_S + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                tmem_st_wait();
                } else { asm volatile("" :: "r"(s[0]), "r"(s[17]), "r"(s[34]), "r"(s[63])); }
                if (trf && r == 0 && j < 32) trf[half * 128 + 4 * j + 3] = clock64();         // P stored to TMEM
                tc_fence_before();
                mbar_arrive(&p_ready[buf]);
                if (tr && r == 0 && j < 32) tr[17 + half * 80 + 2 * j] = clock64();
            }
            // epilogue: O / l -> ctx (bf16), only rows that belong to the sequence
            mbar_wait(&pv_done[half], (nkb - 1) & 1);
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
            if (tr && r == 0) tr[80 + half * 80] = clock64();
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
inline cudaError_t launch_attn7_t(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  int nwork, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn7_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    attn7_kernel<POLY><<<nwork * nheads, ATT_THREADS, A4_SMEM_BYTES, st>>>(tmQKV, ctx, H, nheads, slopes, work, attn_trace_ptr());
    return cudaGetLastError();
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e).
inline cudaError_t launch_attn7_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                    const int4* work, int nwork, int poly, cudaStream_t st) {
    switch (poly) {
        case 0: return launch_attn7_t<0>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 2: return launch_attn7_t<2>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 3: return launch_attn7_t<3>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 4: return launch_attn7_t<4>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        // timing ablations (results are WRONG; tools/ only): 101 no exp2, 102 no TMEM score loads, 104 no P store, 107 all three
        case 101: return launch_attn7_t<101>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 102: return launch_attn7_t<102>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 104: return launch_attn7_t<104>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 107: return launch_attn7_t<107>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
