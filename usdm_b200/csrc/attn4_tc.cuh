// Fused masked self-attention, version 4 (default).  Same math and CTA shape as attn_tc.cuh (one CTA = one sequence, one head,
// 256 queries as two halves of 128; thread == query row == TMEM lane; P kept in tensor memory), restructured after the
// round-1 ncu profile (profiles/r01): the v3 kernel ran at 2.2x its MUFU bound because (1) each half's
// softmax -> PV -> QK -> softmax chain exposed the MMA latency (27 % of softmax-warp samples waiting for S), (2) ~6
// issue slots per score element went to the bias/scale arithmetic, (3) the running-max rescale of O ran on most blocks.
//
//  1. THREE rotating S buffers in TMEM shared by the two halves: the QK MMA of step n+3 is issued right after the PV MMA of
//     step n, so the scores of a half's next key block are complete before its softmax threads finish the current one.
//  2. The ALiBi term is added BY THE TENSOR CORE: scores come out of the MMA already in the log2 domain (log2(e) and the
//     head_dim^-0.5 scale are folded into W_q at weight-finalise time) and one extra K=16 MMA step per off-diagonal block
//     multiplies a constant [+-sigma_hi, +-sigma_mid, +-sigma_lo] query-side tile with a constant [jrel, jrel, jrel]
//     key-side tile (jrel = 0..127 exact in bf16, sigma split into three bf16 terms), i.e. adds -+sigma*(j - k0) exactly;
//     the remaining per-row constant -+sigma*(i - k0) is folded into the exponent offset.  Only the single diagonal block
//     of each half evaluates |i - j| in registers.  Per element the softmax is: max, add, exp2, add, convert.
//  3. Lazy rescale: the reference maximum of a row only moves when the block maximum exceeds it by more than 2^16
//     (P stays <= 65536: harmless in bf16 / fp32), so the O read-modify-write is rare.
//  4. Optional exp2 offload: every POLY-th element pair is exponentiated on the FMA pipe (round-to-nearest split +
//     degree-3 polynomial, max rel. error 7.5e-5, far below the bf16 rounding of P) instead of MUFU.EX2.
//
// Reference semantics (networks.py:319-341 bias + mask, :174-193 scores/softmax/context): additive term for (query i, key j)
//   -slope_h |i - j| for j != 0, 0 for key 0 (time token), keys j >= len masked out.
// TMEM (512 columns): S buffers [0,128) [128,256) [256,384), O_A [384,448), O_B [448,512).
#pragma once
#include "attn_tc.cuh"

namespace tvb {

constexpr int A4_KV_STAGES = 4;
constexpr int A4_SMEM_Q = 0;                                                   // 2 x 16 KB
constexpr int A4_SMEM_KV = 2 * ATT_TILE_BYTES;                                 // stages x (K, V)
constexpr int A4_SMEM_EXT = A4_SMEM_KV + A4_KV_STAGES * 2 * ATT_TILE_BYTES;    // 16 KB: bias operand tile
constexpr int A4_SMEM_BAR = A4_SMEM_EXT + ATT_TILE_BYTES;
constexpr int A4_SMEM_BYTES = A4_SMEM_BAR + 256;
constexpr float A4_RESCALE_LOG2 = 16.0f;

// 2^x for x <= 0 on the FMA pipe, two elements at a time (Cody-Waite split at round-to-nearest, degree-3 minimax on [-.5,.5]).
__device__ __forceinline__ uint64_t exp2_poly2(uint64_t x2) {
    float x0, x1;
    unpack2f(x2, x0, x1);
    x0 = fmaxf(x0, -125.0f);
    x1 = fmaxf(x1, -125.0f);
    const uint64_t magic = pack2f(12582912.0f, 12582912.0f);                   // 1.5 * 2^23: x + magic rounds x to an integer
    const uint64_t xc = pack2f(x0, x1);
    const uint64_t xf = fadd2(xc, magic);
    const uint64_t xr = fadd2(xf, pack2f(-12582912.0f, -12582912.0f));
    uint64_t nxr;
    {
        float a, b;
        unpack2f(xr, a, b);
        nxr = pack2f(-a, -b);
    }
    const uint64_t f = fadd2(xc, nxr);                                         // in [-0.5, 0.5]
    uint64_t p = ffma2(pack2f(0.055171459913253784f, 0.055171459913253784f), f, pack2f(0.2426108568906784f, 0.2426108568906784f));
    p = ffma2(p, f, pack2f(0.6932609677314758f, 0.6932609677314758f));
    p = ffma2(p, f, pack2f(0.9999281167984009f, 0.9999281167984009f));
    float p0, p1, n0, n1;
    unpack2f(p, p0, p1);
    unpack2f(xf, n0, n1);
    // bits(xf) = bits(magic) + n and bits(magic) << 23 == 0 (mod 2^32): (bits(xf) << 23) + bits(p) scales p by 2^n
    const uint32_t r0 = (__float_as_uint(n0) << 23) + __float_as_uint(p0);
    const uint32_t r1 = (__float_as_uint(n1) << 23) + __float_as_uint(p1);
    return pack2f(__uint_as_float(r0), __uint_as_float(r1));
}

// POLY: 0 = all exponentials on MUFU; n > 0 = one element pair in every n goes to the FMA-pipe polynomial.
template <int POLY>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn4_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
             const float* __restrict__ slopes, const int4* __restrict__ work, long long* __restrict__ trace) {
    // trace (test hook, normally nullptr): clock64() stamps of the first 16 CTAs, 256 slots each: [0] entry, [1] set-up done,
    // half A [16+2j] scores of block j visible / [17+2j] P of block j handed over, [80] output stored; half B the same at +80;
    // MMA thread [176] Q landed, [177+n] P of step n received.
    long long* tr = (trace != nullptr && blockIdx.x < 16) ? trace + blockIdx.x * 256 : nullptr;
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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(pv_done + 2);

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
        for (int s = 0; s < 3; ++s) { mbar_init(&s_full[s], 1); mbar_init(&p_ready[s], 128); }
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
                if (tr && n < 70) tr[177 + n] = clock64();
                const uint32_t sV = smem_u32(smem + A4_SMEM_KV + slot * 2 * ATT_TILE_BYTES + ATT_TILE_BYTES);
                const uint32_t tP = tmem_base + (n % 3) * 128;
                const uint32_t tO = tmem_base + 384 + half * 64;
#pragma unroll
                for (int k = 0; k < ATT_BK / 16; ++k) {
                    const uint64_t db = make_smem_desc_sw128(sV + k * 2048, 8192, 1024);   // V rows [16k, 16k+16)
                    umma_ts(tO, tP + k * 8, db, idesc_pv, (j | k) != 0);                   // P: 8 columns (16 bf16) per K step
                }
                umma_commit(&pv_done[half]);
                if (half == nh - 1) umma_commit(&kv_empty[slot]);                         // every MMA reading this stage has been issued
            };
            mbar_wait(q_full, 0);
            tc_fence_after();
            if (tr) tr[176] = clock64();
            for (int n = 0; n < 3 && n < total; ++n) issue_qk(n);
            for (int n = 0; n < total; ++n) {
                issue_pv(n);
                if (n + 3 < total) issue_qk(n + 3);
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
                l_run += lsum;
                // ---- bf16 pairs of P into the first 64 columns of this S buffer (A operand of the PV MMA)
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_S + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                tmem_st_wait();
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

inline long long*& attn_trace_ptr() {      // test hook (tvb_debug_set_trace): device buffer of 16 x 256 int64, or nullptr
    static long long* p = nullptr;
    return p;
}

template <int POLY>
inline cudaError_t launch_attn4_t(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  int nwork, cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn4_kernel<POLY>, cudaFuncAttributeMaxDynamicSharedMemorySize, A4_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    attn4_kernel<POLY><<<nwork * nheads, ATT_THREADS, A4_SMEM_BYTES, st>>>(tmQKV, ctx, H, nheads, slopes, work, attn_trace_ptr());
    return cudaGetLastError();
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e).
inline cudaError_t launch_attn4_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes,
                                    const int4* work, int nwork, int poly, cudaStream_t st) {
    switch (poly) {
        case 0: return launch_attn4_t<0>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 2: return launch_attn4_t<2>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 3: return launch_attn4_t<3>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        case 4: return launch_attn4_t<4>(tmQKV, ctx, H, nheads, slopes, work, nwork, st);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
