// Fused masked self-attention, version 9.  Same roles as attn8_tc.cuh (TMA producer, PV issuer, QK^T issuer, bias-tile writer, 8 softmax
// warps with one thread per query row), restructured around what the ncu captures of v8 showed (profiles/r01/ncu_attn8_final_summary.txt and
// the source-level stall samples: each softmax warp issues in only 25 % of its cycles; 15 % of the fast path is branch / instruction-fetch
// stall around the diag / partial / guard branches; 9 % is the wait for the next score block; the first block of every item runs the
// 994-instruction "careful" path):
//
//  (1) Fixed reference instead of a running maximum.  Key 0 of every sequence is the time token, whose additive bias is 0
//      (networks.py:335-341), so for query i   m_i = s(i, 0) + shift   is a valid softmax reference as soon as block 0 arrives:
//      every other score is bounded by  s(i,j) - s(i,0) <= |q_i| (|k_j| + |k_0|) <= B  with B = 2 max|q| max|k| over the item
//      (Cauchy-Schwarz on the bf16 operands the MMA really multiplies; the row-block maxima of |q|^2, |k|^2 come out of the QKV GEMM's
//      epilogue, gemm_tc.cuh).  shift = max(0, B - 100) keeps every exponent in [-shift, 100]: no overflow, no maximum search, no
//      guard, no rescale of O, and the first block costs what any other block costs.  B > 200 (raw logits beyond +-69 -- no softmax
//      survives that) raises the handle's sticky error flag instead of producing a wrong result (k_attn9_plan).
//  (2) ALiBi window.  With that reference p(i,j) / l_i <= 2^(B - sigma * |i - j|): a key block whose nearest key is further than
//      (B + 40) / sigma from every query row of the half contributes less than 2^-40 of the row sum per key -- below the rounding of the
//      fp32 accumulation of the blocks that are kept -- and is skipped by every role (block 0 is always kept: it holds the time token).
//      Steep heads at S = 1134 keep 4-6 of 9 blocks, at 30 s (S = 2856) 4-8 of 23.
//  (3) One S buffer per half, P in its own tensor-memory region: S_A [0,128) S_B [128,256) P_A [256,320) P_B [320,384) O_A [384,448)
//      O_B [448,512).  A half hands its S buffer back as soon as the last 32 columns are in registers, so the next QK^T overlaps the
//      exponentials of the last chunk and the P hand-over, and the halves no longer wait for each other's PV (v8: three S buffers for
//      two halves left half B zero slack).
//  (4) Straight-line block body.  The body is instantiated twice: `plain` (no diagonal, no masked keys, not block 0 -- 6 of 9 blocks at
//      S = 1134) has no data-dependent branch at all; `general` handles the rest.
// Item parameters (window, shift) are computed per (tile, head) by k_attn9_plan from the row-block norms.
#pragma once
#include <type_traits>

#include "attn8_tc.cuh"

namespace tvb {

constexpr float A9_SKIP_LOG2 = 40.0f;      // blocks whose keys are all below 2^-40 of the row sum are skipped
constexpr float A9_REF_TOP = 100.0f;       // largest exponent handed to exp2 (p <= 2^100; sums of thousands of them stay finite in fp32)
constexpr float A9_LOGIT_LIMIT = 200.0f;   // B beyond this: unsupported (the reference exponent of key 0 would fall below -100)
constexpr int A9_SMEM_BYTES = A6_SMEM_BYTES;

__host__ __device__ __forceinline__ int a9_head(int t, int nheads) {
    // heads rotate with the tile index: a CTA's items (stride gridDim.x) then mix steep (cheap, narrow window) and flat (full) heads
    return (t + t / nheads) % nheads;
}

struct A9Item {
    int q0, start, len, head, nh, qb0;
    int lo0, hi0, lo1, hi1;   // per half: key blocks [lo, hi] besides block 0 (lo >= 1; lo > hi = none); scalars, not arrays: a
                              // run-time half index would put the item in local memory
    int jlo, jhi;          // union over the halves
    float shift;
};
__device__ __forceinline__ A9Item a9_load_item(const int4* __restrict__ work, const int4* __restrict__ win, int t, int nheads) {
    const int4 w = work[t / nheads];
    const int4 v = win[t];
    A9Item it;
    it.head = a9_head(t, nheads);
    it.q0 = w.y; it.start = w.z; it.len = w.w;
    it.nh = (w.y + 128 < w.w) ? 2 : 1;
    it.qb0 = w.y / ATT_BK;
    it.lo0 = v.x & 0xffff; it.hi0 = v.x >> 16;
    it.lo1 = v.y & 0xffff; it.hi1 = v.y >> 16;
    it.jlo = v.w & 0xffff; it.jhi = v.w >> 16;
    it.shift = __int_as_float(v.z);
    return it;
}
__device__ __forceinline__ bool a9_needed(const A9Item& w, int half, int j) {
    const int lo = half ? w.lo1 : w.lo0, hi = half ? w.hi1 : w.hi0;
    return j == 0 || (j >= lo && j <= hi);
}
__device__ __forceinline__ int a9_nblk(const A9Item& w) { return 1 + (w.jhi >= w.jlo ? w.jhi - w.jlo + 1 : 0); }
__device__ __forceinline__ int a9_block(const A9Item& w, int jj) { return jj == 0 ? 0 : w.jlo + jj - 1; }

// ------------------------------------------------------------------------------------------------------------------------
// Row-block norms straight from the packed [rows, 3H] bf16 q|k|v rows: the stand-alone producer of the table the QKV GEMM epilogue
// normally writes (used by the kernel-level test hook and by GEMM variants without that epilogue).  normtab[blk][slot], blk = row / 32,
// slot = head (q) or nheads + head (k): max over the 32 rows of the squared norm of the 64 head columns.
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attn9_rownorm(const __nv_bfloat16* __restrict__ qkv, int rows, int H, int nheads, float* __restrict__ normtab) {
    const int blk = blockIdx.x, lane = threadIdx.x & 31;
    const int row = blk * 32 + lane;
    for (int slot = threadIdx.x >> 5; slot < 2 * nheads; slot += blockDim.x >> 5) {
        float ss = 0.f;
        if (row < rows) {
            const uint4* p = reinterpret_cast<const uint4*>(qkv + static_cast<size_t>(row) * 3 * H + slot * 64);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const uint4 v = p[c];
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float a = __uint_as_float(w[q] << 16), b = __uint_as_float(w[q] & 0xffff0000u);
                    ss = fmaf(a, a, fmaf(b, b, ss));
                }
            }
        }
        const uint32_t m = __reduce_max_sync(0xffffffffu, __float_as_uint(ss));      // ss >= 0: float order == unsigned order
        if (lane == 0) normtab[static_cast<size_t>(blk) * 2 * nheads + slot] = __uint_as_float(m);
    }
}

// ------------------------------------------------------------------------------------------------------------------------
// Per item (tile of <= 256 queries x head): the reference shift and the key-block windows of the two halves.  One warp per item.
// err_flag (mapped host word, may be nullptr) is set to 1 when an item's logit bound exceeds A9_LOGIT_LIMIT (plain store: atomics on
// mapped host memory need PCIe atomics).
// ------------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attn9_plan(const int4* __restrict__ work, int n_items, int nheads, const float* __restrict__ slopes,
                                                     const float* __restrict__ normtab, int use_window, int4* __restrict__ win,
                                                     int* __restrict__ err_flag) {
    pdl_launch_dependents();                           // the attention kernel's set-up may overlap this kernel
    pdl_wait();                                        // (launched with programmatic serialization itself: the QKV GEMM must have finished)
    const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (t >= n_items) return;
    const int4 w = work[t / nheads];
    const int head = a9_head(t, nheads);
    const int q0 = w.y, start = w.z, len = w.w;
    const int nkb = (len + ATT_BK - 1) / ATT_BK;
    const int nh = (q0 + 128 < len) ? 2 : 1;
    const int ld = 2 * nheads;
    auto blkmax = [&](int r0, int r1, int slot) {       // max of normtab[.][slot] over the row blocks touching rows [r0, r1]
        float m = 0.f;
        for (int b = (r0 >> 5) + lane; b <= (r1 >> 5); b += 32) m = fmaxf(m, normtab[static_cast<size_t>(b) * ld + slot]);
        return __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(m)));
    };
    const float k2 = blkmax(start, start + len - 1, nheads + head);
    float bound[2] = {0.f, 0.f};
    for (int h = 0; h < nh; ++h) {
        const int qa = q0 + 128 * h, qb = min(qa + 127, len - 1);
        const float q2 = blkmax(start + qa, start + qb, head);
        // |q.k| <= |q||k|; 1.02 covers the bf16 rounding of the operands (the norms are taken before it) and the fp32 accumulation
        bound[h] = 2.0f * sqrtf(q2) * sqrtf(k2) * 1.02f + 1e-3f;
    }
    if (lane != 0) return;
    const float slope2 = slopes[head] * LOG2E;
    const float bmax = fmaxf(bound[0], bound[1]);
    if (!(bmax <= A9_LOGIT_LIMIT) && err_flag) *err_flag = 1;            // also catches NaN / inf operands
    int lo[2] = {1, 1}, hi[2] = {0, 0};
    for (int h = 0; h < nh; ++h) {
        const int qa = q0 + 128 * h, qb = min(qa + 127, len - 1);
        if (!use_window) { lo[h] = 1; hi[h] = nkb - 1; continue; }
        int l = nkb, r = 0;
        for (int j = 1; j < nkb; ++j) {
            const int ka = j * ATT_BK, kb = min(ka + ATT_BK - 1, len - 1);
            const int d = ka > qb ? ka - qb : (qa > kb ? qa - kb : 0);
            if (bound[h] - slope2 * static_cast<float>(d) > -A9_SKIP_LOG2) { l = min(l, j); r = max(r, j); }
        }
        if (l <= r) { lo[h] = l; hi[h] = r; }
    }
    // every block of the item's list [jlo, jhi] must be wanted by at least one half (the block's K/V stage is released by the last
    // half that uses it): close a gap between the two windows, should one ever appear
    if (nh == 2 && lo[0] <= hi[0] && lo[1] <= hi[1]) {
        if (hi[0] + 1 < lo[1]) hi[0] = lo[1] - 1;
        if (hi[1] + 1 < lo[0]) hi[1] = lo[0] - 1;
    }
    int jlo = 1, jhi = 0;
    for (int h = 0; h < nh; ++h)
        if (lo[h] <= hi[h]) {
            if (jlo > jhi) { jlo = lo[h]; jhi = hi[h]; }
            else { jlo = min(jlo, lo[h]); jhi = max(jhi, hi[h]); }
        }
    const float shift = fminf(fmaxf(bmax - A9_REF_TOP, 0.f), A9_LOGIT_LIMIT - A9_REF_TOP);
    win[t] = make_int4(lo[0] | (hi[0] << 16), lo[1] | (hi[1] << 16), __float_as_int(shift), jlo | (jhi << 16));
}

// ------------------------------------------------------------------------------------------------------------------------
// Event trace for tools/attn_lab.cu (compiled in only there): CTA 0 stamps (tag << 48 | clock) per role into g_a9_trace[role][1024].
#ifdef A9_LAB_TRACE
__device__ long long* g_a9_trace;
#define A9_STAMP(role, tag)                                                                                                  \
    do {                                                                                                                     \
        if ((LAB & 32) && blockIdx.x == 0 && lane == 0 && tr_n < 1024)                                                       \
            g_a9_trace[(role) * 1024 + tr_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xffffffffffffll);       \
    } while (0)
#else
#define A9_STAMP(role, tag) do { } while (0)
#endif

// LAB (tools/attn_lab.cu only; the library always instantiates LAB = 0): timing ablations that switch pieces of the softmax body
// off -- bit 0: no exponential, bit 1: no row sum, bit 2: no offset add, bit 3: no bf16 pack, bit 4: the P hand-over does not wait for
// the previous PV.  Results are meaningless with LAB != 0.
template <int POLY, int LAB = 0>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn9_kernel(const __grid_constant__ CUtensorMap tmQKV, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
             const float* __restrict__ slopes, const int4* __restrict__ work, const int4* __restrict__ win, int n_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A6_SMEM_BAR);
    uint64_t* q_full = bars + 0;                        // [2] per Q buffer
    uint64_t* q_empty = q_full + 2;                     // [2]
    uint64_t* ext_full = q_empty + 2;                   // [2] per bias-tile buffer
    uint64_t* ext_empty = ext_full + 2;                 // [2]
    uint64_t* kv_full = ext_empty + 2;                  // [4]
    uint64_t* kv_empty = kv_full + A4_KV_STAGES;        // [4]
    uint64_t* s_full = kv_empty + A4_KV_STAGES;         // [2] per half: scores of the half's next block are in S_half
    uint64_t* s_free = s_full + 2;                      // [2] per half: the 128 softmax threads hold the whole block in registers
    uint64_t* p_ready = s_free + 2;                     // [2] per half: P_half is written
    uint64_t* pv_done = p_ready + 2;                    // [2 halves][2]: PV number c of a half arrives on [half][c & 1] (phase c >> 1)
    uint64_t* o_free = pv_done + 4;                     // [2] per half: O has been read out by the 128 softmax threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    [[maybe_unused]] int tr_n = 0;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&ext_full[s], 1); mbar_init(&ext_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 128); mbar_init(&p_ready[s], 128);
            mbar_init(&pv_done[2 * s], 1); mbar_init(&pv_done[2 * s + 1], 1); mbar_init(&o_free[s], 128);
        }
        for (int s = 0; s < A4_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
    pdl_launch_dependents();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                      // set-up above overlapped the tail of the planning kernel

    if (warp_idx < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp_idx == 0) {
        // ------------------------------------------ TMA producer ------------------------------------------
        if (elect_one()) {
            int slot = 0; uint32_t phase = 0;
            int itc = 0;
            for (int t = blockIdx.x; t < n_items; t += G, ++itc) {
                const A9Item w = a9_load_item(work, win, t, nheads);
                const int qb = itc & 1;
                mbar_wait(&q_empty[qb], ((itc >> 1) & 1) ^ 1);
                uint8_t* sq = smem + A6_SMEM_Q + qb * 2 * ATT_TILE_BYTES;
                mbar_arrive_expect_tx(&q_full[qb], w.nh * ATT_TILE_BYTES);
                tma_load_2d(sq, &tmQKV, &q_full[qb], w.head * ATT_D, w.start + w.q0);
                if (w.nh == 2) tma_load_2d(sq + ATT_TILE_BYTES, &tmQKV, &q_full[qb], w.head * ATT_D, w.start + w.q0 + 128);
                const int nblk = a9_nblk(w);
                for (int jj = 0; jj < nblk; ++jj) {
                    const int j = a9_block(w, jj);
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
            const float slope2 = slopes[a9_head(t, nheads)] * LOG2E;
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
        int slot = 0;
        int cnt[2] = {0, 0};                                // items in which the half has taken part so far
        int pvc[2] = {0, 0};                                // PV MMAs issued so far per half
        A9Item w = a9_load_item(work, win, blockIdx.x, nheads);
        for (int t = blockIdx.x; t < n_items; t += G) {
            A9Item wn = w;
            if (t + G < n_items) wn = a9_load_item(work, win, t + G, nheads);            // prefetch the next item's parameters
            const int nblk = a9_nblk(w);
            for (int jj = 0; jj < nblk; ++jj) {
                const int j = a9_block(w, jj);
                const int last_half = (w.nh == 2 && a9_needed(w, 1, j)) ? 1 : 0;        // half 0 takes part whenever half 1 does not
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half >= w.nh || !a9_needed(w, half, j)) continue;
                    const int c = pvc[half]++;
                    mbar_wait(&p_ready[half], c & 1);
                    A9_STAMP(2, 20 + half);
                    if (jj == 0) {                                                        // this PV overwrites O: the previous item's
                        if (cnt[half] > 0) mbar_wait(&o_free[half], (cnt[half] - 1) & 1); // epilogue must have read it out
                        ++cnt[half];
                    }
                    tc_fence_after();
                    const uint32_t tP = tmem_base + 256 + half * 64;
                    const uint32_t tO = tmem_base + 384 + half * 64;
                    const uint64_t dv = dv0 + static_cast<uint64_t>(slot * (2 * ATT_TILE_BYTES >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < ATT_BK / 16; ++k)                             // V rows [16k, 16k+16); P: 8 columns per K step
                            umma_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc_pv, (jj | k) != 0);
                        umma_commit(&pv_done[2 * half + (c & 1)]);
                        if (half == last_half) umma_commit(&kv_empty[slot]);              // every MMA reading this stage has been issued
                    }
                    __syncwarp();
                    A9_STAMP(2, 22 + half);
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
        int slot = 0, itc = 0;
        uint32_t kvpar = 0;
        int qkc[2] = {0, 0};                                // QK^T blocks issued so far per half
        A9Item w = a9_load_item(work, win, blockIdx.x, nheads);
        for (int t = blockIdx.x; t < n_items; t += G, ++itc) {
            A9Item wn = w;
            if (t + G < n_items) wn = a9_load_item(work, win, t + G, nheads);
            const int qb = itc & 1;
            mbar_wait(&q_full[qb], (itc >> 1) & 1);
            mbar_wait(&ext_full[qb], (itc >> 1) & 1);
            const uint64_t dext = de0 + static_cast<uint64_t>(qb * (ATT_TILE_BYTES >> 4));
            const int nblk = a9_nblk(w);
            for (int jj = 0; jj < nblk; ++jj) {
                const int j = a9_block(w, jj);
                mbar_wait(&kv_full[slot], kvpar);
                const int last_half = (w.nh == 2 && a9_needed(w, 1, j)) ? 1 : 0;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half >= w.nh || !a9_needed(w, half, j)) continue;
                    const int c = qkc[half]++;
                    mbar_wait(&s_free[half], (c & 1) ^ 1);  // block c - 1 of this half is in its threads' registers (fresh barrier: passes)
                    A9_STAMP(3, 30 + half);
                    tc_fence_after();
                    const uint32_t d = tmem_base + half * 128;
                    const uint64_t da = dq0 + static_cast<uint64_t>((qb * 2 + half) * (ATT_TILE_BYTES >> 4));
                    const uint64_t db = dk0 + static_cast<uint64_t>(slot * (2 * ATT_TILE_BYTES >> 4));
                    const int jd = w.qb0 + half;
                    const bool last = (jj == nblk - 1) && (half == last_half);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                        if (j != jd) umma_ss(d, dext + (j < jd ? 2 : 4), dext, idesc_qk, 1);   // +sigma*jrel left of the diagonal, -sigma*jrel right
                        umma_commit(&s_full[half]);
                        if (last) {                                                       // last QK of the item: Q and bias buffers are free
                            umma_commit(&q_empty[qb]);
                            umma_commit(&ext_empty[qb]);
                        }
                    }
                    __syncwarp();
                    A9_STAMP(3, 32 + half);
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
        const uint32_t tmem_S = tmem_base + half * 128 + lane_off;
        const uint32_t tmem_P = tmem_base + 256 + half * 64 + lane_off;
        const uint32_t tmem_O = tmem_base + 384 + half * 64 + lane_off;
        int cblk = 0;                                       // blocks this half has processed so far (all items)
        // PV(c) of this half retired?  Barrier [c & 1] is in the phase of PV(c) or past it whenever this is called: every block waits
        // for the PV before it (below), so PV(c - 2) has been observed.
        auto wait_pv = [&](int c) { mbar_wait(&pv_done[2 * half + (c & 1)], (c >> 1) & 1); };
        A9Item wnext = a9_load_item(work, win, blockIdx.x, nheads);
        float slope_next = slopes[wnext.head];
        for (int t = blockIdx.x; t < n_items; t += G) {
            const A9Item w = wnext;
            const float slope2 = slope_next * LOG2E;
            if (t + G < n_items) {                          // prefetch the next item's parameters: consumed one item later
                wnext = a9_load_item(work, win, t + G, nheads);
                slope_next = slopes[wnext.head];
            }
            if (half >= w.nh) continue;                     // no rows for this half (the barriers of a half are its own: no rendezvous needed)
            const int len = w.len;
            const int qt = w.q0 + half * 128;               // first query of this half
            const int jd = w.qb0 + half;                    // diagonal key block
            const int i = qt + r;                           // query position in the sequence
            const float a0d = -slope2 * static_cast<float>(r);   // diagonal block: -sigma |r - e|, a = sigma e - sigma r
            const int nblk = a9_nblk(w);
            float m_ref = 0.f, l_run = 0.f;
            // s[0,64): bf16 pairs of P (output); s[64,96) / s[96,128) double-buffer 32-column chunks of S
            uint32_t s[128];
            uint32_t(&ca)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[64]);
            uint32_t(&cb)[32] = *reinterpret_cast<uint32_t(*)[32]>(&s[96]);
            // the half's first block of the item (always key block 0): wait for its scores, request the first chunk
            mbar_wait(&s_full[half], cblk & 1);
            tc_fence_after();
            tmem_ld_32x32(tmem_S, ca);
            int jnext = 0;                                  // block whose first chunk is in flight
            for (int jj = 0; jj < nblk;) {
                const int j = jnext;
                // the half's next block of this item, if any (its own window; the item's list is the union of both halves')
                int jj2 = jj + 1;
                while (jj2 < nblk && !a9_needed(w, half, a9_block(w, jj2))) ++jj2;
                const bool more = jj2 < nblk;
                const int k0 = j * ATT_BK;
                auto body = [&](auto gen_tag) {
                    constexpr bool GEN = decltype(gen_tag)::value;
                    const bool diag = GEN && (j == jd);
                    const bool partial = GEN && (k0 + ATT_BK > len);
                    // additive term of this block: per-row constant (off-diagonal; the key-dependent part came out of the MMA) or none
                    const float c_row = diag ? 0.f : -slope2 * fabsf(static_cast<float>(i - k0));
                    tmem_ld_wait_dep(ca);                    // chunk 0 (requested when the block's scores arrived)
                    if (quad == 0) A9_STAMP(half, 2);
                    if (GEN && j == 0) m_ref = __uint_as_float(ca[0]) + w.shift;        // key 0: the time token, bias 0
                    tmem_ld_32x32(tmem_S + 32, cb);
                    const float off = c_row - m_ref;
                    const uint64_t off2 = pack2f(off, off);
                    uint64_t acc2 = pack2f(0.f, 0.f), accp2 = pack2f(0.f, 0.f);
                    auto proc = [&](uint32_t(&cur)[32], const int c) {
                        if constexpr (GEN) {
                            if (diag) {
                                const float s0raw = __uint_as_float(cur[0]);
                                float a0v = a0d;
                                asm volatile("" : "+f"(a0v));                   // keep the 128 bias values from being hoisted out of the block loop
#pragma unroll
                                for (int e = 0; e < 32; ++e) {
                                    const float a = fmaf(slope2, static_cast<float>(c * 32 + e), a0v);
                                    cur[e] = __float_as_uint(__uint_as_float(cur[e]) - fabsf(a));
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
                        }
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            uint64_t x2 = pack2f(__uint_as_float(cur[2 * u]), __uint_as_float(cur[2 * u + 1]));
                            if (!(LAB & 4)) x2 = fadd2(x2, off2);
                            uint64_t p2;
                            if (POLY > 0 && (u % (POLY > 0 ? POLY : 1)) == (POLY > 0 ? POLY : 1) - 1) {
                                p2 = (LAB & 1) ? x2 : exp2_poly2(x2);
                                if (!(LAB & 2)) accp2 = fadd2(accp2, p2);
                            } else {
                                float lo, hi;
                                unpack2f(x2, lo, hi);
                                if (!(LAB & 1)) {
                                    lo = ex2_approx(lo);
                                    hi = ex2_approx(hi);
                                }
                                p2 = pack2f(lo, hi);
                                if (!(LAB & 2)) acc2 = fadd2(acc2, p2);
                            }
                            float lo, hi;
                            unpack2f(p2, lo, hi);
                            s[c * 16 + u] = (LAB & 8) ? __float_as_uint(lo) ^ __float_as_uint(hi) : pack_bf16x2(lo, hi);
                        }
                    };
                    proc(ca, 0);
                    tmem_ld_wait_dep(cb);
                    if (quad == 0) A9_STAMP(half, 3);
                    tmem_ld_32x32(tmem_S + 64, ca);
                    proc(cb, 1);
                    tmem_ld_wait_dep(ca);
                    if (quad == 0) A9_STAMP(half, 4);
                    tmem_ld_32x32(tmem_S + 96, cb);
                    proc(ca, 2);
                    tmem_ld_wait_dep(cb);
                    if (quad == 0) A9_STAMP(half, 5);
                    // the whole block is in registers: hand S back, the next QK^T overlaps the last chunk and the P hand-over
                    tc_fence_before();
                    mbar_arrive(&s_free[half]);
                    proc(cb, 3);
                    float a0, a1;
                    unpack2f(fadd2(acc2, accp2), a0, a1);
                    l_run += a0 + a1;
                };
                if (j != 0 && j != jd && k0 + ATT_BK <= len) body(std::false_type{});
                else body(std::true_type{});
                if (quad == 0) A9_STAMP(half, 6);
                // ---- P_half must have been read by the previous PV of this half before it is overwritten
                if (cblk > 0 && !(LAB & 16)) wait_pv(cblk - 1);
                tc_fence_after();
                if (quad == 0) A9_STAMP(half, 7);
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_P + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_ready[half]);
                if (quad == 0) A9_STAMP(half, 8);
                ++cblk;
                if (more) {                                  // scores of the half's next block: wait, then request their first 32 columns
                    mbar_wait(&s_full[half], cblk & 1);
                    if (quad == 0) A9_STAMP(half, 1);
                    tc_fence_after();
                    tmem_ld_32x32(tmem_S, ca);
                    jnext = a9_block(w, jj2);
                }
                jj = jj2;
            }
            // epilogue: read O, hand the accumulator back, then O / l -> ctx (bf16), only rows that belong to the sequence
            wait_pv(cblk - 1);
            tc_fence_after();
            uint32_t o[64];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld_32x32(tmem_O + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&o[c * 32]));
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(&o_free[half]);
            const float inv = 1.0f / ((LAB & 2) ? 1.0f : l_run);
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
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

inline cudaError_t launch_attn9_plan(const int4* work, int n_items, int nheads, const float* slopes, const float* normtab, int use_window, int4* win,
                                     int* err_flag, cudaStream_t st, bool pdl) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((n_items + 7) / 8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, k_attn9_plan, work, n_items, nheads, slopes, normtab, use_window, win, err_flag);
}

template <int POLY, int LAB = 0>
inline cudaError_t launch_attn9_t(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  const int4* win, int nwork, int num_sms, cudaStream_t st, bool pdl) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn9_kernel<POLY, LAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, A9_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int n_items = nwork * nheads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ATT_THREADS); cfg.dynamicSmemBytes = A9_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, attn9_kernel<POLY, LAB>, tmQKV, ctx, H, nheads, slopes, work, win, n_items);
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e); normtab holds the row-block norms of
// the same rows (QKV GEMM epilogue or k_attn9_rownorm); win is scratch for nwork * nheads int4.
inline cudaError_t launch_attn9_raw(const CUtensorMap& tmQKV, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                    int nwork, const float* normtab, int4* win, int use_window, int* err_flag, int poly, int num_sms,
                                    cudaStream_t st, bool pdl = false) {
    const int n_items = nwork * nheads;
    cudaError_t e = launch_attn9_plan(work, n_items, nheads, slopes, normtab, use_window, win, err_flag, st, pdl);
    if (e != cudaSuccess) return e;
    switch (poly) {
        case 0: return launch_attn9_t<0>(tmQKV, ctx, H, nheads, slopes, work, win, nwork, num_sms, st, pdl);
        case 2: return launch_attn9_t<2>(tmQKV, ctx, H, nheads, slopes, work, win, nwork, num_sms, st, pdl);
        case 3: return launch_attn9_t<3>(tmQKV, ctx, H, nheads, slopes, work, win, nwork, num_sms, st, pdl);
        case 4: return launch_attn9_t<4>(tmQKV, ctx, H, nheads, slopes, work, win, nwork, num_sms, st, pdl);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
