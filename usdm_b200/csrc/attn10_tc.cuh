// Fused masked self-attention, version 10 = attn9_tc.cuh with the WHOLE additive term produced by the tensor core.
//
// The lab ablations of v9 (tools/attn_lab.cu, profiles/r02/attn_lab.txt) show the softmax warps bound by issue / pipe cycles while two of
// them share a sub-partition, and the per-pair offset add (x + c_row - m_ref, 64 FADD2 per warp and block, at the head of the
// ld -> add -> exp2 -> pack chain) worth 10 % of the kernel.  Here neither the offset nor a per-row reference exists:
//   * no reference: with B = 2 max|q| max|k| <= 200 (k_attn9_plan raises the handle's error flag otherwise) every logit is below 100 in
//     the log2 domain and the time-token key keeps every row sum above 2^-100, so p = 2^(s + bias) is used as it is;
//   * bias(i, j) = -sigma |i - j| = +-(sigma pos_j - sigma i) is bilinear once the side of the diagonal is known, so one extra K = 16 MMA
//     step per score block adds all of it:   A (query side, written per item by warp 3, one column group per half and side)
//         { s_hi, s_mid, s_lo, s_hi, s_mid, s_lo, -(s i)_hi, -(s i)_mid, -(s i)_lo, 0 ... }          (s = sigma log2 e as three bf16 terms)
//     times B (key side) = rows of a static position table  { p_lo x3, p_hi x3, 1 x3, 0 ... }  with p = p_hi + p_lo the key's position
//     (p_lo < 128, p_hi a multiple of 128: both exact in bf16), which the TMA producer loads next to K and V of every block (32-byte
//     swizzle, 4 KB).  Blocks right of the diagonal use the negated A group.  Only the diagonal block (|r - e| is not bilinear), the
//     time token (bias 0) and masked keys of a sequence's last block are still patched element-wise, in the `general` body.
// Shared memory: Q 2 x 2 x 16 KB, three K / V / position stages of 36 KB, two 16 KB query-side tiles = 204 KB.
#pragma once
#include "attn9_tc.cuh"

namespace tvb {

constexpr int A10_KV_STAGES = 3;
constexpr int A10_POS_BYTES = 128 * 16 * 2;                                    // 4 KB: 128 keys x 16 bf16
constexpr int A10_STAGE_BYTES = 2 * ATT_TILE_BYTES + A10_POS_BYTES;            // K, V, positions
constexpr int A10_SMEM_Q = 0;                                                  // 2 items x 2 halves x 16 KB
constexpr int A10_SMEM_KV = 4 * ATT_TILE_BYTES;
constexpr int A10_SMEM_EXT = A10_SMEM_KV + A10_KV_STAGES * A10_STAGE_BYTES;    // 2 x 16 KB query-side tiles
constexpr int A10_SMEM_BAR = A10_SMEM_EXT + 2 * ATT_TILE_BYTES;
constexpr int A10_SMEM_BYTES = A10_SMEM_BAR + 256;
static_assert(A10_STAGE_BYTES % 1024 == 0 && A10_SMEM_EXT % 1024 == 0, "swizzled tiles need 1024-byte alignment");

// Shared-memory matrix descriptor, 32-byte swizzle, K-major: rows of 16 bf16 (one K = 16 step), 8-row groups 256 B apart.
__device__ __forceinline__ uint64_t make_smem_desc_sw32(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);
    d |= static_cast<uint64_t>(1) << 16;                    // LBO (unused for a single K step)
    d |= static_cast<uint64_t>(256 >> 4) << 32;             // SBO
    d |= 1ull << 46;                                        // descriptor version
    d |= 6ull << 61;                                        // SWIZZLE_32B
    return d;
}

// Static key-side operand of the bias MMA: one row of 16 bf16 per packed row.  row_pos = position inside the sequence (0 = time token);
// rows outside every sequence are zero.
__global__ void k_attn10_pos(const int* __restrict__ row_seq, const int* __restrict__ row_pos, int rows, __nv_bfloat16* __restrict__ pos) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    uint4 a = make_uint4(0, 0, 0, 0), b = make_uint4(0, 0, 0, 0);
    if (row_seq[r] >= 0) {
        const int p = row_pos[r];
        const uint32_t lo = __bfloat16_as_ushort(__float2bfloat16_rn(static_cast<float>(p & 127)));
        const uint32_t hi = __bfloat16_as_ushort(__float2bfloat16_rn(static_cast<float>(p & ~127)));
        const uint32_t one = 0x3F80u;
        a = make_uint4(lo | (lo << 16), lo | (hi << 16), hi | (hi << 16), one | (one << 16));
        b = make_uint4(one, 0, 0, 0);
    }
    uint4* o = reinterpret_cast<uint4*>(pos + static_cast<size_t>(r) * 16);
    o[0] = a;
    o[1] = b;
}

// Item order.  The windows make item costs differ by up to 3 x between steep and flat heads, and a static round-robin over the launch
// order leaves the slowest CTA 7-12 % above the mean (simulated with the real windows; profiles/r02).  k_attn10_order sorts the items by
// their planned cost (key blocks of the union window + 1) in descending order -- a counting sort in one CTA, launched after the plan
// kernel -- and CTA b takes order[k G + b] on even rounds, order[k G + G - 1 - b] on odd ones ("snake"), which is within 1-2 % of LPT.
// The order only decides which CTA computes which item: results do not depend on it.  Measured (tools/attn_lab.cu): 16 x 2856 tokens
// 540 -> 443 us per launch; the sort itself costs ~10 us, so the host only asks for it where windows are narrow against the sequence
// (long or ragged batches, tvb.cu: attention()).
// order == nullptr: launch order, plain stride (the head rotation of a9_head already mixes steep and flat heads per CTA, and measured on
// configs[1] the sorted order buys nothing there: 227.5 against 227.7 ms per solve once the 10 us sort is paid, profiles/r02).
__device__ __forceinline__ int a10_index(int k, int b, int G, const int* order) { return k * G + ((order != nullptr && (k & 1)) ? G - 1 - b : b); }
#define A10_ORDER(order, idx) ((order) != nullptr ? (order)[idx] : (idx))

constexpr int A10_COST_BINS = 1024;
__global__ void __launch_bounds__(1024) k_attn10_order(const int4* __restrict__ win, int n_items, int sort, int* __restrict__ order) {
    pdl_launch_dependents();
    __shared__ int hist[A10_COST_BINS];
    __shared__ int warp_tot[32];
    const int tid = threadIdx.x;
    if (!sort) {
        for (int t = tid; t < n_items; t += blockDim.x) order[t] = t;
        return;
    }
    auto cost = [&](int t) {
        const int4 v = win[t];
        const int jlo = v.w & 0xffff, jhi = v.w >> 16;
        const int c = 1 + (jhi >= jlo ? jhi - jlo + 1 : 0);
        return A10_COST_BINS - 1 - min(c, A10_COST_BINS - 1);             // descending cost = ascending bin
    };
    hist[tid] = 0;
    __syncthreads();
    for (int t = tid; t < n_items; t += blockDim.x) atomicAdd(&hist[cost(t)], 1);
    __syncthreads();
    // exclusive scan of the 1024 bins: one bin per thread
    const int mine = hist[tid];
    int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(0xffffffffu, incl, o);
        if ((tid & 31) >= o) incl += n;
    }
    if ((tid & 31) == 31) warp_tot[tid >> 5] = incl;
    __syncthreads();
    if (tid < 32) {
        int w = warp_tot[tid];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int n = __shfl_up_sync(0xffffffffu, w, o);
            if (tid >= o) w += n;
        }
        warp_tot[tid] = w;
    }
    __syncthreads();
    hist[tid] = incl - mine + ((tid >> 5) ? warp_tot[(tid >> 5) - 1] : 0);
    __syncthreads();
    for (int t = tid; t < n_items; t += blockDim.x) order[atomicAdd(&hist[cost(t)], 1)] = t;
}

// Event trace for tools/attn_lab.cu (compiled in only there): CTA 0 stamps (tag << 48 | clock) per role into g_a10_trace[role][1024].
#ifdef A10_LAB_TRACE
__device__ long long* g_a10_trace;
#define A10_STAMP(role, tag)                                                                                                  \
    do {                                                                                                                     \
        if ((LAB & 32) && blockIdx.x == 0 && lane == 0 && tr_n < 1024)                                                       \
            g_a10_trace[(role) * 1024 + tr_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xffffffffffffll);       \
    } while (0)
// finer trace (LAB bit 6): roles 0-3 = the four softmax warps of half A, 4 = QK issuer, 5 = PV issuer; g_a10_trace holds 8 x 1024 stamps
#define A10_STAMP2(role, tag)                                                                                                 \
    do {                                                                                                                     \
        if ((LAB & 64) && blockIdx.x == 0 && lane == 0 && tr_n < 1024)                                                       \
            g_a10_trace[(role) * 1024 + tr_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xffffffffffffll);       \
    } while (0)
#else
#define A10_STAMP(role, tag) do { } while (0)
#define A10_STAMP2(role, tag) do { } while (0)
#endif

// LAB (tools/attn_lab.cu only; the library always instantiates LAB = 0): timing ablations that switch pieces of the softmax body
// off -- bit 0: no exponential, bit 1: no row sum, bit 3: no bf16 pack, bit 4: the P hand-over does not wait for the previous PV.
// Results are meaningless with LAB != 0.
template <int POLY, int LAB = 0>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn10_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPos, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
             const float* __restrict__ slopes, const int4* __restrict__ work, const int4* __restrict__ win, const int* __restrict__ order,
             int n_items) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A10_SMEM_BAR);
    uint64_t* q_full = bars + 0;                        // [2] per Q buffer
    uint64_t* q_empty = q_full + 2;                     // [2]
    uint64_t* ext_full = q_empty + 2;                   // [2] per bias-tile buffer
    uint64_t* ext_empty = ext_full + 2;                 // [2]
    uint64_t* kv_full = ext_empty + 2;                  // [3]
    uint64_t* kv_empty = kv_full + A10_KV_STAGES;       // [3]
    uint64_t* s_full = kv_empty + A10_KV_STAGES;         // [2] per half: scores of the half's next block are in S_half
    uint64_t* s_free = s_full + 2;                      // [2] per half: the 128 softmax threads hold the whole block in registers
    uint64_t* p_ready = s_free + 2;                     // [2] per half: P_half is written
    uint64_t* pv_done = p_ready + 2;                    // [2 halves][2]: PV number c of a half arrives on [half][c & 1] (phase c >> 1)
    uint64_t* o_free = pv_done + 4;                     // [2] per half: O has been read out by the 128 softmax threads
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    [[maybe_unused]] int tr_n = 0;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) { tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmPos); }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&ext_full[s], 1); mbar_init(&ext_empty[s], 1);
            mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 128); mbar_init(&p_ready[s], 128);
            mbar_init(&pv_done[2 * s], 1); mbar_init(&pv_done[2 * s + 1], 1); mbar_init(&o_free[s], 128);
        }
        for (int s = 0; s < A10_KV_STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
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
            for (; a10_index(itc, blockIdx.x, G, order) < n_items; ++itc) {
                const int t = A10_ORDER(order, a10_index(itc, blockIdx.x, G, order));
                const A9Item w = a9_load_item(work, win, t, nheads);
                const int qb = itc & 1;
                mbar_wait(&q_empty[qb], ((itc >> 1) & 1) ^ 1);
                uint8_t* sq = smem + A10_SMEM_Q + qb * 2 * ATT_TILE_BYTES;
                mbar_arrive_expect_tx(&q_full[qb], w.nh * ATT_TILE_BYTES);
                tma_load_2d(sq, &tmQKV, &q_full[qb], w.head * ATT_D, w.start + w.q0);
                if (w.nh == 2) tma_load_2d(sq + ATT_TILE_BYTES, &tmQKV, &q_full[qb], w.head * ATT_D, w.start + w.q0 + 128);
                const int nblk = a9_nblk(w);
                for (int jj = 0; jj < nblk; ++jj) {
                    const int j = a9_block(w, jj);
                    mbar_wait(&kv_empty[slot], phase ^ 1);
                    uint8_t* sk = smem + A10_SMEM_KV + slot * A10_STAGE_BYTES;
                    mbar_arrive_expect_tx(&kv_full[slot], A10_STAGE_BYTES);
                    tma_load_2d(sk, &tmQKV, &kv_full[slot], H + w.head * ATT_D, w.start + j * ATT_BK);
                    tma_load_2d(sk + ATT_TILE_BYTES, &tmQKV, &kv_full[slot], 2 * H + w.head * ATT_D, w.start + j * ATT_BK);
                    tma_load_2d(sk + 2 * ATT_TILE_BYTES, &tmPos, &kv_full[slot], 0, w.start + j * ATT_BK);     // key positions of the block
                    if (++slot == A10_KV_STAGES) { slot = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 3) {
        // ------------------------------------------ query-side bias tile writer ------------------------------------------
        // K-major, 128-B swizzle, one 128 x 64 bf16 tile per item buffer; column group g = 2 * half + side (side 0: key blocks left of the
        // half's diagonal, 1: right = every entry negated) holds for row r (query i = q0 + 128 half + r)
        //   { s_hi, s_mid, s_lo, s_hi, s_mid, s_lo, a_hi, a_mid | a_lo, 0 x 7 },   s = sigma log2 e,  a = -(s i), each as three bf16 terms.
        // 16-B chunk c of row r lives at chunk position c ^ (r & 7).
        int itc = 0;
        for (; a10_index(itc, blockIdx.x, G, order) < n_items; ++itc) {
            const int t = A10_ORDER(order, a10_index(itc, blockIdx.x, G, order));
            const int eb = itc & 1;
            const float slope2 = slopes[a9_head(t, nheads)] * LOG2E;
            const int q0 = work[t / nheads].y;
            mbar_wait(&ext_empty[eb], ((itc >> 1) & 1) ^ 1);
            auto bits = [](__nv_bfloat16 v) { return static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&v)); };
            auto split3 = [&](float x, uint32_t& b0, uint32_t& b1, uint32_t& b2) {
                const __nv_bfloat16 hi = __float2bfloat16_rn(x);
                const float r1 = x - __bfloat162float(hi);
                const __nv_bfloat16 mid = __float2bfloat16_rn(r1);
                const __nv_bfloat16 lo = __float2bfloat16_rn(r1 - __bfloat162float(mid));
                b0 = bits(hi); b1 = bits(mid); b2 = bits(lo);
            };
            uint32_t s0, s1, s2;
            split3(slope2, s0, s1, s2);
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int r = lane + 32 * rr;
                uint8_t* rowp = smem + A10_SMEM_EXT + eb * ATT_TILE_BYTES + r * 128;
                const int x7 = r & 7;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    uint32_t a0, a1, a2;
                    split3(-(slope2 * static_cast<float>(q0 + 128 * half + r)), a0, a1, a2);
                    const uint4 c0 = make_uint4(s0 | (s1 << 16), s2 | (s0 << 16), s1 | (s2 << 16), a0 | (a1 << 16));
                    const uint4 c1 = make_uint4(a2, 0, 0, 0);
                    const uint32_t neg = 0x80008000u;
                    const uint4 n0 = make_uint4(c0.x ^ neg, c0.y ^ neg, c0.z ^ neg, c0.w ^ neg);
                    const uint4 n1 = make_uint4(c1.x ^ 0x8000u, 0, 0, 0);
                    *reinterpret_cast<uint4*>(rowp + (((4 * half + 0) ^ x7) << 4)) = c0;
                    *reinterpret_cast<uint4*>(rowp + (((4 * half + 1) ^ x7) << 4)) = c1;
                    *reinterpret_cast<uint4*>(rowp + (((4 * half + 2) ^ x7) << 4)) = n0;
                    *reinterpret_cast<uint4*>(rowp + (((4 * half + 3) ^ x7) << 4)) = n1;
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&ext_full[eb]);
        }
    } else if (warp_idx == 1) {
        // ------------------------------------------ PV issuer (converged warp) ------------------------------------------
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);          // B (= V) MN-major
        const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_KV + ATT_TILE_BYTES), 8192, 1024);
        int slot = 0;
        int cnt[2] = {0, 0};                                // items in which the half has taken part so far
        int pvc[2] = {0, 0};                                // PV MMAs issued so far per half
        A9Item w = a9_load_item(work, win, A10_ORDER(order, blockIdx.x), nheads);
        for (int kk = 0; a10_index(kk, blockIdx.x, G, order) < n_items; ++kk) {
            A9Item wn = w;
            if (a10_index(kk + 1, blockIdx.x, G, order) < n_items)                               // prefetch the next item's parameters
                wn = a9_load_item(work, win, A10_ORDER(order, a10_index(kk + 1, blockIdx.x, G, order)), nheads);
            const int nblk = a9_nblk(w);
            for (int jj = 0; jj < nblk; ++jj) {
                const int j = a9_block(w, jj);
                const int last_half = (w.nh == 2 && a9_needed(w, 1, j)) ? 1 : 0;        // half 0 takes part whenever half 1 does not
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (half >= w.nh || !a9_needed(w, half, j)) continue;
                    const int c = pvc[half]++;
                    mbar_wait(&p_ready[half], c & 1);
                    A10_STAMP(2, 20 + half);
                    A10_STAMP2(5, 20 + half);
                    if (jj == 0) {                                                        // this PV overwrites O: the previous item's
                        if (cnt[half] > 0) mbar_wait(&o_free[half], (cnt[half] - 1) & 1); // epilogue must have read it out
                        ++cnt[half];
                    }
                    tc_fence_after();
                    const uint32_t tP = tmem_base + 256 + half * 64;
                    const uint32_t tO = tmem_base + 384 + half * 64;
                    const uint64_t dv = dv0 + static_cast<uint64_t>(slot * (A10_STAGE_BYTES >> 4));
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < ATT_BK / 16; ++k)                             // V rows [16k, 16k+16); P: 8 columns per K step
                            umma_ts(tO, tP + k * 8, dv + k * (2048 >> 4), idesc_pv, (jj | k) != 0);
                        umma_commit(&pv_done[2 * half + (c & 1)]);
                        if (half == last_half) umma_commit(&kv_empty[slot]);              // every MMA reading this stage has been issued
                    }
                    A10_STAMP2(5, 50 + half);
                    __syncwarp();
                    A10_STAMP(2, 22 + half);
                    A10_STAMP2(5, 22 + half);
                }
                if (++slot == A10_KV_STAGES) slot = 0;
            }
            w = wn;
        }
    } else if (warp_idx == 2) {
        // ------------------------------------------ QK issuer (converged warp) ------------------------------------------
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, ATT_BK);                      // A, B K-major
        const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_Q), 16, 1024);
        const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_KV), 16, 1024);
        const uint64_t de0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_EXT), 16, 1024);
        const uint64_t dp0 = make_smem_desc_sw32(smem_u32(smem + A10_SMEM_KV + 2 * ATT_TILE_BYTES));
        int slot = 0, itc = 0;
        uint32_t kvpar = 0;
        int qkc[2] = {0, 0};                                // QK^T blocks issued so far per half
        A9Item w = a9_load_item(work, win, A10_ORDER(order, blockIdx.x), nheads);
        for (; a10_index(itc, blockIdx.x, G, order) < n_items; ++itc) {
            A9Item wn = w;
            if (a10_index(itc + 1, blockIdx.x, G, order) < n_items) wn = a9_load_item(work, win, A10_ORDER(order, a10_index(itc + 1, blockIdx.x, G, order)), nheads);
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
                    A10_STAMP(3, 30 + half);
                    A10_STAMP2(4, 30 + half);
                    tc_fence_after();
                    const uint32_t d = tmem_base + half * 128;
                    const uint64_t da = dq0 + static_cast<uint64_t>((qb * 2 + half) * (ATT_TILE_BYTES >> 4));
                    const uint64_t db = dk0 + static_cast<uint64_t>(slot * (A10_STAGE_BYTES >> 4));
                    const uint64_t dpos = dp0 + static_cast<uint64_t>(slot * (A10_STAGE_BYTES >> 4));
                    const int jd = w.qb0 + half;
                    const bool last = (jj == nblk - 1) && (half == last_half);
                    A10_STAMP2(4, 40 + half);
                    if (elect_one()) {
#pragma unroll
                        for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                        // the whole additive term of an off-diagonal block: query-side group (half, side) x the key positions of the stage
                        if (j != jd) umma_ss(d, dext + 2 * (2 * half + (j < jd ? 0 : 1)), dpos, idesc_qk, 1);
                        umma_commit(&s_full[half]);
                        if (last) {                                                       // last QK of the item: Q and bias buffers are free
                            umma_commit(&q_empty[qb]);
                            umma_commit(&ext_empty[qb]);
                        }
                    }
                    A10_STAMP2(4, 42 + half);
                    __syncwarp();
                    A10_STAMP(3, 32 + half);
                    A10_STAMP2(4, 32 + half);
                }
                if (++slot == A10_KV_STAGES) { slot = 0; kvpar ^= 1; }
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
        A9Item wnext = a9_load_item(work, win, A10_ORDER(order, blockIdx.x), nheads);
        float slope_next = slopes[wnext.head];
        for (int kk = 0; a10_index(kk, blockIdx.x, G, order) < n_items; ++kk) {
            const A9Item w = wnext;
            const float slope2 = slope_next * LOG2E;
            if (a10_index(kk + 1, blockIdx.x, G, order) < n_items) {   // prefetch the next item's parameters: consumed one item later
                wnext = a9_load_item(work, win, A10_ORDER(order, a10_index(kk + 1, blockIdx.x, G, order)), nheads);
                slope_next = slopes[wnext.head];
            }
            if (half >= w.nh) continue;                     // no rows for this half (the barriers of a half are its own: no rendezvous needed)
            const int len = w.len;
            const int qt = w.q0 + half * 128;               // first query of this half
            const int jd = w.qb0 + half;                    // diagonal key block
            const int i = qt + r;                           // query position in the sequence
            const float a0d = -slope2 * static_cast<float>(r);   // diagonal block: -sigma |r - e|, a = sigma e - sigma r
            const float s_i = slope2 * static_cast<float>(i);    // what the bias MMA subtracted from the time-token key (bias 0) of block 0
            const int nblk = a9_nblk(w);
            float l_run = 0.f;
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
                    tmem_ld_wait_dep(ca);                    // chunk 0 (requested when the block's scores arrived)
                    if (quad == 0) A10_STAMP(half, 2);
                    if (half == 0) A10_STAMP2(quad, 2);
                    tmem_ld_32x32(tmem_S + 32, cb);
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
                                cur[0] = __float_as_uint(__uint_as_float(cur[0]) + s_i);      // key 0 (time token): no distance penalty
                            }
                            if (partial) {
#pragma unroll
                                for (int e = 0; e < 32; ++e)
                                    if (k0 + c * 32 + e >= len) cur[e] = __float_as_uint(-INFINITY);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < 16; ++u) {
                            const uint64_t x2 = pack2f(__uint_as_float(cur[2 * u]), __uint_as_float(cur[2 * u + 1]));
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
                    if (quad == 0) A10_STAMP(half, 3);
                    tmem_ld_32x32(tmem_S + 64, ca);
                    proc(cb, 1);
                    tmem_ld_wait_dep(ca);
                    if (quad == 0) A10_STAMP(half, 4);
                    tmem_ld_32x32(tmem_S + 96, cb);
                    proc(ca, 2);
                    tmem_ld_wait_dep(cb);
                    if (quad == 0) A10_STAMP(half, 5);
                    if (half == 0) A10_STAMP2(quad, 5);
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
                if (quad == 0) A10_STAMP(half, 6);
                // ---- P_half must have been read by the previous PV of this half before it is overwritten
                if (cblk > 0 && !(LAB & 16)) wait_pv(cblk - 1);
                tc_fence_after();
                if (quad == 0) A10_STAMP(half, 7);
#pragma unroll
                for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_P + c * 16, *reinterpret_cast<uint32_t(*)[16]>(&s[c * 16]));
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_ready[half]);
                if (quad == 0) A10_STAMP(half, 8);
                if (half == 0) A10_STAMP2(quad, 8);
                ++cblk;
                if (more) {                                  // scores of the half's next block: wait, then request their first 32 columns
                    mbar_wait(&s_full[half], cblk & 1);
                    if (quad == 0) A10_STAMP(half, 1);
                    if (half == 0) A10_STAMP2(quad, 1);
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

template <int POLY, int LAB = 0>
inline cudaError_t launch_attn10_t(const CUtensorMap& tmQKV, const CUtensorMap& tmPos, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  const int4* win, const int* order, int nwork, int num_sms, cudaStream_t st, bool pdl) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn10_kernel<POLY, LAB>, cudaFuncAttributeMaxDynamicSharedMemorySize, A10_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int n_items = nwork * nheads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ATT_THREADS); cfg.dynamicSmemBytes = A10_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, attn10_kernel<POLY, LAB>, tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, n_items);
}
// The q columns of tmQKV must have been produced with W_q, b_q scaled by head_dim^-0.5 * log2(e); tmPos maps the position table of
// k_attn10_pos ([rows, 16] bf16, box 16 x 128, 32-byte swizzle); normtab holds the row-block norms of the same rows (QKV GEMM epilogue
// or k_attn9_rownorm); win is scratch for nwork * nheads int4.
inline cudaError_t launch_attn10_raw(const CUtensorMap& tmQKV, const CUtensorMap& tmPos, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                    int nwork, const float* normtab, int4* win, int* order, int use_window, int sort, int* err_flag, int poly,
                                    int num_sms, cudaStream_t st, bool pdl = false) {
    const int n_items = nwork * nheads;
    cudaError_t e = launch_attn9_plan(work, n_items, nheads, slopes, normtab, use_window, win, err_flag, st, pdl);
    if (e != cudaSuccess) return e;
    if (!(sort && use_window)) order = nullptr;
    if (order != nullptr) k_attn10_order<<<1, 1024, 0, st>>>(win, n_items, 1, order);
    e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    switch (poly) {
        case 0: return launch_attn10_t<0>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        case 2: return launch_attn10_t<2>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        case 3: return launch_attn10_t<3>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        case 4: return launch_attn10_t<4>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
