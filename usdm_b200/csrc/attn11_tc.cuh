// Fused masked self-attention, version 11 = attn10_tc.cuh pipelined at 64-key granularity.
//
// What v10's lab trace showed (profiles/r02/attn_lab_v9_trace.txt): with every bit of softmax arithmetic removed the
// kernel still needs 1240 of its 1728 cycles per 128 x 128 score block -- the per-half chain  score-ready -> TMEM loads -> S handed back
// -> QK^T issue -> score-ready  (the QK^T of block c + 1 cannot start before the LAST column of block c is in registers) plus the wait
// for the previous PV, the P store and its completion wait, all serial in each softmax warp.  A second 128-column score buffer per half
// does not fit the 512 tensor-memory columns.  Here the same columns are handed over in halves ("pieces" of 64 keys):
//   * S_half = [piece 0 | piece 1] of 64 columns each, P_half likewise 2 x 32 columns, each piece with its own full / free barriers;
//   * QK^T is issued per piece (M = 128, N = 64, K = 64 + 16): piece 0 of block c + 1 is computed while the softmax warps are still
//     reading / exponentiating piece 1 of block c, i.e. every half is double-buffered at piece granularity inside the old footprint;
//   * PV is issued per piece (K = 64): the tensor core consumes P piece 0 while piece 1 is being exponentiated;
//   * the softmax warps run ONE rolling chunk pipeline (32 columns per tcgen05.ld) across pieces and blocks: the next chunk is always
//     requested before the current one is processed, the P piece store is issued when its second chunk is done and its completion
//     wait + hand-over are deferred behind the next chunk's arithmetic (the store has retired by then);
//   * barrier arrivals are one per warp (count 4) instead of one per thread (count 128).
// K / V / position stages, the bias operands, the plan (windows) and the item order are v10's.
#pragma once
#include "attn10_tc.cuh"

namespace tvb {

constexpr int A11_SMEM_BYTES = A10_SMEM_BAR + 512;     // v10's tiles; 32 barriers + the tensor-memory slot

#ifdef A11_LAB_TRACE
__device__ long long* g_a11_trace;
#define A11_STAMP(role, tag)                                                                                                  \
    do {                                                                                                                     \
        if ((LAB & 32) && blockIdx.x == 0 && lane == 0 && tr_n < 1024)                                                       \
            g_a11_trace[(role) * 1024 + tr_n++] = (static_cast<long long>(tag) << 48) | (clock64() & 0xffffffffffffll);       \
    } while (0)
#else
#define A11_STAMP(role, tag) do { } while (0)
#endif

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
        "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
        "r"(r[31])
        : "memory");
}

// LAB (tools/attn_lab.cu only; the library always instantiates LAB = 0): bit 0: no exponential, bit 1: no row sum, bit 3: no bf16 pack,
// bit 5: event trace of CTA 0.  Results are meaningless with LAB & 11.
template <int POLY, int LAB = 0, int SPLIT = 1>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attn11_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmPos, __nv_bfloat16* __restrict__ ctx, int H, int nheads,
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
    uint64_t* s_full = kv_empty + A10_KV_STAGES;        // [2 halves][2 pieces]: the scores of the piece are in tensor memory
    uint64_t* s_free = s_full + 4;                      // [2][2]: the 4 softmax warps of the half hold the piece in registers
    uint64_t* p_ready = s_free + 4;                     // [2][2]: P piece written
    uint64_t* pv_done = p_ready + 4;                    // [2][2]: the PV MMAs reading the P piece have retired
    uint64_t* o_free = pv_done + 4;                     // [2] per half: O has been read out by the half's softmax warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 2);
    static_assert((10 + 2 * A10_KV_STAGES + 16) * 8 + 8 <= A11_SMEM_BYTES - A10_SMEM_BAR, "barrier area");

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = gridDim.x;
    [[maybe_unused]] int tr_n = 0;

    if (threadIdx.x == 0 && (smem_u32(smem) & 1023u) != 0) __trap();   // swizzle-128B tiles need 1024-B alignment
    if (warp_idx == 0 && lane == 0) { tma_prefetch_desc(&tmQKV); tma_prefetch_desc(&tmPos); }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&q_full[s], 1); mbar_init(&q_empty[s], 1); mbar_init(&ext_full[s], 1); mbar_init(&ext_empty[s], 1);
            mbar_init(&o_free[s], 4);
        }
        for (int s = 0; s < 4; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 4); mbar_init(&p_ready[s], 4); mbar_init(&pv_done[s], 1); }
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
        // ------------------------------------------ TMA producer (as v10) ------------------------------------------
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
        // ------------------------------------------ query-side bias tile writer (as v10) ------------------------------------------
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
        // ------------------------------------------ PV issuer (converged warp): one K = 64 chain per P piece ------------------------------------------
        constexpr uint32_t idesc_pv = make_idesc_bf16(128, ATT_D, false, true);          // B (= V) MN-major
        const uint64_t dv0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_KV + ATT_TILE_BYTES), 8192, 1024);
        int slot = 0;
        int cnt[2] = {0, 0};                                // items in which the half has taken part so far
        int pvc[2] = {0, 0};                                // blocks issued so far per half
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
                    const uint32_t tO = tmem_base + 384 + half * 64;
                    const uint64_t dv = dv0 + static_cast<uint64_t>(slot * (A10_STAGE_BYTES >> 4));
#pragma unroll
                    for (int piece = 0; piece < 2; ++piece) {
                        mbar_wait(&p_ready[2 * half + piece], c & 1);
                        A11_STAMP(2, 20 + 2 * half + piece);
                        if (jj == 0 && piece == 0) {                                          // this PV overwrites O: the previous item's
                            if (cnt[half] > 0) mbar_wait(&o_free[half], (cnt[half] - 1) & 1); // epilogue must have read it out
                            ++cnt[half];
                        }
                        tc_fence_after();
                        const uint32_t tP = tmem_base + 256 + half * 64 + piece * 32;
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < 4; ++k)                                       // V rows [64 piece + 16k, +16); P: 8 columns per K step
                                umma_ts(tO, tP + k * 8, dv + (4 * piece + k) * (2048 >> 4), idesc_pv, (jj | piece | k) != 0);
                            umma_commit(&pv_done[2 * half + piece]);
                            if (half == last_half && piece == 1) umma_commit(&kv_empty[slot]);  // every MMA reading this stage has been issued
                        }
                        __syncwarp();
                    }
                }
                if (++slot == A10_KV_STAGES) slot = 0;
            }
            w = wn;
        }
    } else if (warp_idx == 2) {
        // ------------------------------------------ QK issuer (converged warp): one N = 64 chain per S piece ------------------------------------------
        constexpr uint32_t idesc_qk = make_idesc_bf16(128, 64);                          // A, B K-major
        const uint64_t dq0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_Q), 16, 1024);
        const uint64_t dk0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_KV), 16, 1024);
        const uint64_t de0 = make_smem_desc_sw128(smem_u32(smem + A10_SMEM_EXT), 16, 1024);
        const uint64_t dp0 = make_smem_desc_sw32(smem_u32(smem + A10_SMEM_KV + 2 * ATT_TILE_BYTES));
        int slot = 0, itc = 0;
        uint32_t kvpar = 0;
        int qkc[2] = {0, 0};                                // blocks issued so far per half
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
                    const uint64_t da = dq0 + static_cast<uint64_t>((qb * 2 + half) * (ATT_TILE_BYTES >> 4));
                    const int jd = w.qb0 + half;
                    const bool last = (jj == nblk - 1) && (half == last_half);
                    if constexpr (SPLIT == 0) {
                        // ablation: one N = 128 chain per block (v10's QK^T), both pieces signalled together
                        mbar_wait(&s_free[2 * half], (c & 1) ^ 1);
                        mbar_wait(&s_free[2 * half + 1], (c & 1) ^ 1);
                        A11_STAMP(3, 30 + 2 * half);
                        tc_fence_after();
                        constexpr uint32_t idesc_qk128 = make_idesc_bf16(128, ATT_BK);
                        const uint32_t d = tmem_base + half * 128;
                        const uint64_t db = dk0 + static_cast<uint64_t>(slot * (A10_STAGE_BYTES >> 4));
                        const uint64_t dpos = dp0 + static_cast<uint64_t>(slot * (A10_STAGE_BYTES >> 4));
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk128, k != 0);
                            if (j != jd) umma_ss(d, dext + 2 * (2 * half + (j < jd ? 0 : 1)), dpos, idesc_qk128, 1);
                            umma_commit(&s_full[2 * half]);
                            umma_commit(&s_full[2 * half + 1]);
                            if (last) {
                                umma_commit(&q_empty[qb]);
                                umma_commit(&ext_empty[qb]);
                            }
                        }
                        __syncwarp();
                    } else {
#pragma unroll
                    for (int piece = 0; piece < 2; ++piece) {
                        mbar_wait(&s_free[2 * half + piece], (c & 1) ^ 1);   // piece of block c - 1 is in registers (fresh barrier: passes)
                        A11_STAMP(3, 30 + 2 * half + piece);
                        tc_fence_after();
                        const uint32_t d = tmem_base + half * 128 + piece * 64;
                        // keys [64 piece, +64) of the stage: 64 rows of 128 B (K) / of 32 B (positions)
                        const uint64_t db = dk0 + static_cast<uint64_t>((slot * A10_STAGE_BYTES + piece * 8192) >> 4);
                        const uint64_t dpos = dp0 + static_cast<uint64_t>((slot * A10_STAGE_BYTES + piece * 2048) >> 4);
                        if (elect_one()) {
#pragma unroll
                            for (int k = 0; k < ATT_D / 16; ++k) umma_ss(d, da + 2 * k, db + 2 * k, idesc_qk, k != 0);
                            // the whole additive term of an off-diagonal block: query-side group (half, side) x the key positions of the stage
                            if (j != jd) umma_ss(d, dext + 2 * (2 * half + (j < jd ? 0 : 1)), dpos, idesc_qk, 1);
                            umma_commit(&s_full[2 * half + piece]);
                            if (last && piece == 1) {                                         // last QK of the item: Q and bias buffers are free
                                umma_commit(&q_empty[qb]);
                                umma_commit(&ext_empty[qb]);
                            }
                        }
                        __syncwarp();
                    }
                    }
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
        uint64_t* const sfull = s_full + 2 * half;
        uint64_t* const sfree = s_free + 2 * half;
        uint64_t* const pready = p_ready + 2 * half;
        uint64_t* const pvdone = pv_done + 2 * half;
        int cblk = 0;                                       // blocks this half has processed so far (all items)
        int pend = -1;                                      // P piece whose store has been issued but not yet handed over
        // one arrival per warp: every lane has passed the warp-wide tcgen05.wait before this is called
        auto warp_arrive = [&](uint64_t* bar) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar);
        };
        auto flush_pend = [&]() {
            if (pend >= 0) {
                tmem_st_wait();
                warp_arrive(&pready[pend]);
                pend = -1;
            }
        };
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
            float l_run = 0.f;                              // row sum; per block in v10's order of additions, so that the two kernels give the same bits
            uint32_t ca[32], cb[32];                        // two 32-column chunks of S in flight
            uint32_t pk[32];                                // bf16 pairs of the P piece being built
            // the half's first block of the item (always key block 0): wait for piece 0, request the first chunk
            mbar_wait(&sfull[0], cblk & 1);
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
                            pk[(c & 1) * 16 + u] = (LAB & 8) ? __float_as_uint(lo) ^ __float_as_uint(hi) : pack_bf16x2(lo, hi);
                        }
                    };
                    // P piece (keys [64 piece, +64) of the block) is complete in pk: its region must have been read by the PV of the
                    // previous block's piece, then the store goes out; completion wait and hand-over are deferred (flush_pend)
                    auto store_piece = [&](const int piece) {
                        if (cblk > 0) mbar_wait(&pvdone[piece], (cblk - 1) & 1);
                        tc_fence_after();
                        tmem_st_32x32(tmem_P + piece * 32, pk);
                        pend = piece;
                    };
                    // ---- chunk 0 (requested when the block's piece 0 arrived)
                    tmem_ld_wait_dep(ca);
                    if (quad == 0) A11_STAMP(half, 2);
                    tmem_ld_32x32(tmem_S + 32, cb);
                    proc(ca, 0);
                    flush_pend();                            // piece 1 of the previous block: its store retired during chunk 0's arithmetic
                    // ---- chunk 1: piece 0 is in registers -> hand its columns back (QK^T piece 0 of the next block may start)
                    tmem_ld_wait_dep(cb);
                    if (quad == 0) A11_STAMP(half, 3);
                    warp_arrive(&sfree[0]);
                    mbar_wait(&sfull[1], cblk & 1);
                    tc_fence_after();
                    tmem_ld_32x32(tmem_S + 64, ca);
                    proc(cb, 1);
                    store_piece(0);
                    // ---- chunk 2
                    tmem_ld_wait_dep(ca);
                    if (quad == 0) A11_STAMP(half, 4);
                    tmem_ld_32x32(tmem_S + 96, cb);
                    proc(ca, 2);
                    flush_pend();                            // piece 0 handed to the PV issuer
                    // ---- chunk 3: piece 1 is in registers
                    tmem_ld_wait_dep(cb);
                    if (quad == 0) A11_STAMP(half, 5);
                    warp_arrive(&sfree[1]);
                    if (more) {                              // first chunk of the half's next block: its piece 0 was issued long ago
                        mbar_wait(&sfull[0], (cblk + 1) & 1);
                        if (quad == 0) A11_STAMP(half, 1);
                        tc_fence_after();
                        tmem_ld_32x32(tmem_S, ca);
                    }
                    proc(cb, 3);
                    store_piece(1);
                    float a0, a1;
                    unpack2f(fadd2(acc2, accp2), a0, a1);
                    l_run += a0 + a1;
                    if (quad == 0) A11_STAMP(half, 6);
                };
                if (j != 0 && j != jd && k0 + ATT_BK <= len) body(std::false_type{});
                else body(std::true_type{});
                ++cblk;
                if (more) jnext = a9_block(w, jj2);
                jj = jj2;
            }
            flush_pend();
            // epilogue: read O, hand the accumulator back, then O / l -> ctx (bf16), only rows that belong to the sequence
            mbar_wait(&pvdone[0], (cblk - 1) & 1);
            mbar_wait(&pvdone[1], (cblk - 1) & 1);
            tc_fence_after();
            uint32_t o[64];
#pragma unroll
            for (int c = 0; c < 2; ++c) tmem_ld_32x32(tmem_O + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&o[c * 32]));
            tmem_ld_wait();
            warp_arrive(&o_free[half]);
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

template <int POLY, int LAB = 0, int SPLIT = 1>
inline cudaError_t launch_attn11_t(const CUtensorMap& tmQKV, const CUtensorMap& tmPos, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
                                  const int4* win, const int* order, int nwork, int num_sms, cudaStream_t st, bool pdl) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(attn11_kernel<POLY, LAB, SPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize, A11_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int n_items = nwork * nheads;
    const int grid = n_items < num_sms ? n_items : num_sms;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(ATT_THREADS); cfg.dynamicSmemBytes = A11_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, attn11_kernel<POLY, LAB, SPLIT>, tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, n_items);
}
// Same contract as launch_attn10_raw.
inline cudaError_t launch_attn11_raw(const CUtensorMap& tmQKV, const CUtensorMap& tmPos, __nv_bfloat16* ctx, int H, int nheads, const float* slopes, const int4* work,
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
        case 0: return launch_attn11_t<0>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        case 2: return launch_attn11_t<2>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        case 3: return launch_attn11_t<3>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        case 4: return launch_attn11_t<4>(tmQKV, tmPos, ctx, H, nheads, slopes, work, win, order, nwork, num_sms, st, pdl);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace tvb
