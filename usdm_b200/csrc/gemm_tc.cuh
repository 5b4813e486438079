// bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores:  C[M,N] = A[M,K] * W[N,K]^T (+ fused epilogue)
//
// Replaces every nn.Linear on the hot path (reference networks.py:162-166 q/k/v, :208 out_proj, :226/:230 FFN,
// :363 skip merge).  Both operands are K-major (nn.Linear weights are stored [out, in], activations [tokens, in]).
//
// Structure (one CTA per SM, persistent over output tiles):
//   warp 0      : TMA producer   (cp.async.bulk.tensor 2D, SWIZZLE_128B boxes of 64 bf16 columns, 4-stage ring)
//   warp 1      : MMA issuer     (one elected thread: tcgen05.mma.cta_group::1.kind::f16, M=128, N=256, K=16)
//   warps 2..5  : epilogue       (tcgen05.ld 32x32b from TMEM, bias / GELU / residual, vectorised global stores)
// TMEM holds two 128x256 fp32 accumulators (512 columns) so the epilogue of tile i overlaps the mainloop of tile i+1.
#pragma once
#include "common.cuh"

namespace tvb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BN = 256;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 4;
constexpr int GEMM_A_BYTES = GEMM_BM * GEMM_BK * 2;                 // 16 KB
constexpr int GEMM_B_BYTES = GEMM_BN * GEMM_BK * 2;                 // 32 KB
constexpr int GEMM_STAGE_BYTES = GEMM_A_BYTES + GEMM_B_BYTES;       // 48 KB
constexpr int GEMM_SMEM_BYTES = GEMM_STAGES * GEMM_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int GEMM_THREADS = 192;
constexpr float GEMM_FOLD_MEAN_LIMIT = 8.0f;

enum GemmEpilogue : int {
    EPI_BF16 = 0,        // out_bf16 = acc + bias
    EPI_GELU_BF16 = 1,   // out_bf16 = gelu_erf(acc + bias)
    EPI_RESID_F32 = 2,   // out_f32  = acc + bias + resid
    EPI_DUAL = 3,        // out_f32  = acc + bias ; out_bf16 = bf16(out_f32)
    EPI_F32 = 4,         // out_f32  = acc + bias
    EPI_RESID_LN = 5,    // out_f32  = acc + bias + LayerNorm(resid)  with per-row (mean, rstd) from ln_stats and per-column ln_g / ln_b:
                         // the post-LN residual stream is re-normalised on the fly instead of being written and re-read in fp32
    // LayerNorm folded into the NEXT GEMM (N == 1024 producers only): the residual epilogue also emits a bf16 copy of its
    // un-normalised result and per-row partial (sum, sum of squares) per 128-column slice; the consumer multiplies the raw rows
    // with W diag(gamma) and applies  rstd * acc - rstd * mean * colsum(W diag(gamma)) + (b + W beta)  in its epilogue.
    EPI_RESID_F32_ST = 6,   // EPI_RESID_F32 + bf16 copy + partial statistics
    EPI_RESID_LN_ST = 7,    // EPI_RESID_LN  + bf16 copy + partial statistics
    EPI_FOLD_GELU = 8,      // out_bf16 = gelu(rstd * acc - rstd * mean * fold_c + bias)  (bias already holds b + W beta)
};

struct GemmParams {
    int M, N, K;          // M % 128 == 0, N % 256 == 0, K % 64 == 0
    int K1;               // columns [0,K1) of the reduction come from A, [K1,K) from A2 (skip-merge concat); K1 == K if unused
    const float* bias;    // [N] or nullptr
    const float* resid;   // [M, ld_out] fp32 (EPI_RESID_F32)
    float* out_f32;       // [M, ld_out]
    __nv_bfloat16* out_bf16;
    int ld_out;           // row stride of outputs / resid, in elements
    const float2* ln_stats;   // [M] (mean, rstd) of the resid rows            (EPI_RESID_LN)
    const float* ln_g;        // [N] LayerNorm weight / bias of the residual's norm
    const float* ln_b;
    float2* part;             // [M, part_n] per-row partial (sum, sum of squares), N == 1024   (*_ST write, FOLD reads).  The unit is a 64-column
                              // group g_q summed in column order.  A 256-column writer stores e_k = g_2k + g_2k+1 (part_n == 8, one entry per
                              // 128 columns: 4 float4 per row for the reader); a 128-column writer owns only one group per warp and stores the
                              // g_q themselves (part_n == 16), the reader then forms e_k.  Every rounding is the same either way: an utterance's
                              // result must not depend on which width its batch size selected (batch invariance, tests/test_gpu_configs.py)
    int part_n;               // 8 or 16, see `part`
    const float* fold_c;      // [N] column sums of the folded weight                                              (EPI_FOLD_GELU)
    float2* stats_out;        // [M] (mean, rstd), written once per row by the n == 0 tiles                        (EPI_FOLD_GELU)
    int* fold_flag;           // EPI_FOLD_GELU: mapped host word set to 1 when a row's |mean| * rstd exceeds GEMM_FOLD_MEAN_LIMIT.  The folded LayerNorm
                              // multiplies the bf16-rounded RAW row and subtracts mean * colsum afterwards, so the operand rounding is amplified by
                              // |mean| / std against the reference's round-after-normalise; rows like that (a large common offset across the hidden
                              // dimension) do not occur with random or ordinary weights, and when they do the caller should run with ln1_fold = 0
    float* rownorm;           // EPI_BF16, CTA-pair kernels: when set, normtab[row / 32][col / 64] = max over the 32 rows of the squared norm of the 64
    int rownorm_cols;         //   output columns (fp32 values before the bf16 rounding), for columns < rownorm_cols: the |q|, |k| row-block maxima
    int rownorm_ld;           //   the attention kernel bounds its logits with (attn9_tc.cuh); row stride rownorm_ld
    int resid_l2;             // CTA-pair residual epilogues (ablation, default 0): 1 = every epilogue warp pulls the residual chunks of its NEXT tile into
                              // the L2 one tile ahead (cp.async.bulk.prefetch.tensor), 2 = the chunk after the next one (the staged TMA load itself runs
                              // one chunk ahead).  Measured (profiles/r02): 1 costs the attention-out GEMM 7 % and FFN-out 4 % -- the kernel is HBM
                              // bandwidth bound (452 MB per 77 GFLOP launch), not latency bound, and the early lines are evicted before they are used
    int a_l2;                 // CTA-pair kernels (ablation, default 0): the TMA producer pulls the activation k-block a_l2 blocks ahead of the one it loads
                              // into the L2 (the n-tiles of a row block, which run on neighbouring clusters at the same time, share the k-blocks
                              // between them).  Measured (profiles/r02): a whole tile ahead costs every GEMM 3-7 % (74 MB of prefetched rows
                              // displace the weights and the live rows in the 126 MB L2)
    int lab;                  // timing ablations of the CTA-pair *_ST residual epilogue (option gemm_lab; results are WRONG with any bit set): bit 0 = the
                              // residual tile is not loaded, bit 1 = no fp32 store, bit 2 = no bf16 copy store
    int reverse;              // CTA-pair kernels: walk the 256-row blocks from the last to the first.  Consumers traverse opposite to
                              // the kernel that produced their input, so the rows they read first are the ones still in the 126 MB L2.
};

template <int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * GEMM_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + GEMM_STAGES;
    uint64_t* tfull_bar = empty_bar + GEMM_STAGES;     // [2] accumulator ready
    uint64_t* tempty_bar = tfull_bar + 2;              // [2] accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp_idx = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int tiles_n = p.N / GEMM_BN;
    const int num_tiles = (p.M / GEMM_BM) * tiles_n;
    const int num_kb = p.K / GEMM_BK;

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4);   // one arrive per epilogue warp
        }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        // ------------------------------ TMA producer ------------------------------
        if (elect_one()) {      // one elected lane: the compiler then keeps descriptors / addresses in uniform registers
            int stage = 0;
            uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int m0 = (tile / tiles_n) * GEMM_BM;
                const int n0 = (tile % tiles_n) * GEMM_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * GEMM_STAGE_BYTES;
                    uint8_t* sb = sa + GEMM_A_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], GEMM_STAGE_BYTES);
                    const int k0 = kb * GEMM_BK;
                    if (k0 < p.K1) tma_load_2d(sa, &tmA, &full_bar[stage], k0, m0);
                    else           tma_load_2d(sa, &tmA2, &full_bar[stage], k0 - p.K1, m0);
                    tma_load_2d(sb, &tmB, &full_bar[stage], k0, n0);
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        // ------------------------------ MMA issuer ------------------------------
        if (elect_one()) {      // one elected lane: the compiler then keeps descriptors / addresses in uniform registers
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, GEMM_BN);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * GEMM_BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * GEMM_STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + GEMM_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // advancing 16 bf16 (32 B) along K inside the 128-B swizzle atom = +2 in the (>>4) address field
                        umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[stage]);          // smem slot reusable once these MMAs retire
                    if (++stage == GEMM_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);                 // accumulator complete
            }
        }
    } else {
        // ------------------------------ epilogue warps ------------------------------
        const int quad = warp_idx & 3;                       // TMEM lane quadrant this warp may access
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / tiles_n) * GEMM_BM;
            const int n0 = (tile % tiles_n) * GEMM_BN;
            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const int row = m0 + quad * 32 + lane;
            const size_t row_off = static_cast<size_t>(row) * p.ld_out;
#pragma unroll 1
            for (int c = 0; c < GEMM_BN / 32; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * GEMM_BN + c * 32, r);
                tmem_ld_wait();
                const int col = n0 + c * 32;
                float v[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                if (p.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = __ldg(b4 + j);
                        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                    }
                }
                if constexpr (EPI == EPI_GELU_BF16) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf(v[j]);
                }
                if constexpr (EPI == EPI_RESID_F32) {
                    const float4* r4 = reinterpret_cast<const float4*>(p.resid + row_off + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = r4[j];
                        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                    }
                }
                if constexpr (EPI == EPI_RESID_LN) {
                    const float2 st = p.ln_stats[row];
                    const float4* r4 = reinterpret_cast<const float4*>(p.resid + row_off + col);
                    const float4* g4 = reinterpret_cast<const float4*>(p.ln_g + col);
                    const float4* b4 = reinterpret_cast<const float4*>(p.ln_b + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 x = r4[j], g = __ldg(g4 + j), b = __ldg(b4 + j);
                        v[4 * j + 0] += fmaf((x.x - st.x) * st.y, g.x, b.x); v[4 * j + 1] += fmaf((x.y - st.x) * st.y, g.y, b.y);
                        v[4 * j + 2] += fmaf((x.z - st.x) * st.y, g.z, b.z); v[4 * j + 3] += fmaf((x.w - st.x) * st.y, g.w, b.w);
                    }
                }
                if constexpr (EPI == EPI_RESID_F32 || EPI == EPI_RESID_LN || EPI == EPI_DUAL || EPI == EPI_F32) {
                    float4* o4 = reinterpret_cast<float4*>(p.out_f32 + row_off + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                if constexpr (EPI == EPI_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_DUAL) {
                    uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + row_off + col);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}


// ================================================================================================================
// CTA-pair version (default): a cluster of 2 CTAs on adjacent SMs computes one 256 x 256 output tile with
// tcgen05.mma.cta_group::2 (M = 256).  Each CTA stages its own 128 rows of A and HALF (128 rows) of the W tile, so a
// k-block costs 32 KB of L2->smem traffic per CTA instead of 48 KB and the smem read traffic of the MMA halves.
//   warp 0 TMA producer (both CTAs; completion bytes credited to the leader's barrier), warp 1 MMA issuer (leader CTA
//   only; commits multicast to both CTAs), warps 2..9 epilogue (8 warps: TMEM lane quadrant x column half), each CTA
//   draining its own 128 x 256 fp32 accumulator.  4-stage operand ring, 2 accumulator stages.
// Epilogue data path (measured: per-row direct global accesses made the K=1024 GEMMs epilogue-bound at 30-65 % tensor
// activity, profiles/r01): every warp owns two 4 KB smem staging tiles (32 rows x 128 B, 128-B swizzle).  The residual
// tile is TMA-loaded into the staging tile ahead of time, the thread that owns a TMEM lane adds bias/residual (or GELU)
// to its row in place, and one elected lane TMA-stores the tile: all global traffic is coalesced by the TMA engine.
// ================================================================================================================
constexpr int G2_STAGES = 4;
#ifndef G2_EXTRA_STAGE
#define G2_EXTRA_STAGE 1      // epilogues without the bf16 copy staging run a 5-deep operand ring (see gemm_pair_body)
#endif
constexpr int G2_A_BYTES = 128 * GEMM_BK * 2;                // 16 KB (this CTA's 128 rows of A)
constexpr int G2_B_BYTES = 128 * GEMM_BK * 2;                // 16 KB (this CTA's half of the 256 W rows)
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_EPI_WARPS = 8;
constexpr int G2_EPI_TILE = 4096;                            // 32 rows x 128 B
constexpr int G2_EPI_BUFS = 2;                               // staging tiles per epilogue warp (a third one measured no gain)
constexpr int G2_SMEM_EPI = G2_STAGES * G2_STAGE_BYTES;
constexpr int G2_EPI_TILE2 = 2048;                           // 32 rows x 64 B: bf16 copy of a 32 x 32 chunk (EPI_RESID_*_ST), 64-B swizzle
constexpr int G2_SMEM_EPI2 = G2_SMEM_EPI + G2_EPI_WARPS * G2_EPI_BUFS * G2_EPI_TILE;
constexpr int G2_SMEM_BAR = G2_SMEM_EPI2 + G2_EPI_WARPS * G2_EPI_BUFS * G2_EPI_TILE2;
constexpr int G2_SMEM_BYTES = G2_SMEM_BAR + 1024 + 512;
constexpr int G2_THREADS = 320;

struct Gemm2Maps {
    CUtensorMap a, a2, b;     // operands
    CUtensorMap out;          // fp32 [M,N] box 32x32 (EPI_RESID_F32 / EPI_DUAL / EPI_F32) or bf16 [M,N] box 64x32 (EPI_BF16 / EPI_GELU_BF16)
    CUtensorMap resid;        // fp32 [M,N] box 32x32 (EPI_RESID_F32)
    CUtensorMap out2;         // bf16 [M,N] box 32x32, 64-B swizzle: un-normalised copy for the LayerNorm-folding consumer (EPI_RESID_*_ST)
};

// NP = CTA pairs per cluster.  NP == 1: the kernel described above.  NP == 2 (cluster of 4): the two pairs work on vertically
// adjacent 256-row tiles of the SAME 256 W rows; every CTA fetches only 64 of the 128 W rows it needs and TMA-multicasts them to
// the CTA of the other pair that needs the same rows, so a k-block costs 24 KB of L2->SM traffic per CTA instead of 32 KB
// (measured: the NP == 1 kernel saturates the L2->SM fabric at ~10 TB/s, i.e. ~1.2-1.3 PFLOP/s at 128 FLOP/B, profiles/r01).
// A stage may be overwritten once BOTH pairs have consumed it: each pair leader's commit arrives on the empty barrier of all
// four CTAs (barrier count NP).
// BN = accumulator columns per pair tile.  256: each CTA stages 128 W rows per k-block.  128 (NP == 1 only; tm.b = the box-64
// weight map): half-width tiles for launches whose 256-wide tiling leaves most CTA pairs idle (M <= a few 256-row blocks, i.e. the
// single-utterance requests of the reference app: N = 1024 gives 4 tiles per row block, 24 pairs of 74 busy at B = 1) -- the
// host picks the width that needs fewer (waves x tile time), see linear().
// DIRECT (fp32-output epilogues, NP == 1; ablation, option epi_direct): the epilogue goes straight between registers and global memory
// instead of through staging tiles.  The staged path moves every epilogue byte through shared memory twice (TMA in, LDS, STS, TMA out:
// 640 KB per attention-out tile of 8192 MMA cycles) on top of what the operand ring and the MMA's own operand reads take from it.  Here
// the accumulator is read with tcgen05.ld.16x256b (the MMA C-fragment layout: a quad of threads holds 8 consecutive columns of a row,
// tools/tmem_layout.cu), so a warp's 8-byte accesses cover whole 32-byte sectors of 8 rows: sector-exact global loads of the residual
// and stores of the result without any staging, and the same bits.  Measured: the K = 4096 GEMM's mainloop gains 1.5 % from the freed
// shared-memory bandwidth, but wherever the epilogue is the critical path (K = 1024, attention-out: 82 -> 120 ms per solve) three 8-byte
// accesses per two columns are LSU-issue bound -- so the staged path stays the default.
template <int EPI, int NP, int BN = 256, int DIRECT = 0>
__device__ __forceinline__ void gemm_pair_body(const Gemm2Maps& tm, const GemmParams& p) {
    static_assert(BN == 256 || (BN == 128 && NP == 1), "tile width");
    static_assert(!DIRECT || (NP == 1 && (EPI == EPI_RESID_F32 || EPI == EPI_RESID_LN || EPI == EPI_RESID_F32_ST || EPI == EPI_RESID_LN_ST ||
                                          EPI == EPI_DUAL || EPI == EPI_F32)), "direct epilogue: fp32-output epilogues of the pair kernel");
    constexpr int kHalf = BN / 2;                     // accumulator columns (= W rows) owned by one CTA / one epilogue column half
    constexpr uint32_t kStageTx = G2_A_BYTES + kHalf * GEMM_BK * 2;   // bytes one CTA receives per k-block
    extern __shared__ uint8_t smem_raw[];
    // identical carve-up in both CTAs (the dynamic smem base offset is the same for every CTA of a kernel)
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // operand ring depth: the *_ST epilogues need the extra bf16 staging tiles (32 KB), everything else spends that memory on a
    // fifth stage (an operand tile comes back ~1-2 k cycles after its TMA is issued when the L2 is busy; 4 stages = 2 k cycles of MMA)
    constexpr int kStages = (EPI == EPI_RESID_F32_ST || EPI == EPI_RESID_LN_ST) ? G2_STAGES : G2_STAGES + G2_EXTRA_STAGE;
    constexpr int kSmemEpi = kStages * G2_STAGE_BYTES;
    constexpr int kSmemEpi2 = kSmemEpi + G2_EPI_WARPS * G2_EPI_BUFS * G2_EPI_TILE;
    constexpr int kSmemBar = (kStages == G2_STAGES) ? kSmemEpi2 + G2_EPI_WARPS * G2_EPI_BUFS * G2_EPI_TILE2 : kSmemEpi2;
    static_assert(kSmemBar + 1024 + 512 <= G2_SMEM_BYTES, "shared memory budget");
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kSmemBar);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tfull_bar = empty_bar + kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint64_t* rbar = tempty_bar + 2;                  // [8 warps][G2_EPI_BUFS] residual-tile arrival
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(rbar + G2_EPI_BUFS * G2_EPI_WARPS);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t crank = cluster_ctarank();
    const uint32_t cta = crank & 1;                  // position inside the pair, 0 = leader
    const int pair_id = static_cast<int>(crank >> 1);
    const int tiles_n = p.N / BN;
    const int num_tiles = (p.M / (256 * NP)) * tiles_n;      // cluster tiles: NP vertically adjacent 256 x BN tiles
    const int num_kb = p.K / GEMM_BK;
    const int cluster_id = blockIdx.x / (2 * NP), num_clusters = gridDim.x / (2 * NP);
    const int m_blocks = p.M / (256 * NP);
    auto tile_m0 = [&](int tile) {
        const int mb = tile / tiles_n;
        return ((p.reverse ? m_blocks - 1 - mb : mb) * NP + pair_id) * 256;
    };

    if (warp_idx == 0 && lane == 0) {
        tma_prefetch_desc(&tm.a); tma_prefetch_desc(&tm.a2); tma_prefetch_desc(&tm.b); tma_prefetch_desc(&tm.out);
        if (EPI == EPI_RESID_F32_ST || EPI == EPI_RESID_LN_ST) tma_prefetch_desc(&tm.out2);
        if (EPI == EPI_RESID_F32 || EPI == EPI_RESID_LN || EPI == EPI_RESID_F32_ST || EPI == EPI_RESID_LN_ST) tma_prefetch_desc(&tm.resid);
    }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], NP); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 2 * G2_EPI_WARPS); }
        for (int s = 0; s < G2_EPI_BUFS * G2_EPI_WARPS; ++s) mbar_init(&rbar[s], 1);
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc_2cta<512>(tmem_slot);
    pdl_launch_dependents();
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();                                      // everything above overlapped the tail of the previous kernel

    if (warp_idx == 0) {
        if (elect_one()) {      // one elected lane: the compiler then keeps descriptors / addresses in uniform registers
            int stage = 0; uint32_t phase = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters) {
                const int m0 = tile_m0(tile) + static_cast<int>(cta) * 128;
                const int n0 = (tile % tiles_n) * BN + static_cast<int>(cta) * kHalf;
                const int ntile = tile + num_clusters;                       // L2 prefetch target (a_l2): this CTA's A rows of its next tile
                const int pf_m0 = ntile < num_tiles ? tile_m0(ntile) + static_cast<int>(cta) * 128 : -1;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * G2_STAGE_BYTES;
                    if (cta == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * kStageTx);   // bytes of BOTH CTAs
                    const int k0 = kb * GEMM_BK;
                    if (k0 < p.K1) tma_load_2d_2cta(sa, &tm.a, &full_bar[stage], k0, m0);
                    else           tma_load_2d_2cta(sa, &tm.a2, &full_bar[stage], k0 - p.K1, m0);
                    if (p.a_l2 > 0) {
                        // a_l2 k-blocks ahead of the load just issued (beyond the shared-memory ring), continuing into the next tile's rows
                        int pk = kb + p.a_l2, pm = m0, pt = tile;
                        if (pk >= num_kb) { pk -= num_kb; pm = pf_m0; pt = ntile; }
                        if (pm >= 0 && (pk % tiles_n) == (pt % tiles_n)) {
                            const int pk0 = pk * GEMM_BK;
                            if (pk0 < p.K1) tma_prefetch_l2_2d(&tm.a, pk0, pm);
                            else            tma_prefetch_l2_2d(&tm.a2, pk0 - p.K1, pm);
                        }
                    }
                    if constexpr (NP == 2) {
                        // 64 of my 128 W rows, delivered to me and to the CTA with my pair position in the other pair
                        tma_load_2d_2cta_mc(sa + G2_A_BYTES + pair_id * (G2_B_BYTES / 2), &tm.b, &full_bar[stage], k0, n0 + pair_id * 64,
                                            static_cast<uint16_t>(0x5u << cta));
                    } else
                    tma_load_2d_2cta(sa + G2_A_BYTES, &tm.b, &full_bar[stage], k0, n0);
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (cta == 0 && elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(256, BN);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int as = it & 1;
                mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 256;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * G2_STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + G2_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) umma_ss_2cta(d_tmem, da + 2 * k, db + 2 * k, idesc, (kb | k) != 0);
                    umma_commit_2cta(&empty_bar[stage], NP == 2 ? 0xF : 0x3);   // every CTA that writes into these stages
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit_2cta(&tfull_bar[as], static_cast<uint16_t>(0x3u << (2 * pair_id)));
            }
        }
    } else {
        // ------------------------------------------ epilogue ------------------------------------------
        constexpr bool kBf16Out = (EPI == EPI_BF16 || EPI == EPI_GELU_BF16 || EPI == EPI_FOLD_GELU);
        constexpr bool kStats = (EPI == EPI_RESID_F32_ST || EPI == EPI_RESID_LN_ST);
        constexpr bool kResidPlain = (EPI == EPI_RESID_F32 || EPI == EPI_RESID_F32_ST);
        constexpr bool kResidLn = (EPI == EPI_RESID_LN || EPI == EPI_RESID_LN_ST);
        constexpr bool kResid = kResidPlain || kResidLn;
        constexpr int kChunkCols = kBf16Out ? 64 : 32;            // one staging tile row = 128 B
        constexpr int kChunks = kHalf / kChunkCols;               // per warp per tile
        static_assert(G2_EPI_BUFS == 2, "the residual prefetch below assumes two staging buffers per warp");
        static_assert(!kStats || (kChunkCols == 32 && kChunks % 2 == 0), "row-statistics groups are pairs of 32-column chunks");
        const int ew = warp_idx - 2;
        const int quad = warp_idx & 3;
        const int chalf = ew >> 2;                                // which half of the BN accumulator columns
        uint8_t* stg = smem + kSmemEpi + ew * G2_EPI_BUFS * G2_EPI_TILE;
        uint8_t* stg2 = smem + kSmemEpi2 + ew * G2_EPI_BUFS * G2_EPI_TILE2;   // only used (and only inside the budget) by the *_ST epilogues
        uint64_t* rb = rbar + ew * G2_EPI_BUFS;
        const int r7 = lane & 7;
        uint32_t kc = 0;                                          // running chunk counter
        int buf = 0;                                              // kc % G2_EPI_BUFS
        uint32_t bpar = 0;                                        // (kc / G2_EPI_BUFS) & 1: parity of this buffer's residual barrier
        auto chunk_coords = [&](int tile, int c, int& col, int& row) {
            row = tile_m0(tile) + static_cast<int>(cta) * 128 + quad * 32;
            col = (tile % tiles_n) * BN + chalf * kHalf + c * kChunkCols;
        };
        auto prefetch_tile_resid = [&](int tile) {                // one lane: this warp's residual chunks of `tile` into the L2
            if (tile < num_tiles) {
#pragma unroll
                for (int c = 0; c < kChunks; ++c) {
                    int col, row;
                    chunk_coords(tile, c, col, row);
                    tma_prefetch_l2_2d(&tm.resid, col, row);
                }
            }
        };
        if constexpr (DIRECT) {
            constexpr bool kCopy = kStats || EPI == EPI_DUAL;          // bf16 copy of the result
            const int lr = lane >> 2;                                  // row inside a 16-row group (and lr + 8)
            const int lc = 2 * (lane & 3);                             // first of this thread's two columns inside an 8-column block
            // residual of (tile, row group hh, chunk c): this thread's 2 rows x 4 column blocks x 2 columns
            auto load_resid = [&](int tile, int hh, int c, float2 (&ra)[4], float2 (&rb)[4]) {
                const int row = tile_m0(tile) + static_cast<int>(cta) * 128 + quad * 32 + hh * 16 + lr;
                const int col = (tile % tiles_n) * BN + chalf * kHalf + c * 32 + lc;
                const float* pa = p.resid + static_cast<size_t>(row) * p.ld_out + col;
                const float* pb = pa + static_cast<size_t>(8) * p.ld_out;
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    ra[b] = *reinterpret_cast<const float2*>(pa + 8 * b);
                    rb[b] = *reinterpret_cast<const float2*>(pb + 8 * b);
                }
            };
            [[maybe_unused]] float2 nra[4], nrb[4];                    // residual of the NEXT (row group, chunk), requested one step ahead
            if constexpr (kResid) {
                if (cluster_id < num_tiles) load_resid(cluster_id, 0, 0, nra, nrb);
            }
            int it = 0;
            for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
                const int as = it & 1;
                const int row0 = tile_m0(tile) + static_cast<int>(cta) * 128 + quad * 32;
                const int colw = (tile % tiles_n) * BN + chalf * kHalf;      // first of this warp's kHalf columns
                [[maybe_unused]] float2 sta[2], stb[2];                       // RESID_LN: (mean, rstd) of this thread's 4 rows
                if constexpr (kResidLn) {
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        sta[hh] = p.ln_stats[row0 + hh * 16 + lr];
                        stb[hh] = p.ln_stats[row0 + hh * 16 + lr + 8];
                    }
                }
                mbar_wait(&tfull_bar[as], (it >> 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int hh = 0; hh < 2; ++hh) {
                    const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32 + hh * 16) << 16) + as * 256 + chalf * kHalf;
                    const int rowa = row0 + hh * 16 + lr;
                    float sa1 = 0.f, sa2 = 0.f, sb1 = 0.f, sb2 = 0.f;        // *_ST: this thread's part of the two rows' (sum, sum of squares) over the current 64-column group
                    [[maybe_unused]] float ga1 = 0.f, ga2 = 0.f, gb1 = 0.f, gb2 = 0.f;   // ... and of the finished first group (256-column tiles)
#pragma unroll 1
                    for (int c = 0; c < kChunks; ++c) {
                        uint32_t racc[16];
                        tmem_ld_16x256b_x4(t_acc + c * 32, racc);
                        const int col = colw + c * 32 + lc;
                        // constants of this thread's 8 columns, and the next step's residual, requested before the waits
                        float2 cb2[4], cg2[4], cl2[4];
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            cb2[b] = p.bias != nullptr ? __ldg(reinterpret_cast<const float2*>(p.bias + col + 8 * b)) : make_float2(0.f, 0.f);
                            if constexpr (kResidLn) {
                                cg2[b] = __ldg(reinterpret_cast<const float2*>(p.ln_g + col + 8 * b));
                                cl2[b] = __ldg(reinterpret_cast<const float2*>(p.ln_b + col + 8 * b));
                            }
                        }
                        float2 ra[4], rb[4];
                        if constexpr (kResid) {
#pragma unroll
                            for (int b = 0; b < 4; ++b) { ra[b] = nra[b]; rb[b] = nrb[b]; }
                            int nt = tile, nh = hh, nc = c + 1;
                            if (nc == kChunks) { nc = 0; if (++nh == 2) { nh = 0; nt = tile + num_clusters; } }
                            if (nt < num_tiles) load_resid(nt, nh, nc, nra, nrb);
                        }
                        tmem_ld_wait_dep16(racc);
                        float* oa = p.out_f32 + static_cast<size_t>(rowa) * p.ld_out + col;
                        float* ob = oa + static_cast<size_t>(8) * p.ld_out;
#pragma unroll
                        for (int b = 0; b < 4; ++b) {
                            float a0 = __uint_as_float(racc[4 * b + 0]) + cb2[b].x, a1 = __uint_as_float(racc[4 * b + 1]) + cb2[b].y;
                            float b0 = __uint_as_float(racc[4 * b + 2]) + cb2[b].x, b1 = __uint_as_float(racc[4 * b + 3]) + cb2[b].y;
                            if constexpr (kResidLn) {
                                a0 += fmaf((ra[b].x - sta[hh].x) * sta[hh].y, cg2[b].x, cl2[b].x); a1 += fmaf((ra[b].y - sta[hh].x) * sta[hh].y, cg2[b].y, cl2[b].y);
                                b0 += fmaf((rb[b].x - stb[hh].x) * stb[hh].y, cg2[b].x, cl2[b].x); b1 += fmaf((rb[b].y - stb[hh].x) * stb[hh].y, cg2[b].y, cl2[b].y);
                            } else if constexpr (kResidPlain) {
                                a0 += ra[b].x; a1 += ra[b].y; b0 += rb[b].x; b1 += rb[b].y;
                            }
                            *reinterpret_cast<float2*>(oa + 8 * b) = make_float2(a0, a1);
                            *reinterpret_cast<float2*>(ob + 8 * b) = make_float2(b0, b1);
                            if constexpr (kCopy) {
                                __nv_bfloat16* ca = p.out_bf16 + static_cast<size_t>(rowa) * p.ld_out + col + 8 * b;
                                *reinterpret_cast<uint32_t*>(ca) = pack_bf16x2(a0, a1);
                                *reinterpret_cast<uint32_t*>(ca + static_cast<size_t>(8) * p.ld_out) = pack_bf16x2(b0, b1);
                            }
                            if constexpr (kStats) {
                                sa1 += a0 + a1; sa2 = fmaf(a0, a0, fmaf(a1, a1, sa2));
                                sb1 += b0 + b1; sb2 = fmaf(b0, b0, fmaf(b1, b1, sb2));
                            }
                        }
                        if constexpr (kStats) {
                            if (c & 1) {                               // a 64-column group is complete: reduce over the quad (fixed order), one lane writes
#pragma unroll
                                for (int o = 1; o <= 2; o <<= 1) {
                                    sa1 += __shfl_xor_sync(0xffffffffu, sa1, o); sa2 += __shfl_xor_sync(0xffffffffu, sa2, o);
                                    sb1 += __shfl_xor_sync(0xffffffffu, sb1, o); sb2 += __shfl_xor_sync(0xffffffffu, sb2, o);
                                }
                                const int gcol = colw + (c - 1) * 32;  // first column of the group
                                if constexpr (BN == 256) {
                                    if (c == 1) { ga1 = sa1; ga2 = sa2; gb1 = sb1; gb2 = sb2; }
                                    else if ((lane & 3) == 0) {
                                        p.part[static_cast<size_t>(rowa) * 8 + (gcol >> 7)] = make_float2(ga1 + sa1, ga2 + sa2);
                                        p.part[static_cast<size_t>(rowa + 8) * 8 + (gcol >> 7)] = make_float2(gb1 + sb1, gb2 + sb2);
                                    }
                                } else if ((lane & 3) == 0) {
                                    p.part[static_cast<size_t>(rowa) * 16 + (gcol >> 6)] = make_float2(sa1, sa2);
                                    p.part[static_cast<size_t>(rowa + 8) * 16 + (gcol >> 6)] = make_float2(sb1, sb2);
                                }
                                sa1 = sa2 = sb1 = sb2 = 0.f;
                            }
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], static_cast<uint32_t>(2 * pair_id));
            }
        } else {
        if constexpr (kResid) {
            if (lane == 0 && cluster_id < num_tiles && !(kStats && (p.lab & 1))) {
                int col, row;
                chunk_coords(cluster_id, 0, col, row);
                mbar_arrive_expect_tx(&rb[0], G2_EPI_TILE);
                tma_load_2d(stg, &tm.resid, &rb[0], col, row);
                if (p.resid_l2 == 1) prefetch_tile_resid(cluster_id);
            }
        }
        [[maybe_unused]] float4 ppn[4];                            // FOLD: the next tile's row-statistics partials of this thread's row
        if constexpr (EPI == EPI_FOLD_GELU) {
            if (p.part_n == 8 && cluster_id < num_tiles) {
                const float4* pq = reinterpret_cast<const float4*>(p.part + static_cast<size_t>(tile_m0(cluster_id) + static_cast<int>(cta) * 128 + quad * 32 + lane) * 8);
#pragma unroll
                for (int q = 0; q < 4; ++q) ppn[q] = pq[q];
            }
        }
        int it = 0;
        for (int tile = cluster_id; tile < num_tiles; tile += num_clusters, ++it) {
            const int as = it & 1;
            if constexpr (kResid) {
                if (p.resid_l2 == 1 && lane == 0) prefetch_tile_resid(tile + num_clusters);
            }
            float fa = 1.f, fb = 0.f;                                 // FOLD: rstd and -rstd * mean of this thread's row
            if constexpr (EPI == EPI_FOLD_GELU) {
                // before the accumulator wait: the partials of the row come through an L2 that the operand ring keeps saturated,
                // and every N tile of the row block re-reads them (16 entries read behind the wait cost FFN-in 16 %, profiles/r01)
                const int frow = tile_m0(tile) + static_cast<int>(cta) * 128 + quad * 32 + lane;
                const float4* pp = reinterpret_cast<const float4*>(p.part + static_cast<size_t>(frow) * p.part_n);
                float t1 = 0.f, t2 = 0.f;
                if (p.part_n == 8) {
                    // the four entries were requested one tile ahead (ppn): behind a finished accumulator their latency was 5 % of this
                    // epilogue's stall samples (profiles/r02)
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = ppn[q];                   // (e_2q, e_2q+1)
                        t1 += a.x + a.z; t2 += a.y + a.w;
                    }
                    const int ntile = tile + num_clusters;
                    if (ntile < num_tiles) {
                        const float4* pq = reinterpret_cast<const float4*>(p.part + static_cast<size_t>(tile_m0(ntile) + static_cast<int>(cta) * 128 + quad * 32 + lane) * 8);
#pragma unroll
                        for (int q = 0; q < 4; ++q) ppn[q] = pq[q];
                    }
                } else {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 a = pp[2 * q], b = pp[2 * q + 1];   // (g_4q, g_4q+1), (g_4q+2, g_4q+3)
                        t1 += (a.x + a.z) + (b.x + b.z); t2 += (a.y + a.w) + (b.y + b.w);
                    }
                }
                const float mean = t1 * (1.0f / 1024.0f);
                const float rstd = rsqrtf(fmaxf(t2 * (1.0f / 1024.0f) - mean * mean, 0.f) + 1e-5f);
                fa = rstd; fb = -rstd * mean;
                if ((tile % tiles_n) == 0 && chalf == 0) {
                    p.stats_out[frow] = make_float2(mean, rstd);
                    if (p.fold_flag != nullptr && !(fabsf(fb) <= GEMM_FOLD_MEAN_LIMIT)) *p.fold_flag = 1;     // also catches NaN
                }
            }
            [[maybe_unused]] float2 st = make_float2(0.f, 1.f);      // RESID_LN: (mean, rstd) of this thread's residual row, once per tile
            if constexpr (kResidLn) st = p.ln_stats[tile_m0(tile) + static_cast<int>(cta) * 128 + quad * 32 + lane];
            mbar_wait(&tfull_bar[as], (it >> 1) & 1);
            tc_fence_after();
            const uint32_t t_acc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 256 + chalf * kHalf;
            float ps1 = 0.f, ps2 = 0.f;                               // *_ST: this thread's partial row statistics over the current 64-column group
            [[maybe_unused]] float pg1 = 0.f, pg2 = 0.f;              // ... and the finished first group (256-column tiles)
#pragma unroll 1
            for (int c = 0; c < kChunks; ++c, ++kc) {
                uint8_t* tile_s = stg + buf * G2_EPI_TILE;
                const int nbuf = (buf + 1 == G2_EPI_BUFS) ? 0 : buf + 1;
                int col, row;
                chunk_coords(tile, c, col, row);
                float v[kChunkCols];
                [[maybe_unused]] float4 cb4[8], cg4[8], cl4[8];     // residual epilogues: bias, LayerNorm gamma / beta of the chunk's 32 columns
                if constexpr (kResid) {
                    // The accumulator chunk is requested first: the staging-buffer bookkeeping below (one lane waiting for a TMA store to
                    // finish reading shared memory) then overlaps the TMEM load instead of preceding it.
                    uint32_t racc[32];
                    tmem_ld_32x32(t_acc + c * kChunkCols, racc);
                    // per-column constants of the chunk, requested before the waits below (ncu source view, profiles/r02: issued next to
                    // their uses, the gamma / beta / bias loads were 29 % of this epilogue's stall samples)
                    if (p.bias != nullptr) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) cb4[j] = __ldg(reinterpret_cast<const float4*>(p.bias + col) + j);
                    }
                    if constexpr (kResidLn) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            cg4[j] = __ldg(reinterpret_cast<const float4*>(p.ln_g + col) + j);
                            cl4[j] = __ldg(reinterpret_cast<const float4*>(p.ln_b + col) + j);
                        }
                    }
                    if (lane == 0) {
                        // The next chunk's residual tile is prefetched into staging buffer nbuf, which (two buffers) was last read by the
                        // TMA store of the PREVIOUS chunk - the most recent bulk group.  Every committed store must therefore have
                        // finished reading its source before the load is issued: wait_group.read 0 (a store reads its 4 + 2 KB within a
                        // few hundred cycles, behind the TMEM load issued above).  <1> here was a write-after-read on nbuf.
                        tma_store_wait_read<G2_EPI_BUFS - 2>();
                        int nt = tile, nc = c + 1;
                        if (nc == kChunks) { nc = 0; nt = tile + num_clusters; }
                        if (nt < num_tiles && !(kStats && (p.lab & 1))) {
                            int ncol, nrow;
                            chunk_coords(nt, nc, ncol, nrow);
                            mbar_arrive_expect_tx(&rb[nbuf], G2_EPI_TILE);
                            tma_load_2d(stg + nbuf * G2_EPI_TILE, &tm.resid, &rb[nbuf], ncol, nrow);
                        }
                        if (p.resid_l2 == 2) {
                            int ft = tile, fc = c + 2;
                            if (fc >= kChunks) { fc -= kChunks; ft = tile + num_clusters; }
                            if (ft < num_tiles) {
                                int fcol, frow;
                                chunk_coords(ft, fc, fcol, frow);
                                tma_prefetch_l2_2d(&tm.resid, fcol, frow);
                            }
                        }
                    }
                    __syncwarp();
                    tmem_ld_wait_dep(racc);
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(racc[j]);
                } else {
                    // (measured, profiles/r02: requesting both 32-column halves before the staging wait, as the residual branch does,
                    // cost the GELU epilogue of FFN-in 8 % and QKV 4 % -- these epilogues are issue bound, not latency bound)
                    if (lane == 0) tma_store_wait_read<G2_EPI_BUFS - 1>();   // the store issued G2_EPI_BUFS chunks ago (same buffer) is done reading
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < kChunkCols / 32; ++q) {
                        uint32_t r[32];
                        tmem_ld_32x32(t_acc + c * kChunkCols + q * 32, r);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[q * 32 + j] = __uint_as_float(r[j]);
                    }
                }
                // (measured, profiles/r02: requesting the per-column constants behind the TMEM load in the non-residual epilogues cost
                // FFN-in 1 %; the residual epilogues, whose stall samples were 29 % gamma / beta / bias loads, gain 7 % from it)
                if constexpr (EPI == EPI_FOLD_GELU) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
                    const float4* c4 = reinterpret_cast<const float4*>(p.fold_c + col);
                    const uint64_t fa2 = f2_dup(fa), fb2 = f2_dup(fb);
#pragma unroll
                    for (int j = 0; j < kChunkCols / 4; ++j) {
                        const float4 b = __ldg(b4 + j), cc = __ldg(c4 + j);
                        // fa * v + (fb * cc + b), two columns per FFMA2 (the scalar roundings)
                        f2_unpack(f2_fma(fa2, f2_pack(v[4 * j + 0], v[4 * j + 1]), f2_fma(fb2, f2_pack(cc.x, cc.y), f2_pack(b.x, b.y))), v[4 * j + 0], v[4 * j + 1]);
                        f2_unpack(f2_fma(fa2, f2_pack(v[4 * j + 2], v[4 * j + 3]), f2_fma(fb2, f2_pack(cc.z, cc.w), f2_pack(b.z, b.w))), v[4 * j + 2], v[4 * j + 3]);
                    }
                } else if (p.bias != nullptr) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.bias + col);
#pragma unroll
                    for (int j = 0; j < kChunkCols / 4; ++j) {
                        float4 b;
                        if constexpr (kResid) b = cb4[j]; else b = __ldg(b4 + j);
                        v[4 * j + 0] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                    }
                }
                if constexpr (EPI == EPI_GELU_BF16 || EPI == EPI_FOLD_GELU) {
#pragma unroll
                    for (int j = 0; j < kChunkCols; j += 2) gelu_erf_fast2(v[j], v[j + 1]);
                }
                uint8_t* rowp = tile_s + lane * 128;               // this thread's row of the staging tile
                if constexpr (kResid) {
                    if (!(kStats && (p.lab & 1)))
                    mbar_wait(&rb[buf], bpar);                     // residual tile landed (async proxy -> visible after wait)
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float4* q4 = reinterpret_cast<float4*>(rowp + ((j ^ r7) << 4));
                        const float4 rr = *q4;
                        if constexpr (kResidLn) {
                            const float4 g = cg4[j], b = cl4[j];
                            v[4 * j] += fmaf((rr.x - st.x) * st.y, g.x, b.x); v[4 * j + 1] += fmaf((rr.y - st.x) * st.y, g.y, b.y);
                            v[4 * j + 2] += fmaf((rr.z - st.x) * st.y, g.z, b.z); v[4 * j + 3] += fmaf((rr.w - st.x) * st.y, g.w, b.w);
                        } else {
                            v[4 * j] += rr.x; v[4 * j + 1] += rr.y; v[4 * j + 2] += rr.z; v[4 * j + 3] += rr.w;
                        }
                        *q4 = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    }
                    if constexpr (kStats) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) { ps1 += v[j]; ps2 = fmaf(v[j], v[j], ps2); }
                        if (c & 1) {                               // a 64-column group is complete (two 32-column chunks, summed in column order)
                            if constexpr (BN == 256) {
                                if (c == 1) { pg1 = ps1; pg2 = ps2; }  // first group of this warp's 128 columns
                                else p.part[static_cast<size_t>(row + lane) * 8 + (col >> 7)] = make_float2(pg1 + ps1, pg2 + ps2);
                            } else {
                                p.part[static_cast<size_t>(row + lane) * 16 + (col >> 6)] = make_float2(ps1, ps2);
                            }
                            ps1 = 0.f; ps2 = 0.f;
                        }
                        // bf16 copy through its own staging tile (64-B rows, 16-B chunk j of row r at chunk j ^ ((r >> 1) & 3)) + TMA store:
                        // per-row direct stores cost +30 % on this GEMM (profiles/r01)
                        uint8_t* row2 = stg2 + buf * G2_EPI_TILE2 + lane * 64;
                        const int x3 = (lane >> 1) & 3;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            *reinterpret_cast<uint4*>(row2 + ((j ^ x3) << 4)) =
                                make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                    }
                } else if constexpr (kBf16Out) {
                    if constexpr (EPI == EPI_BF16) {
                        if (p.rownorm != nullptr && col < p.rownorm_cols) {      // one 64-column chunk = one head of q or k
                            uint64_t ss2 = 0;                                      // two partial sums, packed (FFMA2)
#pragma unroll
                            for (int j = 0; j < kChunkCols; j += 2) {
                                uint64_t x2;
                                asm("mov.b64 %0, {%1, %2};" : "=l"(x2) : "f"(v[j]), "f"(v[j + 1]));
                                asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(ss2) : "l"(x2));
                            }
                            float sa, sb;
                            asm("mov.b64 {%0, %1}, %2;" : "=f"(sa), "=f"(sb) : "l"(ss2));
                            const float ss = sa + sb;
                            const uint32_t mx = __reduce_max_sync(0xffffffffu, __float_as_uint(ss));
                            if (lane == 0) p.rownorm[static_cast<size_t>(row >> 5) * p.rownorm_ld + (col >> 6)] = __uint_as_float(mx);
                        }
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<uint4*>(rowp + ((j ^ r7) << 4)) =
                            make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                       pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(rowp + ((j ^ r7) << 4)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    if constexpr (EPI == EPI_DUAL) {               // bf16 copy for the next GEMM (12 launches / pass, K = 2048): direct stores
                        // (staging it like *_ST measured 8 % slower here: the longer mainloop hides the per-row stores)
                        uint4* o4 = reinterpret_cast<uint4*>(p.out_bf16 + static_cast<size_t>(row + lane) * p.ld_out + col);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                               pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                    }
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    if (!(kStats && (p.lab & 2))) tma_store_2d(&tm.out, tile_s, col, row);
                    if constexpr (kStats) { if (!(p.lab & 4)) tma_store_2d(&tm.out2, stg2 + buf * G2_EPI_TILE2, col, row); }
                    tma_store_commit();
                }
                if (++buf == G2_EPI_BUFS) { buf = 0; bpar ^= 1; }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_cluster(&tempty_bar[as], static_cast<uint32_t>(2 * pair_id));   // the pair leader's MMA thread owns accumulator reuse
        }
        if (lane == 0) tma_store_wait_read<0>();
        }   // staged epilogue
    }

    tc_fence_before();
    cluster_sync_all();                                               // nobody exits while the peer may still signal it
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc_2cta<512>(tmem_base);
    }
}

template <int EPI, int BN = 256, int DIRECT = 0>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_tc_kernel(const __grid_constant__ Gemm2Maps tm, const GemmParams p) {
    gemm_pair_body<EPI, 1, BN, DIRECT>(tm, p);
}
// Cluster of 4 = two CTA pairs sharing the W tile by TMA multicast; tm.b must be the box-64 weight map; M % 512 == 0.
template <int EPI>
__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm4_tc_kernel(const __grid_constant__ Gemm2Maps tm, const GemmParams p) {
    gemm_pair_body<EPI, 2>(tm, p);
}

}  // namespace tvb
