// Positional convolution (reference networks.py:67-95, called at :343-346) as an implicit GEMM on tcgen05:
// grouped Conv1d(H, H, k=taps, pad=taps/2, groups=H/64), weight-norm folded, followed by erf-GELU and the row mask.
//
// For one (128-row tile, group g) the accumulator is   sum_tap  X[rows + tap - taps/2, g*64:(g+1)*64] . W[g][tap]^T
// i.e. `taps` MMAs groups of M=128, N=64, K=64 whose A tiles are the same activation matrix shifted by one row per
// tap.  The shift is done by the TMA coordinate (rows outside [0, Mp) are zero-filled by TMA; rows between sequences
// are the zero gap rows of the packed layout), so no im2col buffer exists.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue; 6-stage smem ring; 2 TMEM accumulators.
#pragma once
#include "common.cuh"

namespace tvb {

constexpr int CV_BM = 128;
constexpr int CV_A_BYTES = 128 * 64 * 2;       // 16 KB
constexpr int CV_B_BYTES = 64 * 64 * 2;        //  8 KB
constexpr int CV_STAGE_BYTES = CV_A_BYTES + CV_B_BYTES;
constexpr int CV_STAGES = 6;
constexpr int CV_SMEM_BYTES = CV_STAGES * CV_STAGE_BYTES + 1024 + 256;
constexpr int CV_THREADS = 192;

// MODE 0: out_bf16 = mask * gelu(conv + b)            MODE 1: out_f32 = mask * gelu(conv + b) + resid
template <int MODE>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
               const int* __restrict__ row_seq, int Mp, int H, int taps, int groups, __nv_bfloat16* __restrict__ out_bf16,
               float* __restrict__ out_f32, const float* __restrict__ resid) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CV_STAGES * CV_STAGE_BYTES);
    uint64_t* empty_bar = full_bar + CV_STAGES;
    uint64_t* tfull_bar = empty_bar + CV_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (Mp / CV_BM) * groups;
    const int halo = taps / 2;

    if (warp_idx == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < CV_STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<128>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int g = tile % groups, m0 = (tile / groups) * CV_BM;
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * CV_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full_bar[stage], CV_STAGE_BYTES);
                    tma_load_2d(sa, &tmX, &full_bar[stage], g * 64, m0 + tap - halo);
                    tma_load_2d(sa + CV_A_BYTES, &tmW, &full_bar[stage], 0, (g * taps + tap) * 64);
                    if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 64);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * 64;
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * CV_STAGE_BYTES);
                    const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + CV_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == CV_STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        const int quad = warp_idx & 3;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const int g = tile % groups, m0 = (tile / groups) * CV_BM;
            mbar_wait(&tfull_bar[as], (it >> 1) & 1);
            tc_fence_after();
            const int row = m0 + quad * 32 + lane;
            const bool valid = row_seq[row] >= 0;
            const size_t off = static_cast<size_t>(row) * H + g * 64;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * 64 + c * 32, r);
                tmem_ld_wait();
                float v[32];
                const float4* b4 = reinterpret_cast<const float4*>(bias + g * 64 + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = __ldg(b4 + j);
                    v[4 * j + 0] = valid ? gelu_erf(__uint_as_float(r[4 * j + 0]) + b.x) : 0.f;
                    v[4 * j + 1] = valid ? gelu_erf(__uint_as_float(r[4 * j + 1]) + b.y) : 0.f;
                    v[4 * j + 2] = valid ? gelu_erf(__uint_as_float(r[4 * j + 2]) + b.z) : 0.f;
                    v[4 * j + 3] = valid ? gelu_erf(__uint_as_float(r[4 * j + 3]) + b.w) : 0.f;
                }
                if constexpr (MODE == 0) {
                    uint4* o4 = reinterpret_cast<uint4*>(out_bf16 + off + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                } else {
                    const float4* r4 = reinterpret_cast<const float4*>(resid + off + c * 32);
                    float4* o4 = reinterpret_cast<float4*>(out_f32 + off + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q = r4[j];
                        o4[j] = make_float4(v[4 * j] + q.x, v[4 * j + 1] + q.y, v[4 * j + 2] + q.z, v[4 * j + 3] + q.w);
                    }
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
        tmem_dealloc<128>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Halo variant (default).  conv_tc_kernel above re-fetches the 16 KB activation tile for every one of the `taps` MMA groups: 24 KB
// of TMA traffic per 64 tensor-core cycles, several times what L2 -> SM can deliver, so it runs at a third of the tensor rate.
// Here one CTA owns 256 rows of one group: the two 128-row halves are fetched ONCE each with their halo (rows m0 - taps/2 ...,
// 160-row boxes), every tap's MMA reads them through a shared-memory descriptor advanced by `tap` rows (the 128-byte swizzle is a
// function of the absolute smem address, so a row-shifted start sees the tile as TMA wrote it), and the two halves share each
// 8 KB weight tile: 288 KB instead of 1.5 MB of fetches per 256 x 64 outputs.
//   warp 0 TMA producer, warp 1 MMA issuer, warps 2..9 epilogue (4 per half); 2 activation slots, 8-deep weight ring,
//   2 x 2 TMEM accumulators (tile parity x half).
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int CH_ROWS = 160;                       // 128 + 2 * 15 halo rows, padded to a multiple of 8
constexpr int CH_A_BYTES = CH_ROWS * 128;          // 20 KB per half
constexpr int CH_A_SLOT = 2 * CH_A_BYTES;          // both halves of one 256-row tile
constexpr int CH_W_STAGES = 8;
constexpr int CH_SMEM_W = 2 * CH_A_SLOT;           // two slots
constexpr int CH_SMEM_BAR = CH_SMEM_W + CH_W_STAGES * CV_B_BYTES;
constexpr int CH_SMEM_BYTES = CH_SMEM_BAR + 1024 + 256;
constexpr int CH_THREADS = 320;

template <int MODE>
__global__ void __launch_bounds__(CH_THREADS, 1)
conv_halo_kernel(const __grid_constant__ CUtensorMap tmX /* box 64 x 160 */, const __grid_constant__ CUtensorMap tmW, const float* __restrict__ bias,
                 const int* __restrict__ row_seq, int Mp, int H, int taps, int groups, __nv_bfloat16* __restrict__ out_bf16,
                 float* __restrict__ out_f32, const float* __restrict__ resid) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + CH_SMEM_BAR);
    uint64_t* a_empty = a_full + 2;
    uint64_t* w_full = a_empty + 2;
    uint64_t* w_empty = w_full + CH_W_STAGES;
    uint64_t* tfull_bar = w_empty + CH_W_STAGES;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_tiles = (Mp / 256) * groups;
    const int halo = taps / 2;

    if (warp_idx == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < 2; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < CH_W_STAGES; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 8); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<256>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        if (elect_one()) {
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int g = tile % groups, m0 = (tile / groups) * 256;
                mbar_wait(&a_empty[as], aph ^ 1);
                mbar_arrive_expect_tx(&a_full[as], CH_A_SLOT);
                tma_load_2d(smem + as * CH_A_SLOT, &tmX, &a_full[as], g * 64, m0 - halo);
                tma_load_2d(smem + as * CH_A_SLOT + CH_A_BYTES, &tmX, &a_full[as], g * 64, m0 + 128 - halo);
                if (++as == 2) { as = 0; aph ^= 1; }
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&w_empty[ws], wph ^ 1);
                    mbar_arrive_expect_tx(&w_full[ws], CV_B_BYTES);
                    tma_load_2d(smem + CH_SMEM_W + ws * CV_B_BYTES, &tmW, &w_full[ws], 0, (g * taps + tap) * 64);
                    if (++ws == CH_W_STAGES) { ws = 0; wph ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 64);
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
                mbar_wait(&a_full[as], aph);
                tc_fence_after();
                const uint32_t sa0 = smem_u32(smem + as * CH_A_SLOT);
                const uint32_t d_tmem = tmem_base + acc * 128;
                for (int tap = 0; tap < taps; ++tap) {
                    mbar_wait(&w_full[ws], wph);
                    tc_fence_after();
                    const uint64_t db = make_smem_desc_sw128(smem_u32(smem + CH_SMEM_W + ws * CV_B_BYTES), 16, 1024);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint64_t da = make_smem_desc_sw128(sa0 + half * CH_A_BYTES + static_cast<uint32_t>(tap) * 128u, 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(d_tmem + half * 64, da + 2 * k, db + 2 * k, idesc, (tap | k) != 0);
                    }
                    umma_commit(&w_empty[ws]);
                    if (++ws == CH_W_STAGES) { ws = 0; wph ^= 1; }
                }
                umma_commit(&a_empty[as]);
                if (++as == 2) { as = 0; aph ^= 1; }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        const int quad = warp_idx & 3, half = (warp_idx - 2) >> 2;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int g = tile % groups, m0 = (tile / groups) * 256 + half * 128;
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            const int row = m0 + quad * 32 + lane;
            const bool valid = row_seq[row] >= 0;
            const size_t off = static_cast<size_t>(row) * H + g * 64;
#pragma unroll 1
            for (int c = 0; c < 2; ++c) {
                uint32_t r[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + acc * 128 + half * 64 + c * 32, r);
                tmem_ld_wait();
                float v[32];
                const float4* b4 = reinterpret_cast<const float4*>(bias + g * 64 + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b = __ldg(b4 + j);
                    v[4 * j + 0] = valid ? gelu_erf(__uint_as_float(r[4 * j + 0]) + b.x) : 0.f;
                    v[4 * j + 1] = valid ? gelu_erf(__uint_as_float(r[4 * j + 1]) + b.y) : 0.f;
                    v[4 * j + 2] = valid ? gelu_erf(__uint_as_float(r[4 * j + 2]) + b.z) : 0.f;
                    v[4 * j + 3] = valid ? gelu_erf(__uint_as_float(r[4 * j + 3]) + b.w) : 0.f;
                }
                if constexpr (MODE == 0) {
                    uint4* o4 = reinterpret_cast<uint4*>(out_bf16 + off + c * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        o4[j] = make_uint4(pack_bf16x2(v[8 * j], v[8 * j + 1]), pack_bf16x2(v[8 * j + 2], v[8 * j + 3]),
                                           pack_bf16x2(v[8 * j + 4], v[8 * j + 5]), pack_bf16x2(v[8 * j + 6], v[8 * j + 7]));
                } else {
                    const float4* r4 = reinterpret_cast<const float4*>(resid + off + c * 32);
                    float4* o4 = reinterpret_cast<float4*>(out_f32 + off + c * 32);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 q = r4[j];
                        o4[j] = make_float4(v[4 * j] + q.x, v[4 * j + 1] + q.y, v[4 * j + 2] + q.z, v[4 * j + 3] + q.w);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<256>(tmem_base);
    }
}

inline cudaError_t launch_conv_halo_raw(const CUtensorMap& tmX160, const CUtensorMap& tmW, const float* bias, const int* row_seq, int Mp, int H,
                                        int taps, int groups, __nv_bfloat16* out_bf16, float* out_f32, const float* resid, int num_sms,
                                        cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(conv_halo_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_halo_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CH_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int tiles = (Mp / 256) * groups;
    const int grid = tiles < num_sms ? tiles : num_sms;
    if (out_f32)
        conv_halo_kernel<1><<<grid, CH_THREADS, CH_SMEM_BYTES, st>>>(tmX160, tmW, bias, row_seq, Mp, H, taps, groups, nullptr, out_f32, resid);
    else
        conv_halo_kernel<0><<<grid, CH_THREADS, CH_SMEM_BYTES, st>>>(tmX160, tmW, bias, row_seq, Mp, H, taps, groups, out_bf16, nullptr, nullptr);
    return cudaGetLastError();
}

inline cudaError_t launch_conv_tc_raw(const CUtensorMap& tmX, const CUtensorMap& tmW, const float* bias, const int* row_seq, int Mp, int H,
                                      int taps, int groups, __nv_bfloat16* out_bf16, float* out_f32, const float* resid, int num_sms,
                                      cudaStream_t st) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(conv_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CV_SMEM_BYTES);
        if (e != cudaSuccess) return e;
        attr.mark();
    }
    const int tiles = (Mp / CV_BM) * groups;
    const int grid = tiles < num_sms ? tiles : num_sms;
    if (out_f32)
        conv_tc_kernel<1><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(tmX, tmW, bias, row_seq, Mp, H, taps, groups, nullptr, out_f32, resid);
    else
        conv_tc_kernel<0><<<grid, CV_THREADS, CV_SMEM_BYTES, st>>>(tmX, tmW, bias, row_seq, Mp, H, taps, groups, out_bf16, nullptr, nullptr);
    return cudaGetLastError();
}

}  // namespace tvb
