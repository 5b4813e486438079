// tcgen05.mma issue-rate microbenchmark (diagnostics, not part of libtvb.so): one CTA per SM issues `n` MMAs of one shape back to back
// from a single thread, commits, and reports cycles per MMA.  Shapes are the ones the attention kernels use.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/mma_bench tools/mma_bench.cu && ./tools/mma_bench
#include <cuda_runtime.h>
#include <cstdio>
#include "../usdm_b200/csrc/attn10_tc.cuh"
using namespace tvb;

// mode 0: SS, A 128 x 16 (K-major sw128), B N x 16 (K-major sw128), 4 K steps per "block" walking the 64-wide tile
// mode 1: TS, A = TMEM (128 x 16 bf16 = 8 columns), B = 16 x 64 MN-major (V), N = 64
// mode 2: SS with the sw32 position operand as B (K = 16 single step)
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) k_mma(int n, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 3 * 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot;
    long long t0 = 0, t1 = 0;
    if (threadIdx.x == 0) {
        const uint64_t da = make_smem_desc_sw128(smem_u32(smem), 16, 1024);
        const uint64_t db = make_smem_desc_sw128(smem_u32(smem + 16384), 16, 1024);
        const uint64_t dv = make_smem_desc_sw128(smem_u32(smem + 16384), 8192, 1024);
        const uint64_t dp = make_smem_desc_sw32(smem_u32(smem + 32768));
        constexpr uint32_t idesc = MODE == 1 ? make_idesc_bf16(128, N, false, true) : make_idesc_bf16(128, N);
        t0 = clock64();
        for (int i = 0; i < n; ++i) {
            const int k = i & 3;
            if (MODE == 0) umma_ss(tb + (i & 1) * 128, da + 2 * k, db + 2 * k, idesc, 1);
            else if (MODE == 3) umma_ss(tb, da + 2 * k, db + 2 * k, idesc, 1);
            else if (MODE == 4) umma_ts(tb + 384 + (i & 1) * 64, tb + 256 + k * 8, dv + k * (2048 >> 4), make_idesc_bf16(128, 64, false, true), 1);
            else if (MODE == 5) {
                if (i & 1) umma_ts(tb + 384, tb + 256 + k * 8, dv + k * (2048 >> 4), make_idesc_bf16(128, 64, false, true), 1);
                else umma_ss(tb, da + 2 * k, db + 2 * k, make_idesc_bf16(128, 128), 1);
            }
            else if (MODE == 1) umma_ts(tb + 384, tb + 256 + k * 8, dv + k * (2048 >> 4), idesc, 1);
            else umma_ss(tb + (i & 1) * 128, da + 2 * k, dp, idesc, 1);
        }
        umma_commit(&bar);
        mbar_wait(&bar, 0);
        t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

// tcgen05.ld throughput: every warp reads its lane quadrant, `n` loads of 32 (or 16) columns, DEPTH loads in flight between waits
template <int COLS, int DEPTH>
__global__ void __launch_bounds__(512, 1) k_ldtm(int n, long long* out, unsigned* sink) {
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot + ((static_cast<uint32_t>((threadIdx.x >> 5) & 3) * 32) << 16);
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < n; i += DEPTH) {
        if constexpr (COLS == 32) {
            uint32_t r[DEPTH][32];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) tmem_ld_32x32(tb + ((i + d) & 7) * 32, r[d]);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) acc ^= r[d][0] ^ r[d][31];
        } else {
            uint32_t r[DEPTH][16];
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) tmem_ld_32x16(tb + ((i + d) & 15) * 16, r[d]);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < DEPTH; ++d) acc ^= r[d][0] ^ r[d][15];
        }
    }
    const long long t1 = clock64();
    __syncthreads();
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc<512>(slot); }
}
template <int COLS, int DEPTH>
static void run_ld(long long* d, unsigned* sink) {
    const int n = 4096;
    for (int warps : {1, 4, 8, 16}) {
        k_ldtm<COLS, DEPTH><<<148, warps * 32>>>(n, d, sink);
        k_ldtm<COLS, DEPTH><<<148, warps * 32>>>(n, d, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        const double bytes = (double)n * warps * 32 * COLS * 4;
        printf("tcgen05.ld 32x32b.x%-2d depth %d  warps %2d: %6.1f cycles per load per warp, %6.1f B/cycle/SM (%s)\n", COLS, DEPTH, warps, (double)c / n, bytes / c,
               cudaGetErrorString(e));
    }
}

template <int MODE, int N>
static void run(const char* name, long long* d) {
    const int n = 4000;
    cudaFuncSetAttribute(k_mma<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 16384 + 1024);
    for (int grid : {1, 148}) {
        k_mma<MODE, N><<<grid, 128, 3 * 16384 + 1024>>>(n, d);
        k_mma<MODE, N><<<grid, 128, 3 * 16384 + 1024>>>(n, d);
        cudaError_t e = cudaDeviceSynchronize();
        long long c = 0;
        cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        printf("%-44s N=%3d grid=%3d: %7.1f cycles per MMA (%s)\n", name, N, grid, (double)c / n, cudaGetErrorString(e));
    }
}
int main() {
    long long* d;
    cudaMalloc(&d, 8);
    run<3, 128>("SS, ONE accumulator (dependent chain)", d);
    run<3, 64>("SS, ONE accumulator (dependent chain)", d);
    run<4, 64>("TS, two accumulators alternating", d);
    run<5, 128>("SS N=128 / TS N=64 interleaved (per MMA avg)", d);
    run<0, 256>("SS K-major sw128 A, B (QK^T)", d);
    run<0, 128>("SS K-major sw128 A, B (QK^T)", d);
    run<0, 64>("SS K-major sw128 A, B (QK^T)", d);
    run<0, 32>("SS K-major sw128 A, B (QK^T)", d);
    run<1, 64>("TS A in TMEM, B MN-major sw128 (PV)", d);
    run<2, 128>("SS sw128 A, sw32 B (bias step)", d);
    run<2, 64>("SS sw128 A, sw32 B (bias step)", d);
    unsigned* sink;
    cudaMalloc(&sink, 4);
    run_ld<32, 1>(d, sink);
    run_ld<32, 2>(d, sink);
    run_ld<16, 2>(d, sink);
    run_ld<16, 4>(d, sink);
    return 0;
}
