// Prints the thread <-> element mapping of tcgen05.ld.16x256b (diagnostics): TMEM is filled with lane * 1000 + column through 32x32b
// stores (lane = row), then read back with 16x256b.x2 by warp 0.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o tools/tmem_layout tools/tmem_layout.cu && ./tools/tmem_layout
#include <cuda_runtime.h>
#include <cstdio>
#include "../usdm_b200/csrc/common.cuh"
using namespace tvb;

__global__ void __launch_bounds__(128, 1) k(float* out) {
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) tmem_alloc<64>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = slot + ((static_cast<uint32_t>(warp * 32)) << 16);
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = __float_as_uint((warp * 32 + lane) * 1000.0f + c);
    tmem_st_32x16(tb, *reinterpret_cast<uint32_t(*)[16]>(&v[0]));
    tmem_st_32x16(tb + 16, *reinterpret_cast<uint32_t(*)[16]>(&v[16]));
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 1) {                 // lanes 32..63: read lanes [32,48) and [48,64), columns [0,16)
        for (int h = 0; h < 2; ++h) {
            uint32_t r[8];
            const uint32_t ta = slot + ((static_cast<uint32_t>(32 + 16 * h)) << 16);
            asm volatile("tcgen05.ld.sync.aligned.16x256b.x2.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(ta) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int i = 0; i < 8; ++i) out[(h * 32 + lane) * 8 + i] = __uint_as_float(r[i]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<64>(slot); }
}
int main() {
    float* d; cudaMalloc(&d, 2 * 32 * 8 * 4);
    k<<<1, 128>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(e));
    float h[512]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    for (int hh = 0; hh < 2; ++hh)
        for (int t = 0; t < 32; ++t) {
            printf("half %d thread %2d:", hh, t);
            for (int i = 0; i < 8; ++i) printf("  r%d=(lane %d, col %d)", i, (int)(h[(hh * 32 + t) * 8 + i] / 1000), (int)h[(hh * 32 + t) * 8 + i] % 1000);
            printf("\n");
        }
    return 0;
}
