// Per-pipe issue cost on sm_100a, measured with clock64(): N dependent-free instructions per thread, W warps per SM sub-partition.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_bench pipe_bench.cu && ./pipe_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
template <int OP>
__global__ void k(long long* out, float seed, int iters) {
    float a[8];
    uint64_t p[8];
    for (int i = 0; i < 8; ++i) { a[i] = seed + i * 0.001f + threadIdx.x * 1e-6f; asm("mov.b64 %0, {%1, %1};" : "=l"(p[i]) : "f"(a[i])); }
    uint64_t c2;
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(seed));
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
            if (OP == 2) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a[i]) : "f"(seed));
            if (OP == 4) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(r); }
            if (OP == 5) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(a[(i + 1) & 7]), "f"(seed));
            if (OP == 6) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(seed));
            // packed half-precision exponentials: two results per lane and instruction
            if (OP == 7) { uint32_t r = __float_as_uint(a[i]); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
            if (OP == 8) { uint32_t r = __float_as_uint(a[i]); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
            if (OP == 9) { uint32_t r; asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a[i]), "f"(a[(i + 1) & 7])); a[i] = __uint_as_float(r); }
            if (OP == 10) { uint32_t r = __float_as_uint(a[i]); asm volatile("fma.rn.f16x2 %0, %0, %1, %1;" : "+r"(r) : "r"(__float_as_uint(seed))); a[i] = __uint_as_float(r); }
            if (OP == 11) {      // MUFU + FFMA2 interleaved 1:1 (do the two pipes overlap?)
                asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
                asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p[i]) : "l"(c2));
            }
            if (OP == 12) {      // f16x2 MUFU + FFMA2 + pack, the mix of a packed-exponential softmax (per 2 elements)
                uint32_t r;
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(c2));
                float lo, hi; asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i]));
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
                asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(r));
                a[i] = __uint_as_float(r);
            }
            if (OP == 13) { uint32_t r = __float_as_uint(a[i]); asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(r)); a[i] = __uint_as_float(r); }
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[i])); s += a[i] + lo + hi; }
    if (threadIdx.x == 0) { out[blockIdx.x * 2] = t1 - t0; out[blockIdx.x * 2 + 1] = (long long)s; }
}
template <int OP>
void run(const char* name) {
    long long* d; cudaMalloc(&d, 1024);
    for (int warps = 4; warps <= 16; warps *= 2) {      // warps per CTA (1 CTA on 1 SM): 1, 2, 4 per sub-partition
        const int iters = 1000;
        k<OP><<<1, warps * 32>>>(d, 0.5f, iters);
        long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        const double per = double(h[0]) / (iters * 8);
        printf("%-28s warps/SMSP %d: %.2f cycles per instruction per warp  -> %.2f cycles per warp-instruction on the sub-partition\n", name, warps / 4, per, per / (warps / 4));
    }
    cudaFree(d);
}
int main() {
    run<0>("MUFU.EX2 (ex2.approx.ftz)");
    run<1>("FADD2 (add.rn.f32x2)");
    run<2>("FFMA2 (fma.rn.f32x2)");
    run<3>("FFMA");
    run<4>("F2FP.BF16 pack");
    run<5>("FMNMX3 (max.f32 3-input)");
    run<6>("FADD");
    run<7>("MUFU.EX2 f16x2");
    run<8>("MUFU.EX2 bf16x2");
    run<9>("F2FP.F16 pack");
    run<10>("HFMA2 (f16x2)");
    run<11>("MUFU.EX2 f32 + FFMA2 pair");
    run<12>("FADD2 + F2FP.F16 + EX2.f16x2");
    run<13>("MUFU.TANH f16x2");
    return 0;
}
