// Attention kernel laboratory (diagnostics, not part of libtvb.so): times attn9_tc.cuh / attn8_tc.cuh on synthetic q|k|v rows of the
// configs[1] shape with CUDA events, including the LAB timing ablations of attn9 (pieces of the softmax body switched off).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o tools/attn_lab tools/attn_lab.cu && ./tools/attn_lab [scale] [len] [nseq]
#include <cuda.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define A9_LAB_TRACE
#define A10_LAB_TRACE
#define A11_LAB_TRACE
#include "../usdm_b200/csrc/attn11_tc.cuh"

using namespace tvb;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

typedef CUresult (*PFN_encode)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__global__ void k_init(__nv_bfloat16* qkv, size_t n, int H, float scale, unsigned seed) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned x = (unsigned)(i * 2654435761u) ^ seed;
        x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
        unsigned y = x * 1664525u + 1013904223u;
        const float u1 = ((x >> 8) + 1) * (1.0f / 16777217.0f), u2 = (y >> 8) * (1.0f / 16777216.0f);
        float v = sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2) * scale;
        if ((int)(i % (3 * H)) < H) v *= 0.125f * 1.4426950408889634f;
        qkv[i] = __float2bfloat16_rn(v);
    }
}

template <typename F>
static float time_ms(F&& fn, int iters) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 3; ++i) fn();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < iters; ++i) fn();
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return ms / iters;
}

int main(int argc, char** argv) {
    setvbuf(stdout, nullptr, _IOLBF, 0);
    const float scale = argc > 1 ? atof(argv[1]) : 0.45f;
    const int len = argc > 2 ? atoi(argv[2]) : 1134;
    const int nseq = argc > 3 ? atoi(argv[3]) : 32;
    const int H = 1024, nh = 16;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    std::vector<int4> work;
    int o = 0;
    for (int s = 0; s < nseq; ++s) {
        for (int q0 = 0; q0 < len; q0 += ATT_BQ) work.push_back(make_int4(s, q0, o, len));
        o += len + 16;
    }
    const int rows = (o + 255) / 256 * 256;
    std::stable_sort(work.begin(), work.end(), [](const int4& a, const int4& b) {
        auto cost = [](const int4& w) { return ((w.w + ATT_BK - 1) / ATT_BK) * (w.y + 128 < w.w ? 2 : 1); };
        return cost(a) > cost(b);
    });
    __nv_bfloat16 *qkv, *ctx;
    CK(cudaMalloc(&qkv, (size_t)rows * 3 * H * 2));
    CK(cudaMalloc(&ctx, (size_t)rows * H * 2));
    k_init<<<148 * 8, 256>>>(qkv, (size_t)rows * 3 * H, H, scale, 12345u);
    CK(cudaGetLastError());
    int4 *d_work, *d_win;
    int* d_order;
    CK(cudaMalloc(&d_order, work.size() * 16 * sizeof(int)));
    CK(cudaMalloc(&d_work, work.size() * sizeof(int4)));
    CK(cudaMalloc(&d_win, work.size() * nh * sizeof(int4)));
    CK(cudaMemcpy(d_work, work.data(), work.size() * sizeof(int4), cudaMemcpyHostToDevice));
    float *normtab, *slopes;
    CK(cudaMalloc(&normtab, (size_t)(rows / 32 + 1) * 2 * nh * sizeof(float)));
    std::vector<float> sl(nh);
    for (int i = 0; i < nh; ++i) sl[i] = powf(2.0f, -0.5f * (i + 1));
    CK(cudaMalloc(&slopes, nh * sizeof(float)));
    CK(cudaMemcpy(slopes, sl.data(), nh * sizeof(float), cudaMemcpyHostToDevice));
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    CUtensorMap tm;
    {
        cuuint64_t dims[2] = {(cuuint64_t)(3 * H), (cuuint64_t)rows};
        cuuint64_t strides[1] = {(cuuint64_t)(3 * H) * 2};
        cuuint32_t box[2] = {64, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((PFN_encode)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("tensor map failed %d\n", (int)r); return 1; }
    }
    // v10: key-position table + its 32-byte-swizzle map
    __nv_bfloat16* pos;
    CUtensorMap tmp;
    {
        std::vector<int> hs(rows, -1), hp(rows, 0);
        int oo = 0;
        for (int s = 0; s < nseq; ++s) { for (int i = 0; i < len; ++i) { hs[oo + i] = s; hp[oo + i] = i; } oo += len + 16; }
        int *rseq, *rpos;
        CK(cudaMalloc(&rseq, rows * sizeof(int))); CK(cudaMalloc(&rpos, rows * sizeof(int)));
        CK(cudaMemcpy(rseq, hs.data(), rows * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(rpos, hp.data(), rows * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&pos, (size_t)rows * 16 * 2));
        k_attn10_pos<<<(rows + 255) / 256, 256>>>(rseq, rpos, rows, pos);
        CK(cudaDeviceSynchronize());
        cuuint64_t dims[2] = {16, (cuuint64_t)rows};
        cuuint64_t strides[1] = {32};
        cuuint32_t box[2] = {16, 128};
        cuuint32_t estr[2] = {1, 1};
        CUresult r = ((PFN_encode)fn)(&tmp, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, pos, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                      CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("pos tensor map failed %d\n", (int)r); return 1; }
    }
    const int nwork = (int)work.size(), n_items = nwork * nh;
    k_attn9_rownorm<<<(rows + 31) / 32, 256>>>(qkv, rows, H, nh, normtab);
    CK(cudaDeviceSynchronize());
    const double flops = 4.0 * 64 * nh * (double)nseq * len * len;
    printf("# attn_lab: %d sequences x %d tokens, %d heads, scale %.2f, %d items, %d SMs\n", nseq, len, nh, scale, n_items, sms);
    const int* lab_order = nullptr;
    for (int use_win = 0; use_win < 2; ++use_win) {
        lab_order = nullptr;
        k_attn9_plan<<<(n_items + 7) / 8, 256>>>(d_work, n_items, nh, slopes, normtab, use_win, d_win, nullptr);
        CK(cudaDeviceSynchronize());
        std::vector<int4> win(n_items);
        CK(cudaMemcpy(win.data(), d_win, n_items * sizeof(int4), cudaMemcpyDeviceToHost));
        long blocks = 0, full = 0;
        float shift_max = 0;
        for (int t = 0; t < n_items; ++t) {
            const int4 w = work[t / nh];
            const int nkb = (w.w + ATT_BK - 1) / ATT_BK, nhv = (w.y + 128 < w.w) ? 2 : 1;
            const int lo[2] = {win[t].x & 0xffff, win[t].y & 0xffff}, hi[2] = {win[t].x >> 16, win[t].y >> 16};
            for (int h = 0; h < nhv; ++h) { blocks += 1 + (hi[h] >= lo[h] ? hi[h] - lo[h] + 1 : 0); full += nkb; }
            float sh; memcpy(&sh, &win[t].z, 4); shift_max = std::max(shift_max, sh);
        }
        printf("window=%d: %ld of %ld block-halves kept (%.1f %%), max shift %.1f\n", use_win, blocks, full, 100.0 * blocks / full, shift_max);
#define RUN9(POLY, LAB, name)                                                                                                     \
        {                                                                                                                         \
            const float ms = time_ms([&] { CK((launch_attn9_t<POLY, LAB>(tm, ctx, H, nh, slopes, d_work, d_win, nwork, sms, 0, false))); }, 20); \
            printf("  v9 win=%d poly=%d lab=%2d %-34s %8.1f us  %7.1f TFLOP/s (algorithmic)  %6.0f cycles/block-half @1.9GHz\n", use_win, POLY, LAB, name, ms * 1e3, \
                   flops / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.9e9 * sms / blocks);                                                 \
        }
#define RUN10(POLY, LAB, name)                                                                                                    \
        {                                                                                                                         \
            const float ms = time_ms([&] { CK((launch_attn10_t<POLY, LAB>(tm, tmp, ctx, H, nh, slopes, d_work, d_win, lab_order, nwork, sms, 0, false))); }, 20); \
            printf("  v10 win=%d poly=%d lab=%2d %-33s %8.1f us  %7.1f TFLOP/s (algorithmic)  %6.0f cycles/block-half @1.9GHz\n", use_win, POLY, LAB, name, ms * 1e3, \
                   flops / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.9e9 * sms / blocks);                                                 \
        }
#define RUN11(POLY, LAB, name)                                                                                                    \
        {                                                                                                                         \
            const float ms = time_ms([&] { CK((launch_attn11_t<POLY, LAB>(tm, tmp, ctx, H, nh, slopes, d_work, d_win, lab_order, nwork, sms, 0, false))); }, 20); \
            printf("  v11 win=%d poly=%d lab=%2d %-33s %8.1f us  %7.1f TFLOP/s (algorithmic)  %6.0f cycles/block-half @1.9GHz\n", use_win, POLY, LAB, name, ms * 1e3, \
                   flops / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.9e9 * sms / blocks);                                                 \
        }
#define RUN11U(POLY, LAB, name)                                                                                                   \
        {                                                                                                                         \
            const float ms = time_ms([&] { CK((launch_attn11_t<POLY, LAB, 0>(tm, tmp, ctx, H, nh, slopes, d_work, d_win, lab_order, nwork, sms, 0, false))); }, 20); \
            printf("  v11u win=%d poly=%d lab=%2d %-32s %8.1f us  %7.1f TFLOP/s (algorithmic)  %6.0f cycles/block-half @1.9GHz\n", use_win, POLY, LAB, name, ms * 1e3, \
                   flops / (ms * 1e-3) / 1e12, ms * 1e-3 * 1.9e9 * sms / blocks);                                                 \
        }
        RUN11U(3, 0, "unsplit QK: baseline");
        RUN11U(4, 0, "unsplit QK: poly 1/4");
        RUN11U(2, 0, "unsplit QK: poly 1/2");
        if (use_win == 0) { RUN11U(3, 1, "unsplit QK: no exponential"); RUN11U(3, 11, "unsplit QK: no exp / sum / pack"); }
        RUN11(3, 0, "baseline");
        RUN11(0, 0, "all exp2 on MUFU");
        RUN11(2, 0, "poly 1/2");
        RUN11(4, 0, "poly 1/4");
        if (use_win == 0) {
            RUN11(3, 1, "no exponential");
            RUN11(3, 2, "no row sum");
            RUN11(3, 8, "no bf16 pack");
            RUN11(3, 11, "no exp / sum / pack");
            // v11 against v10, bit for bit (same plan, same order of additions)
            __nv_bfloat16* ctx2;
            CK(cudaMalloc(&ctx2, (size_t)rows * H * 2));
            CK(cudaMemset(ctx, 0, (size_t)rows * H * 2)); CK(cudaMemset(ctx2, 0, (size_t)rows * H * 2));
            CK((launch_attn10_t<3, 0>(tm, tmp, ctx, H, nh, slopes, d_work, d_win, lab_order, nwork, sms, 0, false)));
            CK((launch_attn11_t<3, 0>(tm, tmp, ctx2, H, nh, slopes, d_work, d_win, lab_order, nwork, sms, 0, false)));
            CK(cudaDeviceSynchronize());
            std::vector<uint16_t> h1((size_t)rows * H), h2((size_t)rows * H);
            CK(cudaMemcpy(h1.data(), ctx, h1.size() * 2, cudaMemcpyDeviceToHost));
            CK(cudaMemcpy(h2.data(), ctx2, h2.size() * 2, cudaMemcpyDeviceToHost));
            size_t diff = 0, nz = 0;
            for (size_t i = 0; i < h1.size(); ++i) { diff += h1[i] != h2[i]; nz += h1[i] != 0; }
            printf("  v11 vs v10: %zu of %zu outputs differ (%zu non-zero)\n", diff, h1.size(), nz);
            cudaFree(ctx2);
        }
        RUN10(3, 0, "baseline");
        RUN10(0, 0, "all exp2 on MUFU");
        RUN10(2, 0, "poly 1/2");
        RUN10(4, 0, "poly 1/4");
        if (use_win == 0) {
            RUN10(3, 1, "no exponential");
            RUN10(3, 2, "no row sum");
            RUN10(3, 8, "no bf16 pack");
            RUN10(3, 11, "no exp / sum / pack");
        }
        if (use_win) {
            k_attn10_order<<<1, 1024>>>(d_win, n_items, 1, d_order);
            CK(cudaDeviceSynchronize());
            lab_order = d_order;
            RUN10(3, 0, "baseline, cost-sorted snake order");
            RUN10(4, 0, "poly 1/4, cost-sorted snake order");
            const float ms = time_ms([&] { k_attn9_plan<<<(n_items + 7) / 8, 256>>>(d_work, n_items, nh, slopes, normtab, 1, d_win, nullptr); k_attn10_order<<<1, 1024>>>(d_win, n_items, 1, d_order); }, 50);
            printf("  plan + order kernels: %.1f us per pair\n", ms * 1e3);
        }
        RUN9(3, 0, "baseline");
        RUN9(0, 0, "all exp2 on MUFU");
        RUN9(2, 0, "poly 1/2");
        RUN9(4, 0, "poly 1/4");
        if (use_win == 0) {
            RUN9(3, 1, "no exponential");
            RUN9(3, 4, "no offset add");
            RUN9(3, 15, "no exp / sum / offset / pack");
        }
    }
    // ---- event trace of CTA 0 (LAB bit 5): per role (softmax half A / B quad 0, PV issuer, QK issuer) tagged clock64() stamps
    for (int lab : {32}) {
        long long* d_tr;
        CK(cudaMalloc(&d_tr, 4 * 1024 * sizeof(long long)));
        CK(cudaMemset(d_tr, 0, 4 * 1024 * sizeof(long long)));
        CK(cudaMemcpyToSymbol(g_a11_trace, &d_tr, sizeof(d_tr)));
        k_attn9_plan<<<(n_items + 7) / 8, 256>>>(d_work, n_items, nh, slopes, normtab, 0, d_win, nullptr);
        if (lab == 32) CK((launch_attn11_t<3, 32>(tm, tmp, ctx, H, nh, slopes, d_work, d_win, nullptr, nwork, sms, 0, false)));
        else { }
        CK(cudaDeviceSynchronize());
        std::vector<long long> tr(4 * 1024);
        CK(cudaMemcpy(tr.data(), d_tr, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        struct Ev { long long t; int role, tag; };
        std::vector<Ev> ev;
        for (int r = 0; r < 4; ++r)
            for (int i = 0; i < 1024; ++i)
                if (tr[r * 1024 + i]) ev.push_back({tr[r * 1024 + i] & 0xffffffffffffll, r, (int)(tr[r * 1024 + i] >> 48)});
        std::sort(ev.begin(), ev.end(), [](const Ev& a, const Ev& b) { return a.t < b.t; });
        printf("# trace lab=%d (CTA 0; roles: 0/1 softmax half A/B warp 0, 2 PV issuer, 3 QK issuer; softmax tags: 1 scores seen, 2-5 chunk 0-3 in registers "
               "(3 = piece 0 released, 5 = piece 1 released), 6 block done; PV: 20+2h+piece p_ready seen; QK: 30+2h+piece s_free seen)\n", lab);
        const long long t0 = ev.empty() ? 0 : ev[0].t;
        int shown = 0;
        for (const Ev& e : ev) {
            if (e.t - t0 < 40000) continue;                     // skip the first item: steady state
            printf("  %8lld role %d tag %2d\n", e.t - t0, e.role, e.tag);
            if (++shown >= 150) break;
        }
        cudaFree(d_tr);
    }
    {   // fine trace of v10 (LAB bit 6): roles 0-3 softmax warps of half A, 4 QK issuer, 5 PV issuer
        long long* d_tr;
        CK(cudaMalloc(&d_tr, 8 * 1024 * sizeof(long long)));
        CK(cudaMemset(d_tr, 0, 8 * 1024 * sizeof(long long)));
        CK(cudaMemcpyToSymbol(g_a10_trace, &d_tr, sizeof(d_tr)));
        k_attn9_plan<<<(n_items + 7) / 8, 256>>>(d_work, n_items, nh, slopes, normtab, 0, d_win, nullptr);
        CK((launch_attn10_t<3, 64>(tm, tmp, ctx, H, nh, slopes, d_work, d_win, nullptr, nwork, sms, 0, false)));
        CK(cudaDeviceSynchronize());
        std::vector<long long> tr(8 * 1024);
        CK(cudaMemcpy(tr.data(), d_tr, tr.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        struct Ev { long long t; int role, tag; };
        std::vector<Ev> ev;
        for (int r = 0; r < 8; ++r)
            for (int i = 0; i < 1024; ++i)
                if (tr[r * 1024 + i]) ev.push_back({tr[r * 1024 + i] & 0xffffffffffffll, r, (int)(tr[r * 1024 + i] >> 48)});
        std::sort(ev.begin(), ev.end(), [](const Ev& a, const Ev& b) { return a.t < b.t; });
        printf("# fine trace v10 (CTA 0): fine 0-3 = softmax warps of half A (1 scores seen, 2 chunk 0 in regs, 5 S released, 8 P handed over), fine 4 = QK issuer "
               "(30+h s_free seen, 40+h before issue, 42+h issued + committed, 32+h after syncwarp), fine 5 = PV issuer (20+h p_ready seen, 50+h issued, 22+h end)\n");
        const long long t0 = ev.empty() ? 0 : ev[0].t;
        int shown = 0;
        for (const Ev& e : ev) {
            if (e.t - t0 < 40000) continue;
            printf("  %8lld fine %d tag %2d\n", e.t - t0, e.role, e.tag);
            if (++shown >= 260) break;
        }
        cudaFree(d_tr);
    }
    {
        const float ms = time_ms([&] { CK(launch_attn8_raw(tm, ctx, H, nh, slopes, d_work, nwork, 3, sms, 0, false, false)); }, 20);
        printf("  v8 poly=3 %-45s %8.1f us  %7.1f TFLOP/s\n", "", ms * 1e3, flops / (ms * 1e-3) / 1e12);
    }
    return 0;
}
