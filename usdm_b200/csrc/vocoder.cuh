// BigVGAN v1 vocoder kernels (reference src/decoder/voicebox/vocoder/models.py:192-222).
//
// Activations are TIME-MAJOR: x[b][t][c] (c contiguous), so a 1-D convolution over time is a GEMM whose A operand is the
// activation matrix shifted by (tap - taps/2) * dilation ROWS per tap -- the shift is a TMA coordinate and the zero padding at
// both ends of every utterance is TMA's out-of-bounds fill on the time axis of a 3-D (channel, time, batch) tensor map, so no
// im2col buffer exists.  A ConvTranspose1d(k = 2u, stride u, pad u/2) (models.py:158-163) is the same kernel: 3 taps and
// N = u * C_out outputs per input step, row [u * C_out] of step t being output steps t*u .. t*u+u-1 (weights regrouped once).
//   voc_conv_halo_kernel: tcgen05 implicit-GEMM conv (default); warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue, persistent;
//                        the activation tile is fetched once with its halo and each tap is a row-shifted smem descriptor;
//                        2 TMEM accumulators; epilogue out = alpha * (acc + bias + resid) (+ out).
//   voc_conv_tc_kernel : same math, one activation fetch per tap (ablation `halo = 0`, and layers whose halo exceeds 32 rows).
//   k_voc_act          : Activation1d (alias_free_torch/act.py:21-27): x2 kaiser-sinc upsample (polyphase, replicate edges),
//                        Snake / SnakeBeta (activations.py:40-50, :107-119), x2 low-pass downsample -- one pass, registers only.
//   k_voc_conv_f32     : CUDA-core fp32 conv with the same epilogue (validation path, <= 1e-5 of the reference).
//   k_voc_post         : conv_post (C -> 1, k = 7) + tanh (models.py:217-219).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace tvb {

// ------------------------------------------------------------------------------------------------------------------------------
// weight preparation (run once)
// ------------------------------------------------------------------------------------------------------------------------------
// weight_norm (dim 0): w[i][...] = g[i] * v[i][...] / ||v[i]||   (torch.nn.utils.weight_norm, models.py:32-50,149,160,185)
__global__ void k_voc_weight_norm(const float* __restrict__ v, const float* __restrict__ g, float* __restrict__ w, int inner) {
    __shared__ float red[32];
    const float* vi = v + static_cast<size_t>(blockIdx.x) * inner;
    float s = 0.f;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) s += vi[i] * vi[i];
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        s = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (threadIdx.x == 0) red[0] = s;
    }
    __syncthreads();
    const float sc = g[blockIdx.x] / sqrtf(red[0]);
    float* wi = w + static_cast<size_t>(blockIdx.x) * inner;
    for (int i = threadIdx.x; i < inner; i += blockDim.x) wi[i] = vi[i] * sc;
}

// Conv1d weight [co][ci][k]  ->  tensor-core operand W2[n][tap][ci_pad] (bf16, zero padded) and CUDA-core operand Wt[tap][ci][n].
__global__ void k_voc_pack_conv(const float* __restrict__ w, int Cout, int Cin, int k, int n_pad, int ci_pad, __nv_bfloat16* __restrict__ w2,
                                float* __restrict__ wt) {
    const size_t total = static_cast<size_t>(n_pad) * k * ci_pad;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % ci_pad), tap = static_cast<int>((i / ci_pad) % k), n = static_cast<int>(i / (static_cast<size_t>(ci_pad) * k));
        float v = 0.f;
        if (n < Cout && ci < Cin) {
            v = w[(static_cast<size_t>(n) * Cin + ci) * k + tap];
            wt[(static_cast<size_t>(tap) * Cin + ci) * Cout + n] = v;
        }
        w2[i] = __float2bfloat16(v);
    }
}
// ConvTranspose1d weight [ci][co][k], stride u, padding (k - u) / 2  ->  3-tap conv with N = u * Cout:
//   out[t*u + r][co] = sum_s sum_ci x[t + s - 1][ci] * W[ci][co][r + pad + (1 - s) * u]     (taps outside [0, k) are zero)
__global__ void k_voc_pack_convT(const float* __restrict__ w, int Cin, int Cout, int k, int u, int n_pad, int ci_pad,
                                 __nv_bfloat16* __restrict__ w2, float* __restrict__ wt) {
    const int N = u * Cout, pad = (k - u) / 2;
    const size_t total = static_cast<size_t>(n_pad) * 3 * ci_pad;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int ci = static_cast<int>(i % ci_pad), s = static_cast<int>((i / ci_pad) % 3), n = static_cast<int>(i / (static_cast<size_t>(ci_pad) * 3));
        float v = 0.f;
        if (n < N && ci < Cin) {
            const int r = n / Cout, co = n % Cout, j = r + pad + (1 - s) * u;
            if (j >= 0 && j < k) v = w[(static_cast<size_t>(ci) * Cout + co) * k + j];
            wt[(static_cast<size_t>(s) * Cin + ci) * N + n] = v;
        }
        w2[i] = __float2bfloat16(v);
    }
}
__global__ void k_voc_tile_bias(const float* __restrict__ b, int Cout, int N, float* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) out[i] = b[i % Cout];
}
// per-channel snake constants: (alpha', 1 / (beta' + 1e-9)) with alpha' = exp(alpha) when log-scale (activations.py:112-117)
__global__ void k_voc_snake_consts(const float* __restrict__ alpha, const float* __restrict__ beta, int C, int logscale, float2* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    float a = alpha[c], b = beta[c];
    if (logscale) { a = expf(a); b = expf(b); }
    out[c] = make_float2(a, 1.0f / (b + 1e-9f));
}

// ------------------------------------------------------------------------------------------------------------------------------
// layout changes
// ------------------------------------------------------------------------------------------------------------------------------
// mel [B][C][T] fp32 (reference layout, models.py:192)  ->  time-major [B][T][ld] in OutT
template <typename OutT>
__global__ void k_voc_mel_in(const float* __restrict__ mel, int C, int T, int ld, OutT* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = mel + static_cast<size_t>(b) * C * T;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? src[static_cast<size_t>(c) * T + t] : 0.f;
    }
    __syncthreads();
    OutT* dst = out + static_cast<size_t>(b) * T * ld;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        if (t < T && c < C) dst[static_cast<size_t>(t) * ld + c] = static_cast<OutT>(tile[threadIdx.x][i]);
    }
}
// time-major [B][T][C] -> [B][C][T] (test hook output)
__global__ void k_voc_tm_to_cm(const float* __restrict__ x, int C, int T, float* __restrict__ out) {
    __shared__ float tile[32][33];
    const int b = blockIdx.z, t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
    const float* src = x + static_cast<size_t>(b) * T * C;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int t = t0 + i, c = c0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < C && t < T) ? src[static_cast<size_t>(t) * C + c] : 0.f;
    }
    __syncthreads();
    float* dst = out + static_cast<size_t>(b) * C * T;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, t = t0 + threadIdx.x;
        if (t < T && c < C) dst[static_cast<size_t>(c) * T + t] = tile[threadIdx.x][i];
    }
}
// fp32 [rows][C] -> bf16 [rows][ld]   (operand copy of a stage output for the next upsampling conv)
__global__ void k_voc_to_bf16(const float* __restrict__ x, size_t rows, int C, int ld, __nv_bfloat16* __restrict__ out) {
    const size_t total = rows * C;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const size_t r = i / C;
        const int c = static_cast<int>(i - r * C);
        out[r * ld + c] = __float2bfloat16(x[i]);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Activation1d: up x2 (12-tap kaiser-sinc, replicate edges, gain 2) -> snake -> low-pass + down x2 (12 taps, replicate edges)
//   u[2q]   = 2 * sum_m f[2m+1] * x[clamp(q + 2 - m)]         u[2q+1] = 2 * sum_m f[2m] * x[clamp(q + 3 - m)]       m = 0..5
//   v       = u + inv_b * sin^2(a * u)
//   y[t]    = sum_j f[j] * v[clamp(2t + j - 5, 0, 2T-1)]                                                            j = 0..11
// (resample.py:22-30 and filter.py:86-94 written out per output phase; checked against the reference form in the CPU test-suite.)
// One thread = one channel x VOC_ACT_TT consecutive outputs; x, (v_even, v_odd) windows live in registers.
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int VOC_ACT_TT = 64;
static_assert((VOC_ACT_TT + 6) % 7 == 0, "a chunk must be a whole number of 7-iteration groups");
struct VocFilt { float up[12]; float dn[12]; float dn_sum; };    // up[] carries the upsampler's x2 gain; dn_sum = sum(dn) (1 for the reference's filters)

// FAST (bf16 path): sin^2 through MUFU.COS -- v = u + hb - hb * cos(2 a u), hb = inv_b / 2; the constant hb passes through the
// unit-gain low-pass unchanged, so it is added once per output instead of once per upsampled sample.  The approximation's absolute
// error (~|arg| * 2^-23 for the |arg| of a few tens met here) is three orders of magnitude below the bf16 rounding of the output.
// The 7-deep x and (v_even, v_odd) windows are rings indexed by (iteration mod 7); the inner loop is unrolled by 7 so every ring
// index is a compile-time constant and the windows never move between registers.  A chunk is exactly 10 groups of 7 iterations
// (6 warm-up iterations + 64 outputs); chunks whose whole footprint lies inside the signal take the INTERIOR instantiation, which
// has no index clamps, no edge branch and no per-iteration range tests (they were half of the issued instructions).
// VEC = 2 (bf16 path, even channel counts): one thread carries TWO adjacent channels as packed fp32 pairs -- 8-byte loads, FFMA2 for
// both FIR filters and the snake update (the filter taps are scalar-broadcast operands of FFMA2), one 4-byte bf16x2 store --
// which halves the issued instructions per output once more (the kernel is issue-bound, not pipe- or HBM-bound).
__device__ __forceinline__ uint64_t vpack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void vunpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t vfma(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t vmul(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}
template <int VEC> struct VocLane;
template <> struct VocLane<1> {
    using V = float;
    static __device__ __forceinline__ V ld(const float* p) { return __ldg(p); }
    static __device__ __forceinline__ V fmac(float c, V x, V acc) { return fmaf(c, x, acc); }
    static __device__ __forceinline__ V zero() { return 0.f; }
};
template <> struct VocLane<2> {
    using V = uint64_t;
    static __device__ __forceinline__ V ld(const float* p) { return __ldg(reinterpret_cast<const unsigned long long*>(p)); }
    static __device__ __forceinline__ V fmac(float c, V x, V acc) { return vfma(vpack(c, c), x, acc); }
    static __device__ __forceinline__ V zero() { return 0ull; }
};

template <typename OutT, bool FAST, bool INTERIOR, int VEC>
__device__ __forceinline__ void voc_act_chunk(const float* __restrict__ xb, OutT* __restrict__ ob, int C, int T, int ld_out, int t0,
                                              const float2* __restrict__ cst, const VocFilt& F) {
    static_assert(VEC == 1 || FAST, "the packed variant exists for the bf16 path only");
    using L = VocLane<VEC>;
    using V = typename L::V;
    auto ldx = [&](int t) {
        if (!INTERIOR) t = t < 0 ? 0 : (t >= T ? T - 1 : t);
        return L::ld(xb + static_cast<size_t>(t) * C);
    };
    const float2 ab = cst[0], ab1 = VEC == 2 ? cst[1] : cst[0];
    const float a2 = 2.f * ab.x, hb = 0.5f * ab.y, a2b = 2.f * ab1.x, hbb = 0.5f * ab1.y;
    V y0;
    if constexpr (VEC == 2) y0 = vpack(hb * F.dn_sum, hbb * F.dn_sum);
    else y0 = FAST ? hb * F.dn_sum : 0.f;
    auto snake = [&](V u) -> V {                     // FAST returns v - hb (see above)
        if constexpr (VEC == 2) {
            float g0, g1;
            vunpack(vmul(u, vpack(a2, a2b)), g0, g1);
            return vfma(vpack(-hb, -hbb), vpack(__cosf(g0), __cosf(g1)), u);
        } else {
            if (FAST) return fmaf(-hb, __cosf(u * a2), u);
            const float s = sinf(u * ab.x);
            return fmaf(ab.y, s * s, u);
        }
    };
    const int t_end = INTERIOR ? t0 + VOC_ACT_TT : min(t0 + VOC_ACT_TT, T);
    const int q_last = t_end + 2;
    V xr[7], ve[7], vo[7];
#pragma unroll
    for (int i = 0; i < 7; ++i) ve[i] = vo[i] = L::zero();
    xr[0] = L::zero();
#pragma unroll
    for (int i = 0; i < 6; ++i) xr[1 + i] = ldx(t0 - 6 + i);          // x[q-3 .. q+2] for the first q = t0 - 3 ("iterations -6 .. -1")
    V nx[7];                                                           // the 7 samples the current group shifts in, loaded one group ahead
#pragma unroll
    for (int r = 0; r < 7; ++r) nx[r] = ldx(t0 + r);
    // running pointers (interior chunks): one 64-bit add per row instead of a 64-bit multiply-add per access
    const float* xp = xb + static_cast<size_t>(INTERIOR ? t0 + 7 : 0) * C;     // row q0 + 10 of the group being prefetched
    OutT* op = ob + static_cast<size_t>(t0) * ld_out;                          // row of the next output
    const size_t xstep = static_cast<size_t>(C), ostep = static_cast<size_t>(ld_out);

    // one group of 7 iterations starting at q0; FIRST = the warm-up group of an interior chunk (only its last iteration has an output)
    auto group = [&](int q0, auto first_tag) {
        constexpr bool FIRST = decltype(first_tag)::value;
        V pf[7];                                                       // next group's loads are in flight while this group computes
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            if (INTERIOR) { pf[r] = L::ld(xp); xp += xstep; }
            else pf[r] = ldx(q0 + 10 + r);
        }
#pragma unroll
        for (int r = 0; r < 7; ++r) {
            const int q = q0 + r;
            if (INTERIOR || q <= q_last) {
                xr[r] = nx[r];                                            // = x[q + 3]; window x[q-3+i] = xr[(r + 1 + i) % 7]
                V e, o;
                if (INTERIOR || (q >= 0 && q < T)) {
                    V ue = L::zero(), uo = L::zero();
#pragma unroll
                    for (int m = 0; m < 6; ++m) {
                        ue = L::fmac(F.up[2 * m + 1], xr[(r + 6 - m) % 7], ue);   // x[q + 2 - m] = window[5 - m]
                        uo = L::fmac(F.up[2 * m], xr[(r + 7 - m) % 7], uo);       // x[q + 3 - m] = window[6 - m]
                    }
                    e = snake(ue);
                    o = snake(uo);
                } else {
                    // replicate padding of v: left of 0 every position takes v[0], right of 2T-1 takes v[2T-1]
                    const int qc = q < 0 ? 0 : T - 1;
                    V uu = L::zero();
#pragma unroll
                    for (int m = 0; m < 6; ++m) uu = L::fmac(q < 0 ? F.up[2 * m + 1] : F.up[2 * m], ldx(qc + (q < 0 ? 2 : 3) - m), uu);
                    e = o = snake(uu);
                }
                ve[r] = e; vo[r] = o;                                     // pair q' = t - 3 + i sits in slot (r + 1 + i) % 7, t = q - 3
                const int t = q - 3;
                if (INTERIOR ? (!FIRST || r == 6) : (t >= t0 && t < t_end)) {
                    V y = L::fmac(F.dn[0], vo[(r + 1) % 7], y0);
#pragma unroll
                    for (int d = 1; d <= 5; ++d) {
                        y = L::fmac(F.dn[2 * d - 1], ve[(r + 1 + d) % 7], y);
                        y = L::fmac(F.dn[2 * d], vo[(r + 1 + d) % 7], y);
                    }
                    y = L::fmac(F.dn[11], ve[r], y);
                    if constexpr (VEC == 2) {
                        float y_lo, y_hi;
                        vunpack(y, y_lo, y_hi);
                        *reinterpret_cast<__nv_bfloat162*>(op) = __floats2bfloat162_rn(y_lo, y_hi);
                    } else {
                        *op = static_cast<OutT>(y);
                    }
                    op += ostep;
                }
            }
        }
#pragma unroll
        for (int r = 0; r < 7; ++r) nx[r] = pf[r];
    };
    if (INTERIOR) {
        group(t0 - 3, std::true_type{});
#pragma unroll 1
        for (int g = 1; g < (VOC_ACT_TT + 6) / 7; ++g) group(t0 - 3 + 7 * g, std::false_type{});
    } else {
        for (int q0 = t0 - 3; q0 <= q_last; q0 += 7) group(q0, std::false_type{});
    }
}

// F.up arrives pre-multiplied by the upsampler's gain of 2 (resample.py:27).  VEC = channels per thread (2 needs even C and ld_out).
template <typename OutT, bool FAST, int VEC>
__global__ void __launch_bounds__(128)
k_voc_act(const float* __restrict__ x, int C, int T, int ld_out, const float2* __restrict__ consts, VocFilt F, OutT* __restrict__ out) {
    // flat (time chunk, channel) -> thread map: consecutive lanes take consecutive channels and wrap into the next chunk, so narrow
    // layers (24 / 48 channels) still fill every lane and a warp's load is at most two contiguous runs
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    const int cv = C / VEC;
    const int c = (gid % cv) * VEC, t0 = (gid / cv) * VOC_ACT_TT, b = blockIdx.y;
    if (t0 >= T) return;
    const float* xb = x + static_cast<size_t>(b) * T * C + c;
    OutT* ob = out + static_cast<size_t>(b) * T * ld_out + c;
    if (t0 >= 6 && t0 + VOC_ACT_TT + 16 <= T) voc_act_chunk<OutT, FAST, true, VEC>(xb, ob, C, T, ld_out, t0, consts + c, F);
    else voc_act_chunk<OutT, FAST, false, VEC>(xb, ob, C, T, ld_out, t0, consts + c, F);
}

// ------------------------------------------------------------------------------------------------------------------------------
// conv epilogue shared by both conv kernels:  out = alpha * (acc + bias[n] + resid) (+ out when accumulate)
// ------------------------------------------------------------------------------------------------------------------------------
struct VocConvParams {
    int B, T, Cin, N;            // T = input (= output) time steps per utterance, N = outputs per step (u * C_out for the upsamplers)
    int taps, dil;               // tap s reads input step t + (s - taps/2) * dil
    int ci_pad, n_pad, bn;       // padded K per tap (multiple of 64), padded N (multiple of bn), N tile
    const float* bias;           // [N]
    const float* resid;          // [B][T][N] or nullptr
    float* out;                  // [B][T][N]
    float alpha;
    int accumulate;
};

// CUDA-core fp32 reference conv (validation path): thread = (4 consecutive time steps, n); Wt[tap][ci][n] coalesced over n.
__global__ void __launch_bounds__(128)
k_voc_conv_f32(const float* __restrict__ x, const float* __restrict__ wt, VocConvParams p) {
    const int n = blockIdx.y * blockDim.x + threadIdx.x;
    const int t0 = blockIdx.x * 4, b = blockIdx.z;
    if (n >= p.N) return;
    const float* xb = x + static_cast<size_t>(b) * p.T * p.Cin;
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int s = 0; s < p.taps; ++s) {
        const int sh = (s - p.taps / 2) * p.dil;
        const float* w = wt + static_cast<size_t>(s) * p.Cin * p.N + n;
        const float* xr[4];
        bool ok[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int t = t0 + i + sh;
            ok[i] = t >= 0 && t < p.T && t0 + i < p.T;
            xr[i] = xb + static_cast<size_t>(ok[i] ? t : 0) * p.Cin;
        }
        for (int ci = 0; ci < p.Cin; ++ci) {
            const float wv = __ldg(w + static_cast<size_t>(ci) * p.N);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i] = fmaf(ok[i] ? __ldg(xr[i] + ci) : 0.f, wv, acc[i]);
        }
    }
    const float bv = p.bias[n];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int t = t0 + i;
        if (t >= p.T) break;
        const size_t off = (static_cast<size_t>(b) * p.T + t) * p.N + n;
        float v = acc[i] + bv;
        if (p.resid) v += p.resid[off];
        v *= p.alpha;
        if (p.accumulate) v += p.out[off];
        p.out[off] = v;
    }
}

// tcgen05 implicit-GEMM conv.  A: 3-D map (Cin, T, B) bf16, box 64 x 128 x 1, 128-byte swizzle.  W: 2-D map [n_pad][taps * ci_pad],
// box 64 x bn.  K loop = taps x (ci_pad / 64) stages; tile = (128 time steps, bn outputs) of one utterance.
constexpr int VC_BM = 128;
constexpr int VC_A_BYTES = VC_BM * 64 * 2;
constexpr int VC_THREADS = 192;
// ACC_COLS = TMEM columns per accumulator (>= the layer's N tile).  Wide layers (N tile > 128) run one CTA per SM with a 4-deep ring;
// narrow layers (the 96 .. 24-channel stages: short K loops, epilogue-bound) run 2 / 3 CTAs per SM with a 3-deep ring each, so
// several epilogues' global-memory latencies overlap.
template <int ACC_COLS> struct VcCfg {
    static constexpr int kStages = ACC_COLS > 128 ? 4 : 3;
    static constexpr int kMinBlocks = ACC_COLS > 128 ? 1 : (ACC_COLS > 64 ? 2 : 3);
    static constexpr int kBBytes = ACC_COLS * 64 * 2;
    static constexpr int kStageBytes = VC_A_BYTES + kBBytes;
    static constexpr int kSmem = kStages * kStageBytes + 1024 + 256;
    static_assert(kMinBlocks * (kSmem + 1024) <= 232448, "shared memory of the co-resident CTAs");
    static_assert(kMinBlocks * 2 * ACC_COLS <= 512, "TMEM of the co-resident CTAs");
};

__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int32_t c0, int32_t c1, int32_t c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}

// 256-bit global accesses (sm_100a): each thread owns one output row, so a warp-wide access touches 32 different rows; 32-byte
// pieces make every transaction a whole sector (16-byte float4 accesses would be two half-sector transactions for the same data).
struct f32x8 { float v[8]; };
__device__ __forceinline__ f32x8 ldg256(const float* p) {
    f32x8 r;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]), "=f"(r.v[6]), "=f"(r.v[7]) : "l"(p));
    return r;
}
__device__ __forceinline__ void stg256(float* p, const f32x8& r) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(r.v[0]), "f"(r.v[1]), "f"(r.v[2]), "f"(r.v[3]), "f"(r.v[4]),
                 "f"(r.v[5]), "f"(r.v[6]), "f"(r.v[7]) : "memory");
}

// One epilogue warp's 32 rows of a (128 x bn) accumulator: out = alpha * (acc + bias + resid) (+ out), 16 columns at a time.
// vec = N % 8 == 0 (every layer of the shipped model): 256-bit accesses; otherwise a scalar path (narrow test configurations).
__device__ __forceinline__ void voc_epilogue_rows(const VocConvParams& p, uint32_t acc_tmem, int quad, int lane, int nt, int mt, int b, bool vec,
                                                  int cset = 0, int nsets = 1) {
    const int t = mt * VC_BM + quad * 32 + lane;
    const bool row_ok = t < p.T;
    const size_t row_off = (static_cast<size_t>(b) * p.T + (row_ok ? t : 0)) * p.N;
#pragma unroll 1
    for (int c = cset; c < p.bn / 16; c += nsets) {       // `nsets` warps share a lane quarter and interleave its 16-column chunks
        const int n0 = nt * p.bn + c * 16;
        const bool live = row_ok && n0 < p.N;
        // the residual / running-sum operands do not depend on the accumulator: get them in flight before waiting on TMEM
        f32x8 rq[2], oq[2];
        if (vec && live) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = n0 + 8 * j;
                if (p.resid && n < p.N) rq[j] = ldg256(p.resid + row_off + n);
                if (p.accumulate && n < p.N) oq[j] = ldg256(p.out + row_off + n);
            }
        }
        uint32_t r[16];
        tmem_ld_32x16(acc_tmem + (static_cast<uint32_t>(quad * 32) << 16) + c * 16, r);
        tmem_ld_wait();
        if (!live) continue;
        if (vec) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int n = n0 + 8 * j;
                if (n >= p.N) break;
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n)), b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                f32x8 v;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float x = __uint_as_float(r[8 * j + e]) + bv[e];
                    if (p.resid) x += rq[j].v[e];
                    x *= p.alpha;
                    if (p.accumulate) x += oq[j].v[e];
                    v.v[e] = x;
                }
                stg256(p.out + row_off + n, v);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int n = n0 + j;
                if (n < p.N) {
                    float v = __uint_as_float(r[j]) + p.bias[n];
                    if (p.resid) v += p.resid[row_off + n];
                    v *= p.alpha;
                    if (p.accumulate) v += p.out[row_off + n];
                    p.out[row_off + n] = v;
                }
            }
        }
    }
}

template <int ACC_COLS>        // 256, 128 or 64
__global__ void __launch_bounds__(VC_THREADS, VcCfg<ACC_COLS>::kMinBlocks)
voc_conv_tc_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW, VocConvParams p) {
    using Cfg = VcCfg<ACC_COLS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
    uint64_t* empty_bar = full_bar + Cfg::kStages;
    uint64_t* tfull_bar = empty_bar + Cfg::kStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.T + VC_BM - 1) / VC_BM, n_tiles = p.n_pad / p.bn;
    const int num_tiles = p.B * m_tiles * n_tiles;
    const int kc = p.ci_pad / 64, ksteps = p.taps * kc;
    const uint32_t stage_tx = VC_A_BYTES + static_cast<uint32_t>(p.bn) * 128u;

    if (warp_idx == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<2 * ACC_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % n_tiles, mt = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const int tap = ks / kc, k0 = (ks - tap * kc) * 64;
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::kStageBytes;
                    mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
                    tma_load_3d(sa, &tmX, &full_bar[stage], k0, mt * VC_BM + (tap - p.taps / 2) * p.dil, b);
                    tma_load_2d(sa + VC_A_BYTES, &tmW, &full_bar[stage], tap * p.ci_pad + k0, nt * p.bn);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, p.bn);
            int stage = 0; uint32_t phase = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                mbar_wait(&tempty_bar[as], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * ACC_COLS;
                for (int ks = 0; ks < ksteps; ++ks) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
                    const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                    const uint64_t db = make_smem_desc_sw128(sa + VC_A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (ks | k) != 0);
                    umma_commit(&empty_bar[stage]);
                    if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull_bar[as]);
            }
        }
    } else {
        const int quad = warp_idx & 3;
        const bool vec = (p.N & 7) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const int nt = tile % n_tiles, mt = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            mbar_wait(&tfull_bar[as], (it >> 1) & 1);
            tc_fence_after();
            voc_epilogue_rows(p, tmem_base + as * ACC_COLS, quad, lane, nt, mt, b, vec);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<2 * ACC_COLS>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------------------------------------------
// Halo variant (default): the activation tile is fetched ONCE per (tile, 64-channel chunk) with its halo --
// rows [t0 - halo, t0 + 128 + halo), halo = (taps / 2) * dilation <= 32 -- and every tap's MMA reads it through a shared-memory
// descriptor whose start address is advanced by tap * dilation rows (128 B each; the 128-byte swizzle is a function of the
// absolute shared-memory address, so a row-shifted start sees the tile exactly as TMA laid it out).  Per tile this replaces
// `taps` 16 KB activation fetches by one <= 24 KB fetch: 3-11x less L2 -> SM traffic, which is what bounds the narrow
// (96 .. 24-channel) stages.  Two rings: 2 activation slots (24 KB) and 3-4 weight slots (bn x 128 B).
// ------------------------------------------------------------------------------------------------------------------------------
constexpr int VH_A_ROWS = 192;
constexpr int VH_A_BYTES = VH_A_ROWS * 128;        // 24 KB
constexpr int VH_A_SLOTS = 2;
template <int ACC_COLS> struct VhCfg {
    // wide tiles run one CTA per SM, so their epilogue gets two warps per TMEM lane quarter (8 epilogue warps); the multi-CTA
    // narrow variants keep 4 (their parallelism comes from the co-resident CTAs, and registers are the limit there)
    static constexpr int kEpiSets = ACC_COLS > 128 ? 2 : 1;
    static constexpr int kThreads = 64 + 128 * kEpiSets;
    static constexpr int kWStages = ACC_COLS > 128 ? 4 : 3;
    static constexpr int kMinBlocks = ACC_COLS > 128 ? 1 : (ACC_COLS > 64 ? 2 : 3);
    static constexpr int kWBytes = ACC_COLS * 64 * 2;
    static constexpr int kSmemW = VH_A_SLOTS * VH_A_BYTES;
    static constexpr int kSmemBar = kSmemW + kWStages * kWBytes;
    static constexpr int kSmem = kSmemBar + 1024 + 256;
    static_assert(kMinBlocks * (kSmem + 1024) <= 232448, "shared memory of the co-resident CTAs");
    static_assert(kMinBlocks * 2 * ACC_COLS <= 512, "TMEM of the co-resident CTAs");
};

template <int ACC_COLS>
__global__ void __launch_bounds__(VhCfg<ACC_COLS>::kThreads, VhCfg<ACC_COLS>::kMinBlocks)
voc_conv_halo_kernel(const __grid_constant__ CUtensorMap tmX /* box 64 x a_rows x 1 */, const __grid_constant__ CUtensorMap tmW, VocConvParams p,
                     int a_rows) {
    using Cfg = VhCfg<ACC_COLS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + Cfg::kSmemBar);
    uint64_t* a_empty = a_full + VH_A_SLOTS;
    uint64_t* w_full = a_empty + VH_A_SLOTS;
    uint64_t* w_empty = w_full + Cfg::kWStages;
    uint64_t* tfull_bar = w_empty + Cfg::kWStages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp_idx = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_tiles = (p.T + VC_BM - 1) / VC_BM, n_tiles = p.n_pad / p.bn;
    const int num_tiles = p.B * m_tiles * n_tiles;
    const int kc = p.ci_pad / 64;
    const int halo = (p.taps / 2) * p.dil;
    const uint32_t a_tx = static_cast<uint32_t>(a_rows) * 128u, w_tx = static_cast<uint32_t>(p.bn) * 128u;

    if (warp_idx == 0 && lane == 0) { tma_prefetch_desc(&tmX); tma_prefetch_desc(&tmW); }
    if (warp_idx == 1 && lane == 0) {
        for (int s = 0; s < VH_A_SLOTS; ++s) { mbar_init(&a_full[s], 1); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < Cfg::kWStages; ++s) { mbar_init(&w_full[s], 1); mbar_init(&w_empty[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], 4 * Cfg::kEpiSets); }
        fence_barrier_init();
    }
    if (warp_idx == 2) tmem_alloc<2 * ACC_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp_idx == 0) {
        if (elect_one()) {
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                const int nt = tile % n_tiles, mt = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
                for (int c = 0; c < kc; ++c) {
                    mbar_wait(&a_empty[as], aph ^ 1);
                    mbar_arrive_expect_tx(&a_full[as], a_tx);
                    tma_load_3d(smem + as * VH_A_BYTES, &tmX, &a_full[as], c * 64, mt * VC_BM - halo, b);
                    if (++as == VH_A_SLOTS) { as = 0; aph ^= 1; }
                    for (int tap = 0; tap < p.taps; ++tap) {
                        mbar_wait(&w_empty[ws], wph ^ 1);
                        mbar_arrive_expect_tx(&w_full[ws], w_tx);
                        tma_load_2d(smem + Cfg::kSmemW + ws * Cfg::kWBytes, &tmW, &w_full[ws], tap * p.ci_pad + c * 64, nt * p.bn);
                        if (++ws == Cfg::kWStages) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp_idx == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(128, p.bn);
            int as = 0, ws = 0; uint32_t aph = 0, wph = 0; int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
                for (int c = 0; c < kc; ++c) {
                    mbar_wait(&a_full[as], aph);
                    const uint32_t sa0 = smem_u32(smem + as * VH_A_BYTES);
                    for (int tap = 0; tap < p.taps; ++tap) {
                        mbar_wait(&w_full[ws], wph);
                        tc_fence_after();
                        const uint32_t sa = sa0 + static_cast<uint32_t>(tap * p.dil) * 128u;       // row shift = tap * dilation
                        // measured on B200: the descriptor's base-offset field must stay 0 here (the pattern start is 1024-aligned);
                        // setting it to the start row's phase breaks the result
                        const uint64_t da = make_smem_desc_sw128(sa, 16, 1024);
                        const uint64_t db = make_smem_desc_sw128(smem_u32(smem + Cfg::kSmemW + ws * Cfg::kWBytes), 16, 1024);
#pragma unroll
                        for (int k = 0; k < 4; ++k) umma_ss(d_tmem, da + 2 * k, db + 2 * k, idesc, (c | tap | k) != 0);
                        umma_commit(&w_empty[ws]);
                        if (++ws == Cfg::kWStages) { ws = 0; wph ^= 1; }
                    }
                    umma_commit(&a_empty[as]);
                    if (++as == VH_A_SLOTS) { as = 0; aph ^= 1; }
                }
                umma_commit(&tfull_bar[acc]);
            }
        }
    } else {
        const int quad = warp_idx & 3;
        const bool vec = (p.N & 7) == 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
            const int acc = it & 1;
            const int nt = tile % n_tiles, mt = (tile / n_tiles) % m_tiles, b = tile / (n_tiles * m_tiles);
            mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
            tc_fence_after();
            voc_epilogue_rows(p, tmem_base + acc * ACC_COLS, quad, lane, nt, mt, b, vec, (warp_idx - 2) >> 2, Cfg::kEpiSets);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp_idx == 2) {
        tc_fence_after();
        tmem_dealloc<2 * ACC_COLS>(tmem_base);
    }
}

// conv_post (C -> 1, k = 7, pad 3) + tanh (models.py:217-219); x is the fp32 output of activation_post, time-major [B][T][C].
__global__ void k_voc_post(const float* __restrict__ x, int C, int T, const float* __restrict__ w /* [C][7] */, const float* __restrict__ bias,
                           float* __restrict__ wav /* [B][T] */) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x, b = blockIdx.y;
    if (t >= T) return;
    const float* xb = x + static_cast<size_t>(b) * T * C;
    float acc = bias[0];
    for (int j = 0; j < 7; ++j) {
        const int tt = t + j - 3;
        if (tt < 0 || tt >= T) continue;
        const float* xr = xb + static_cast<size_t>(tt) * C;
        for (int c = 0; c < C; ++c) acc = fmaf(xr[c], __ldg(w + c * 7 + j), acc);
    }
    wav[static_cast<size_t>(b) * T + t] = tanhf(acc);
}

}  // namespace tvb
