// CUDA-core kernels of the token-Voicebox estimator: the row/elementwise tails of the bf16 path and the complete
// fp32 validation path (TVB_PRECISION_FP32_VALIDATE).  All of them work on the packed row layout described in
// one row per token, sequences separated by >= 15 zero "gap" rows so that the k=31 positional convolution
// sees zeros beyond a sequence (reference networks.py:333 zeroes padded rows before the convs).
#pragma once
#include "common.cuh"

namespace tvb {

struct RowLayout {
    int Mp;                 // rows in every [rows, C] activation buffer (multiple of 128)
    int nseq;               // sequences in this pass (2B with CFG: [uncond ; cond], voicebox.py:61-65)
    const int* row_seq;     // [Mp] sequence id of the row, -1 for gap rows
    const int* row_pos;     // [Mp] position inside the sequence; 0 = time token, tau + 1 = mel frame tau
    const int* seq_start;   // [nseq] first row
    const int* seq_len;     // [nseq] valid tokens (lengths + 1, networks.py:310)
    const int* seq_utt;     // [nseq] utterance index into the caller's [B, ...] tensors
    const int* seq_uncond;  // [nseq] 1 = unconditional branch (unit id := n_tokens, cond := 0)
};

template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f32(float v);
template <> __device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ------------------------------------------------------------------------------------------------------------
// Frame projection: the K = n_feats slice of proj_in (networks.py:305-307), split by linearity into
//   PREP (once per solve):  base[row] = G[unit] + b_in + Wc . cond[:, tau]         (step-invariant part)
//   STEP (every estimator call): h_in[row] = base[row] + Wy . y[:, tau]; row 0 of a sequence = time token (:19-28,:312)
// G = sqrt(E) * embed @ W_e^T is precomputed at weight-finalise time.  Wt is the [n_feats][H] transposed weight slice.
// ------------------------------------------------------------------------------------------------------------
constexpr int FP_ROWS = 16;
template <bool PREP>
__global__ void __launch_bounds__(256)
k_frame_proj(RowLayout L, int T, int nf, int H, const float* __restrict__ Wt, const float* __restrict__ src,
             const int64_t* __restrict__ x, const float* __restrict__ G, const float* __restrict__ b_in, int uncond_id,
             int src_is_zero, const float* __restrict__ base, const float* __restrict__ t_seq, float t_uniform,
             const float* __restrict__ freqs, float* __restrict__ out_f32, __nv_bfloat16* __restrict__ out_bf16,
             int* __restrict__ err_flag = nullptr) {
    extern __shared__ float s_src[];   // [nf][FP_ROWS]
    __shared__ int s_seq[FP_ROWS], s_pos[FP_ROWS];
    const int r0 = blockIdx.x * FP_ROWS;
    if (threadIdx.x < FP_ROWS) {
        s_seq[threadIdx.x] = L.row_seq[r0 + threadIdx.x];
        s_pos[threadIdx.x] = L.row_pos[r0 + threadIdx.x];
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < nf * FP_ROWS; idx += blockDim.x) {
        const int c = idx / FP_ROWS, r = idx % FP_ROWS;
        const int seq = s_seq[r], pos = s_pos[r];
        float v = 0.f;
        if (seq >= 0 && pos >= 1 && !src_is_zero && !(PREP && L.seq_uncond[seq]))
            v = src[(static_cast<size_t>(L.seq_utt[seq]) * nf + c) * T + (pos - 1)];
        s_src[idx] = v;
    }
    __syncthreads();
    for (int n = threadIdx.x * 4; n < H; n += blockDim.x * 4) {
        float4 acc[FP_ROWS];
#pragma unroll
        for (int r = 0; r < FP_ROWS; ++r) {
            const int seq = s_seq[r], pos = s_pos[r];
            const int row = r0 + r;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
            if (seq >= 0) {
                if (PREP) {
                    if (pos >= 1) {
                        int64_t id = L.seq_uncond[seq] ? uncond_id : x[static_cast<size_t>(L.seq_utt[seq]) * T + (pos - 1)];
                        // nn.Embedding asserts device-side on an id outside its table (networks.py:305); here the lookup stays inside
                        // G (the unconditional row stands in) and the handle's sticky flag reports the call as invalid
                        if (id < 0 || id > uncond_id) {
                            if (err_flag && n == 0) *err_flag = 1;
                            id = uncond_id;
                        }
                        const float4 g = *reinterpret_cast<const float4*>(G + static_cast<size_t>(id) * H + n);
                        const float4 b = *reinterpret_cast<const float4*>(b_in + n);
                        a = make_float4(g.x + b.x, g.y + b.y, g.z + b.z, g.w + b.w);
                    }
                } else {
                    if (pos >= 1) {
                        a = *reinterpret_cast<const float4*>(base + static_cast<size_t>(row) * H + n);
                    } else {
                        const float t = t_seq ? t_seq[seq] : t_uniform;
                        const float st = 1000.0f * t;               // (scale * x) * emb, networks.py:26
                        const int half = H / 2;
                        float e[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int k = n + j;
                            e[j] = (k < half) ? sinf(st * freqs[k]) : cosf(st * freqs[k - half]);
                        }
                        a = make_float4(e[0], e[1], e[2], e[3]);
                    }
                }
            }
            acc[r] = a;
        }
        for (int c = 0; c < nf; ++c) {
            const float4 w = *reinterpret_cast<const float4*>(Wt + static_cast<size_t>(c) * H + n);
#pragma unroll
            for (int r = 0; r < FP_ROWS; ++r) {
                const float s = s_src[c * FP_ROWS + r];
                acc[r].x = fmaf(w.x, s, acc[r].x);
                acc[r].y = fmaf(w.y, s, acc[r].y);
                acc[r].z = fmaf(w.z, s, acc[r].z);
                acc[r].w = fmaf(w.w, s, acc[r].w);
            }
        }
#pragma unroll
        for (int r = 0; r < FP_ROWS; ++r) {
            const size_t o = static_cast<size_t>(r0 + r) * H + n;
            *reinterpret_cast<float4*>(out_f32 + o) = acc[r];
            if (out_bf16)
                *reinterpret_cast<uint2*>(out_bf16 + o) = make_uint2(pack_bf16x2(acc[r].x, acc[r].y), pack_bf16x2(acc[r].z, acc[r].w));
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// bf16-path prologue of every estimator pass (replaces k_frame_proj<false> when the tensor-core frame projection is on):
// gathers the ODE state y [B][nf][T] (networks.py:302-309) into the packed-row GEMM operand
//     A[row][0:nf] = hi(y),  A[row][nf:2nf] = lo(y)   (y = hi + lo to 2^-17 in two bf16 planes),  zero elsewhere,
// so that A . [W_y | W_y | 0]^T + base is the frame projection on the tensor core, and writes this pass's time-token rows
// (raw sinusoid of 1000 t, networks.py:22-26) into `base`, the GEMM's residual operand.
// ------------------------------------------------------------------------------------------------------------
constexpr int PY_ROWS = 32;
__global__ void __launch_bounds__(256)
k_pack_y(RowLayout L, int T, int nf, int H, int ldA, const float* __restrict__ y, const float* __restrict__ t_seq, float t_uniform,
         const float* __restrict__ freqs, float* __restrict__ base, __nv_bfloat16* __restrict__ A) {
    extern __shared__ float s_y[];       // [PY_ROWS][nf + 1]
    __shared__ int s_seq[PY_ROWS], s_pos[PY_ROWS];
    const int r0 = blockIdx.x * PY_ROWS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < PY_ROWS) {
        s_seq[threadIdx.x] = L.row_seq[r0 + threadIdx.x];
        s_pos[threadIdx.x] = L.row_pos[r0 + threadIdx.x];
    }
    __syncthreads();
    {   // lane = row: consecutive rows are consecutive frames of one utterance, so each channel's read is contiguous
        const int seq = s_seq[lane], pos = s_pos[lane];
        const bool live = seq >= 0 && pos >= 1;
        const float* src = live ? y + static_cast<size_t>(L.seq_utt[seq]) * nf * T + (pos - 1) : y;
        for (int c = warp; c < nf; c += 8) s_y[lane * (nf + 1) + c] = live ? __ldg(src + static_cast<size_t>(c) * T) : 0.f;
    }
    __syncthreads();
    const int pairs = ldA / 2;
    for (int idx = threadIdx.x; idx < PY_ROWS * pairs; idx += blockDim.x) {
        const int r = idx / pairs, k = (idx - r * pairs) * 2;
        float v[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int kk = k + j;
            if (kk < nf) v[j] = __bfloat162float(__float2bfloat16(s_y[r * (nf + 1) + kk]));
            else if (kk < 2 * nf) { const float f = s_y[r * (nf + 1) + kk - nf]; v[j] = f - __bfloat162float(__float2bfloat16(f)); }
            else v[j] = 0.f;
        }
        *reinterpret_cast<uint32_t*>(A + static_cast<size_t>(r0 + r) * ldA + k) = pack_bf16x2(v[0], v[1]);
    }
    for (int r = 0; r < PY_ROWS; ++r) {
        if (s_seq[r] < 0 || s_pos[r] != 0) continue;             // block-uniform
        const float st = 1000.0f * (t_seq ? t_seq[s_seq[r]] : t_uniform);      // (scale * x) * emb, networks.py:26
        const int half = H / 2;
        for (int n = threadIdx.x; n < H; n += blockDim.x)
            base[static_cast<size_t>(r0 + r) * H + n] = (n < half) ? sinf(st * freqs[n]) : cosf(st * freqs[n - half]);
    }
}
// W2[n][k] = W_y[n][k] for k < nf, W_y[n][k - nf] for k < 2 nf, 0 up to ld   (Wy_t is W_y transposed: [nf][H])
__global__ void k_build_wy2(const float* __restrict__ Wy_t, int nf, int H, int ld, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= H * ld) return;
    const int n = i / ld, k = i % ld;
    out[i] = k < nf ? Wy_t[static_cast<size_t>(k) * H + n] : (k < 2 * nf ? Wy_t[static_cast<size_t>(k - nf) * H + n] : 0.f);
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm over the hidden dim (eps 1e-5; networks.py:348, :259, :263) + the row mask that follows it.
// One warp per row, the row held in registers (H <= 1024, H % 128 == 0), two-pass mean / variance in fp32.
// Writes the fp32 residual stream and, on the bf16 path, the bf16 copy the next GEMM consumes through TMA.
// ------------------------------------------------------------------------------------------------------------
template <typename TB>
__global__ void __launch_bounds__(256)
k_layernorm(const float* __restrict__ x, float* __restrict__ y, TB* __restrict__ yb, const float* __restrict__ gamma,
            const float* __restrict__ beta, const int* __restrict__ row_seq, int Mp, int H, float eps, float2* __restrict__ stats,
            int reverse) {
    // y == nullptr (with stats != nullptr): the fp32 result is not materialised; consumers that need it (the residual add of
    // the next GEMM epilogue, EPI_RESID_LN) recompute it from x and the (mean, rstd) written to stats[row].
    pdl_launch_dependents();
    pdl_wait();
    const int row_fwd = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row_fwd >= Mp) return;
    const int row = reverse ? Mp - 1 - row_fwd : row_fwd;      // reverse: start with the rows the producer wrote last (still in L2)
    const int nv = H / 128;                     // float4 per lane
    const bool valid = row_seq[row] >= 0;
    float4 v[8];
    const float4* x4 = reinterpret_cast<const float4*>(x + static_cast<size_t>(row) * H);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv) {
            v[j] = x4[lane + 32 * j];
            sum += (v[j].x + v[j].y) + (v[j].z + v[j].w);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / H;
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv) {
            const float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
            sq += (a * a + b * b) + (c * c + d * d);
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / H + eps);
    if (stats != nullptr && lane == 0) stats[row] = make_float2(mean, rstd);
    float4* y4 = reinterpret_cast<float4*>(y + static_cast<size_t>(row) * H);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (j < nv) {
            const int c4 = lane + 32 * j;
            const float4 g = reinterpret_cast<const float4*>(gamma)[c4];
            const float4 b = reinterpret_cast<const float4*>(beta)[c4];
            float4 o;
            o.x = valid ? (v[j].x - mean) * rstd * g.x + b.x : 0.f;
            o.y = valid ? (v[j].y - mean) * rstd * g.y + b.y : 0.f;
            o.z = valid ? (v[j].z - mean) * rstd * g.z + b.z : 0.f;
            o.w = valid ? (v[j].w - mean) * rstd * g.w + b.w : 0.f;
            if (y != nullptr) y4[c4] = o;
            if (yb) {
                if constexpr (sizeof(TB) == 2) {
                    reinterpret_cast<uint2*>(yb + static_cast<size_t>(row) * H)[c4] =
                        make_uint2(pack_bf16x2(o.x, o.y), pack_bf16x2(o.z, o.w));
                } else {
                    reinterpret_cast<float4*>(yb + static_cast<size_t>(row) * H)[c4] = o;
                }
            }
        }
}

// ------------------------------------------------------------------------------------------------------------
// Positional convolution (networks.py:87-95): grouped Conv1d(k, pad k/2, groups) -> erf-GELU -> row mask,
// direct form on CUDA cores.  Wp is the weight-norm-folded weight packed [group][tap][ic][oc] (fp32).
//   MODE 0: out = mask * gelu(conv + b)                       (first conv, feeds the second)
//   MODE 1: out = mask * gelu(conv + b) + resid               (second conv + residual, networks.py:346; pre-LN0)
// ------------------------------------------------------------------------------------------------------------
constexpr int PC_ROWS = 32;
template <typename TIn, typename TOut, int MODE>
__global__ void __launch_bounds__(256)
k_posconv_simt(const TIn* __restrict__ in, TOut* __restrict__ out, const float* __restrict__ resid,
               const float* __restrict__ Wp, const float* __restrict__ bias, const int* __restrict__ row_seq, int Mp, int H,
               int taps) {
    extern __shared__ float s_in[];   // [PC_ROWS + taps - 1][64]
    const int g = blockIdx.y;
    const int r0 = blockIdx.x * PC_ROWS;
    const int halo = taps / 2;
    const int nrows = PC_ROWS + taps - 1;
    for (int idx = threadIdx.x; idx < nrows * 64; idx += blockDim.x) {
        const int rr = idx >> 6, ic = idx & 63;
        const int row = r0 - halo + rr;
        s_in[idx] = (row >= 0 && row < Mp) ? to_f32<TIn>(in[static_cast<size_t>(row) * H + g * 64 + ic]) : 0.f;
    }
    __syncthreads();
    const int oc = threadIdx.x & 63;
    const int rs = (threadIdx.x >> 6) * 8;
    float acc[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = 0.f;
    const float* w = Wp + static_cast<size_t>(g) * taps * 64 * 64 + oc;
    for (int tap = 0; tap < taps; ++tap) {
#pragma unroll 8
        for (int ic = 0; ic < 64; ++ic) {
            const float wv = w[(tap * 64 + ic) * 64];
#pragma unroll
            for (int r = 0; r < 8; ++r) acc[r] = fmaf(wv, s_in[(rs + r + tap) * 64 + ic], acc[r]);
        }
    }
    const float b = bias[g * 64 + oc];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = r0 + rs + r;
        if (row >= Mp) continue;
        const size_t o = static_cast<size_t>(row) * H + g * 64 + oc;
        float v = (row_seq[row] >= 0) ? gelu_erf(acc[r] + b) : 0.f;
        if (MODE == 1) v += resid[o];
        out[o] = from_f32<TOut>(v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// fp32 GEMM on CUDA cores: C[M,N] = alpha * A[M,K] . W[N,K]^T + bias (+gelu | +resid).  Validation path and the
// small N = n_feats output projection (networks.py:372).  128x128x16 tiles, 8x8 micro-tiles, 256 threads.
// ------------------------------------------------------------------------------------------------------------
enum SgemmEpilogue : int { SG_PLAIN = 0, SG_GELU = 1, SG_RESID = 2 };
struct SgemmParams {
    int M, N, K, K1;           // reduction columns [0,K1) from A, [K1,K) from A2
    const float* A; int lda;
    const float* A2; int lda2;
    const float* W; int ldw;
    const float* bias;
    const float* resid; int ldr;
    float* C; int ldc;
    float alpha;
};
template <int EPI>
__global__ void __launch_bounds__(256)
k_sgemm(const SgemmParams p) {
    __shared__ float As[16][132];
    __shared__ float Bs[16][132];
    const int m0 = blockIdx.y * 128, n0 = blockIdx.x * 128;
    const int tid = threadIdx.x;
    const int ty = tid >> 4, tx = tid & 15;
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
        const float* Ab = (k0 < p.K1) ? p.A : p.A2;
        const int lda = (k0 < p.K1) ? p.lda : p.lda2;
        const int ka = (k0 < p.K1) ? k0 : k0 - p.K1;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int f = tid + i * 256;
            const int r = f >> 2, kq = (f & 3) * 4;
            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
            if (m0 + r < p.M) a = *reinterpret_cast<const float4*>(Ab + static_cast<size_t>(m0 + r) * lda + ka + kq);
            if (n0 + r < p.N) b = *reinterpret_cast<const float4*>(p.W + static_cast<size_t>(n0 + r) * p.ldw + k0 + kq);
            As[kq + 0][r] = a.x; As[kq + 1][r] = a.y; As[kq + 2][r] = a.z; As[kq + 3][r] = a.w;
            Bs[kq + 0][r] = b.x; Bs[kq + 1][r] = b.y; Bs[kq + 2][r] = b.z; Bs[kq + 3][r] = b.w;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int row = m0 + ty * 8 + i;
        if (row >= p.M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = n0 + tx * 8 + j;
            if (col >= p.N) continue;
            float v = p.alpha * acc[i][j];
            if (p.bias) v += p.bias[col];
            if (EPI == SG_GELU) v = gelu_erf(v);
            if (EPI == SG_RESID) v += p.resid[static_cast<size_t>(row) * p.ldr + col];
            p.C[static_cast<size_t>(row) * p.ldc + col] = v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Masked self-attention with the reference's additive term (networks.py:319-341, :174-193), CUDA-core version:
//   P = softmax_j( q_i . k_j - slope_h |i - j| [j != 0] ),  keys j >= len masked,  ctx_i = sum_j P_ij v_j.
// q is pre-scaled (0.125 folded into W_q/b_q at finalise).  One block = 16 queries of one (sequence, head); online
// softmax in fp32.  Used by the fp32 validation path and as the on-device cross-check of the tcgen05 kernel.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
k_attn_simt(const T* __restrict__ qkv, T* __restrict__ ctx, RowLayout L, int H, const float* __restrict__ slopes,
            const int2* __restrict__ work, float score_scale) {
    __shared__ float Qs[16][64];
    __shared__ float Ks[32][65];
    __shared__ float Vs[32][64];
    const int seq = work[blockIdx.x].x, q0 = work[blockIdx.x].y;
    const int head = blockIdx.y;
    const int Ls = L.seq_len[seq], start = L.seq_start[seq];
    const int ld = 3 * H;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float slope = slopes[head];
    for (int idx = threadIdx.x; idx < 16 * 64; idx += 128) {
        const int qi = idx >> 6, d = idx & 63;
        Qs[qi][d] = (q0 + qi < Ls) ? to_f32<T>(qkv[static_cast<size_t>(start + q0 + qi) * ld + head * 64 + d]) : 0.f;
    }
    float m[4], l[4], o0[4], o1[4];
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) { m[qq] = -INFINITY; l[qq] = 0.f; o0[qq] = 0.f; o1[qq] = 0.f; }
    for (int kt = 0; kt < Ls; kt += 32) {
        __syncthreads();
        for (int idx = threadIdx.x; idx < 32 * 64; idx += 128) {
            const int key = idx >> 6, d = idx & 63;
            const bool ok = kt + key < Ls;
            const size_t base = static_cast<size_t>(start + kt + key) * ld + head * 64 + d;
            Ks[key][d] = ok ? to_f32<T>(qkv[base + H]) : 0.f;
            Vs[key][d] = ok ? to_f32<T>(qkv[base + 2 * H]) : 0.f;
        }
        __syncthreads();
        const int j = kt + lane;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) {
            const int qi = warp * 4 + qq;
            const int i = q0 + qi;
            float s = 0.f;
#pragma unroll 16
            for (int d = 0; d < 64; ++d) s = fmaf(Qs[qi][d], Ks[lane][d], s);
            s *= score_scale;          // 1 (fp32 path) or ln 2 (bf16 path: q carries log2(e) for the tcgen05 kernel)
            if (j != 0) s -= slope * fabsf(static_cast<float>(i - j));
            if (j >= Ls) s = -INFINITY;
            float mx = s;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float mn = fmaxf(m[qq], mx);
            const float alpha = expf(m[qq] - mn);
            const float pj = (j < Ls) ? expf(s - mn) : 0.f;
            float ps = pj;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
            l[qq] = l[qq] * alpha + ps;
            float a0 = o0[qq] * alpha, a1 = o1[qq] * alpha;
#pragma unroll 8
            for (int k = 0; k < 32; ++k) {
                const float pk = __shfl_sync(0xffffffffu, pj, k);
                a0 = fmaf(pk, Vs[k][lane], a0);
                a1 = fmaf(pk, Vs[k][lane + 32], a1);
            }
            o0[qq] = a0; o1[qq] = a1; m[qq] = mn;
        }
    }
#pragma unroll
    for (int qq = 0; qq < 4; ++qq) {
        const int i = q0 + warp * 4 + qq;
        if (i < Ls) {
            const size_t o = static_cast<size_t>(start + i) * H + head * 64;
            ctx[o + lane] = from_f32<T>(o0[qq] / l[qq]);
            ctx[o + lane + 32] = from_f32<T>(o1[qq] / l[qq]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Output side: vbuf[row, c] holds W_out h + b for every token row (networks.py:372).
// k_scatter_velocity: plain estimator output  v[s, c, tau] (zero on padded frames, time token dropped, :372-374).
// k_ode_update: CFG combine (voicebox.py:69-71) + Euler / Heun state update (:86-87, :110, :121) + prompt re-noise
//               (:89-92, :115-118, :126-129), one thread per (utterance, mel bin, frame).
// ------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_scatter_velocity(const float* __restrict__ vbuf, int ldv, RowLayout L, int nf, int T, float* __restrict__ out) {
    const size_t n = static_cast<size_t>(L.nseq) * nf * T;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n; idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int tau = idx % T;
        const int c = (idx / T) % nf;
        const int s = idx / (static_cast<size_t>(T) * nf);
        float v = 0.f;
        if (tau + 1 < L.seq_len[s]) v = vbuf[static_cast<size_t>(L.seq_start[s] + 1 + tau) * ldv + c];
        out[idx] = v;
    }
}

struct OdeParams {
    int B, nf, T;
    int cfg;                 // 1: sequences [0,B) are uncond, [B,2B) cond
    float g;                 // gradient_scale
    float dt, t_next;
    int P;                   // prompt frames to re-noise (0 = none)
    float one_minus_sigma;   // (1 - sigma_min) as fp32
    const float* z_base;     // z the update is applied to
    const float* v_prev;     // Heun corrector: predictor velocity, else nullptr
    const float* eps;        // fresh noise [B,nf,T] (P > 0)
    const float* cond;
    float* z_out;
    float* v_out;            // optional: guided velocity of this evaluation
};
__global__ void __launch_bounds__(256)
k_ode_update(const float* __restrict__ vbuf, int ldv, RowLayout L, OdeParams p) {
    const size_t n = static_cast<size_t>(p.B) * p.nf * p.T;
    for (size_t idx = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; idx < n; idx += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int tau = idx % p.T;
        const int c = (idx / p.T) % p.nf;
        const int b = idx / (static_cast<size_t>(p.T) * p.nf);
        const int sc = p.cfg ? b + p.B : b;
        float v = 0.f;
        if (tau + 1 < L.seq_len[sc]) {
            const float vc = vbuf[static_cast<size_t>(L.seq_start[sc] + 1 + tau) * ldv + c];
            v = vc;
            if (p.cfg) {
                const float vu = vbuf[static_cast<size_t>(L.seq_start[b] + 1 + tau) * ldv + c];
                v = vc + p.g * (vc - vu);
            }
        }
        if (p.v_out) p.v_out[idx] = v;
        float z;
        if (p.v_prev) z = p.z_base[idx] + (p.dt * (p.v_prev[idx] + v)) / 2.0f;
        else          z = p.z_base[idx] + p.dt * v;
        if (tau < p.P) z = (1.0f - p.one_minus_sigma * p.t_next) * p.eps[idx] + p.t_next * p.cond[idx];
        p.z_out[idx] = z;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Weight packing kernels (run once in tvb_finalize_weights)
// ------------------------------------------------------------------------------------------------------------
// ---- LayerNorm folding (tvb_finalize_weights): W' = W diag(gamma) ; c[n] = sum_k bf16(W'[n,k]) ; b'[n] = b[n] + sum_k W[n,k] beta[k]
__global__ void k_fold_cols(const float* __restrict__ W, const float* __restrict__ gamma, float* __restrict__ out, int N, int K) {
    const size_t n = static_cast<size_t>(N) * K;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = W[i] * gamma[i % K];
}
// one warp per output row
__global__ void k_fold_rows(const float* __restrict__ W, const __nv_bfloat16* __restrict__ Wf16, const float* __restrict__ beta,
                            const float* __restrict__ bias, float* __restrict__ c_out, float* __restrict__ b_out, int N, int K) {
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= N) return;
    float c = 0.f, b = 0.f;
    for (int k = lane; k < K; k += 32) {
        c += __bfloat162float(Wf16[static_cast<size_t>(row) * K + k]);
        b = fmaf(W[static_cast<size_t>(row) * K + k], beta[k], b);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { c += __shfl_xor_sync(0xffffffffu, c, o); b += __shfl_xor_sync(0xffffffffu, b, o); }
    if (lane == 0) { c_out[row] = c; b_out[row] = bias[row] + b; }
}
__global__ void k_fill_f32(float* __restrict__ out, int n, float v) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = v;
}
__global__ void k_f32_to_bf16(const float* __restrict__ in, __nv_bfloat16* __restrict__ out, size_t n) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = __float2bfloat16_rn(in[i]);
}
__global__ void k_scale_copy(const float* __restrict__ in, float* __restrict__ out, size_t n, float s) {
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        out[i] = in[i] * s;
}
// out[c][n] = W[n, col0 + c]   (W row-major [H, ldw])
__global__ void k_transpose_slice(const float* __restrict__ W, int ldw, int col0, int ncols, int H, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < ncols * H) {
        const int c = i / H, n = i % H;
        out[i] = W[static_cast<size_t>(n) * ldw + col0 + c];
    }
}
// weight_norm(dim=2) (networks.py:78-82): one L2 norm per tap over (out, in) of v [H, cpg, taps].
__global__ void k_conv_tap_norm(const float* __restrict__ v, int H, int cpg, int taps, float* __restrict__ norms) {
    __shared__ double red[256];
    const int tap = blockIdx.x;
    double s = 0.0;
    for (int i = threadIdx.x; i < H * cpg; i += blockDim.x) {
        const double a = v[static_cast<size_t>(i) * taps + tap];
        s += a * a;
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) norms[tap] = static_cast<float>(sqrt(red[0]));
}
// Folded conv weight in two packings: fp32 [g][tap][ic][oc] (CUDA-core conv) and bf16 [g][tap][oc][ic] (tcgen05 conv, K-major B).
__global__ void k_conv_pack(const float* __restrict__ v, const float* __restrict__ gain, const float* __restrict__ norms, int H,
                            int cpg, int taps, float* __restrict__ wp_f32, __nv_bfloat16* __restrict__ wp_bf16) {
    const size_t n = static_cast<size_t>(H) * cpg * taps;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int tap = i % taps;
        const int ic = (i / taps) % cpg;
        const int ocg = i / (static_cast<size_t>(taps) * cpg);   // global out channel
        const int g = ocg / cpg, oc = ocg % cpg;
        const float w = v[i] * (gain[tap] / norms[tap]);
        wp_f32[((static_cast<size_t>(g) * taps + tap) * cpg + ic) * cpg + oc] = w;
        wp_bf16[((static_cast<size_t>(g) * taps + tap) * cpg + oc) * cpg + ic] = __float2bfloat16_rn(w);
    }
}

}  // namespace tvb
