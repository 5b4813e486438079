// Decoder glue on the device: the two steps either side of the sampler in reference
// src/decoder/voicebox/util/model_util.py reconstruct_speech():
//   k_process_unit      : process_unit (:50-54)  50 Hz units -> one unit per mel frame (repeat x441, majority vote over 256)
//   k_mel_spectrogram   : get_mel (:24-38) -> vocoder/meldataset.py mel_spectrogram (:55-78): reflect pad, 1024-point Hann STFT
//                         (hop 256, center=False), sqrt(re^2 + im^2 + 1e-9), mel filterbank, log(clamp(., 1e-5)), then the
//                         (mel - mean) / std normalisation of model_util.py:82.
// Both are small HBM/latency-bound kernels that run once per utterance, outside the ODE loop.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace tvb {

// out[f] = mode of rep-times-repeated units over samples [f*hop, (f+1)*hop); ties -> smallest value (torch.mode).
// General in rep / hop: a window may cover several units; one thread per output frame walks them.
__global__ void __launch_bounds__(256)
k_process_unit(const int64_t* __restrict__ units, int64_t n_units, int rep, int hop, int64_t* __restrict__ out, int64_t n_out) {
    const int64_t f = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (f >= n_out) return;
    const int64_t s0 = f * hop, s1 = s0 + hop;                 // sample window in the repeated sequence
    const int64_t u0 = s0 / rep, u1 = (s1 - 1) / rep;           // units touching it
    int64_t best = 0;
    int64_t best_cnt = -1;
    for (int64_t u = u0; u <= u1; ++u) {
        const int64_t val = units[u];
        // total count of `val` inside the window (it may occur in several runs); only evaluate at its first occurrence
        bool first = true;
        for (int64_t v = u0; v < u; ++v)
            if (units[v] == val) { first = false; break; }
        if (!first) continue;
        int64_t cnt = 0;
        for (int64_t v = u; v <= u1; ++v)
            if (units[v] == val) {
                const int64_t a = v * rep > s0 ? v * rep : s0;
                const int64_t b = (v + 1) * rep < s1 ? (v + 1) * rep : s1;
                cnt += b - a;
            }
        if (cnt > best_cnt || (cnt == best_cnt && val < best)) { best = val; best_cnt = cnt; }
    }
    out[f] = best;
}

// One CTA (512 threads) per frame: windowed frame -> radix-2 DIT FFT of NFFT complex points in shared memory -> magnitudes ->
// mel projection -> log -> normalise.  NFFT is a power of two <= 2048; win == NFFT (the reference's win_size == n_fft).
template <int NFFT>
__global__ void __launch_bounds__(512)
k_mel_spectrogram(const float* __restrict__ wav, int64_t n, int hop, int n_mels, const float* __restrict__ basis /*[n_mels, NFFT/2+1]*/,
                  float mean, float inv_std, float* __restrict__ out /*[n_mels, n_frames]*/, int64_t n_frames) {
    __shared__ float re[NFFT], im[NFFT];
    __shared__ float mag[NFFT / 2 + 1];
    constexpr int LOG = (NFFT == 2048) ? 11 : (NFFT == 1024) ? 10 : (NFFT == 512) ? 9 : 8;
    const int64_t f = blockIdx.x;
    const int pad = (NFFT - hop) / 2;
    for (int k = threadIdx.x; k < NFFT; k += blockDim.x) {
        int64_t src = f * hop + k - pad;                        // index into the unpadded signal, reflect (no edge repeat)
        if (src < 0) src = -src;
        if (src >= n) src = 2 * (n - 1) - src;
        const float wv = 0.5f - 0.5f * cospif(2.0f * static_cast<float>(k) / NFFT);     // periodic Hann (torch.hann_window)
        const int rk = __brev(static_cast<unsigned>(k)) >> (32 - LOG);                  // bit-reversed slot for the DIT FFT
        re[rk] = wav[src] * wv;
        im[rk] = 0.f;
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s <= LOG; ++s) {
        const int half = 1 << (s - 1);
        for (int t = threadIdx.x; t < NFFT / 2; t += blockDim.x) {
            const int grp = t / half, pos = t % half;
            const int i0 = grp * 2 * half + pos, i1 = i0 + half;
            float sn, cs;
            sincospif(-static_cast<float>(pos) / half, &sn, &cs);                       // e^{-i pi pos / half}
            const float xr = re[i1] * cs - im[i1] * sn, xi = re[i1] * sn + im[i1] * cs;
            const float ar = re[i0], ai = im[i0];
            re[i0] = ar + xr; im[i0] = ai + xi;
            re[i1] = ar - xr; im[i1] = ai - xi;
        }
        __syncthreads();
    }
    for (int k = threadIdx.x; k <= NFFT / 2; k += blockDim.x) mag[k] = sqrtf(re[k] * re[k] + im[k] * im[k] + 1e-9f);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    for (int m = warp; m < n_mels; m += nw) {
        const float* b = basis + static_cast<size_t>(m) * (NFFT / 2 + 1);
        float acc = 0.f;
        for (int k = lane; k <= NFFT / 2; k += 32) acc = fmaf(b[k], mag[k], acc);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (lane == 0) out[static_cast<size_t>(m) * n_frames + f] = (logf(fmaxf(acc, 1e-5f)) - mean) * inv_std;
    }
}

}  // namespace tvb
