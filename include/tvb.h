/* tvb.h — C ABI of the B200-native token-Voicebox mel generator (libtvb.so).
 *
 * Drop-in boundary for ONE path of ishine/usdm: the flow-matching ODE sampler of src/decoder
 * (reference files: src/decoder/voicebox/model/voicebox.py, src/decoder/voicebox/model/networks.py).
 * Plain pointers and sizes only; no torch / C++ types cross this boundary.  Every function returns
 * 0 on success or a negative tvb_status; tvb_last_error() gives a human-readable message for the
 * calling thread's most recent failure.  No exception crosses the ABI.
 *
 * Conventions
 *  - "dev" pointers are CUDA device pointers on the handle's device; "host" pointers are ordinary memory.
 *  - `stream` is a cudaStream_t passed as void* (torch: torch.cuda.current_stream().cuda_stream).  All work is
 *    enqueued on it; no call synchronises the device except tvb_load_weight/tvb_finalize_weights/tvb_destroy.
 *  - A handle is not thread-safe.  It owns its packed weights and an internal workspace; inputs are borrowed
 *    for the duration of the enqueued work and never written.
 *  - Tensor layouts are the reference's: x int64 [B,T]; y/z/cond/noise/out fp32 [B,n_feats,T] contiguous.
 *  - Padded frames (tau >= lengths[b]) of every output are exactly 0, as in the reference (networks.py:372).
 */
#ifndef TVB_H_
#define TVB_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvb_handle tvb_handle;

typedef enum tvb_status {
    TVB_OK = 0,
    TVB_ERR_INVALID = -1,     /* bad argument / unsupported configuration (Python wrapper raises ValueError) */
    TVB_ERR_CUDA = -2,        /* CUDA runtime / driver failure            (RuntimeError) */
    TVB_ERR_STATE = -3,       /* call order violated, e.g. weights not finalized (RuntimeError) */
    TVB_ERR_NOMEM = -4
} tvb_status;

typedef enum tvb_precision {
    TVB_PRECISION_BF16 = 0,           /* tcgen05 bf16 tensor-core path, fp32 accumulate / residual / softmax / LN */
    TVB_PRECISION_FP32_VALIDATE = 1   /* all-fp32 CUDA-core path used to validate against the fp32 reference   */
} tvb_precision;

typedef enum tvb_solver { TVB_SOLVER_EULER = 0, TVB_SOLVER_HEUN = 1 } tvb_solver;

/* Mirrors the keyword arguments of reference Voicebox.__init__ (voicebox.py:154-156) that shape the network. */
typedef struct tvb_config {
    int32_t n_feats;              /* 80   */
    int32_t n_tokens;             /* 10000 (the estimator embeds n_tokens + 1 ids; id n_tokens = unconditional) */
    int32_t embedding_dim;        /* 1280 */
    int32_t hidden_size;          /* 1024 */
    int32_t intermediate_size;    /* 4096 */
    int32_t num_attention_heads;  /* 16   */
    int32_t num_hidden_layers;    /* 24   */
    int32_t convpos_width;        /* 31   */
    int32_t convpos_groups;       /* 16   */
    int32_t convpos_depth;        /* 2    */
    float sigma_min;              /* 1e-4 */
} tvb_config;

/* Replaces: Voicebox.__init__ (voicebox.py:153-176) -> Transformer.__init__ (networks.py:270-300). */
int tvb_create(const tvb_config* cfg, int device, tvb_handle** out);
int tvb_destroy(tvb_handle* h);

/* Replaces: nn.Module.load_state_dict / PyTorchModelHubMixin.from_pretrained (voicebox.py:10,153).
 * `ref_key` is the reference state-dict key (e.g. "estimator.layers.3.attention.q_proj.weight",
 * "estimator.pos_conv_embeds.0.conv.parametrizations.weight.original1"); data is fp32, contiguous, host or device. */
int tvb_load_weight(tvb_handle* h, const char* ref_key, const void* data, const int64_t* shape, int32_t ndim);
/* Folds weight-norm and the q scaling, builds the unit-embedding->hidden gather table, packs bf16 copies. */
int tvb_finalize_weights(tvb_handle* h);

int tvb_set_precision(tvb_handle* h, int32_t precision);
int tvb_get_precision(const tvb_handle* h);
/* Ablation / cross-check switches of the bf16 path: "attn_tc" and "conv_tc" (1 = tcgen05 kernel, default;
 * 0 = the CUDA-core kernel of the same op; both are GPU kernels of this library, there is no CPU path),
 * "gemm_2cta" (1 = CTA-pair tcgen05 GEMM, default; 0 = single-CTA tcgen05 GEMM), "gemm_mc" (1 = two CTA pairs per cluster
 * sharing the W tile by TMA multicast; 0 = independent pairs, default), "gemm_bn" (1 = the CTA-pair GEMM switches to 128-column tiles when
 * the 256-column tiling would leave CTA pairs idle -- small batches, default; 0 = always 256-column tiles), "attn_v" (10 = attn10_tc.cuh, default; 9 = attn9_tc.cuh, 8 = attn8_tc.cuh (running-maximum softmax, any logit range), 7 .. 4 = attnN_tc.cuh,
 * 3 = attn_tc.cuh), "attn_win" (v9+: 1 = skip key blocks whose probabilities are provably below 2^-40 of the row sum, default; 0 = process every
 * block), "attn_sort" (v10: cost-sorted item order, 0 never / 1 for long or ragged plans, default / 2 always), "qkv_norm" (1 = the QKV GEMM
 * epilogue emits the row-block norms attention v9+ bounds its logits with, default; 0 = a separate kernel), "attn_poly" (attention v4+: 0 = every
 * exp2 on MUFU; 2/3/4 = one element pair in n on the FMA-pipe polynomial, default 4), "resid_l2" (residual GEMM epilogues pull their next
 * chunks into the L2: 0 off, default -- measured slower; 1 / 2), "ln_fuse" (2 = LayerNorm computed inside the residual GEMM epilogue by 8-CTA clusters; 1 = LayerNorm
 * kernel emits bf16 + row statistics and the residual GEMM epilogue re-normalises the fp32 stream; 0 = separate fp32 LayerNorm output), "ln1_fold" (with ln_fuse = 1: 1 = the LayerNorm between attention and FFN is folded into the FFN-in GEMM, default;
 * 0 = it runs as a kernel), "zigzag" (1 = producer / consumer kernels walk the rows in opposite directions for L2 reuse,
 * default), "graph" (1 = tvb_solve replays the estimator pass of small batches as a CUDA graph, default; 0 = direct launches),
 * "pdl" (1 = GEMM / attention / LayerNorm kernels use programmatic dependent launch, default), "attn_pace" (attention v8 MUFU
 * pacing ablation, default 0), "frame_tc" / "out_tc" (1 = the per-pass frame projection of the ODE state / the output
 * projection run as tcgen05 GEMMs, default; 0 = the CUDA-core kernels), "conv_halo" (1 = positional conv fetches each
 * activation tile once with its halo, default; 0 = one fetch per tap), "a_l2" (pair GEMM producers prefetch activation k-blocks this far
 * ahead into the L2; 0 = off, default -- measured slower), "epi_direct" (fp32-output GEMM epilogues straight between registers and global memory instead of staged
 * through shared memory: 0 = never, default; 1 = FFN-out only; 2 = all -- same bits, measured no faster), "gemm_lab" (TIMING ablations of the attention-out epilogue: bit 0 no residual
 * load, bit 1 no fp32 store, bit 2 no bf16 copy; any non-zero value gives WRONG results, default 0), and "profile" (see tvb_profile_read). */
int tvb_set_option(tvb_handle* h, const char* name, int32_t value);

/* Replaces: Transformer.forward(x, y, cond, t, lengths) (networks.py:302-374) — the per-step velocity estimator.
 * x_dev int64 [B,T]; y_dev, cond_dev fp32 [B,n_feats,T]; t_host fp32 [B]; lengths_host int64 [B];
 * out_v_dev fp32 [B,n_feats,T].  T <= 32767 (every compute call; TVB_ERR_INVALID beyond). */
int tvb_estimator_forward(tvb_handle* h, const int64_t* x_dev, const float* y_dev, const float* cond_dev,
                          const float* t_host, const int64_t* lengths_host, int32_t B, int32_t T,
                          float* out_v_dev, void* stream);

/* Replaces: CFM.sample (voicebox.py:51-72): classifier-free-guidance doubling [uncond ; cond] in one pass and
 * the combine v_c + g (v_c - v_u).  t is uniform over the batch (as in the reference).  gradient_scale <= 0
 * runs the conditional branch only; speech_prompt == 0 zeroes cond (voicebox.py:56-58). */
int tvb_sample_cfg(tvb_handle* h, const int64_t* x_dev, const float* z_dev, const float* cond_dev,
                   const int64_t* lengths_host, int32_t B, int32_t T, float t, float gradient_scale,
                   int32_t speech_prompt, float* out_v_dev, void* stream);

/* Replaces: CFM.solve_euler / CFM.solve_heun (voicebox.py:74-138), i.e. the whole ODE loop of CFM.generate.
 *  t_span_host      fp32 [n_steps + 1]  (generate(): linspace(0, 1, n + 1), voicebox.py:145)
 *  z0_dev           fp32 [B,n_feats,T]  initial noise (generate(): randn_like(cond), voicebox.py:142)
 *  prompt_noise_dev fp32 [n_noise,B,n_feats,T] or NULL: the fresh randn_like(cond) draws the reference makes after
 *                   each (sub)step when speech_prompt (voicebox.py:90,116,127), in call order;
 *                   n_noise = n_steps (Euler) or 2*n_steps-1 (Heun).  Required iff speech_prompt && prompt_len > 0.
 *  prompt_len       prompt_lengths[0] of the reference (batch-uniform, voicebox.py:92)
 *  out_mel_dev      fp32 [B,n_feats,T]  (the reference returns the full length, prompt region included)
 *  step_v_dev       optional fp32 [n_evals,B,n_feats,T]: guided velocity of every estimator evaluation
 *                   (parity hook for the per-step velocity test), or NULL. */
int tvb_solve(tvb_handle* h, int32_t solver, const int64_t* x_dev, const float* z0_dev, const float* cond_dev,
              const int64_t* lengths_host, int32_t B, int32_t T, int32_t prompt_len, int32_t n_steps,
              const float* t_span_host, float gradient_scale, int32_t speech_prompt,
              const float* prompt_noise_dev, float* out_mel_dev, float* step_v_dev, void* stream);

/* Data errors that only a kernel can see.  The reference's nn.Embedding asserts device-side when a unit id is outside its
 * table (networks.py:305); here the lookup is kept in range and a sticky per-handle flag is raised instead.  Because no compute
 * call synchronises the device, the flag is reported by the NEXT tvb_estimator_forward / tvb_sample_cfg / tvb_solve on the
 * handle, or by this function (call it after synchronising the stream to learn about the call just made):
 * TVB_ERR_INVALID once, then cleared; TVB_OK otherwise.  Never synchronises.
 * The same channel carries two weight-dependent conditions of the bf16 fast paths (one report each; tvb_last_error() names the option
 * that selects the safe sibling): attention logits beyond the range of the default attention kernel ("attn_v" = 8), and rows whose
 * mean exceeds 8 standard deviations in the LayerNorm folded into the FFN-in GEMM ("ln1_fold" = 0).  A caller should synchronise and
 * call this once after its first call with new weights and switch the option if told to (usdm_b200/voicebox.py: Voicebox._guarded). */
int tvb_async_status(tvb_handle* h);

/* ---- decoder glue either side of the sampler; stateless, no handle ----
 * Replaces: process_unit (util/model_util.py:50-54): repeat every unit `rep` (= sampling_rate // 50 = 441) times, cut into
 * windows of `hop` (= 256) samples, majority vote per window (ties -> smallest value, as torch.mode).
 * units_dev int64 [n_units] -> out_dev int64 [n_out], n_out = n_units * rep / hop. */
int tvb_process_unit(int device, const int64_t* units_dev, int64_t n_units, int32_t rep, int32_t hop, int64_t* out_dev, int64_t n_out,
                     void* stream);
/* Replaces: get_mel (util/model_util.py:24-38) -> mel_spectrogram (vocoder/meldataset.py:55-78) + the (mel - mean) / std
 * normalisation of model_util.py:82.  wav_dev fp32 [n_samples] already clamped to [-1, 1]; reflect padding (n_fft - hop) / 2,
 * periodic Hann window of n_fft, hop `hop`, center = False; mel_basis_dev fp32 [n_mels, n_fft / 2 + 1];
 * out_dev fp32 [n_mels, n_frames] with n_frames = (n_samples + n_fft - hop - n_fft) / hop + 1. */
int tvb_mel_spectrogram(int device, const float* wav_dev, int64_t n_samples, int32_t n_fft, int32_t hop, int32_t n_mels,
                        const float* mel_basis_dev, float mean, float std, float* out_dev, int64_t n_frames, void* stream);

/* Per-kernel timing (bench.py's roofline leg).  With tvb_set_option(h, "profile", 1) every kernel launch is bracketed by
 * CUDA events on the launching stream; tvb_profile_read synchronises the device, sums the elapsed milliseconds and launch
 * counts by category into two arrays of 16 entries and resets the log.
 * Categories: 0 linear-layer GEMMs (total), 1 attention, 2 positional conv, 3 LayerNorm, 4 frame projection / time token,
 * 5 output projection + CFG/ODE update, 6-7 unused, 8-12 the GEMMs by call site (QKV, attention out, FFN in, FFN out,
 * skip merge; these also count into 0), 13-15 unused. */
int tvb_profile_read(tvb_handle* h, double* ms_by_cat, int64_t* launches_by_cat);

/* Number of this library's kernels launched by the handle since creation (bench.py's gpu_launches). */
int64_t tvb_kernel_launches(const tvb_handle* h);
/* Bytes of device memory currently held by the handle (weights + workspace). */
int64_t tvb_device_bytes(const tvb_handle* h);

/* Kernel-level test hooks (used by tests/ only): a single tcgen05 GEMM  C = A * W^T + bias with a chosen epilogue.
 * a_dev bf16 [M,K], w_dev bf16 [N,K], bias_dev fp32 [N] or NULL, resid_dev fp32 [M,N] or NULL,
 * out_f32_dev fp32 [M,N] or NULL, out_bf16_dev bf16 [M,N] or NULL.  epilogue: 0 bf16, 1 gelu->bf16, 2 +resid->f32,
 * 3 f32+bf16, 4 f32; OR 0x100 to run the CTA-pair (cta_group::2) kernel (then M % 256 == 0), OR 0x200 for the
 * cluster-of-4 multicast kernel (then M % 512 == 0), OR 0x400 for the CTA-pair kernel with 128-column tiles (M % 256 == 0,
 * N % 128 == 0).
 * M % 128 == 0, N % 256 == 0, K % 64 == 0. */
int tvb_debug_gemm_bf16(int device, const void* a_dev, const void* w_dev, const float* bias_dev, const float* resid_dev,
                        float* out_f32_dev, void* out_bf16_dev, int32_t M, int32_t N, int32_t K, int32_t epilogue,
                        void* stream);

/* Host-logic test hook (no device needed): the tile width (256 or 128 columns) the CTA-pair GEMM uses for an output of
 * `rows` x N on a device with `num_sms` SMs -- 128 only when the 256-column tiling does not fill one wave of CTA pairs and the
 * narrow tiles need fewer (waves x width).  rows % 256 == 0, N % 128 == 0; negative = error. */
int tvb_debug_pair_gemm_bn(int32_t num_sms, int32_t rows, int32_t N);

/* Kernel-level test hook (tests/ only): the tcgen05 attention kernel alone on packed [rows, 3H] bf16 q|k|v rows (q pre-scaled
 * by head_dim^-0.5 * log2(e)); sequence s occupies rows [seq_start[s], seq_start[s] + seq_len[s]).  ctx_dev bf16 [rows, H].
 * version: 10 = attn10_tc.cuh (100: the same without the ALiBi window), 9 = attn9_tc.cuh (90: without the window), 8 = attn8_tc.cuh (persistent + two issue warps), 7 = attn7_tc.cuh (two issue warps), 6 = attn6_tc.cuh (persistent), 5 = attn5_tc.cuh, 4 = attn4_tc.cuh, 3 = attn_tc.cuh (P in tensor memory), 2 = attn_tc.cuh (P through shared memory).  Synchronises. */
int tvb_debug_attention(int device, const void* qkv_dev, void* ctx_dev, const int32_t* seq_start_host, const int32_t* seq_len_host,
                        int32_t nseq, int32_t rows, int32_t H, int32_t nheads, const float* slopes_dev, int32_t version,
                        int32_t poly, void* stream);

/* Test hook: attention v4 launches write clock64() stamps of their first 16 CTAs into this device buffer (2 x 16 x 256 int64;
 * layout in attn4_tc.cuh); NULL (default) disables. */
int tvb_debug_set_trace(void* trace_dev);

const char* tvb_last_error(void);
const char* tvb_version(void);

#ifdef __cplusplus
}
#endif
#endif /* TVB_H_ */
