/* tvb_vocoder.h — C ABI of the B200-native BigVGAN v1 vocoder forward (same libtvb.so).
 *
 * Boundary for the step AFTER the mel generator in ishine/usdm's decoder: `vocoder.forward(y_dec)`
 * (src/decoder/voicebox/util/model_util.py:104) on the generator of src/decoder/voicebox/vocoder/models.py:132-222.
 * Same conventions as tvb.h: plain pointers and sizes, 0 or a negative tvb_status, tvb_last_error() for the message,
 * "dev" = CUDA device pointer on the handle's device, `stream` = cudaStream_t as void*, a handle is not thread-safe.
 */
#ifndef TVB_VOCODER_H_
#define TVB_VOCODER_H_

#include <stdint.h>

#include "tvb.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tvb_voc_handle tvb_voc_handle;

#define TVB_VOC_MAX_UPSAMPLES 8
#define TVB_VOC_MAX_KERNELS 4

/* Mirrors the fields of the BigVGAN hyper-parameter dict `h` the constructor reads (models.py:145-186). Only resblock "1"
 * (AMPBlock1, three dilations per kernel size) is supported: it is what nvidia/bigvgan_22khz_80band ships (model_util.py:64-67). */
typedef struct tvb_voc_config {
    int32_t num_mels;                                        /* 80   */
    int32_t upsample_initial_channel;                        /* 1536 */
    int32_t num_upsamples;                                   /* 6    */
    int32_t upsample_rates[TVB_VOC_MAX_UPSAMPLES];           /* 4 4 2 2 2 2 */
    int32_t upsample_kernel_sizes[TVB_VOC_MAX_UPSAMPLES];    /* 8 8 4 4 4 4  (must equal 2 * rate) */
    int32_t num_kernels;                                     /* 3    */
    int32_t resblock_kernel_sizes[TVB_VOC_MAX_KERNELS];      /* 3 7 11 */
    int32_t resblock_dilation_sizes[TVB_VOC_MAX_KERNELS][3]; /* 1 3 5 each */
    int32_t snake_logscale;                                  /* 1    */
    int32_t snakebeta;                                       /* 1 = "snakebeta" (alpha and beta), 0 = "snake" (alpha only) */
} tvb_voc_config;

/* Replaces: BigVGAN.__init__ (models.py:143-190). */
int tvb_voc_create(const tvb_voc_config* cfg, int device, tvb_voc_handle** out);
void tvb_voc_destroy(tvb_voc_handle* h);

/* Replaces: load_state_dict / BigVGAN._from_pretrained (models.py:244-327).  `key` is the reference state-dict key
 * ("conv_pre.weight", "ups.0.0.weight_g", "resblocks.3.convs1.1.bias", "resblocks.3.activations.2.act.alpha",
 * "activation_post.upsample.filter", ...); data is fp32, host or device.  Both layouts are accepted: plain `weight`
 * (after remove_weight_norm(), model_util.py:68) or `weight_g` + `weight_v` (folded once in tvb_voc_finalize_weights). */
int tvb_voc_load_weight(tvb_voc_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim);
int tvb_voc_finalize_weights(tvb_voc_handle* h);

/* tvb_precision: bf16 tcgen05 convolutions (default) or the all-fp32 CUDA-core validation path. */
int tvb_voc_set_precision(tvb_voc_handle* h, int32_t precision);

/* Ablation switches: "halo" = 1 (default: activation tile fetched once per tile, taps are row-shifted shared-memory descriptors),
 * 0 (one activation fetch per tap). */
int tvb_voc_set_option(tvb_voc_handle* h, const char* name, int32_t value);

/* Replaces: BigVGAN.forward (models.py:192-222).  mel_dev fp32 [B, num_mels, T] (reference layout) ->
 * wav_dev fp32 [B, 1, T * prod(upsample_rates)].  Enqueues on `stream`; synchronises only when the workspace must grow. */
int tvb_voc_forward(tvb_voc_handle* h, const float* mel_dev, int32_t B, int32_t T, float* wav_dev, void* stream);

/* Test hook: one Activation1d (alias_free_torch/act.py:21-27) of the loaded model on x_dev fp32 [B, C, T] -> out_dev [B, C, T].
 * `name` is the module prefix ("resblocks.0.activations.0", "activation_post"). */
int tvb_voc_debug_activation(tvb_voc_handle* h, const char* name, const float* x_dev, int32_t B, int32_t T, float* out_dev, void* stream);

/* Measurement: kernels launched so far; per-category device milliseconds of the last forward when profiling is on
 * (ms[0] = convolutions, ms[1] = activations, ms[2] = layout / conv_post, then when n allows ms[3 + i] = convolutions of
 * upsampling stage i and ms[3 + TVB_VOC_MAX_UPSAMPLES + i] = its activations); device bytes held (weights + workspace). */
int64_t tvb_voc_kernel_launches(const tvb_voc_handle* h);
int tvb_voc_set_profile(tvb_voc_handle* h, int32_t on);
int tvb_voc_profile_read(tvb_voc_handle* h, float* ms, int32_t n);
int64_t tvb_voc_device_bytes(const tvb_voc_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* TVB_VOCODER_H_ */
