// Host side of the BigVGAN vocoder path (include/tvb_vocoder.h).  Included at the end of tvb.cu: shares its error plumbing,
// the TMA descriptor encoder and RawTensor.  Orchestration only -- every sample is computed by the kernels of vocoder.cuh.
// Reference walk: src/decoder/voicebox/vocoder/models.py:192-222 (BigVGAN.forward), :70-79 (AMPBlock1.forward).

struct VocConv {              // one Conv1d / ConvTranspose1d as a shifted-row GEMM (see vocoder.cuh)
    int Cin = 0, N = 0, taps = 0, dil = 1, ci_pad = 0, n_pad = 0, bn = 0;
    bf16* w2 = nullptr;       // [n_pad][taps][ci_pad] bf16
    float* wt = nullptr;      // [taps][Cin][N] fp32
    float* bias = nullptr;    // [N]
    CUtensorMap tmW;
};
struct VocActW {
    int C = 0;
    float2* consts = nullptr;
    VocFilt F;
};

struct tvb_voc_handle {
    tvb_voc_config cfg;
    int device = 0;
    int precision = TVB_PRECISION_BF16;
    bool finalized = false;
    int num_sms = 148;
    int64_t launches = 0;
    size_t dev_bytes = 0, ws_bytes = 0;
    std::map<std::string, RawTensor> raw;
    std::vector<void*> raw_owned, owned, ws_owned;
    VocConv conv_pre;
    std::vector<VocConv> ups, convs1, convs2;     // convs*[n * 3 + m], n = stage * num_kernels + j
    std::map<std::string, VocActW> acts;          // by module prefix
    float* post_w = nullptr;                      // conv_post weight [C][7]
    float* post_b = nullptr;
    int post_C = 0;
    // workspace (time-major activations)
    size_t ws_elems = 0;
    float *X = nullptr, *XS = nullptr, *U = nullptr, *R = nullptr, *Hb = nullptr, *AF = nullptr;
    bf16* A16 = nullptr;
    // profiling
    int opt_act_vec2 = 1;                         // 1 = two channels per thread with packed fp32x2 math in the activation kernel (bf16 path, even C)
    int opt_halo = 1;                             // 1 = halo kernel (activation tile fetched once per tile), 0 = one fetch per tap
    bool profile = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_cat;
    size_t ev_used = 0;
    int cur_stage = -1;                           // upsampling stage being enqueued (profiling split)
    float prof_ms[3 + 2 * TVB_VOC_MAX_UPSAMPLES] = {};   // [0..2] conv / activation / other, then conv per stage, activation per stage
};

struct VocProf {
    tvb_voc_handle* h;
    cudaStream_t st;
    VocProf(tvb_voc_handle* h_, cudaStream_t st_, int cat) : h(h_), st(st_) {
        ++h->launches;
        if (!h->profile) return;
        while (h->ev_pool.size() < h->ev_used + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            h->ev_pool.push_back(e);
        }
        h->ev_cat.push_back(cat | ((h->cur_stage + 1) << 8));
        cudaEventRecord(h->ev_pool[h->ev_used], st);
    }
    ~VocProf() {
        if (!h->profile) return;
        cudaEventRecord(h->ev_pool[h->ev_used + 1], st);
        h->ev_used += 2;
    }
};

static int voc_alloc(tvb_voc_handle* h, std::vector<void*>& bag, void** p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(TVB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    bag.push_back(*p);
    if (&bag == &h->ws_owned) h->ws_bytes += bytes;
    else if (&bag == &h->owned) h->dev_bytes += bytes;
    return TVB_OK;
}
template <typename T>
static int voc_alloc_t(tvb_voc_handle* h, std::vector<void*>& bag, T** p, size_t n) {
    return voc_alloc(h, bag, reinterpret_cast<void**>(p), n * sizeof(T));
}
static void voc_free_bag(std::vector<void*>& bag) {
    for (void* p : bag) cudaFree(p);
    bag.clear();
}

// 3-D (channel, time, batch) bf16 map over a time-major activation buffer [B][T][ld]; box 64 x 128 x 1, 128-byte swizzle.
// Rows outside [0, T) and channels outside [0, C) read as zero: this IS the convolution's zero padding.
static int make_tmap_act3(CUtensorMap* m, const void* ptr, int C, int ld, int T, int B, int box_rows = 128) {
    TRY(resolve_driver());
    cuuint64_t dims[3] = {static_cast<cuuint64_t>(C), static_cast<cuuint64_t>(T), static_cast<cuuint64_t>(B)};
    cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * 2, static_cast<cuuint64_t>(T) * ld * 2};
    cuuint32_t box[3] = {64, static_cast<cuuint32_t>(box_rows), 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TVB_ERR_CUDA, "cuTensorMapEncodeTiled (3-D activation map) failed (%d) C=%d ld=%d T=%d B=%d", (int)r, C, ld, T, B);
    return TVB_OK;
}

static inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

extern "C" int tvb_voc_create(const tvb_voc_config* cfg, int device, tvb_voc_handle** out) {
    if (!cfg || !out) return fail(TVB_ERR_INVALID, "null argument to tvb_voc_create");
    const tvb_voc_config& c = *cfg;
    if (c.num_mels <= 0 || c.upsample_initial_channel <= 0 || c.num_upsamples <= 0 || c.num_upsamples > TVB_VOC_MAX_UPSAMPLES ||
        c.num_kernels <= 0 || c.num_kernels > TVB_VOC_MAX_KERNELS)
        return fail(TVB_ERR_INVALID, "bad vocoder configuration");
    if ((c.upsample_initial_channel >> c.num_upsamples) < 1 || (c.upsample_initial_channel & ((1 << c.num_upsamples) - 1)))
        return fail(TVB_ERR_INVALID, "upsample_initial_channel must be divisible by 2^num_upsamples");
    for (int i = 0; i < c.num_upsamples; ++i)
        if (c.upsample_rates[i] < 1 || c.upsample_kernel_sizes[i] != 2 * c.upsample_rates[i] || (c.upsample_rates[i] & 1))
            return fail(TVB_ERR_INVALID, "upsampler %d: kernel size must be 2 * rate with an even rate (the shipped checkpoint's 8/4 and 4/2)", i);
    for (int j = 0; j < c.num_kernels; ++j) {
        if (c.resblock_kernel_sizes[j] < 1 || !(c.resblock_kernel_sizes[j] & 1)) return fail(TVB_ERR_INVALID, "resblock kernel sizes must be odd");
        for (int m = 0; m < 3; ++m)
            if (c.resblock_dilation_sizes[j][m] < 1) return fail(TVB_ERR_INVALID, "bad dilation");
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(TVB_ERR_CUDA, "no CUDA device: libtvb has no CPU path");
    if (device < 0 || device >= ndev) return fail(TVB_ERR_INVALID, "device %d out of range", device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(TVB_ERR_CUDA, "device %d is sm_%d%d; libtvb is built for sm_100a only", device, prop.major, prop.minor);
    auto* h = new tvb_voc_handle();
    h->cfg = c;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    *out = h;
    return TVB_OK;
}

extern "C" void tvb_voc_destroy(tvb_voc_handle* h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    voc_free_bag(h->raw_owned);
    voc_free_bag(h->owned);
    voc_free_bag(h->ws_owned);
    for (auto e : h->ev_pool) cudaEventDestroy(e);
    delete h;
}

extern "C" int tvb_voc_load_weight(tvb_voc_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim) {
    if (!h || !key || !data || !shape || ndim < 1 || ndim > 4) return fail(TVB_ERR_INVALID, "bad argument to tvb_voc_load_weight");
    if (h->finalized) return fail(TVB_ERR_STATE, "weights already finalized");
    CK(cudaSetDevice(h->device));
    RawTensor t;
    t.n = 1;
    for (int i = 0; i < ndim; ++i) {
        if (shape[i] <= 0) return fail(TVB_ERR_INVALID, "bad shape for %s", key);
        t.shape.push_back(shape[i]);
        t.n *= static_cast<size_t>(shape[i]);
    }
    auto it = h->raw.find(key);
    if (it != h->raw.end()) {
        if (it->second.n != t.n) return fail(TVB_ERR_INVALID, "weight %s reloaded with a different size", key);
        t.p = it->second.p;
    } else {
        TRY(voc_alloc_t(h, h->raw_owned, &t.p, t.n));
    }
    CK(cudaMemcpy(t.p, data, t.n * sizeof(float), cudaMemcpyDefault));
    h->raw[key] = t;
    return TVB_OK;
}

static int voc_need(tvb_voc_handle* h, const std::string& key, std::vector<int64_t> shape, const float** out) {
    auto it = h->raw.find(key);
    if (it == h->raw.end()) return fail(TVB_ERR_STATE, "missing weight %s", key.c_str());
    if (it->second.shape != shape) {
        std::string s;
        for (auto d : it->second.shape) s += std::to_string(d) + ",";
        return fail(TVB_ERR_INVALID, "weight %s has shape [%s], unexpected", key.c_str(), s.c_str());
    }
    *out = it->second.p;
    return TVB_OK;
}

// plain `weight`, or weight_g * weight_v / ||weight_v|| (norm over all dims but 0)
static int voc_weight(tvb_voc_handle* h, const std::string& name, std::vector<int64_t> shape, const float** out) {
    if (h->raw.count(name + ".weight")) return voc_need(h, name + ".weight", shape, out);
    const float *v = nullptr, *g = nullptr;
    TRY(voc_need(h, name + ".weight_v", shape, &v));
    TRY(voc_need(h, name + ".weight_g", {shape[0], 1, 1}, &g));
    float* w = nullptr;
    const size_t inner = static_cast<size_t>(shape[1] * shape[2]);
    TRY(voc_alloc_t(h, h->raw_owned, &w, static_cast<size_t>(shape[0]) * inner));
    k_voc_weight_norm<<<static_cast<unsigned>(shape[0]), 256>>>(v, g, w, static_cast<int>(inner));
    CKL();
    *out = w;
    return TVB_OK;
}

static void voc_pick_tile(VocConv* L) {
    const int n16 = round_up(L->N, 16);
    if (n16 <= 256) { L->bn = n16; L->n_pad = n16; return; }
    for (int bn : {256, 192, 128})
        if (L->N % bn == 0) { L->bn = bn; L->n_pad = L->N; return; }
    L->bn = 256;
    L->n_pad = round_up(L->N, 256);
}

static int voc_make_conv(tvb_voc_handle* h, const std::string& name, int Cout, int Cin, int k, int dil, VocConv* L) {
    const float *w = nullptr, *b = nullptr;
    TRY(voc_weight(h, name, {Cout, Cin, k}, &w));
    TRY(voc_need(h, name + ".bias", {Cout}, &b));
    L->Cin = Cin; L->N = Cout; L->taps = k; L->dil = dil; L->ci_pad = round_up(Cin, 64);
    voc_pick_tile(L);
    const size_t n2 = static_cast<size_t>(L->n_pad) * k * L->ci_pad;
    TRY(voc_alloc_t(h, h->owned, &L->w2, n2));
    TRY(voc_alloc_t(h, h->owned, &L->wt, static_cast<size_t>(k) * Cin * Cout));
    TRY(voc_alloc_t(h, h->owned, &L->bias, static_cast<size_t>(Cout)));
    k_voc_pack_conv<<<grid_for(n2), 256>>>(w, Cout, Cin, k, L->n_pad, L->ci_pad, L->w2, L->wt);
    CKL();
    CK(cudaMemcpy(L->bias, b, Cout * sizeof(float), cudaMemcpyDeviceToDevice));
    return make_tmap(&L->tmW, L->w2, L->n_pad, static_cast<int64_t>(k) * L->ci_pad, static_cast<int64_t>(k) * L->ci_pad, L->bn);
}

static int voc_make_convT(tvb_voc_handle* h, const std::string& name, int Cin, int Cout, int k, int u, VocConv* L) {
    const float *w = nullptr, *b = nullptr;
    TRY(voc_weight(h, name, {Cin, Cout, k}, &w));
    TRY(voc_need(h, name + ".bias", {Cout}, &b));
    L->Cin = Cin; L->N = u * Cout; L->taps = 3; L->dil = 1; L->ci_pad = round_up(Cin, 64);
    voc_pick_tile(L);
    const size_t n2 = static_cast<size_t>(L->n_pad) * 3 * L->ci_pad;
    TRY(voc_alloc_t(h, h->owned, &L->w2, n2));
    TRY(voc_alloc_t(h, h->owned, &L->wt, static_cast<size_t>(3) * Cin * L->N));
    TRY(voc_alloc_t(h, h->owned, &L->bias, static_cast<size_t>(L->N)));
    k_voc_pack_convT<<<grid_for(n2), 256>>>(w, Cin, Cout, k, u, L->n_pad, L->ci_pad, L->w2, L->wt);
    CKL();
    k_voc_tile_bias<<<grid_for(L->N), 256>>>(b, Cout, L->N, L->bias);
    CKL();
    return make_tmap(&L->tmW, L->w2, L->n_pad, 3ll * L->ci_pad, 3ll * L->ci_pad, L->bn);
}

// torch.kaiser_window(12, beta, periodic=False) * sinc low-pass, normalised (alias_free_torch/filter.py:24-50); used only when the
// state dict does not carry the registered `filter` buffers.
static void voc_default_filter(float* f) {
    const int k = 12, half = 6;
    const double cutoff = 0.25, half_width = 0.3;
    const double delta_f = 4 * half_width, A = 2.285 * (half - 1) * M_PI * delta_f + 7.95;
    const double beta = A > 50. ? 0.1102 * (A - 8.7) : (A >= 21. ? 0.5842 * std::pow(A - 21., 0.4) + 0.07886 * (A - 21.) : 0.);
    auto i0 = [](double x) {
        double s = 1., term = 1.;
        for (int n = 1; n < 64; ++n) { term *= (x / (2. * n)) * (x / (2. * n)); s += term; }
        return s;
    };
    double sum = 0., v[12];
    for (int n = 0; n < k; ++n) {
        const double r = (n - (k - 1) / 2.) / ((k - 1) / 2.);
        const double win = i0(beta * std::sqrt(std::max(0., 1. - r * r))) / i0(beta);
        const double t = (n - half) + 0.5, x = 2 * cutoff * t;
        const double sinc = x == 0. ? 1. : std::sin(M_PI * x) / (M_PI * x);
        v[n] = 2 * cutoff * win * sinc;
        sum += v[n];
    }
    for (int n = 0; n < k; ++n) f[n] = static_cast<float>(v[n] / sum);
}

static int voc_filter(tvb_voc_handle* h, const std::string& key, float* f) {
    auto it = h->raw.find(key);
    if (it == h->raw.end()) { voc_default_filter(f); return TVB_OK; }
    if (it->second.n != 12) return fail(TVB_ERR_INVALID, "%s: only the 12-tap x2 filters of Activation1d's defaults are supported", key.c_str());
    CK(cudaMemcpy(f, it->second.p, 12 * sizeof(float), cudaMemcpyDeviceToHost));
    return TVB_OK;
}

static int voc_make_act(tvb_voc_handle* h, const std::string& name, int C) {
    VocActW a;
    a.C = C;
    const float *alpha = nullptr, *beta = nullptr;
    TRY(voc_need(h, name + ".act.alpha", {C}, &alpha));
    if (h->cfg.snakebeta) TRY(voc_need(h, name + ".act.beta", {C}, &beta));
    else beta = alpha;
    TRY(voc_alloc_t(h, h->owned, &a.consts, static_cast<size_t>(C)));
    k_voc_snake_consts<<<(C + 127) / 128, 128>>>(alpha, beta, C, h->cfg.snake_logscale, a.consts);
    CKL();
    TRY(voc_filter(h, name + ".upsample.filter", a.F.up));
    TRY(voc_filter(h, name + ".downsample.lowpass.filter", a.F.dn));
    a.F.dn_sum = 0.f;
    for (int i = 0; i < 12; ++i) {
        a.F.up[i] *= 2.f;                 // UpSample1d multiplies by its ratio (resample.py:27); exact in fp32
        a.F.dn_sum += a.F.dn[i];
    }
    h->acts[name] = a;
    return TVB_OK;
}

extern "C" int tvb_voc_finalize_weights(tvb_voc_handle* h) {
    if (!h) return fail(TVB_ERR_INVALID, "null handle");
    if (h->finalized) return fail(TVB_ERR_STATE, "weights already finalized");
    CK(cudaSetDevice(h->device));
    const tvb_voc_config& c = h->cfg;
    const int C0 = c.upsample_initial_channel;
    TRY(voc_make_conv(h, "conv_pre", C0, c.num_mels, 7, 1, &h->conv_pre));
    h->ups.resize(c.num_upsamples);
    h->convs1.resize(static_cast<size_t>(c.num_upsamples) * c.num_kernels * 3);
    h->convs2.resize(h->convs1.size());
    int ch = C0;
    for (int i = 0; i < c.num_upsamples; ++i) {
        TRY(voc_make_convT(h, "ups." + std::to_string(i) + ".0", C0 >> i, C0 >> (i + 1), c.upsample_kernel_sizes[i], c.upsample_rates[i], &h->ups[i]));
        ch = C0 >> (i + 1);
        for (int j = 0; j < c.num_kernels; ++j) {
            const int n = i * c.num_kernels + j, k = c.resblock_kernel_sizes[j];
            const std::string rb = "resblocks." + std::to_string(n);
            for (int m = 0; m < 3; ++m) {
                TRY(voc_make_conv(h, rb + ".convs1." + std::to_string(m), ch, ch, k, c.resblock_dilation_sizes[j][m], &h->convs1[n * 3 + m]));
                TRY(voc_make_conv(h, rb + ".convs2." + std::to_string(m), ch, ch, k, 1, &h->convs2[n * 3 + m]));
            }
            for (int m = 0; m < 6; ++m) TRY(voc_make_act(h, rb + ".activations." + std::to_string(m), ch));
        }
    }
    TRY(voc_make_act(h, "activation_post", ch));
    const float *pw = nullptr, *pb = nullptr;
    TRY(voc_weight(h, "conv_post", {1, ch, 7}, &pw));
    TRY(voc_need(h, "conv_post.bias", {1}, &pb));
    TRY(voc_alloc_t(h, h->owned, &h->post_w, static_cast<size_t>(ch) * 7));
    TRY(voc_alloc_t(h, h->owned, &h->post_b, 1));
    CK(cudaMemcpy(h->post_w, pw, static_cast<size_t>(ch) * 7 * sizeof(float), cudaMemcpyDeviceToDevice));
    CK(cudaMemcpy(h->post_b, pb, sizeof(float), cudaMemcpyDeviceToDevice));
    h->post_C = ch;
    CK(cudaDeviceSynchronize());
    voc_free_bag(h->raw_owned);          // the state-dict copies are not needed once packed
    h->raw.clear();
    static PerDeviceOnce attr;
    if (!attr.done()) {
        CK(cudaFuncSetAttribute(voc_conv_tc_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, VcCfg<256>::kSmem));
        CK(cudaFuncSetAttribute(voc_conv_tc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, VcCfg<128>::kSmem));
        CK(cudaFuncSetAttribute(voc_conv_tc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, VcCfg<64>::kSmem));
        CK(cudaFuncSetAttribute(voc_conv_halo_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, VhCfg<256>::kSmem));
        CK(cudaFuncSetAttribute(voc_conv_halo_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, VhCfg<128>::kSmem));
        CK(cudaFuncSetAttribute(voc_conv_halo_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, VhCfg<64>::kSmem));
        attr.mark();
    }
    h->finalized = true;
    return TVB_OK;
}

extern "C" int tvb_voc_set_precision(tvb_voc_handle* h, int32_t precision) {
    if (!h) return fail(TVB_ERR_INVALID, "null handle");
    if (precision != TVB_PRECISION_BF16 && precision != TVB_PRECISION_FP32_VALIDATE) return fail(TVB_ERR_INVALID, "unknown precision %d", precision);
    h->precision = precision;
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// forward
// ----------------------------------------------------------------------------------------------------------------
static int voc_workspace(tvb_voc_handle* h, int B, int T) {
    const tvb_voc_config& c = h->cfg;
    size_t need = static_cast<size_t>(B) * T * std::max(c.upsample_initial_channel, round_up(c.num_mels, 8));
    long long Ts = T;
    for (int i = 0; i < c.num_upsamples; ++i) {
        Ts *= c.upsample_rates[i];
        need = std::max(need, static_cast<size_t>(static_cast<size_t>(B) * Ts * round_up(c.upsample_initial_channel >> (i + 1), 8)));
    }
    if (Ts >= (1ll << 31) / 4) return fail(TVB_ERR_INVALID, "utterance too long for the vocoder (T * hop = %lld)", Ts);
    if (need <= h->ws_elems) return TVB_OK;
    CK(cudaDeviceSynchronize());
    voc_free_bag(h->ws_owned);
    h->ws_bytes = 0;
    h->ws_elems = 0;
    for (float** p : {&h->X, &h->XS, &h->U, &h->R, &h->Hb, &h->AF}) TRY(voc_alloc_t(h, h->ws_owned, p, need));
    TRY(voc_alloc_t(h, h->ws_owned, &h->A16, need));
    h->ws_elems = need;
    return TVB_OK;
}

static int voc_act(tvb_voc_handle* h, const std::string& name, const float* x, int B, int T, cudaStream_t st, bool force_f32 = false) {
    auto it = h->acts.find(name);
    if (it == h->acts.end()) return fail(TVB_ERR_STATE, "no activation %s", name.c_str());
    const VocActW& a = it->second;
    const int chunks = (T + VOC_ACT_TT - 1) / VOC_ACT_TT;
    const bool tc = h->precision == TVB_PRECISION_BF16 && !force_f32;
    const int vec = (tc && h->opt_act_vec2 && (a.C & 1) == 0) ? 2 : 1;
    dim3 block(128), grid(static_cast<unsigned>((static_cast<long long>(chunks) * (a.C / vec) + 127) / 128), B);
    VocProf ps(h, st, 1);
    if (tc && vec == 2)
        k_voc_act<bf16, true, 2><<<grid, block, 0, st>>>(x, a.C, T, round_up(a.C, 8), a.consts, a.F, h->A16);
    else if (tc)
        k_voc_act<bf16, true, 1><<<grid, block, 0, st>>>(x, a.C, T, round_up(a.C, 8), a.consts, a.F, h->A16);
    else
        k_voc_act<float, false, 1><<<grid, block, 0, st>>>(x, a.C, T, a.C, a.consts, a.F, h->AF);
    CKL();
    return TVB_OK;
}

// out = alpha * (conv(in) + bias + resid) (+ out).  Input: h->A16 (bf16 path, ld = round_up(Cin, 8)) or `in32` (fp32 path).
static int voc_conv(tvb_voc_handle* h, const VocConv& L, const float* in32, int B, int T, const float* resid, float* out, float alpha,
                    bool accumulate, cudaStream_t st) {
    VocConvParams p{};
    p.B = B; p.T = T; p.Cin = L.Cin; p.N = L.N; p.taps = L.taps; p.dil = L.dil; p.ci_pad = L.ci_pad; p.n_pad = L.n_pad; p.bn = L.bn;
    p.bias = L.bias; p.resid = resid; p.out = out; p.alpha = alpha; p.accumulate = accumulate ? 1 : 0;
    VocProf ps(h, st, 0);
    if (h->precision == TVB_PRECISION_BF16) {
        CUtensorMap tmX;
        const int tiles = B * ((T + VC_BM - 1) / VC_BM) * (L.n_pad / L.bn);
        const int a_rows = round_up(VC_BM + 2 * (L.taps / 2) * L.dil, 8);
        if (h->opt_halo && a_rows <= VH_A_ROWS) {
            TRY(make_tmap_act3(&tmX, h->A16, L.Cin, round_up(L.Cin, 8), T, B, a_rows));
            if (L.bn > 128)
                voc_conv_halo_kernel<256><<<std::min(tiles, h->num_sms * VhCfg<256>::kMinBlocks), VhCfg<256>::kThreads, VhCfg<256>::kSmem, st>>>(tmX, L.tmW, p, a_rows);
            else if (L.bn > 64)
                voc_conv_halo_kernel<128><<<std::min(tiles, h->num_sms * VhCfg<128>::kMinBlocks), VhCfg<128>::kThreads, VhCfg<128>::kSmem, st>>>(tmX, L.tmW, p, a_rows);
            else
                voc_conv_halo_kernel<64><<<std::min(tiles, h->num_sms * VhCfg<64>::kMinBlocks), VhCfg<64>::kThreads, VhCfg<64>::kSmem, st>>>(tmX, L.tmW, p, a_rows);
        } else {
            TRY(make_tmap_act3(&tmX, h->A16, L.Cin, round_up(L.Cin, 8), T, B));
            if (L.bn > 128)
                voc_conv_tc_kernel<256><<<std::min(tiles, h->num_sms * VcCfg<256>::kMinBlocks), VC_THREADS, VcCfg<256>::kSmem, st>>>(tmX, L.tmW, p);
            else if (L.bn > 64)
                voc_conv_tc_kernel<128><<<std::min(tiles, h->num_sms * VcCfg<128>::kMinBlocks), VC_THREADS, VcCfg<128>::kSmem, st>>>(tmX, L.tmW, p);
            else
                voc_conv_tc_kernel<64><<<std::min(tiles, h->num_sms * VcCfg<64>::kMinBlocks), VC_THREADS, VcCfg<64>::kSmem, st>>>(tmX, L.tmW, p);
        }
    } else {
        dim3 grid((T + 3) / 4, (L.N + 127) / 128, B);
        k_voc_conv_f32<<<grid, 128, 0, st>>>(in32, L.wt, p);
    }
    CKL();
    return TVB_OK;
}

extern "C" int tvb_voc_forward(tvb_voc_handle* h, const float* mel_dev, int32_t B, int32_t T, float* wav_dev, void* stream) {
    if (!h || !mel_dev || !wav_dev) return fail(TVB_ERR_INVALID, "null argument to tvb_voc_forward");
    if (!h->finalized) return fail(TVB_ERR_STATE, "weights not finalized");
    if (B <= 0 || T <= 0 || B > 65535) return fail(TVB_ERR_INVALID, "bad B / T");
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const tvb_voc_config& c = h->cfg;
    TRY(voc_workspace(h, B, T));
    const bool tc = h->precision == TVB_PRECISION_BF16;
    h->ev_used = 0;
    h->ev_cat.clear();
    {   // mel [B][num_mels][T] -> time-major operand
        dim3 grid((T + 31) / 32, (c.num_mels + 31) / 32, B), block(32, 8);
        VocProf ps(h, st, 2);
        if (tc) k_voc_mel_in<bf16><<<grid, block, 0, st>>>(mel_dev, c.num_mels, T, round_up(c.num_mels, 8), h->A16);
        else k_voc_mel_in<float><<<grid, block, 0, st>>>(mel_dev, c.num_mels, T, c.num_mels, h->AF);
        CKL();
    }
    float *X = h->X, *XS = h->XS;
    TRY(voc_conv(h, h->conv_pre, h->AF, B, T, nullptr, X, 1.f, false, st));                    // models.py:194
    int Tc = T, ch = c.upsample_initial_channel;
    const float inv_k = 1.0f / static_cast<float>(c.num_kernels);
    for (int i = 0; i < c.num_upsamples; ++i) {
        h->cur_stage = i;
        if (tc) {
            const size_t rows = static_cast<size_t>(B) * Tc;
            VocProf ps(h, st, 2);
            k_voc_to_bf16<<<grid_for(rows * ch), 256, 0, st>>>(X, rows, ch, round_up(ch, 8), h->A16);
            CKL();
        }
        TRY(voc_conv(h, h->ups[i], X, B, Tc, nullptr, h->U, 1.f, false, st));                    // models.py:198-200
        Tc *= c.upsample_rates[i];
        ch >>= 1;
        for (int j = 0; j < c.num_kernels; ++j) {                                                // models.py:202-208
            const int n = i * c.num_kernels + j;
            const std::string rb = "resblocks." + std::to_string(n) + ".activations.";
            for (int m = 0; m < 3; ++m) {                                                        // models.py:72-77
                const float* src = m == 0 ? h->U : h->R;
                const bool last = m == 2;
                TRY(voc_act(h, rb + std::to_string(2 * m), src, B, Tc, st));
                TRY(voc_conv(h, h->convs1[n * 3 + m], h->AF, B, Tc, nullptr, h->Hb, 1.f, false, st));
                TRY(voc_act(h, rb + std::to_string(2 * m + 1), h->Hb, B, Tc, st));
                TRY(voc_conv(h, h->convs2[n * 3 + m], h->AF, B, Tc, src, last ? XS : h->R, last ? inv_k : 1.f, last && j > 0, st));
            }
        }
        std::swap(X, XS);                                                                        // x = xs / num_kernels (models.py:209)
    }
    h->cur_stage = -1;
    TRY(voc_act(h, "activation_post", X, B, Tc, st, true));                                      // models.py:212
    {
        VocProf ps(h, st, 2);
        dim3 grid((Tc + 255) / 256, B);
        k_voc_post<<<grid, 256, 0, st>>>(h->AF, h->post_C, Tc, h->post_w, h->post_b, wav_dev);   // models.py:213-214
        CKL();
    }
    if (h->profile) {
        CK(cudaStreamSynchronize(st));
        for (float& v : h->prof_ms) v = 0.f;
        for (size_t e = 0; e < h->ev_cat.size(); ++e) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, h->ev_pool[2 * e], h->ev_pool[2 * e + 1]));
            const int cat = h->ev_cat[e] & 0xff, stage = (h->ev_cat[e] >> 8) - 1;
            h->prof_ms[cat] += ms;
            if (stage >= 0 && cat < 2) h->prof_ms[3 + cat * TVB_VOC_MAX_UPSAMPLES + stage] += ms;
        }
    }
    return TVB_OK;
}

extern "C" int tvb_voc_debug_activation(tvb_voc_handle* h, const char* name, const float* x_dev, int32_t B, int32_t T, float* out_dev, void* stream) {
    if (!h || !name || !x_dev || !out_dev || B <= 0 || T <= 0) return fail(TVB_ERR_INVALID, "bad argument to tvb_voc_debug_activation");
    if (!h->finalized) return fail(TVB_ERR_STATE, "weights not finalized");
    auto it = h->acts.find(name);
    if (it == h->acts.end()) return fail(TVB_ERR_INVALID, "no activation named %s", name);
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int C = it->second.C;
    const size_t need = static_cast<size_t>(B) * T * C;
    if (need > h->ws_elems) return fail(TVB_ERR_INVALID, "run tvb_voc_forward once at a size >= this one first (workspace)");
    dim3 grid((T + 31) / 32, (C + 31) / 32, B), block(32, 8);
    k_voc_mel_in<float><<<grid, block, 0, st>>>(x_dev, C, T, C, h->X);
    CKL();
    TRY(voc_act(h, name, h->X, B, T, st, true));
    k_voc_tm_to_cm<<<grid, block, 0, st>>>(h->AF, C, T, out_dev);
    CKL();
    h->launches += 2;
    return TVB_OK;
}

extern "C" int tvb_voc_set_option(tvb_voc_handle* h, const char* name, int32_t value) {
    if (!h || !name) return fail(TVB_ERR_INVALID, "null argument");
    if (!strcmp(name, "halo")) h->opt_halo = value;
    else if (!strcmp(name, "act_vec2")) h->opt_act_vec2 = value;
    else return fail(TVB_ERR_INVALID, "unknown vocoder option %s", name);
    return TVB_OK;
}
extern "C" int64_t tvb_voc_kernel_launches(const tvb_voc_handle* h) { return h ? h->launches : 0; }
extern "C" int64_t tvb_voc_device_bytes(const tvb_voc_handle* h) { return h ? static_cast<int64_t>(h->dev_bytes + h->ws_bytes) : 0; }
extern "C" int tvb_voc_set_profile(tvb_voc_handle* h, int32_t on) {
    if (!h) return fail(TVB_ERR_INVALID, "null handle");
    h->profile = on != 0;
    return TVB_OK;
}
extern "C" int tvb_voc_profile_read(tvb_voc_handle* h, float* ms, int32_t n) {
    if (!h || !ms || n < 3) return fail(TVB_ERR_INVALID, "need room for at least 3 categories");
    for (int i = 0; i < n && i < 3 + 2 * TVB_VOC_MAX_UPSAMPLES; ++i) ms[i] = h->prof_ms[i];
    return TVB_OK;
}
