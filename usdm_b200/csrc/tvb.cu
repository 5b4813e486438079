// libtvb: C-ABI engine of the B200-native token-Voicebox mel generator (see include/tvb.h).
// Host-side orchestration only lives here: weight packing, the packed row layout ("plan"), the per-step kernel
// sequence and the ODE loop.  All arithmetic runs in the kernels of gemm_tc.cuh / attn_tc.cuh / conv_tc.cuh / simt.cuh.
// There is deliberately no CPU implementation of anything: without a CUDA device every compute entry point fails.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/tvb.h"
#include "../../include/tvb_vocoder.h"
#include "attn_tc.cuh"
#include "attn4_tc.cuh"
#include "attn5_tc.cuh"
#include "attn6_tc.cuh"
#include "attn7_tc.cuh"
#include "attn8_tc.cuh"
#include "attn9_tc.cuh"
#include "attn10_tc.cuh"
#include <algorithm>
#include "conv_tc.cuh"
#include "gemm_tc.cuh"
#include "gemmln_tc.cuh"
#include "glue.cuh"
#include "simt.cuh"
#include "vocoder.cuh"

using namespace tvb;
typedef __nv_bfloat16 bf16;

// ----------------------------------------------------------------------------------------------------------------
// error plumbing
// ----------------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static int fail(int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CK(call)                                                                                              \
    do {                                                                                                      \
        cudaError_t e__ = (call);                                                                             \
        if (e__ != cudaSuccess) return fail(TVB_ERR_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
    } while (0)
#define CKL() CK(cudaGetLastError())
#define TRY(expr)                  \
    do {                           \
        int rc__ = (expr);         \
        if (rc__ != TVB_OK) return rc__; \
    } while (0)

// ----------------------------------------------------------------------------------------------------------------
// TMA descriptor creation (driver entry point resolved at run time: no link-time dependency on libcuda)
// ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                        const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                        CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled g_encode = nullptr;
static int resolve_driver() {
    if (g_encode) return TVB_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    if (!fn || q != cudaDriverEntryPointSuccess) return fail(TVB_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    g_encode = reinterpret_cast<PFN_tmapEncodeTiled>(fn);
    return TVB_OK;
}
// 2-D row-major [rows, cols] (leading dimension ld elements), box = box_rows x box_cols (128 bytes wide), 128-byte swizzle.
// Default element type bf16 (box 64 columns); f32 = true makes an fp32 map (box 32 columns).
static int make_tmap(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows, int box_cols = 64,
                     bool f32 = false, bool sw64 = false, bool sw32 = false) {
    TRY(resolve_driver());
    const int esz = f32 ? 4 : 2;
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * esz};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(m, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, sw32 ? CU_TENSOR_MAP_SWIZZLE_32B : sw64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(TVB_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld", (int)r,
                                       (long long)rows, (long long)cols, (long long)ld);
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// handle
// ----------------------------------------------------------------------------------------------------------------
struct RawTensor {
    float* p = nullptr;
    std::vector<int64_t> shape;
    size_t n = 0;
};
struct Lin {                 // one nn.Linear: y = x W^T + b
    const float* w32 = nullptr;
    bf16* w16 = nullptr;
    const float* bias = nullptr;
    const float* bias_tc = nullptr;   // bias of the bf16 tensor-core path (differs from `bias` only for the fused QKV projection)
    const float* fold_c = nullptr;    // column sums of a LayerNorm-folded weight (EPI_FOLD_GELU)
    int N = 0, K = 0;
    CUtensorMap tm;          // B operand map: [N, K], box 256 x 64 (single-CTA kernel)
    CUtensorMap tm2;         // B operand map: [N, K], box 128 x 64 (CTA-pair kernel: each CTA stages half of the W tile)
    CUtensorMap tm3;         // B operand map: [N, K], box  64 x 64 (cluster-of-4 kernel: each CTA fetches a quarter and multicasts it)
};
struct Act {                 // an activation buffer [Mp, C] in the precision of the active path
    float* f32 = nullptr;
    bf16* b16 = nullptr;
    int C = 0;
    CUtensorMap tm;          // A operand map (bf16 path): box 128 x 64
    CUtensorMap tm_out;      // epilogue TMA-store map (bf16 path): box 32 rows x 64 columns
    CUtensorMap tm_out32;    // epilogue TMA-store map, box 32 rows x 32 columns, 64-B swizzle (hbB: bf16 copy written by EPI_RESID_*_ST)
};
struct LayerW {
    Lin qkv, o, f1, f2;
    Lin f1_fold;                       // FFN-in with LN1 folded in: W1 diag(gamma1) in bf16, bias_tc = b1 + W1 beta1
    float* f1_fold_c = nullptr;        // [F] column sums of the folded bf16 weight
    const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

struct tvb_handle {
    tvb_config cfg;
    int device = 0;
    int precision = TVB_PRECISION_BF16;
    bool finalized = false;
    int num_sms = 148;
    int64_t launches = 0;
    int* err_host = nullptr;    // mapped pinned word that kernels set when they meet invalid DATA (unit ids outside the table); host-visible
    int* err_dev = nullptr;     //   without a device synchronisation, checked and cleared by tvb_async_status / the next compute call
    size_t dev_bytes = 0;       // weights
    size_t ws_bytes = 0;        // workspace
    int opt_attn_tc = 1;        // bf16 path: 1 = tcgen05 attention, 0 = CUDA-core attention (ablation / cross-check)
    int opt_conv_tc = 1;        // bf16 path: 1 = tcgen05 positional conv, 0 = CUDA-core conv
    int opt_gemm_2cta = 1;      // bf16 path: 1 = CTA-pair (cta_group::2) GEMM, 0 = single-CTA GEMM
    int opt_ln1_fold = 1;       // bf16 path with ln_fuse == 1: LN1 folded into the FFN-in GEMM (no LN1 kernel, attention-out epilogue emits
                                // a bf16 copy + row statistics); 0 = LN1 as a kernel
    int opt_conv_halo = 1;      // tcgen05 positional conv: 1 = halo kernel (activation tile fetched once per 256-row tile, taps as row-shifted descriptors), 0 = one fetch per tap
    int opt_out_tc = 1;         // bf16 path: 1 = output projection (hidden -> n_feats) as a tcgen05 GEMM on the last LayerNorm's bf16 output, 0 = fp32 CUDA-core GEMM
    int opt_frame_tc = 1;       // bf16 path: 1 = per-pass frame projection (y -> hidden) as a tcgen05 GEMM over a hi/lo bf16 split of y, 0 = CUDA-core kernel
    int opt_pdl = 1;            // 1 = GEMM / attention / LayerNorm kernels are launched with programmatic stream serialization: their set-up
                                // overlaps the tail of the previous kernel (griddepcontrol), 0 = ordinary launches
    int opt_graph = 1;          // 1 = small batches (<= 16384 packed rows) replay one estimator pass as a CUDA graph inside tvb_solve
                                // (the pass is ~230 launches; at B = 1 the host cannot feed them fast enough), 0 = always launch directly
    int opt_zigzag = 1;         // bf16 path: row-block traversal alternates between producer and consumer kernels (L2 reuse)
    int opt_resid_l2 = 0;       // CTA-pair residual GEMM epilogues prefetch residual chunks into the L2 ahead of their staged load (GemmParams::resid_l2; ablation)
    int opt_a_l2 = 0;           // CTA-pair GEMM producers prefetch activation k-blocks this far ahead into the L2 (GemmParams::a_l2; ablation)
    int opt_epi_direct = 0;     // CTA-pair GEMM, fp32-output epilogues (ablation): registers <-> global memory directly (tcgen05.ld.16x256b fragments,
                                // sector-exact 8-byte accesses; gemm_tc.cuh: DIRECT) instead of staged through shared-memory tiles and TMA.  0 = never
                                // (default), 1 = the FFN-out GEMM only, 2 = every fp32-output epilogue.  Same bits as the staged path.  Measured
                                // (profiles/r02/bench_epi_direct_*.json): FFN-out 185.0 -> 182.5 ms per solve (K = 4096 hides the epilogue and the mainloop
                                // gets the shared-memory bandwidth the staging took) but nothing at the step level; attention-out 82 -> 120 ms (where the
                                // epilogue is the critical path the 8-byte accesses are LSU-issue bound)
    int opt_gemm_lab = 0;       // timing ablations of the *_ST residual epilogue (GemmParams::lab): any non-zero value gives WRONG results
    int opt_gemm_bn = 1;        // CTA-pair GEMM: 1 = 128-column tiles when the 256-column tiling needs more (waves x tile width) -- small batches; 0 = always 256
    int ln_part_n = 8;          // entries per row the last *_ST GEMM wrote into ln_part (8: 256-column tiles, 16: 128-column tiles)
    int opt_gemm_mc = 0;        // CTA-pair GEMM: 1 = clusters of two pairs sharing the W tile by TMA multicast, 0 = independent pairs (default:
                                // measured 7 % slower with multicast, profiles/r01: the pair kernel is not L2->SM bound after all)
    int opt_attn_ptmem = 1;     // tcgen05 attention v3: 1 = P kept in tensor memory (TS MMA), 0 = P staged through shared memory
    int opt_attn_v = 10;        // tcgen05 attention kernel generation: 10 = attn10_tc.cuh (default), 9 .. 4 = attnN_tc.cuh, 3 = attn_tc.cuh
    int opt_qkv_norm = 1;       // attention v9+: 1 = the QKV GEMM epilogue emits the row-block norms of q and k (default), 0 = a separate kernel reads them back
    int opt_attn_sort = 1;      // attention v10: items handed to the CTAs in descending planned cost, snake order (k_attn10_order): 0 = never (launch
                                // order, stride), 2 = always, 1 = when the plan has long (>= 16 key blocks) or ragged sequences (default)
    bool plan_sort = false;     // the rule of opt_attn_sort == 1 evaluated for the current plan
    int* d_order = nullptr;
    bf16* d_pos = nullptr;      // attention v10: [rows, 16] key-side operand of the bias MMA (key positions, k_attn10_pos), rebuilt with the plan
    CUtensorMap tm_pos;
    int opt_attn_win = 1;       // attention v9: 1 = key blocks whose probabilities are provably below 2^-40 of the row sum are skipped (ALiBi
                                // window, attn9_tc.cuh), 0 = every key block is processed (ablation / exactness check)
    float* normtab = nullptr;   // attention v9: [rows / 32][2 heads] row-block maxima of |q|^2, |k|^2, written by the QKV GEMM epilogue
    int4* d_win = nullptr;      //   and the per-item (tile, head) windows / reference shifts derived from them (k_attn9_plan)
    bool norm_from_gemm = false;   // the last QKV GEMM filled normtab (otherwise attention() runs k_attn9_rownorm)
    int opt_attn_pace = 0;      // attention v8: 1 = MUFU pairs paced by an earlier pair's tail (forces ptxas to interleave the pipes in program
                                // order; measured 2-15 % SLOWER: the added dependency latency outweighs it), 0 = free order (default)
    int opt_attn_poly = 4;      // attention v4+: 0 = every exp2 on MUFU, n = one element pair in n on the FMA-pipe polynomial (default 4: v9 measured
                                // 244 ms per solve against 253 ms with 3, profiles/r02)
    int opt_ln_fuse = 1;        // bf16 path: 2 = LayerNorm fused into the residual GEMM epilogue (8-CTA clusters, gemmln_tc.cuh);
                                //            1 = LayerNorm writes bf16 + (mean, rstd) only and the residual GEMM epilogue re-normalises
                                //            the fp32 stream on the fly (EPI_RESID_LN); 0 = LayerNorm materialises the fp32 result
    int profile = 0;            // 1: bracket every kernel with CUDA events on the launching stream (bench.py roofline leg)
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<int> ev_cat;
    std::vector<void*> owned;                    // weight-side allocations
    std::map<std::string, RawTensor> raw;

    // packed weights
    float *G = nullptr, *b_in = nullptr, *Wy_t = nullptr, *Wc_t = nullptr, *freqs = nullptr, *slopes = nullptr;
    std::vector<float*> conv_wp32;
    std::vector<bf16*> conv_wp16;
    std::vector<const float*> conv_bias;
    std::vector<CUtensorMap> conv_tm;            // [G*taps*64, 64] bf16 packed weights, box 64 x 64
    const float *ln0_g = nullptr, *ln0_b = nullptr;
    std::vector<LayerW> layers;
    std::vector<Lin> skip;
    const float *w_out = nullptr, *b_out = nullptr;
    bf16* w_out16 = nullptr;                     // [2][128 rows (n_feats padded), H] for the fused output kernel

    // plan (row layout) + workspace
    std::vector<void*> ws_owned;
    int ws_cap_rows = 0, ws_precision = -1;
    size_t ws_cap_state = 0;
    RowLayout L{};
    int *d_row_seq = nullptr, *d_row_pos = nullptr, *d_seq = nullptr;   // d_seq: 4 * nseq_cap ints
    int seq_cap = 0;
    int2* d_work16 = nullptr; int work16_cap = 0, n_work16 = 0;        // 16-query tiles (CUDA-core attention)
    int4* d_work128 = nullptr; int work128_cap = 0, n_work128 = 0;     // ATT_BQ-query tiles (tcgen05 attention)
    float* d_tseq = nullptr;
    // CUDA graphs of one estimator pass, keyed by the state pointer they read (tvb_solve: z_work, z_hat); dropped whenever the
    // plan, the precision or an option changes (graph_epoch)
    struct PassGraph { const float* y = nullptr; cudaGraphExec_t exec = nullptr; int64_t launches = 0; };
    PassGraph graphs[2];
    cudaStream_t cap_stream = nullptr;                  // capture happens on a private stream (the caller's may be the legacy default
                                                        // stream, which cannot be captured); the graph is launched on the caller's
    std::vector<int64_t> plan_lengths;
    int plan_precision = -1;
    int plan_B = 0, plan_T = 0, plan_cfg = 0, plan_half = 0;
    float* base = nullptr;
    Act hin, p1, slots[32], hbA, hbB, qkv, ctx, mid;
    float *xpre = nullptr, *xalt = nullptr, *hres = nullptr, *vbuf = nullptr;
    float2 *statsA = nullptr, *statsB = nullptr;       // per-row (mean, rstd) of the two LayerNorms in flight (ln_fuse)
    float2* ln_part = nullptr;                         // [Mp, <= 16] partial row statistics (ln1_fold), see GemmParams::part
    float *z_work = nullptr, *z_hat = nullptr, *v_pred = nullptr;
    CUtensorMap tm_hin_halo, tm_p1_halo;         // positional-conv A operand maps with the halo (box 160 rows x 64 columns)
    CUtensorMap tm_base, tm_hin32;               // fp32 maps of `base` (residual of the frame-projection GEMM) and of hin.f32 (its output)
    Lin lin_out;                                 // output projection padded to one 256-column tile: [256, H] (rows >= n_feats are zero)
    float *wout2 = nullptr, *bout2 = nullptr;
    int vbuf_ld = 0;                             // row stride of vbuf as written by the last estimator pass (n_feats, or 256 on the tensor-core path)
    CUtensorMap tm_vbuf;
    Lin lin_y;                                   // frame projection of the ODE state: [H, KY] = [W_y | W_y | 0] in bf16 (KY = 2 n_feats rounded up to 64)
    Act yop;                                     // its A operand: hi / lo bf16 planes of y in packed rows
    float *wy2 = nullptr, *zero_bias = nullptr;
    CUtensorMap tm_xpre, tm_xalt, tm_hres;       // fp32 [Mp, H] epilogue maps (box 32 rows x 32 columns): TMA store / residual load
};

static void drop_graphs(tvb_handle* h);

// per-kernel timing scopes (only active when option "profile" is set)
enum ProfCat { PC_GEMM = 0, PC_ATTN = 1, PC_CONV = 2, PC_NORM = 3, PC_FRAME = 4, PC_OUT = 5,
               PC_GEMM_QKV = 8, PC_GEMM_O = 9, PC_GEMM_F1 = 10, PC_GEMM_F2 = 11, PC_GEMM_SKIP = 12, PC_NCAT = 16 };
struct ProfScope {
    tvb_handle* h;
    cudaStream_t st;
    bool on;
    ProfScope(tvb_handle* h_, cudaStream_t st_, int cat) : h(h_), st(st_), on(h_ && h_->profile) {
        if (!on) return;
        while (h->ev_pool.size() < h->ev_used + 2) {
            cudaEvent_t e;
            cudaEventCreate(&e);
            h->ev_pool.push_back(e);
        }
        h->ev_cat.push_back(cat);
        cudaEventRecord(h->ev_pool[h->ev_used], st);
    }
    ~ProfScope() {
        if (!on) return;
        cudaEventRecord(h->ev_pool[h->ev_used + 1], st);
        h->ev_used += 2;
    }
};

static int dev_alloc(tvb_handle* h, std::vector<void*>& bag, void** p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(TVB_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    bag.push_back(*p);
    if (&bag == &h->ws_owned) h->ws_bytes += bytes;
    else h->dev_bytes += bytes;
    return TVB_OK;
}
template <typename T>
static int dev_alloc_t(tvb_handle* h, std::vector<void*>& bag, T** p, size_t n) {
    return dev_alloc(h, bag, reinterpret_cast<void**>(p), n * sizeof(T));
}

static int grid_for(size_t n, int block = 256) {
    size_t g = (n + block - 1) / block;
    if (g > 148 * 16) g = 148 * 16;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

// ----------------------------------------------------------------------------------------------------------------
// create / destroy / weights
// ----------------------------------------------------------------------------------------------------------------
extern "C" const char* tvb_last_error(void) { return g_err.c_str(); }
extern "C" const char* tvb_version(void) { return "tvb-b200 0.1 (sm_100a)"; }

extern "C" int tvb_create(const tvb_config* cfg, int device, tvb_handle** out) {
    if (!cfg || !out) return fail(TVB_ERR_INVALID, "null argument");
    const tvb_config& c = *cfg;
    if (c.hidden_size <= 0 || c.num_attention_heads <= 0 || c.hidden_size / c.num_attention_heads != 64 ||
        c.hidden_size % c.num_attention_heads != 0)
        return fail(TVB_ERR_INVALID, "unsupported: head_dim must be 64 (hidden %d, heads %d)", c.hidden_size, c.num_attention_heads);
    if (c.hidden_size % 256 != 0 || c.hidden_size > 1024)
        return fail(TVB_ERR_INVALID, "unsupported hidden_size %d (need multiple of 256, <= 1024)", c.hidden_size);
    if (c.intermediate_size % 256 != 0) return fail(TVB_ERR_INVALID, "unsupported intermediate_size %d", c.intermediate_size);
    if (c.embedding_dim % 16 != 0) return fail(TVB_ERR_INVALID, "unsupported embedding_dim %d", c.embedding_dim);
    if (c.convpos_groups <= 0 || c.hidden_size / c.convpos_groups != 64 || c.hidden_size % c.convpos_groups != 0)
        return fail(TVB_ERR_INVALID, "unsupported: conv channels per group must be 64");
    if (c.convpos_width % 2 != 1 || c.convpos_width > 33) return fail(TVB_ERR_INVALID, "unsupported convpos_width %d (odd, <= 33)", c.convpos_width);
    if (c.num_hidden_layers % 2 != 0 || c.num_hidden_layers > 64 || c.num_hidden_layers < 2)
        return fail(TVB_ERR_INVALID, "unsupported num_hidden_layers %d", c.num_hidden_layers);
    if (c.n_feats <= 0 || c.n_feats > 128 || c.n_feats % 16 != 0) return fail(TVB_ERR_INVALID, "unsupported n_feats %d", c.n_feats);
    if (c.convpos_depth < 1 || c.convpos_depth > 2) return fail(TVB_ERR_INVALID, "unsupported convpos_depth %d (1 or 2)", c.convpos_depth);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0) {
        cudaGetLastError();
        return fail(TVB_ERR_CUDA, "no CUDA device available (%s); libtvb has no CPU path", cudaGetErrorString(e));
    }
    if (device < 0 || device >= ndev) return fail(TVB_ERR_INVALID, "device %d out of range", device);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(TVB_ERR_INVALID, "device %d is sm_%d%d; libtvb is built for sm_100a only", device, prop.major, prop.minor);
    tvb_handle* h = new tvb_handle();
    h->cfg = c;
    h->device = device;
    h->num_sms = prop.multiProcessorCount;
    if (cudaHostAlloc(reinterpret_cast<void**>(&h->err_host), 64, cudaHostAllocMapped) != cudaSuccess ||
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&h->err_dev), h->err_host, 0) != cudaSuccess) {
        cudaGetLastError();
        if (h->err_host) cudaFreeHost(h->err_host);
        delete h;
        return fail(TVB_ERR_NOMEM, "cannot allocate the mapped status word");
    }
    for (int i = 0; i < 16; ++i) h->err_host[i] = 0;
    *out = h;
    return TVB_OK;
}

// Data errors found by kernels of EARLIER calls (no call synchronises the device to look for them): reported once, then cleared.
static int take_async_error(tvb_handle* h) {
    volatile int* f = h->err_host;
    if (f && f[0]) {
        f[0] = 0;
        return fail(TVB_ERR_INVALID, "a previous call was given unit ids outside [0, n_tokens = %d]; its output is invalid "
                                     "(the reference's nn.Embedding asserts on such ids, networks.py:305)", h->cfg.n_tokens);
    }
    if (f && f[1]) {
        f[1] = 0;
        return fail(TVB_ERR_INVALID, "a previous call produced attention logits beyond the range the tcgen05 attention kernel supports "
                                     "(2 max|q| max|k| > %g in the log2 domain, or non-finite activations); its output is invalid -- "
                                     "set option attn_v = 8 for such weights", static_cast<double>(A9_LOGIT_LIMIT));
    }
    if (f && f[2]) {
        f[2] = 0;
        return fail(TVB_ERR_INVALID, "a previous call met rows whose mean exceeds %g standard deviations in the LayerNorm that is folded into the "
                                     "FFN-in GEMM (the raw-row operand loses precision there); its output is less accurate than the tolerance -- "
                                     "set option ln1_fold = 0 for such weights", static_cast<double>(GEMM_FOLD_MEAN_LIMIT));
    }
    return TVB_OK;
}
extern "C" int tvb_async_status(tvb_handle* h) {
    if (!h) return fail(TVB_ERR_INVALID, "null handle");
    return take_async_error(h);
}

extern "C" int tvb_destroy(tvb_handle* h) {
    if (!h) return TVB_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    for (void* p : h->owned) cudaFree(p);
    for (void* p : h->ws_owned) cudaFree(p);
    for (cudaEvent_t e : h->ev_pool) cudaEventDestroy(e);
    if (h->err_host) cudaFreeHost(h->err_host);
    drop_graphs(h);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    delete h;
    return TVB_OK;
}

extern "C" int tvb_load_weight(tvb_handle* h, const char* key, const void* data, const int64_t* shape, int32_t ndim) {
    if (!h || !key || !data || !shape || ndim < 1 || ndim > 4) return fail(TVB_ERR_INVALID, "bad argument to tvb_load_weight");
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
        TRY(dev_alloc_t(h, h->owned, &t.p, t.n));
    }
    CK(cudaMemcpy(t.p, data, t.n * sizeof(float), cudaMemcpyDefault));
    h->raw[key] = t;
    return TVB_OK;
}

static int need(tvb_handle* h, const std::string& key, std::vector<int64_t> shape, const float** out) {
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

// w32_tc / bias_tc: the weights the bf16 tensor-core path uses when they differ from the fp32 path's (fused QKV: the q rows
// additionally carry log2(e) so that the tcgen05 attention kernel gets scores in the log2 domain straight out of the MMA).
static int make_lin(tvb_handle* h, Lin* lin, const float* w32, const float* bias, int N, int K, const float* w32_tc = nullptr,
                    const float* bias_tc = nullptr) {
    lin->w32 = w32;
    lin->bias = bias;
    lin->bias_tc = bias_tc ? bias_tc : bias;
    lin->N = N;
    lin->K = K;
    TRY(dev_alloc_t(h, h->owned, &lin->w16, static_cast<size_t>(N) * K));
    k_f32_to_bf16<<<grid_for(static_cast<size_t>(N) * K), 256>>>(w32_tc ? w32_tc : w32, lin->w16, static_cast<size_t>(N) * K);
    CKL();
    TRY(make_tmap(&lin->tm, lin->w16, N, K, K, GEMM_BN));
    TRY(make_tmap(&lin->tm2, lin->w16, N, K, K, 128));
    TRY(make_tmap(&lin->tm3, lin->w16, N, K, K, 64));
    return TVB_OK;
}

template <int EPI>
static int launch_sgemm(tvb_handle* h, const SgemmParams& p, cudaStream_t st) {
    dim3 grid((p.N + 127) / 128, (p.M + 127) / 128);
    ProfScope ps(h, st, h->precision == TVB_PRECISION_BF16 ? PC_OUT : PC_GEMM);
    k_sgemm<EPI><<<grid, 256, 0, st>>>(p);
    h->launches++;
    CKL();
    return TVB_OK;
}

extern "C" int tvb_finalize_weights(tvb_handle* h) {
    if (!h) return fail(TVB_ERR_INVALID, "null handle");
    if (h->finalized) return TVB_OK;
    CK(cudaSetDevice(h->device));
    const tvb_config& c = h->cfg;
    const int H = c.hidden_size, E = c.embedding_dim, F = c.intermediate_size, nf = c.n_feats, nl = c.num_hidden_layers;
    const int taps = c.convpos_width, cpg = H / c.convpos_groups;
    const std::string p = "estimator.";
    const float *embed, *w_in, *b_in;
    TRY(need(h, p + "embed.weight", {c.n_tokens + 1, E}, &embed));
    TRY(need(h, p + "proj_in.weight", {H, 2 * nf + E, 1}, &w_in));
    TRY(need(h, p + "proj_in.bias", {H}, &b_in));
    h->b_in = const_cast<float*>(b_in);
    const int ldin = 2 * nf + E;
    // unit table: G[id] = sqrt(E) * embed[id] @ W_in[:, :E]^T                                  (networks.py:305-307)
    TRY(dev_alloc_t(h, h->owned, &h->G, static_cast<size_t>(c.n_tokens + 1) * H));
    {
        SgemmParams sp{};
        sp.M = c.n_tokens + 1; sp.N = H; sp.K = E; sp.K1 = E;
        sp.A = embed; sp.lda = E; sp.A2 = embed; sp.lda2 = E;
        sp.W = w_in; sp.ldw = ldin; sp.bias = nullptr; sp.resid = nullptr; sp.ldr = 0;
        sp.C = h->G; sp.ldc = H; sp.alpha = sqrtf(static_cast<float>(E));
        TRY(launch_sgemm<SG_PLAIN>(h, sp, 0));
    }
    TRY(dev_alloc_t(h, h->owned, &h->Wy_t, static_cast<size_t>(nf) * H));
    TRY(dev_alloc_t(h, h->owned, &h->Wc_t, static_cast<size_t>(nf) * H));
    k_transpose_slice<<<(nf * H + 255) / 256, 256>>>(w_in, ldin, E, nf, H, h->Wy_t);
    k_transpose_slice<<<(nf * H + 255) / 256, 256>>>(w_in, ldin, E + nf, nf, H, h->Wc_t);
    CKL();
    {   // tensor-core frame projection operand: [W_y | W_y | 0] (the two halves meet the hi and lo bf16 planes of y)
        const int KY = (2 * nf + 63) / 64 * 64;
        TRY(dev_alloc_t(h, h->owned, &h->wy2, static_cast<size_t>(H) * KY));
        TRY(dev_alloc_t(h, h->owned, &h->zero_bias, static_cast<size_t>(H)));
        CK(cudaMemset(h->zero_bias, 0, sizeof(float) * H));
        k_build_wy2<<<(H * KY + 255) / 256, 256>>>(h->Wy_t, nf, H, KY, h->wy2);
        CKL();
        TRY(make_lin(h, &h->lin_y, h->wy2, h->zero_bias, H, KY));
    }
    // time-token frequencies f_k = exp(-k ln(1e4)/(H/2-1))                                    (networks.py:22-25)
    {
        std::vector<float> f(H / 2);
        const double s = std::log(10000.0) / (H / 2 - 1);
        const float sf = static_cast<float>(s);
        for (int k = 0; k < H / 2; ++k) f[k] = static_cast<float>(std::exp(static_cast<double>(static_cast<float>(k) * -sf)));
        TRY(dev_alloc_t(h, h->owned, &h->freqs, f.size()));
        CK(cudaMemcpy(h->freqs, f.data(), f.size() * sizeof(float), cudaMemcpyHostToDevice));
        // ALiBi slopes (networks.py:99-115), heads = power of two or not
        std::vector<float> sl;
        auto pow2 = [](int n) {
            std::vector<double> v;
            const double start = std::pow(2.0, -std::pow(2.0, -(std::log2((double)n) - 3)));
            for (int i = 0; i < n; ++i) v.push_back(start * std::pow(start, i));
            return v;
        };
        const int nh = c.num_attention_heads;
        const double l2 = std::log2((double)nh);
        if (std::floor(l2) == l2) {
            for (double v : pow2(nh)) sl.push_back(static_cast<float>(v));
        } else {
            const int cp = 1 << static_cast<int>(std::floor(l2));
            for (double v : pow2(cp)) sl.push_back(static_cast<float>(v));
            std::vector<double> ext = pow2(2 * cp);
            for (int i = 0; i < nh - cp; ++i) sl.push_back(static_cast<float>(ext[2 * i]));
        }
        TRY(dev_alloc_t(h, h->owned, &h->slopes, sl.size()));
        CK(cudaMemcpy(h->slopes, sl.data(), sl.size() * sizeof(float), cudaMemcpyHostToDevice));
    }
    // positional convs: fold weight norm, pack                                                (networks.py:67-95)
    float* norms = nullptr;
    TRY(dev_alloc_t(h, h->owned, &norms, 64));
    for (int i = 0; i < c.convpos_depth; ++i) {
        const std::string cp = p + "pos_conv_embeds." + std::to_string(i) + ".conv.";
        const float *g, *v, *b;
        TRY(need(h, cp + "parametrizations.weight.original0", {1, 1, taps}, &g));
        TRY(need(h, cp + "parametrizations.weight.original1", {H, cpg, taps}, &v));
        TRY(need(h, cp + "bias", {H}, &b));
        float* w32; bf16* w16;
        const size_t n = static_cast<size_t>(H) * cpg * taps;
        TRY(dev_alloc_t(h, h->owned, &w32, n));
        TRY(dev_alloc_t(h, h->owned, &w16, n));
        k_conv_tap_norm<<<taps, 256>>>(v, H, cpg, taps, norms);
        k_conv_pack<<<grid_for(n), 256>>>(v, g, norms, H, cpg, taps, w32, w16);
        CKL();
        CK(cudaDeviceSynchronize());
        h->conv_wp32.push_back(w32);
        h->conv_wp16.push_back(w16);
        h->conv_bias.push_back(b);
        CUtensorMap tm;
        TRY(make_tmap(&tm, w16, static_cast<int64_t>(c.convpos_groups) * taps * cpg, cpg, cpg, 64));
        h->conv_tm.push_back(tm);
    }
    TRY(need(h, p + "layer_norm.weight", {H}, &h->ln0_g));
    TRY(need(h, p + "layer_norm.bias", {H}, &h->ln0_b));
    // encoder layers
    h->layers.resize(nl);
    float* wqkv_tc = nullptr;                                  // scratch, reused by every layer (stream-ordered), freed below
    CK(cudaMalloc(&wqkv_tc, (3 * static_cast<size_t>(H) > static_cast<size_t>(F) ? 3 * static_cast<size_t>(H) : static_cast<size_t>(F)) * H * sizeof(float)));
    for (int l = 0; l < nl; ++l) {
        const std::string lp = p + "layers." + std::to_string(l) + ".";
        const float *wq, *wk, *wv, *bq, *bk, *bv, *wo, *bo, *w1, *b1, *w2, *b2;
        TRY(need(h, lp + "attention.q_proj.weight", {H, H}, &wq));
        TRY(need(h, lp + "attention.k_proj.weight", {H, H}, &wk));
        TRY(need(h, lp + "attention.v_proj.weight", {H, H}, &wv));
        TRY(need(h, lp + "attention.q_proj.bias", {H}, &bq));
        TRY(need(h, lp + "attention.k_proj.bias", {H}, &bk));
        TRY(need(h, lp + "attention.v_proj.bias", {H}, &bv));
        TRY(need(h, lp + "attention.out_proj.weight", {H, H}, &wo));
        TRY(need(h, lp + "attention.out_proj.bias", {H}, &bo));
        TRY(need(h, lp + "feed_forward.intermediate_dense.weight", {F, H}, &w1));
        TRY(need(h, lp + "feed_forward.intermediate_dense.bias", {F}, &b1));
        TRY(need(h, lp + "feed_forward.output_dense.weight", {H, F}, &w2));
        TRY(need(h, lp + "feed_forward.output_dense.bias", {H}, &b2));
        LayerW& lw = h->layers[l];
        TRY(need(h, lp + "layer_norm.weight", {H}, &lw.ln1_g));
        TRY(need(h, lp + "layer_norm.bias", {H}, &lw.ln1_b));
        TRY(need(h, lp + "final_layer_norm.weight", {H}, &lw.ln2_g));
        TRY(need(h, lp + "final_layer_norm.bias", {H}, &lw.ln2_b));
        // fused [q*0.125 ; k ; v]: the q scaling (head_dim^-0.5, applied after the bias, networks.py:162) is a power of two -> exact
        float *wqkv, *bqkv, *bqkv_tc;
        const size_t hh = static_cast<size_t>(H) * H;
        TRY(dev_alloc_t(h, h->owned, &wqkv, 3 * hh));
        TRY(dev_alloc_t(h, h->owned, &bqkv, 3 * static_cast<size_t>(H)));
        TRY(dev_alloc_t(h, h->owned, &bqkv_tc, 3 * static_cast<size_t>(H)));
        k_scale_copy<<<grid_for(hh), 256>>>(wq, wqkv, hh, 0.125f);
        k_scale_copy<<<grid_for(hh), 256>>>(wk, wqkv + hh, hh, 1.0f);
        k_scale_copy<<<grid_for(hh), 256>>>(wv, wqkv + 2 * hh, hh, 1.0f);
        k_scale_copy<<<grid_for(H), 256>>>(bq, bqkv, H, 0.125f);
        k_scale_copy<<<grid_for(H), 256>>>(bk, bqkv + H, H, 1.0f);
        k_scale_copy<<<grid_for(H), 256>>>(bv, bqkv + 2 * H, H, 1.0f);
        // tensor-core copy: q rows times 0.125 * log2(e) (the softmax of the tcgen05 kernels works in base 2)
        k_scale_copy<<<grid_for(hh), 256>>>(wq, wqkv_tc, hh, 0.125f * LOG2E);
        k_scale_copy<<<grid_for(2 * hh), 256>>>(wqkv + hh, wqkv_tc + hh, 2 * hh, 1.0f);
        k_scale_copy<<<grid_for(H), 256>>>(bq, bqkv_tc, H, 0.125f * LOG2E);
        k_scale_copy<<<grid_for(2 * H), 256>>>(bqkv + H, bqkv_tc + H, 2 * H, 1.0f);
        CKL();
        TRY(make_lin(h, &lw.qkv, wqkv, bqkv, 3 * H, H, wqkv_tc, bqkv_tc));
        TRY(make_lin(h, &lw.o, wo, bo, H, H));
        TRY(make_lin(h, &lw.f1, w1, b1, F, H));
        TRY(make_lin(h, &lw.f2, w2, b2, H, F));
        {   // LN1 folded into FFN-in (bf16 path, option ln1_fold): see gemm_tc.cuh EPI_FOLD_GELU
            float* bfold = nullptr;
            TRY(dev_alloc_t(h, h->owned, &bfold, static_cast<size_t>(F)));
            TRY(dev_alloc_t(h, h->owned, &lw.f1_fold_c, static_cast<size_t>(F)));
            k_fold_cols<<<grid_for(static_cast<size_t>(F) * H), 256>>>(w1, lw.ln1_g, wqkv_tc, F, H);     // scratch holds 3 H^2 >= F H floats? checked below
            CKL();
            TRY(make_lin(h, &lw.f1_fold, w1, b1, F, H, wqkv_tc, bfold));
            k_fold_rows<<<(F + 7) / 8, 256>>>(w1, lw.f1_fold.w16, lw.ln1_b, b1, lw.f1_fold_c, bfold, F, H);
            CKL();
            lw.f1_fold.fold_c = lw.f1_fold_c;
        }
    }
    h->skip.resize(nl / 2);
    for (int m = 0; m < nl / 2; ++m) {
        const std::string sp = p + "skip_connections_layers." + std::to_string(m) + ".";
        const float *w, *b;
        TRY(need(h, sp + "weight", {H, 2 * H}, &w));
        TRY(need(h, sp + "bias", {H}, &b));
        TRY(make_lin(h, &h->skip[m], w, b, H, 2 * H));
    }
    TRY(need(h, p + "proj_out.weight", {nf, H, 1}, &h->w_out));
    TRY(need(h, p + "proj_out.bias", {nf}, &h->b_out));
    if (nf <= 256) {   // tensor-core output projection operand
        TRY(dev_alloc_t(h, h->owned, &h->wout2, static_cast<size_t>(256) * H));
        TRY(dev_alloc_t(h, h->owned, &h->bout2, static_cast<size_t>(256)));
        CK(cudaMemset(h->wout2, 0, sizeof(float) * 256 * H));
        CK(cudaMemset(h->bout2, 0, sizeof(float) * 256));
        CK(cudaMemcpy(h->wout2, h->w_out, sizeof(float) * nf * H, cudaMemcpyDeviceToDevice));
        CK(cudaMemcpy(h->bout2, h->b_out, sizeof(float) * nf, cudaMemcpyDeviceToDevice));
        TRY(make_lin(h, &h->lin_out, h->wout2, h->bout2, 256, H));
    }
    CK(cudaDeviceSynchronize());
    cudaFree(wqkv_tc);
    h->finalized = true;
    return TVB_OK;
}

extern "C" int tvb_profile_read(tvb_handle* h, double* ms_by_cat, int64_t* launches_by_cat) {
    if (!h || !ms_by_cat || !launches_by_cat) return fail(TVB_ERR_INVALID, "null argument");
    CK(cudaSetDevice(h->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < PC_NCAT; ++i) { ms_by_cat[i] = 0.0; launches_by_cat[i] = 0; }
    for (size_t i = 0; i < h->ev_cat.size(); ++i) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, h->ev_pool[2 * i], h->ev_pool[2 * i + 1]));
        ms_by_cat[h->ev_cat[i]] += ms;
        launches_by_cat[h->ev_cat[i]] += 1;
        if (h->ev_cat[i] >= PC_GEMM_QKV) {              // per-shape GEMM categories also count into the GEMM total
            ms_by_cat[PC_GEMM] += ms;
            launches_by_cat[PC_GEMM] += 1;
        }
    }
    h->ev_used = 0;
    h->ev_cat.clear();
    return TVB_OK;
}

extern "C" int tvb_set_precision(tvb_handle* h, int32_t precision) {
    if (!h) return fail(TVB_ERR_INVALID, "null handle");
    if (precision != TVB_PRECISION_BF16 && precision != TVB_PRECISION_FP32_VALIDATE) return fail(TVB_ERR_INVALID, "unknown precision %d", precision);
    h->precision = precision;
    drop_graphs(h);
    return TVB_OK;
}
extern "C" int tvb_get_precision(const tvb_handle* h) { return h ? h->precision : TVB_ERR_INVALID; }
extern "C" int64_t tvb_kernel_launches(const tvb_handle* h) { return h ? h->launches : 0; }
extern "C" int64_t tvb_device_bytes(const tvb_handle* h) { return h ? static_cast<int64_t>(h->dev_bytes + h->ws_bytes) : 0; }
extern "C" int tvb_set_option(tvb_handle* h, const char* name, int32_t value) {
    if (!h || !name) return fail(TVB_ERR_INVALID, "null argument");
    drop_graphs(h);
    if (!strcmp(name, "attn_tc")) h->opt_attn_tc = value ? 1 : 0;
    else if (!strcmp(name, "conv_tc")) h->opt_conv_tc = value ? 1 : 0;
    else if (!strcmp(name, "gemm_2cta")) h->opt_gemm_2cta = value ? 1 : 0;
    else if (!strcmp(name, "gemm_mc")) h->opt_gemm_mc = value ? 1 : 0;
    else if (!strcmp(name, "gemm_bn")) h->opt_gemm_bn = value ? 1 : 0;
    else if (!strcmp(name, "resid_l2")) { if (value < 0 || value > 2) return fail(TVB_ERR_INVALID, "resid_l2 must be 0, 1 or 2"); h->opt_resid_l2 = value; }
    else if (!strcmp(name, "gemm_lab")) h->opt_gemm_lab = value & 7;
    else if (!strcmp(name, "epi_direct")) { if (value < 0 || value > 2) return fail(TVB_ERR_INVALID, "epi_direct must be 0, 1 or 2"); h->opt_epi_direct = value; }
    else if (!strcmp(name, "a_l2")) { if (value < 0 || value > 64) return fail(TVB_ERR_INVALID, "a_l2 must be 0 .. 64"); h->opt_a_l2 = value; }
    else if (!strcmp(name, "zigzag")) h->opt_zigzag = value ? 1 : 0;
    else if (!strcmp(name, "graph")) h->opt_graph = value ? 1 : 0;
    else if (!strcmp(name, "pdl")) h->opt_pdl = value ? 1 : 0;
    else if (!strcmp(name, "frame_tc")) h->opt_frame_tc = value ? 1 : 0;
    else if (!strcmp(name, "out_tc")) h->opt_out_tc = value ? 1 : 0;
    else if (!strcmp(name, "conv_halo")) h->opt_conv_halo = value ? 1 : 0;
    else if (!strcmp(name, "ln1_fold")) h->opt_ln1_fold = value ? 1 : 0;
    else if (!strcmp(name, "attn_ptmem")) h->opt_attn_ptmem = value ? 1 : 0;
    else if (!strcmp(name, "attn_v")) { if (value < 3 || value > 10) return fail(TVB_ERR_INVALID, "attn_v must be 3 .. 10"); h->opt_attn_v = value; }
    else if (!strcmp(name, "attn_win")) h->opt_attn_win = value ? 1 : 0;
    else if (!strcmp(name, "qkv_norm")) h->opt_qkv_norm = value ? 1 : 0;
    else if (!strcmp(name, "attn_sort")) { if (value < 0 || value > 2) return fail(TVB_ERR_INVALID, "attn_sort must be 0, 1 or 2"); h->opt_attn_sort = value; }
    else if (!strcmp(name, "attn_poly")) { if (value != 0 && value != 2 && value != 3 && value != 4 && !(value >= 100 && value < 108)) return fail(TVB_ERR_INVALID, "attn_poly must be 0, 2, 3 or 4"); h->opt_attn_poly = value; }
    else if (!strcmp(name, "attn_pace")) h->opt_attn_pace = value ? 1 : 0;
    else if (!strcmp(name, "ln_fuse")) { if (value < 0 || value > 2) return fail(TVB_ERR_INVALID, "ln_fuse must be 0, 1 or 2"); h->opt_ln_fuse = value; }
    else if (!strcmp(name, "profile")) { h->profile = value ? 1 : 0; h->ev_used = 0; h->ev_cat.clear(); }
    else return fail(TVB_ERR_INVALID, "unknown option %s", name);
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// plan: packed row layout + workspace
// ----------------------------------------------------------------------------------------------------------------
static int alloc_act(tvb_handle* h, Act* a, int rows, int C, bool bf) {
    a->C = C;
    a->f32 = nullptr;
    a->b16 = nullptr;
    if (bf) {
        TRY(dev_alloc_t(h, h->ws_owned, &a->b16, static_cast<size_t>(rows) * C));
        CK(cudaMemset(a->b16, 0, static_cast<size_t>(rows) * C * sizeof(bf16)));
    } else {
        TRY(dev_alloc_t(h, h->ws_owned, &a->f32, static_cast<size_t>(rows) * C));
        CK(cudaMemset(a->f32, 0, static_cast<size_t>(rows) * C * sizeof(float)));
    }
    return TVB_OK;
}
static int alloc_f32(tvb_handle* h, float** p, size_t n) {
    TRY(dev_alloc_t(h, h->ws_owned, p, n));
    CK(cudaMemset(*p, 0, n * sizeof(float)));
    return TVB_OK;
}

static void drop_graphs(tvb_handle* h) {
    for (auto& g : h->graphs) {
        if (g.exec) cudaGraphExecDestroy(g.exec);
        g = tvb_handle::PassGraph();
    }
}

static int build_plan(tvb_handle* h, int B, int T, const int64_t* lengths, int cfg_double, cudaStream_t st) {
    // same batch layout as the previous call (the usual serving loop): keep the plan, the workspace and the captured graphs
    if (h->finalized && B == h->plan_B && T == h->plan_T && cfg_double == h->plan_cfg && h->plan_precision == h->precision &&
        static_cast<int>(h->plan_lengths.size()) == B && std::equal(h->plan_lengths.begin(), h->plan_lengths.end(), lengths))
        return TVB_OK;
    drop_graphs(h);
    const tvb_config& c = h->cfg;
    if (!h->finalized) return fail(TVB_ERR_STATE, "tvb_finalize_weights has not been called");
    if (B <= 0 || T <= 0) return fail(TVB_ERR_INVALID, "B and T must be positive");
    // key positions reach the attention kernel as two bf16 terms, p = (p & ~127) + (p & 127) (attn10_tc.cuh: k_attn10_pos): exact below 2^15
    if (T + 1 > 32768) return fail(TVB_ERR_INVALID, "T = %d: sequences longer than 32767 frames (6 min 20 s at hop 256 / 22.05 kHz) are not supported", T);
    const int H = c.hidden_size, F = c.intermediate_size, nf = c.n_feats, nl = c.num_hidden_layers;
    const int GAP = 16;
    std::vector<int> off(B);
    int64_t o = 0;
    for (int b = 0; b < B; ++b) {
        if (lengths[b] < 0 || lengths[b] > T) return fail(TVB_ERR_INVALID, "lengths[%d] = %lld outside [0, T=%d]", b, (long long)lengths[b], T);
        off[b] = static_cast<int>(o);
        o += lengths[b] + 1 + GAP;
    }
    if (o > (1ll << 30)) return fail(TVB_ERR_INVALID, "batch too large");
    const int nhalf = cfg_double ? 2 : 1;
    const int64_t halign = cfg_double ? 256 : 512;          // Mp must be a multiple of 512 (two vertically adjacent CTA-pair GEMM tiles per cluster)
    const int half = static_cast<int>((o + halign - 1) / halign * halign);
    const int Mp = half * nhalf;
    const int nseq = B * nhalf;
    std::vector<int> row_seq(Mp, -1), row_pos(Mp, 0), seqinfo(4 * static_cast<size_t>(nseq));
    std::vector<int2> work16;
    std::vector<int4> work128;
    for (int s = 0; s < nseq; ++s) {
        const int b = s % B;
        const int uncond = cfg_double ? (s < B) : 0;
        const int start = (s / B) * half + off[b];
        const int len = static_cast<int>(lengths[b]) + 1;
        seqinfo[s] = start;
        seqinfo[nseq + s] = len;
        seqinfo[2 * nseq + s] = b;
        seqinfo[3 * nseq + s] = uncond;
        for (int i = 0; i < len; ++i) { row_seq[start + i] = s; row_pos[start + i] = i; }
        for (int q0 = 0; q0 < len; q0 += 16) work16.push_back(make_int2(s, q0));
        for (int q0 = 0; q0 < len; q0 += ATT_BQ) work128.push_back(make_int4(s, q0, start, len));
    }
    // heaviest query tiles first (cost ~ key blocks x query halves): the persistent attention kernel deals items round-robin,
    // the non-persistent ones are scheduled in launch order, either way the short tiles fill the tail
    std::stable_sort(work128.begin(), work128.end(), [](const int4& a, const int4& b) {
        auto cost = [](const int4& w) { return ((w.w + ATT_BK - 1) / ATT_BK) * (w.y + 128 < w.w ? 2 : 1); };
        return cost(a) > cost(b);
    });
    // ---- workspace (re)allocation
    const bool bf = h->precision == TVB_PRECISION_BF16;
    const size_t state_n = static_cast<size_t>(B) * nf * T;
    if (Mp > h->ws_cap_rows || h->ws_precision != h->precision || state_n > h->ws_cap_state || nseq > h->seq_cap ||
        (int)work16.size() > h->work16_cap || (int)work128.size() > h->work128_cap) {
        CK(cudaStreamSynchronize(st));
        CK(cudaDeviceSynchronize());
        for (void* p : h->ws_owned) cudaFree(p);
        h->ws_owned.clear();
        h->ws_bytes = 0;
        const int rows = Mp > h->ws_cap_rows ? Mp : h->ws_cap_rows;
        const size_t sn = state_n > h->ws_cap_state ? state_n : h->ws_cap_state;
        const int sc = nseq > h->seq_cap ? nseq : h->seq_cap;
        const int w16c = (int)work16.size() > h->work16_cap ? (int)work16.size() : h->work16_cap;
        const int w128c = (int)work128.size() > h->work128_cap ? (int)work128.size() : h->work128_cap;
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_row_seq, rows));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_row_pos, rows));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_seq, 4 * static_cast<size_t>(sc)));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_work16, w16c));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_work128, w128c));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_tseq, sc));
        TRY(alloc_f32(h, &h->normtab, static_cast<size_t>(rows / 32 + 1) * 2 * c.num_attention_heads));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_win, static_cast<size_t>(w128c) * c.num_attention_heads));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_pos, static_cast<size_t>(rows) * 16));
        TRY(dev_alloc_t(h, h->ws_owned, &h->d_order, static_cast<size_t>(w128c) * c.num_attention_heads));
        TRY(alloc_f32(h, &h->base, static_cast<size_t>(rows) * H));
        TRY(alloc_f32(h, &h->xpre, static_cast<size_t>(rows) * H));
        TRY(alloc_f32(h, &h->xalt, static_cast<size_t>(rows) * H));
        TRY(alloc_f32(h, reinterpret_cast<float**>(&h->statsA), 2 * static_cast<size_t>(rows)));
        TRY(alloc_f32(h, reinterpret_cast<float**>(&h->statsB), 2 * static_cast<size_t>(rows)));
        TRY(alloc_f32(h, reinterpret_cast<float**>(&h->ln_part), 32 * static_cast<size_t>(rows)));
        TRY(alloc_f32(h, &h->hres, static_cast<size_t>(rows) * H));
        TRY(alloc_f32(h, &h->vbuf, static_cast<size_t>(rows) * std::max(nf, 256)));
        TRY(alloc_f32(h, &h->z_work, sn));
        TRY(alloc_f32(h, &h->z_hat, sn));
        TRY(alloc_f32(h, &h->v_pred, sn));
        TRY(alloc_act(h, &h->hin, rows, H, false));                    // fp32 conv residual, always
        if (bf) {
            TRY(dev_alloc_t(h, h->ws_owned, &h->hin.b16, static_cast<size_t>(rows) * H));
            CK(cudaMemset(h->hin.b16, 0, static_cast<size_t>(rows) * H * sizeof(bf16)));
        }
        if (bf) TRY(alloc_act(h, &h->yop, rows, h->lin_y.K, true));
        TRY(alloc_act(h, &h->p1, rows, H, bf));
        for (int i = 0; i < nl / 2; ++i) TRY(alloc_act(h, &h->slots[i], rows, H, bf));
        TRY(alloc_act(h, &h->hbA, rows, H, bf));
        TRY(alloc_act(h, &h->hbB, rows, H, bf));
        TRY(alloc_act(h, &h->qkv, rows, 3 * H, bf));
        TRY(alloc_act(h, &h->ctx, rows, H, bf));
        TRY(alloc_act(h, &h->mid, rows, F, bf));
        h->ws_cap_rows = rows;
        h->ws_cap_state = sn;
        h->seq_cap = sc;
        h->work16_cap = w16c;
        h->work128_cap = w128c;
        h->ws_precision = h->precision;
    }
    CK(cudaMemcpyAsync(h->d_row_seq, row_seq.data(), sizeof(int) * Mp, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_row_pos, row_pos.data(), sizeof(int) * Mp, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_seq, seqinfo.data(), sizeof(int) * seqinfo.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_work16, work16.data(), sizeof(int2) * work16.size(), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(h->d_work128, work128.data(), sizeof(int4) * work128.size(), cudaMemcpyHostToDevice, st));
    k_attn10_pos<<<(Mp + 255) / 256, 256, 0, st>>>(h->d_row_seq, h->d_row_pos, Mp, h->d_pos);
    CKL();
    CK(cudaStreamSynchronize(st));   // host vectors go out of scope
    h->n_work16 = static_cast<int>(work16.size());
    h->n_work128 = static_cast<int>(work128.size());
    h->L.Mp = Mp;
    h->L.nseq = nseq;
    h->L.row_seq = h->d_row_seq;
    h->L.row_pos = h->d_row_pos;
    h->L.seq_start = h->d_seq;
    h->L.seq_len = h->d_seq + nseq;
    h->L.seq_utt = h->d_seq + 2 * nseq;
    h->L.seq_uncond = h->d_seq + 3 * nseq;
    {
        int64_t lmin = lengths[0], lmax = lengths[0];
        for (int b = 1; b < B; ++b) { lmin = std::min<int64_t>(lmin, lengths[b]); lmax = std::max<int64_t>(lmax, lengths[b]); }
        h->plan_sort = static_cast<int>(work128.size()) * c.num_attention_heads >= 2 * h->num_sms && (lmax + 1 >= 16 * ATT_BK || lmin != lmax);
    }
    h->plan_B = B; h->plan_T = T; h->plan_cfg = cfg_double; h->plan_half = half;
    h->plan_lengths.assign(lengths, lengths + B);
    h->plan_precision = h->precision;
    if (bf) {   // A-operand maps cover exactly the Mp live rows: TMA zero-fills anything outside
        TRY(make_tmap(&h->hin.tm, h->hin.b16, Mp, H, H, 128));
        TRY(make_tmap(&h->p1.tm, h->p1.b16, Mp, H, H, 128));
        for (int i = 0; i < nl / 2; ++i) TRY(make_tmap(&h->slots[i].tm, h->slots[i].b16, Mp, H, H, 128));
        TRY(make_tmap(&h->hbA.tm, h->hbA.b16, Mp, H, H, 128));
        TRY(make_tmap(&h->hbB.tm, h->hbB.b16, Mp, H, H, 128));
        TRY(make_tmap(&h->qkv.tm, h->qkv.b16, Mp, 3 * H, 3 * H, 128));
        TRY(make_tmap(&h->ctx.tm, h->ctx.b16, Mp, H, H, 128));
        TRY(make_tmap(&h->mid.tm, h->mid.b16, Mp, F, F, 128));
        TRY(make_tmap(&h->qkv.tm_out, h->qkv.b16, Mp, 3 * H, 3 * H, 32));
        TRY(make_tmap(&h->mid.tm_out, h->mid.b16, Mp, F, F, 32));
        TRY(make_tmap(&h->hbB.tm_out32, h->hbB.b16, Mp, H, H, 32, 32, false, true));
        TRY(make_tmap(&h->yop.tm, h->yop.b16, Mp, h->lin_y.K, h->lin_y.K, 128));
        TRY(make_tmap(&h->tm_pos, h->d_pos, Mp, 16, 16, 128, 16, false, false, true));
        TRY(make_tmap(&h->tm_hin_halo, h->hin.b16, Mp, H, H, CH_ROWS));
        TRY(make_tmap(&h->tm_p1_halo, h->p1.b16, Mp, H, H, CH_ROWS));
        TRY(make_tmap(&h->hin.tm_out32, h->hin.b16, Mp, H, H, 32, 32, false, true));
        TRY(make_tmap(&h->tm_base, h->base, Mp, H, H, 32, 32, true));
        TRY(make_tmap(&h->tm_vbuf, h->vbuf, Mp, 256, 256, 32, 32, true));
        TRY(make_tmap(&h->tm_hin32, h->hin.f32, Mp, H, H, 32, 32, true));
        TRY(make_tmap(&h->tm_xpre, h->xpre, Mp, H, H, 32, 32, true));
        TRY(make_tmap(&h->tm_hres, h->hres, Mp, H, H, 32, 32, true));
        TRY(make_tmap(&h->tm_xalt, h->xalt, Mp, H, H, 32, 32, true));
    }
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// kernel launch helpers
// ----------------------------------------------------------------------------------------------------------------
template <int EPI>
static int launch_gemm_tc_t(tvb_handle* h, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB, const GemmParams& p,
                            cudaStream_t st, int num_sms, int cat = PC_GEMM) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        CK(cudaFuncSetAttribute(gemm_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_SMEM_BYTES));
        attr.mark();
    }
    const int tiles = (p.M / GEMM_BM) * (p.N / GEMM_BN);
    const int grid = tiles < num_sms ? tiles : num_sms;
    ProfScope ps(h, st, cat);
    gemm_tc_kernel<EPI><<<grid, GEMM_THREADS, GEMM_SMEM_BYTES, st>>>(tmA, tmA2, tmB, p);
    if (h) h->launches++;
    CKL();
    return TVB_OK;
}
static int launch_gemm_tc(tvb_handle* h, int epi, const CUtensorMap& tmA, const CUtensorMap& tmA2, const CUtensorMap& tmB,
                          const GemmParams& p, cudaStream_t st, int num_sms, int cat = PC_GEMM) {
    if (p.M % GEMM_BM || p.N % GEMM_BN || p.K % GEMM_BK || p.K1 % GEMM_BK) return fail(TVB_ERR_INVALID, "gemm shape %dx%dx%d not tileable", p.M, p.N, p.K);
    switch (epi) {
        case EPI_BF16: return launch_gemm_tc_t<EPI_BF16>(h, tmA, tmA2, tmB, p, st, num_sms, cat);
        case EPI_GELU_BF16: return launch_gemm_tc_t<EPI_GELU_BF16>(h, tmA, tmA2, tmB, p, st, num_sms, cat);
        case EPI_RESID_F32: return launch_gemm_tc_t<EPI_RESID_F32>(h, tmA, tmA2, tmB, p, st, num_sms, cat);
        case EPI_DUAL: return launch_gemm_tc_t<EPI_DUAL>(h, tmA, tmA2, tmB, p, st, num_sms, cat);
        case EPI_F32: return launch_gemm_tc_t<EPI_F32>(h, tmA, tmA2, tmB, p, st, num_sms, cat);
        case EPI_RESID_LN: return launch_gemm_tc_t<EPI_RESID_LN>(h, tmA, tmA2, tmB, p, st, num_sms, cat);
    }
    return fail(TVB_ERR_INVALID, "unknown epilogue %d", epi);
}

template <int EPI, int BN = 256, int DIRECT = 0>
static int launch_gemm2_tc_t(tvb_handle* h, const Gemm2Maps& maps, const GemmParams& p, cudaStream_t st, int num_sms, int cat = PC_GEMM) {
    static PerDeviceOnce attr;
    if (!attr.done()) {
        CK(cudaFuncSetAttribute(gemm2_tc_kernel<EPI, BN, DIRECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
        attr.mark();
    }
    const int tiles = (p.M / 256) * (p.N / BN);
    int grid = 2 * tiles < (num_sms & ~1) ? 2 * tiles : (num_sms & ~1);
    ProfScope ps(h, st, cat);
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(G2_THREADS); cfg.dynamicSmemBytes = G2_SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = (h && h->opt_pdl) ? 1 : 0;
    CK(cudaLaunchKernelEx(&cfg, gemm2_tc_kernel<EPI, BN, DIRECT>, maps, p));
    if (h) h->launches++;
    return TVB_OK;
}
template <int EPI>
static int launch_gemm4_tc_t(tvb_handle* h, const Gemm2Maps& maps, const GemmParams& p, cudaStream_t st, int cat) {
    static PerDeviceInt max_clusters_dev;
    int max_clusters = max_clusters_dev.get();
    if (max_clusters == 0) {
        CK(cudaFuncSetAttribute(gemm4_tc_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(4 * 64);
        cfg.blockDim = dim3(G2_THREADS);
        cfg.dynamicSmemBytes = G2_SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 4; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        CK(cudaOccupancyMaxActiveClusters(&n, gemm4_tc_kernel<EPI>, &cfg));
        if (n < 1) return fail(TVB_ERR_CUDA, "no 4-CTA cluster fits on this device");
        max_clusters = n;
        max_clusters_dev.set(n);
        if (getenv("TVB_DEBUG")) fprintf(stderr, "[tvb] gemm4_tc_kernel: %d resident 4-CTA clusters\n", n);
    }
    const int tiles = (p.M / 512) * (p.N / 256);
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    ProfScope ps(h, st, cat);
    gemm4_tc_kernel<EPI><<<4 * clusters, G2_THREADS, G2_SMEM_BYTES, st>>>(maps, p);
    if (h) h->launches++;
    CKL();
    return TVB_OK;
}
// Cluster-of-4 multicast GEMM; maps.b must be the box-64 weight map.
static int launch_gemm4_tc(tvb_handle* h, int epi, const Gemm2Maps& maps, const GemmParams& p, cudaStream_t st, int cat = PC_GEMM) {
    if (p.M % 512 || p.N % 256 || p.K % GEMM_BK || p.K1 % GEMM_BK) return fail(TVB_ERR_INVALID, "gemm shape %dx%dx%d not tileable by 4-CTA clusters", p.M, p.N, p.K);
    switch (epi) {
        case EPI_BF16: return launch_gemm4_tc_t<EPI_BF16>(h, maps, p, st, cat);
        case EPI_GELU_BF16: return launch_gemm4_tc_t<EPI_GELU_BF16>(h, maps, p, st, cat);
        case EPI_RESID_F32: return launch_gemm4_tc_t<EPI_RESID_F32>(h, maps, p, st, cat);
        case EPI_DUAL: return launch_gemm4_tc_t<EPI_DUAL>(h, maps, p, st, cat);
        case EPI_F32: return launch_gemm4_tc_t<EPI_F32>(h, maps, p, st, cat);
        case EPI_RESID_LN: return launch_gemm4_tc_t<EPI_RESID_LN>(h, maps, p, st, cat);
    }
    return fail(TVB_ERR_INVALID, "unknown epilogue %d", epi);
}
// CTA-pair GEMM; maps.b must be the box-128 weight map (bn == 256) or the box-64 one (bn == 128), maps.out / maps.resid the epilogue
// maps of the destination.
static int launch_gemm2_tc(tvb_handle* h, int epi, const Gemm2Maps& maps, const GemmParams& p, cudaStream_t st, int num_sms, int cat = PC_GEMM,
                           int bn = 256, int direct = 0) {
    if (p.M % 256 || p.N % bn || p.K % GEMM_BK || p.K1 % GEMM_BK) return fail(TVB_ERR_INVALID, "gemm shape %dx%dx%d not tileable by CTA pairs", p.M, p.N, p.K);
#define TVB_G2_CASE(E) case E: return bn == 128 ? launch_gemm2_tc_t<E, 128>(h, maps, p, st, num_sms, cat) : launch_gemm2_tc_t<E, 256>(h, maps, p, st, num_sms, cat);
    // fp32-output epilogues: direct (register <-> global) or staged through shared memory
#define TVB_G2_CASE_D(E) case E: if (direct) return bn == 128 ? launch_gemm2_tc_t<E, 128, 1>(h, maps, p, st, num_sms, cat) : launch_gemm2_tc_t<E, 256, 1>(h, maps, p, st, num_sms, cat); \
                                 return bn == 128 ? launch_gemm2_tc_t<E, 128>(h, maps, p, st, num_sms, cat) : launch_gemm2_tc_t<E, 256>(h, maps, p, st, num_sms, cat);
    switch (epi) {
        TVB_G2_CASE(EPI_BF16)
        TVB_G2_CASE(EPI_GELU_BF16)
        TVB_G2_CASE_D(EPI_RESID_F32)
        TVB_G2_CASE_D(EPI_DUAL)
        TVB_G2_CASE_D(EPI_F32)
        TVB_G2_CASE_D(EPI_RESID_LN)
        TVB_G2_CASE_D(EPI_RESID_F32_ST)
        TVB_G2_CASE_D(EPI_RESID_LN_ST)
        TVB_G2_CASE(EPI_FOLD_GELU)
    }
#undef TVB_G2_CASE
#undef TVB_G2_CASE_D
    return fail(TVB_ERR_INVALID, "unknown epilogue %d", epi);
}

// Tile width of the CTA-pair GEMM for an [Mp, N] output: 256 columns, except when that tiling does not even fill one wave of
// CTA pairs and 128-column tiles finish in fewer (waves x tile width) -- the single-utterance launches (B = 1, 5 s: 24 tiles
// for N = 1024 on 74 pairs; attention-out 21.9 -> 17.8 us, FFN-out 36.0 -> 29.1 us, 84.6 -> 71.5 ms per 32-NFE solve).
// Once every pair is busy the 128-column tile is the slower one (half the MMA work per A-tile fetch: FFN-in at B = 16 ran
// 1.5 x longer with it, profiles/r01), so full waves always use 256.
static int pair_gemm_bn_rule(int num_sms, int Mp, int N) {
    if (N % 256) return 128;
    const int pairs = num_sms / 2;
    const int tiles256 = (Mp / 256) * (N / 256);
    auto cost = [&](int bn) { return (((Mp / 256) * (N / bn) + pairs - 1) / pairs) * bn; };
    return (tiles256 <= pairs && cost(128) < cost(256)) ? 128 : 256;
}
static int pair_gemm_bn(const tvb_handle* h, int Mp, int N) {
    if (N % 256 == 0 && !h->opt_gemm_bn) return 256;
    return pair_gemm_bn_rule(h->num_sms, Mp, N);
}

// out_f32 / out (bf16) = LayerNorm(a W^T + b + resid; g, beta): residual GEMM with the norm in its epilogue (bf16 path, N == 1024).
static int linear_ln(tvb_handle* h, const Lin& lin, const Act& a, const float* resid, float* out_f32, Act* out, const float* g,
                     const float* beta, cudaStream_t st, int cat) {
    static PerDeviceInt max_clusters_dev;
    int max_clusters = max_clusters_dev.get();
    if (max_clusters == 0) {
        CK(cudaFuncSetAttribute(gemmln_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GL_SMEM_BYTES));
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(8 * 64);
        cfg.blockDim = dim3(GL_THREADS);
        cfg.dynamicSmemBytes = GL_SMEM_BYTES;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeClusterDimension;
        at[0].val.clusterDim.x = 8; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        cfg.attrs = at; cfg.numAttrs = 1;
        int n = 0;
        CK(cudaOccupancyMaxActiveClusters(&n, gemmln_tc_kernel, &cfg));
        if (n < 1) return fail(TVB_ERR_CUDA, "no 8-CTA cluster of the fused GEMM+LayerNorm kernel fits on this device");
        max_clusters = n;
        max_clusters_dev.set(n);
        if (getenv("TVB_DEBUG")) fprintf(stderr, "[tvb] gemmln_tc_kernel: %d resident 8-CTA clusters\n", n);
    }
    const int Mp = h->L.Mp;
    if (lin.N != GL_N || Mp % 256 || lin.K % GEMM_BK) return fail(TVB_ERR_INVALID, "fused GEMM+LayerNorm needs N == 1024, M %% 256 == 0");
    GemmLnMaps maps;
    maps.a = a.tm; maps.b = lin.tm2;
    maps.out = (out_f32 == h->xpre) ? h->tm_xpre : (out_f32 == h->xalt) ? h->tm_xalt : h->tm_hres;
    maps.resid = (resid == h->xpre) ? h->tm_xpre : (resid == h->xalt) ? h->tm_xalt : h->tm_hres;
    GemmLnParams p{};
    p.M = Mp; p.K = lin.K; p.bias = lin.bias_tc; p.ln_g = g; p.ln_b = beta; p.eps = 1e-5f; p.out_bf16 = out->b16;
    const int tiles = Mp / 256;
    const int clusters = tiles < max_clusters ? tiles : max_clusters;
    ProfScope ps(h, st, cat);
    gemmln_tc_kernel<<<8 * clusters, GL_THREADS, GL_SMEM_BYTES, st>>>(maps, p);
    h->launches++;
    CKL();
    return TVB_OK;
}

enum LinEpi { LE_PLAIN, LE_GELU, LE_RESID, LE_DUAL, LE_RESID_ST /* + bf16 copy + row statistics */, LE_FOLD_GELU /* LN folded in */, LE_F32 /* fp32 out, no residual */ };
// LayerNorm whose fp32 result is not materialised: rows of `x` with per-row (mean, rstd) in `stats` and the norm's affine.
struct LnRef {
    const float* x = nullptr;
    const float2* stats = nullptr;
    const float *g = nullptr, *b = nullptr;
};
static const CUtensorMap& f32_map(tvb_handle* h, const float* ptr) {
    return ptr == h->xpre ? h->tm_xpre : ptr == h->xalt ? h->tm_xalt : ptr == h->base ? h->tm_base : ptr == h->hin.f32 ? h->tm_hin32 : ptr == h->vbuf ? h->tm_vbuf : h->tm_hres;
}
// y = [a | a2] W^T + b with the epilogue the call site needs, in the active precision.
//   LE_PLAIN: -> out (Act)          LE_GELU: -> gelu -> out (Act)
//   LE_RESID: -> + resid -> out_f32 (resid = fp32 rows, or LayerNorm(ln->x) re-normalised in the epilogue when ln != nullptr)
//   LE_DUAL:  -> out_f32 and out (Act)
static int linear(tvb_handle* h, const Lin& lin, const Act& a, const Act* a2, LinEpi epi, Act* out, const float* resid, float* out_f32,
                  cudaStream_t st, int cat = PC_GEMM, const LnRef* ln = nullptr, bool rev = false) {
    const int Mp = h->L.Mp;
    if (h->precision == TVB_PRECISION_BF16) {
        GemmParams p{};
        p.M = Mp; p.N = lin.N; p.K = lin.K; p.K1 = a2 ? a.C : lin.K;
        p.bias = lin.bias_tc; p.resid = ln ? ln->x : resid; p.out_f32 = out_f32; p.out_bf16 = out ? out->b16 : nullptr; p.ld_out = lin.N;
        if (ln) { p.ln_stats = ln->stats; p.ln_g = ln->g; p.ln_b = ln->b; }
        p.reverse = (rev && h->opt_zigzag) ? 1 : 0;
        p.resid_l2 = h->opt_resid_l2;
        p.lab = h->opt_gemm_lab;
        p.a_l2 = h->opt_a_l2;
        if (cat == PC_GEMM_QKV) {         // attention v9 bounds its logits with row-block norms of q and k: the CTA-pair epilogue emits them
            h->norm_from_gemm = h->opt_qkv_norm && h->opt_attn_v >= 9 && h->opt_gemm_2cta && !h->opt_gemm_mc && epi == LE_PLAIN && lin.N == 3 * h->cfg.hidden_size;
            if (h->norm_from_gemm) { p.rownorm = h->normtab; p.rownorm_cols = 2 * h->cfg.hidden_size; p.rownorm_ld = 2 * h->cfg.num_attention_heads; }
        }
        const int e = epi == LE_PLAIN ? EPI_BF16 : epi == LE_GELU ? EPI_GELU_BF16 : epi == LE_RESID ? (ln ? EPI_RESID_LN : EPI_RESID_F32)
                    : epi == LE_RESID_ST ? (ln ? EPI_RESID_LN_ST : EPI_RESID_F32_ST) : epi == LE_FOLD_GELU ? EPI_FOLD_GELU : epi == LE_F32 ? EPI_F32 : EPI_DUAL;
        const int bn = (h->opt_gemm_2cta && !h->opt_gemm_mc) ? pair_gemm_bn(h, Mp, lin.N) : 256;
        if (epi == LE_RESID_ST) { p.part = h->ln_part; p.part_n = h->ln_part_n = (bn == 128) ? 16 : 8; }
        if (epi == LE_FOLD_GELU) { p.part = h->ln_part; p.part_n = h->ln_part_n; p.stats_out = h->statsB; p.fold_c = lin.fold_c; p.fold_flag = h->err_dev + 2; }
        if ((epi == LE_RESID_ST || epi == LE_FOLD_GELU) && !(h->opt_gemm_2cta && !h->opt_gemm_mc && lin.N % 256 == 0))
            return fail(TVB_ERR_STATE, "internal: folded LayerNorm needs the CTA-pair GEMM");
        if (epi == LE_RESID_ST && lin.N != 1024) return fail(TVB_ERR_STATE, "internal: partial row statistics are laid out for N == 1024");
        if (h->opt_gemm_2cta) {
            Gemm2Maps maps;
            maps.a = a.tm; maps.a2 = a2 ? a2->tm : a.tm; maps.b = lin.tm2;
            if (epi == LE_PLAIN || epi == LE_GELU || epi == LE_FOLD_GELU) maps.out = out->tm_out;
            else maps.out = f32_map(h, out_f32);
            maps.resid = f32_map(h, p.resid);
            if (epi == LE_RESID_ST || epi == LE_DUAL) maps.out2 = out->tm_out32;
            if (h->opt_gemm_mc && Mp % 512 == 0) {
                maps.b = lin.tm3;
                return launch_gemm4_tc(h, e, maps, p, st, cat);
            }
            if (bn == 128) maps.b = lin.tm3;
            const int direct = h->opt_epi_direct == 2 || (h->opt_epi_direct == 1 && cat == PC_GEMM_F2 && e == EPI_RESID_LN);
            return launch_gemm2_tc(h, e, maps, p, st, h->num_sms, cat, bn, direct);
        }
        return launch_gemm_tc(h, e, a.tm, a2 ? a2->tm : a.tm, lin.tm, p, st, h->num_sms, cat);
    }
    if (ln) return fail(TVB_ERR_STATE, "internal: fused LayerNorm residual is a bf16-path feature");
    SgemmParams p{};
    p.M = Mp; p.N = lin.N; p.K = lin.K; p.K1 = a2 ? a.C : lin.K;
    p.A = a.f32; p.lda = a.C; p.A2 = a2 ? a2->f32 : a.f32; p.lda2 = a2 ? a2->C : a.C;
    p.W = lin.w32; p.ldw = lin.K; p.bias = lin.bias; p.resid = resid; p.ldr = lin.N; p.alpha = 1.0f;
    p.C = (epi == LE_RESID || epi == LE_DUAL) ? out_f32 : out->f32; p.ldc = lin.N;
    if (epi == LE_GELU) TRY(launch_sgemm<SG_GELU>(h, p, st));
    else if (epi == LE_RESID) TRY(launch_sgemm<SG_RESID>(h, p, st));
    else TRY(launch_sgemm<SG_PLAIN>(h, p, st));
    if (epi == LE_DUAL && out && out->f32 != out_f32)
        CK(cudaMemcpyAsync(out->f32, out_f32, sizeof(float) * static_cast<size_t>(Mp) * lin.N, cudaMemcpyDeviceToDevice, st));
    return TVB_OK;
}

// y (fp32, may be nullptr when stats != nullptr) and yb (operand copy in the active precision) = LayerNorm(x); eps 1e-5.
static int layernorm(tvb_handle* h, const float* x, float* y, Act* yb, const float* g, const float* b, cudaStream_t st,
                     float2* stats = nullptr, bool rev = false) {
    const int Mp = h->L.Mp, H = h->cfg.hidden_size;
    ProfScope ps(h, st, PC_NORM);
    if (h->precision == TVB_PRECISION_BF16)
    {
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3((Mp + 7) / 8); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 0; cfg.stream = st;
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = at; cfg.numAttrs = h->opt_pdl ? 1 : 0;
        bf16* ybp = yb ? yb->b16 : nullptr;
        const int* rs = h->L.row_seq;
        CK(cudaLaunchKernelEx(&cfg, k_layernorm<bf16>, x, y, ybp, g, b, rs, Mp, H, 1e-5f, stats, (rev && h->opt_zigzag) ? 1 : 0));
    }
    else
        k_layernorm<float><<<(Mp + 7) / 8, 256, 0, st>>>(x, y, yb ? yb->f32 : nullptr, g, b, h->L.row_seq, Mp, H, 1e-5f, stats, 0);
    h->launches++;
    CKL();
    return TVB_OK;
}

static int attention(tvb_handle* h, cudaStream_t st) {
    const int H = h->cfg.hidden_size, nh = h->cfg.num_attention_heads;
    ProfScope ps(h, st, PC_ATTN);
    // bf16 path: q carries head_dim^-0.5 * log2(e) (tvb_finalize_weights), so tensor-core scores are already in the log2 domain
    if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc && (h->opt_attn_v == 9 || h->opt_attn_v == 10)) {
        if (!h->norm_from_gemm) {         // QKV GEMM variant without the row-norm epilogue: take the norms from its output
            k_attn9_rownorm<<<(h->L.Mp + 31) / 32, 256, 0, st>>>(h->qkv.b16, h->L.Mp, H, nh, h->normtab);
            h->launches++;
            CKL();
        }
        if (h->opt_attn_v == 10)
        {
            const int sort = h->opt_attn_sort == 2 || (h->opt_attn_sort == 1 && h->plan_sort);
            CK(launch_attn10_raw(h->qkv.tm, h->tm_pos, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->normtab, h->d_win, h->d_order,
                                 h->opt_attn_win, sort, h->err_dev + 1, h->opt_attn_poly, h->num_sms, st, h->opt_pdl != 0));
            if (sort && h->opt_attn_win) h->launches++;                // k_attn10_order
        }
        else
            CK(launch_attn9_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->normtab, h->d_win, h->opt_attn_win,
                                h->err_dev + 1, h->opt_attn_poly, h->num_sms, st, h->opt_pdl != 0));
        h->launches++;                    // k_attn9_plan
    } else if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc && h->opt_attn_v == 8) {
        CK(launch_attn8_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->opt_attn_poly, h->num_sms, st, h->opt_attn_pace != 0, h->opt_pdl != 0));
    } else if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc && h->opt_attn_v == 7) {
        CK(launch_attn7_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->opt_attn_poly, st));
    } else if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc && h->opt_attn_v == 6) {
        CK(launch_attn6_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->opt_attn_poly, h->num_sms, st));
    } else if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc && h->opt_attn_v == 5) {
        CK(launch_attn5_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->opt_attn_poly, st));
    } else if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc && h->opt_attn_v == 4) {
        CK(launch_attn4_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->opt_attn_poly, st));
    } else if (h->precision == TVB_PRECISION_BF16 && h->opt_attn_tc) {
        CK(launch_attn_tc_raw(h->qkv.tm, h->ctx.b16, H, nh, h->slopes, h->d_work128, h->n_work128, h->opt_attn_ptmem != 0, 1.0f, st));
    } else if (h->precision == TVB_PRECISION_BF16) {
        dim3 grid(h->n_work16, nh);
        k_attn_simt<bf16><<<grid, 128, 0, st>>>(h->qkv.b16, h->ctx.b16, h->L, H, h->slopes, h->d_work16, 0.6931471805599453f);
    } else {
        dim3 grid(h->n_work16, nh);
        k_attn_simt<float><<<grid, 128, 0, st>>>(h->qkv.f32, h->ctx.f32, h->L, H, h->slopes, h->d_work16, 1.0f);
    }
    h->launches++;
    CKL();
    return TVB_OK;
}

static int posconv(tvb_handle* h, int i, bool last, cudaStream_t st) {
    const tvb_config& c = h->cfg;
    const int Mp = h->L.Mp, H = c.hidden_size, taps = c.convpos_width;
    const bool bf = h->precision == TVB_PRECISION_BF16;
    const bool first = (i == 0);
    ProfScope ps(h, st, PC_CONV);
    if (bf && h->opt_conv_tc) {
        // in: hin (first) or p1; out: p1 bf16 (not last) or xpre fp32 = gelu*mask + hin (last)
        if (h->opt_conv_halo && Mp % 256 == 0 && taps / 2 <= (CH_ROWS - 128) / 2 && H == c.convpos_groups * 64) {
            CK(launch_conv_halo_raw(first ? h->tm_hin_halo : h->tm_p1_halo, h->conv_tm[i], h->conv_bias[i], h->L.row_seq, Mp, H, taps, c.convpos_groups,
                                    last ? nullptr : h->p1.b16, last ? h->xpre : nullptr, last ? h->hin.f32 : nullptr, h->num_sms, st));
            h->launches++;
            return TVB_OK;
        }
        const CUtensorMap& tin = first ? h->hin.tm : h->p1.tm;
        CK(launch_conv_tc_raw(tin, h->conv_tm[i], h->conv_bias[i], h->L.row_seq, Mp, H, taps, c.convpos_groups,
                              last ? nullptr : h->p1.b16, last ? h->xpre : nullptr, last ? h->hin.f32 : nullptr, h->num_sms, st));
        h->launches++;
        return TVB_OK;
    }
    dim3 grid((Mp + PC_ROWS - 1) / PC_ROWS, c.convpos_groups);
    const size_t smem = static_cast<size_t>(PC_ROWS + taps - 1) * 64 * sizeof(float);
    if (bf) {
        const bf16* inb = first ? h->hin.b16 : h->p1.b16;
        if (last)
            k_posconv_simt<bf16, float, 1><<<grid, 256, smem, st>>>(inb, h->xpre, h->hin.f32, h->conv_wp32[i], h->conv_bias[i], h->L.row_seq, Mp, H, taps);
        else
            k_posconv_simt<bf16, bf16, 0><<<grid, 256, smem, st>>>(inb, h->p1.b16, nullptr, h->conv_wp32[i], h->conv_bias[i], h->L.row_seq, Mp, H, taps);
        h->launches++;
        CKL();
        return TVB_OK;
    }
    const float* in = first ? h->hin.f32 : h->p1.f32;
    if (last)
        k_posconv_simt<float, float, 1><<<grid, 256, smem, st>>>(in, h->xpre, h->hin.f32, h->conv_wp32[i], h->conv_bias[i], h->L.row_seq, Mp, H, taps);
    else
        k_posconv_simt<float, float, 0><<<grid, 256, smem, st>>>(in, h->p1.f32, nullptr, h->conv_wp32[i], h->conv_bias[i], h->L.row_seq, Mp, H, taps);
    h->launches++;
    CKL();
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// one estimator pass over the planned rows (networks.py:302-374). Leaves W_out h + b in vbuf[row, :].
// ----------------------------------------------------------------------------------------------------------------
static int prep_base(tvb_handle* h, const int64_t* x, const float* cond, int cond_is_zero, cudaStream_t st) {
    const tvb_config& c = h->cfg;
    const size_t smem = static_cast<size_t>(c.n_feats) * FP_ROWS * sizeof(float);
    ProfScope ps(h, st, PC_FRAME);
    k_frame_proj<true><<<h->L.Mp / FP_ROWS, 256, smem, st>>>(h->L, h->plan_T, c.n_feats, c.hidden_size, h->Wc_t, cond, x, h->G, h->b_in,
                                                             c.n_tokens, cond_is_zero, nullptr, nullptr, 0.f, nullptr, h->base, nullptr, h->err_dev);
    h->launches++;
    CKL();
    return TVB_OK;
}

static int run_estimator(tvb_handle* h, const float* y, const float* t_seq_dev, float t_uniform, cudaStream_t st) {
    const tvb_config& c = h->cfg;
    const int Mp = h->L.Mp, H = c.hidden_size, nl = c.num_hidden_layers, nf = c.n_feats;
    const bool bf = h->precision == TVB_PRECISION_BF16;
    const size_t smem = static_cast<size_t>(nf) * FP_ROWS * sizeof(float);
    if (bf && h->opt_frame_tc && h->opt_gemm_2cta && !h->opt_gemm_mc && H % 256 == 0) {
        // hin = [hi(y) | lo(y)] . [W_y | W_y]^T + base on the tensor core (the CUDA-core kernel below needs 3 GFMA per pass, ~110 us
        // of FMA pipe at best); the epilogue emits the fp32 stream and the bf16 operand copy like the attention-out GEMM does
        {
            ProfScope ps(h, st, PC_FRAME);
            k_pack_y<<<Mp / PY_ROWS, 256, static_cast<size_t>(PY_ROWS) * (nf + 1) * sizeof(float), st>>>(
                h->L, h->plan_T, nf, H, h->lin_y.K, y, t_seq_dev, t_uniform, h->freqs, h->base, h->yop.b16);
            h->launches++;
            CKL();
        }
        TRY(linear(h, h->lin_y, h->yop, nullptr, LE_RESID_ST, &h->hin, h->base, h->hin.f32, st, PC_FRAME));
    } else {
        ProfScope ps(h, st, PC_FRAME);
        k_frame_proj<false><<<Mp / FP_ROWS, 256, smem, st>>>(h->L, h->plan_T, nf, H, h->Wy_t, y, nullptr, nullptr, nullptr, 0, 0, h->base,
                                                             t_seq_dev, t_uniform, h->freqs, h->hin.f32, bf ? h->hin.b16 : nullptr);
        h->launches++;
        CKL();
    }
    for (int i = 0; i < c.convpos_depth; ++i) TRY(posconv(h, i, i == c.convpos_depth - 1, st));
    // ln_fuse (bf16 path): the fp32 post-LN stream is never written.  Each LayerNorm emits the bf16 GEMM operand and per-row
    // (mean, rstd); the residual GEMM that needs LayerNorm(x) in fp32 re-normalises x in its epilogue.  The pre-norm sums
    // ping-pong between xpre and xalt: O-proj reads xpre / writes xalt, FFN-out reads xalt / writes xpre.
    const bool epi_ln = bf && h->opt_ln_fuse == 2 && h->opt_gemm_2cta && H == GL_N;   // LayerNorm inside the residual GEMM epilogue
    const bool fuse = bf && h->opt_ln_fuse && !epi_ln;
    const bool out_tc = fuse && h->opt_out_tc && h->opt_gemm_2cta && !h->opt_gemm_mc && h->lin_out.N == 256;
    LnRef cur_ln;                      // LayerNorm that produced the current layer input (valid when fuse && !plain_resid)
    bool plain_resid = !fuse;          // the current layer input exists as fp32 rows in hres
    // LN1 folded into the FFN-in GEMM (no LN1 kernel): needs the CTA-pair GEMM epilogues
    const bool fold1 = fuse && h->opt_ln1_fold && h->opt_gemm_2cta && !h->opt_gemm_mc && c.intermediate_size % 256 == 0 && H == 1024;
    // zigzag traversal: every kernel walks the row blocks opposite to the kernel that produced its main input (L2 reuse)
    const bool ln0_rev = !fold1, skip_rev = fold1, f1_rev = !fold1, f2_rev = fold1, ln2_rev = !fold1;
    if (fuse) {
        TRY(layernorm(h, h->xpre, nullptr, &h->slots[0], h->ln0_g, h->ln0_b, st, h->statsA, ln0_rev));
        cur_ln.x = h->xpre; cur_ln.stats = h->statsA; cur_ln.g = h->ln0_g; cur_ln.b = h->ln0_b;
    } else {
        TRY(layernorm(h, h->xpre, h->hres, &h->slots[0], h->ln0_g, h->ln0_b, st));
    }

    int sp = 1;                       // skip stack pointer: slots[0..sp) are live
    const Act* cur = &h->slots[0];    // bf16/fp32 copy of the current layer input
    for (int l = 0; l < nl; ++l) {
        const LayerW& lw = h->layers[l];
        if (l >= nl / 2) {            // U-Net merge: Linear(2H -> H)([h ; skip]) as a two-operand split-K GEMM (networks.py:361-363)
            const Act* skip = &h->slots[--sp];
            // traversal directions (zigzag): LayerNorms that feed a layer run last-row-first, so the consumer of their output
            // (skip merge or QKV) runs first-row-first; a QKV fed by the skip merge runs last-row-first
            TRY(linear(h, h->skip[l - nl / 2], *cur, skip, LE_DUAL, &h->hbB, nullptr, h->hres, st, PC_GEMM_SKIP, nullptr, skip_rev));
            cur = &h->hbB;
            plain_resid = true;
        }
        TRY(linear(h, lw.qkv, *cur, nullptr, LE_PLAIN, &h->qkv, nullptr, nullptr, st, PC_GEMM_QKV, nullptr, fold1 ? (l < nl / 2) : (l >= nl / 2)));
        TRY(attention(h, st));
        const bool last = (l == nl - 1);
        if (epi_ln) {
            // hres holds the layer input (fp32, normalised); O-proj + LN1 -> xalt / hbB; FFN; FFN-out + LN2 -> hres / dst
            TRY(linear_ln(h, lw.o, h->ctx, h->hres, h->xalt, &h->hbB, lw.ln1_g, lw.ln1_b, st, PC_GEMM_O));
            TRY(linear(h, lw.f1, h->hbB, nullptr, LE_GELU, &h->mid, nullptr, nullptr, st, PC_GEMM_F1));
            Act* dst = (l < nl / 2 - 1) ? &h->slots[sp++] : &h->hbA;
            TRY(linear_ln(h, lw.f2, h->mid, h->xalt, h->hres, dst, lw.ln2_g, lw.ln2_b, st, PC_GEMM_F2));
            cur = dst;
        } else if (fuse) {
            // x1 = attn + (layer input) -> xalt ; LN1 -> hbB + statsB
            // attention wrote ctx first-row-first: O-proj walks back (rev), LN1 forward, FFN-in back, FFN-out forward, LN2 back
            if (fold1) {
                // x1 = attn + input -> xalt (fp32) and hbB (bf16, un-normalised) + partial row statistics; FFN-in consumes the raw
                // rows with W1 diag(gamma1) and applies LN1 in its epilogue (it also leaves (mean, rstd) in statsB for FFN-out)
                TRY(linear(h, lw.o, h->ctx, nullptr, LE_RESID_ST, &h->hbB, h->hres, h->xalt, st, PC_GEMM_O, plain_resid ? nullptr : &cur_ln, true));
                TRY(linear(h, lw.f1_fold, h->hbB, nullptr, LE_FOLD_GELU, &h->mid, nullptr, nullptr, st, PC_GEMM_F1, nullptr, f1_rev));
            } else {
                TRY(linear(h, lw.o, h->ctx, nullptr, LE_RESID, nullptr, h->hres, h->xalt, st, PC_GEMM_O, plain_resid ? nullptr : &cur_ln, true));
                TRY(layernorm(h, h->xalt, nullptr, &h->hbB, lw.ln1_g, lw.ln1_b, st, h->statsB, false));
                TRY(linear(h, lw.f1, h->hbB, nullptr, LE_GELU, &h->mid, nullptr, nullptr, st, PC_GEMM_F1, nullptr, f1_rev));
            }
            LnRef ln1;
            ln1.x = h->xalt; ln1.stats = h->statsB; ln1.g = lw.ln1_g; ln1.b = lw.ln1_b;
            TRY(linear(h, lw.f2, h->mid, nullptr, LE_RESID, nullptr, nullptr, h->xpre, st, PC_GEMM_F2, &ln1, f2_rev));
            Act* dst = (l < nl / 2 - 1) ? &h->slots[sp++] : &h->hbA;
            // the last layer's output feeds the fp32 output projection: materialise it
            TRY(layernorm(h, h->xpre, (last && !out_tc) ? h->hres : nullptr, dst, lw.ln2_g, lw.ln2_b, st, h->statsA, ln2_rev));
            cur_ln.x = h->xpre; cur_ln.stats = h->statsA; cur_ln.g = lw.ln2_g; cur_ln.b = lw.ln2_b;
            plain_resid = false;
            cur = dst;
        } else {
            TRY(linear(h, lw.o, h->ctx, nullptr, LE_RESID, nullptr, h->hres, h->xpre, st, PC_GEMM_O));
            TRY(layernorm(h, h->xpre, h->hres, &h->hbB, lw.ln1_g, lw.ln1_b, st));
            TRY(linear(h, lw.f1, h->hbB, nullptr, LE_GELU, &h->mid, nullptr, nullptr, st, PC_GEMM_F1));
            TRY(linear(h, lw.f2, h->mid, nullptr, LE_RESID, nullptr, h->hres, h->xpre, st, PC_GEMM_F2));
            Act* dst = (l < nl / 2 - 1) ? &h->slots[sp++] : &h->hbA;
            TRY(layernorm(h, h->xpre, h->hres, dst, lw.ln2_g, lw.ln2_b, st));
            cur = dst;
        }
    }
    if (out_tc) {
        // output projection H -> n_feats on the tensor core: A = the last LayerNorm's bf16 output, one 256-column tile (n_feats
        // live columns), fp32 result with row stride 256
        h->vbuf_ld = 256;
        return linear(h, h->lin_out, *cur, nullptr, LE_F32, nullptr, nullptr, h->vbuf, st, PC_OUT);
    }
    h->vbuf_ld = nf;
    // output projection H -> n_feats on the fp32 residual stream (networks.py:372)
    SgemmParams p{};
    p.M = Mp; p.N = nf; p.K = H; p.K1 = H;
    p.A = h->hres; p.lda = H; p.A2 = h->hres; p.lda2 = H;
    p.W = h->w_out; p.ldw = H; p.bias = h->b_out; p.resid = nullptr; p.ldr = 0; p.C = h->vbuf; p.ldc = nf; p.alpha = 1.0f;
    TRY(launch_sgemm<SG_PLAIN>(h, p, st));
    return TVB_OK;
}

// One estimator pass at uniform time t inside the ODE loop.  Small batches are launch-bound (a pass is ~230 kernels of a few
// microseconds each): the pass is captured once per state pointer into a CUDA graph that reads t from device memory and is
// replayed afterwards; large batches launch directly.
static int run_estimator_step(tvb_handle* h, const float* y, float t, cudaStream_t st) {
    const bool use_graph = h->opt_graph && !h->profile && h->L.Mp <= 16384;
    if (!use_graph) return run_estimator(h, y, nullptr, t, st);
    k_fill_f32<<<1, 128, 0, st>>>(h->d_tseq, h->L.nseq, t);
    h->launches++;
    CKL();
    tvb_handle::PassGraph* g = nullptr;
    for (auto& c : h->graphs)
        if (c.y == y) g = &c;                              // captured, or reserved by the first (direct) pass with this pointer
    if (!g) {
        for (auto& c : h->graphs)
            if (c.y == nullptr) { g = &c; break; }
        if (!g) { drop_graphs(h); g = &h->graphs[0]; }
        // first pass with this state pointer: run it directly (one-time attribute set-up happens here), capture the second
        g->y = y;
        return run_estimator(h, y, h->d_tseq, 0.f, st);
    }
    if (!g->exec) {
        const int64_t l0 = h->launches;
        cudaGraph_t graph = nullptr;
        if (!h->cap_stream) CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        CK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
        const int rc = run_estimator(h, y, h->d_tseq, 0.f, h->cap_stream);
        const cudaError_t ce = cudaStreamEndCapture(h->cap_stream, &graph);
        if (rc != TVB_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (ce != cudaSuccess) return fail(TVB_ERR_CUDA, "graph capture of the estimator pass failed: %s", cudaGetErrorString(ce));
        g->launches = h->launches - l0;
        h->launches = l0;
        const cudaError_t ie = cudaGraphInstantiate(&g->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ie != cudaSuccess) { g->exec = nullptr; return fail(TVB_ERR_CUDA, "cudaGraphInstantiate failed: %s", cudaGetErrorString(ie)); }
    }
    CK(cudaGraphLaunch(g->exec, st));
    h->launches += g->launches;
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// public compute entry points
// ----------------------------------------------------------------------------------------------------------------
extern "C" int tvb_estimator_forward(tvb_handle* h, const int64_t* x_dev, const float* y_dev, const float* cond_dev, const float* t_host,
                                     const int64_t* lengths_host, int32_t B, int32_t T, float* out_v_dev, void* stream) {
    if (!h || !x_dev || !y_dev || !cond_dev || !t_host || !lengths_host || !out_v_dev) return fail(TVB_ERR_INVALID, "null argument");
    TRY(take_async_error(h));
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    TRY(build_plan(h, B, T, lengths_host, 0, st));
    CK(cudaMemcpyAsync(h->d_tseq, t_host, sizeof(float) * B, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    TRY(prep_base(h, x_dev, cond_dev, 0, st));
    TRY(run_estimator(h, y_dev, h->d_tseq, 0.f, st));
    const size_t n = static_cast<size_t>(B) * h->cfg.n_feats * T;
    k_scatter_velocity<<<grid_for(n), 256, 0, st>>>(h->vbuf, h->vbuf_ld, h->L, h->cfg.n_feats, T, out_v_dev);
    h->launches++;
    CKL();
    return TVB_OK;
}

static int ode_update(tvb_handle* h, const OdeParams& p, cudaStream_t st) {
    const size_t n = static_cast<size_t>(p.B) * p.nf * p.T;
    ProfScope ps(h, st, PC_OUT);
    k_ode_update<<<grid_for(n), 256, 0, st>>>(h->vbuf, h->vbuf_ld, h->L, p);
    h->launches++;
    CKL();
    return TVB_OK;
}

extern "C" int tvb_sample_cfg(tvb_handle* h, const int64_t* x_dev, const float* z_dev, const float* cond_dev, const int64_t* lengths_host,
                              int32_t B, int32_t T, float t, float gradient_scale, int32_t speech_prompt, float* out_v_dev, void* stream) {
    if (!h || !x_dev || !z_dev || !cond_dev || !lengths_host || !out_v_dev) return fail(TVB_ERR_INVALID, "null argument");
    TRY(take_async_error(h));
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cfg = gradient_scale > 0.f ? 1 : 0;
    TRY(build_plan(h, B, T, lengths_host, cfg, st));
    TRY(prep_base(h, x_dev, cond_dev, speech_prompt ? 0 : 1, st));
    TRY(run_estimator(h, z_dev, nullptr, t, st));
    OdeParams p{};
    p.B = B; p.nf = h->cfg.n_feats; p.T = T; p.cfg = cfg; p.g = gradient_scale; p.dt = 0.f; p.t_next = 0.f; p.P = 0;
    p.one_minus_sigma = static_cast<float>(1.0 - static_cast<double>(h->cfg.sigma_min));
    p.z_base = z_dev; p.v_prev = nullptr; p.eps = nullptr; p.cond = cond_dev; p.z_out = h->z_hat; p.v_out = out_v_dev;
    return ode_update(h, p, st);
}

extern "C" int tvb_solve(tvb_handle* h, int32_t solver, const int64_t* x_dev, const float* z0_dev, const float* cond_dev,
                         const int64_t* lengths_host, int32_t B, int32_t T, int32_t prompt_len, int32_t n_steps, const float* t_span_host,
                         float gradient_scale, int32_t speech_prompt, const float* prompt_noise_dev, float* out_mel_dev, float* step_v_dev,
                         void* stream) {
    if (!h || !x_dev || !z0_dev || !cond_dev || !lengths_host || !t_span_host || !out_mel_dev) return fail(TVB_ERR_INVALID, "null argument");
    if (solver != TVB_SOLVER_EULER && solver != TVB_SOLVER_HEUN) return fail(TVB_ERR_INVALID, "unknown solver %d", solver);
    if (n_steps < 1) return fail(TVB_ERR_INVALID, "n_steps must be >= 1");
    if (prompt_len < 0 || prompt_len > T) return fail(TVB_ERR_INVALID, "prompt_len %d outside [0, T]", prompt_len);
    const int P = speech_prompt ? prompt_len : 0;
    if (P > 0 && !prompt_noise_dev) return fail(TVB_ERR_INVALID, "prompt_noise_dev is required when speech_prompt and prompt_len > 0");
    TRY(take_async_error(h));
    CK(cudaSetDevice(h->device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int cfg = gradient_scale > 0.f ? 1 : 0;
    TRY(build_plan(h, B, T, lengths_host, cfg, st));
    TRY(prep_base(h, x_dev, cond_dev, speech_prompt ? 0 : 1, st));
    const size_t n = static_cast<size_t>(B) * h->cfg.n_feats * T;
    CK(cudaMemcpyAsync(h->z_work, z0_dev, n * sizeof(float), cudaMemcpyDeviceToDevice, st));

    OdeParams p{};
    p.B = B; p.nf = h->cfg.n_feats; p.T = T; p.cfg = cfg; p.g = gradient_scale; p.P = P;
    p.one_minus_sigma = static_cast<float>(1.0 - static_cast<double>(h->cfg.sigma_min));
    p.cond = cond_dev;
    // fp32 host arithmetic on t / dt exactly as the reference's 0-dim fp32 tensors (voicebox.py:75, :87, :96)
    volatile float t = t_span_host[0];
    volatile float dt = t_span_host[1] - t_span_host[0];
    int noise_i = 0, eval_i = 0;
    for (int step = 1; step <= n_steps; ++step) {
        TRY(run_estimator_step(h, h->z_work, t, st));
        const float t_next = t + dt;
        const bool corrector = (solver == TVB_SOLVER_HEUN) && (step < n_steps);
        p.dt = dt; p.t_next = t_next;
        p.z_base = h->z_work; p.v_prev = nullptr;
        p.eps = P > 0 ? prompt_noise_dev + static_cast<size_t>(noise_i++) * n : nullptr;
        p.z_out = corrector ? h->z_hat : h->z_work;
        p.v_out = corrector ? h->v_pred : (step_v_dev ? step_v_dev + static_cast<size_t>(eval_i) * n : nullptr);
        TRY(ode_update(h, p, st));
        if (corrector && step_v_dev)
            CK(cudaMemcpyAsync(step_v_dev + static_cast<size_t>(eval_i) * n, h->v_pred, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
        ++eval_i;
        t = t_next;
        if (corrector) {
            TRY(run_estimator_step(h, h->z_hat, t, st));
            p.z_base = h->z_work; p.v_prev = h->v_pred;
            p.eps = P > 0 ? prompt_noise_dev + static_cast<size_t>(noise_i++) * n : nullptr;
            p.z_out = h->z_work;
            p.v_out = step_v_dev ? step_v_dev + static_cast<size_t>(eval_i) * n : nullptr;
            TRY(ode_update(h, p, st));
            ++eval_i;
        }
        if (step < n_steps) dt = t_span_host[step + 1] - t;
    }
    CK(cudaMemcpyAsync(out_mel_dev, h->z_work, n * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// test hooks
// ----------------------------------------------------------------------------------------------------------------
extern "C" int tvb_debug_pair_gemm_bn(int32_t num_sms, int32_t rows, int32_t N) {
    if (num_sms < 2 || rows <= 0 || rows % 256 || N <= 0 || N % 128) return fail(TVB_ERR_INVALID, "bad argument");
    return pair_gemm_bn_rule(num_sms, rows, N);
}

extern "C" int tvb_debug_gemm_bf16(int device, const void* a_dev, const void* w_dev, const float* bias_dev, const float* resid_dev,
                                   float* out_f32_dev, void* out_bf16_dev, int32_t M, int32_t N, int32_t K, int32_t epilogue, void* stream) {
    if (!a_dev || !w_dev) return fail(TVB_ERR_INVALID, "null argument");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(TVB_ERR_INVALID, "sm_100a device required");
    const bool mc = (epilogue & 0x200) != 0;
    const bool bn128 = (epilogue & 0x400) != 0 && !mc;          // CTA-pair kernel with 128-column tiles
    const bool pair = (epilogue & 0x100) != 0 || mc || bn128;
    CUtensorMap tmA, tmB;
    TRY(make_tmap(&tmA, a_dev, M, K, K, 128));
    TRY(make_tmap(&tmB, w_dev, N, K, K, (mc || bn128) ? 64 : pair ? 128 : GEMM_BN));
    GemmParams p{};
    p.M = M; p.N = N; p.K = K; p.K1 = K; p.bias = bias_dev; p.resid = resid_dev; p.out_f32 = out_f32_dev;
    p.out_bf16 = static_cast<bf16*>(out_bf16_dev); p.ld_out = N;
    if (pair) {
        const int e = epilogue & 0xff;
        Gemm2Maps maps;
        maps.a = tmA; maps.a2 = tmA; maps.b = tmB;
        if (e == EPI_BF16 || e == EPI_GELU_BF16) TRY(make_tmap(&maps.out, out_bf16_dev, M, N, N, 32));
        else TRY(make_tmap(&maps.out, out_f32_dev, M, N, N, 32, 32, true));
        if (e == EPI_RESID_F32) TRY(make_tmap(&maps.resid, resid_dev, M, N, N, 32, 32, true));
        else maps.resid = maps.out;
        if (e == EPI_DUAL) TRY(make_tmap(&maps.out2, out_bf16_dev, M, N, N, 32, 32, false, true));
        if (mc) return launch_gemm4_tc(nullptr, e, maps, p, static_cast<cudaStream_t>(stream));
        const int direct = (epilogue & 0x800) != 0 && (e == EPI_RESID_F32 || e == EPI_DUAL || e == EPI_F32);   // direct fp32 epilogue
        return launch_gemm2_tc(nullptr, e, maps, p, static_cast<cudaStream_t>(stream), prop.multiProcessorCount, PC_GEMM, bn128 ? 128 : 256, direct);
    }
    return launch_gemm_tc(nullptr, epilogue, tmA, tmA, tmB, p, static_cast<cudaStream_t>(stream), prop.multiProcessorCount);
}

// Kernel-level test hook: the tcgen05 attention kernels alone on caller-provided packed [rows, 3H] bf16 q|k|v
// (q pre-scaled by head_dim^-0.5 * log2(e)); sequences occupy rows [seq_start[s], seq_start[s] + seq_len[s]).
extern "C" int tvb_debug_attention(int device, const void* qkv_dev, void* ctx_dev, const int32_t* seq_start_host, const int32_t* seq_len_host,
                                   int32_t nseq, int32_t rows, int32_t H, int32_t nheads, const float* slopes_dev, int32_t version,
                                   int32_t poly, void* stream) {
    if (!qkv_dev || !ctx_dev || !seq_start_host || !seq_len_host || !slopes_dev || nseq <= 0) return fail(TVB_ERR_INVALID, "bad argument");
    if (H != nheads * 64) return fail(TVB_ERR_INVALID, "head_dim must be 64");
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(TVB_ERR_INVALID, "sm_100a device required");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    std::vector<int4> work;
    for (int s = 0; s < nseq; ++s) {
        if (seq_len_host[s] <= 0 || seq_start_host[s] < 0 || seq_start_host[s] + seq_len_host[s] > rows) return fail(TVB_ERR_INVALID, "bad sequence %d", s);
        for (int q0 = 0; q0 < seq_len_host[s]; q0 += ATT_BQ) work.push_back(make_int4(s, q0, seq_start_host[s], seq_len_host[s]));
    }
    std::stable_sort(work.begin(), work.end(), [](const int4& a, const int4& b) {
        auto cost = [](const int4& w) { return ((w.w + ATT_BK - 1) / ATT_BK) * (w.y + 128 < w.w ? 2 : 1); };
        return cost(a) > cost(b);
    });
    int4* d_work = nullptr;
    CK(cudaMalloc(&d_work, work.size() * sizeof(int4)));
    CK(cudaMemcpyAsync(d_work, work.data(), work.size() * sizeof(int4), cudaMemcpyHostToDevice, st));
    CUtensorMap tm;
    int rc = make_tmap(&tm, qkv_dev, rows, 3 * H, 3 * H, 128);
    if (rc == TVB_OK) {
        cudaError_t e;
        if (version == 9 || version == 90 || version == 10 || version == 100) {     // 90 / 100: every key block (no ALiBi window)
            float* normtab = nullptr; int4* win = nullptr; bf16* pos = nullptr; int *rseq = nullptr, *rpos = nullptr, *order = nullptr;
            e = cudaMalloc(&normtab, static_cast<size_t>(rows / 32 + 1) * 2 * nheads * sizeof(float));
            if (e == cudaSuccess) e = cudaMalloc(&win, work.size() * nheads * sizeof(int4));
            if (e == cudaSuccess) e = cudaMalloc(&pos, static_cast<size_t>(rows) * 16 * sizeof(bf16));
            if (e == cudaSuccess) e = cudaMalloc(&rseq, static_cast<size_t>(rows) * sizeof(int));
            if (e == cudaSuccess) e = cudaMalloc(&rpos, static_cast<size_t>(rows) * sizeof(int));
            if (e == cudaSuccess) e = cudaMalloc(&order, work.size() * nheads * sizeof(int));
            CUtensorMap tmp;
            if (e == cudaSuccess) {
                std::vector<int> hs(rows, -1), hp(rows, 0);
                for (int s = 0; s < nseq; ++s)
                    for (int i = 0; i < seq_len_host[s]; ++i) { hs[seq_start_host[s] + i] = s; hp[seq_start_host[s] + i] = i; }
                cudaMemcpyAsync(rseq, hs.data(), sizeof(int) * rows, cudaMemcpyHostToDevice, st);
                cudaMemcpyAsync(rpos, hp.data(), sizeof(int) * rows, cudaMemcpyHostToDevice, st);
                k_attn10_pos<<<(rows + 255) / 256, 256, 0, st>>>(rseq, rpos, rows, pos);
                cudaStreamSynchronize(st);
                if (make_tmap(&tmp, pos, rows, 16, 16, 128, 16, false, false, true) != TVB_OK) e = cudaErrorInvalidValue;
            }
            if (e == cudaSuccess) {
                k_attn9_rownorm<<<(rows + 31) / 32, 256, 0, st>>>(static_cast<const bf16*>(qkv_dev), rows, H, nheads, normtab);
                if (version >= 10 && version != 90)
                    e = launch_attn10_raw(tm, tmp, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), normtab, win, order,
                                          version == 10, 1, nullptr, poly, prop.multiProcessorCount, st);
                else
                    e = launch_attn9_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), normtab, win, version == 9, nullptr,
                                         poly, prop.multiProcessorCount, st);
            }
            cudaStreamSynchronize(st);
            cudaFree(normtab); cudaFree(win); cudaFree(pos); cudaFree(rseq); cudaFree(rpos); cudaFree(order);
        }
        else if (version == 8) e = launch_attn8_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), poly, prop.multiProcessorCount, st);
        else if (version == 7) e = launch_attn7_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), poly, st);
        else if (version == 6) e = launch_attn6_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), poly, prop.multiProcessorCount, st);
        else if (version == 5) e = launch_attn5_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), poly, st);
        else if (version == 4) e = launch_attn4_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), poly, st);
        else e = launch_attn_tc_raw(tm, static_cast<bf16*>(ctx_dev), H, nheads, slopes_dev, d_work, (int)work.size(), version != 2, 1.0f, st);
        if (e != cudaSuccess) rc = fail(TVB_ERR_CUDA, "attention launch: %s", cudaGetErrorString(e));
    }
    cudaError_t e2 = cudaStreamSynchronize(st);
    cudaFree(d_work);
    if (rc == TVB_OK && e2 != cudaSuccess) return fail(TVB_ERR_CUDA, "attention kernel: %s", cudaGetErrorString(e2));
    return rc;
}

// Test hook: device buffer (16 x 256 int64) that attention v4 launches fill with clock64() stamps of their first 16 CTAs; nullptr disables.
extern "C" int tvb_debug_set_trace(void* trace_dev) {
    attn_trace_ptr() = static_cast<long long*>(trace_dev);
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// decoder glue: stateless device functions either side of the sampler
// ----------------------------------------------------------------------------------------------------------------
extern "C" int tvb_process_unit(int device, const int64_t* units_dev, int64_t n_units, int32_t rep, int32_t hop, int64_t* out_dev,
                                int64_t n_out, void* stream) {
    if (!units_dev || !out_dev) return fail(TVB_ERR_INVALID, "null argument");
    if (rep <= 0 || hop <= 0 || n_units < 0) return fail(TVB_ERR_INVALID, "bad rep / hop / n_units");
    if (n_out != n_units * rep / hop) return fail(TVB_ERR_INVALID, "n_out must be n_units * rep / hop = %lld", (long long)(n_units * rep / hop));
    if (n_out == 0) return TVB_OK;
    CK(cudaSetDevice(device));
    k_process_unit<<<static_cast<unsigned>((n_out + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(units_dev, n_units, rep, hop, out_dev, n_out);
    CKL();
    return TVB_OK;
}

extern "C" int tvb_mel_spectrogram(int device, const float* wav_dev, int64_t n_samples, int32_t n_fft, int32_t hop, int32_t n_mels,
                                   const float* mel_basis_dev, float mean, float std, float* out_dev, int64_t n_frames, void* stream) {
    if (!wav_dev || !mel_basis_dev || !out_dev) return fail(TVB_ERR_INVALID, "null argument");
    if (hop <= 0 || hop > n_fft || (n_fft - hop) % 2 != 0 || n_mels <= 0 || std == 0.f) return fail(TVB_ERR_INVALID, "bad STFT / mel parameters");
    const int pad = (n_fft - hop) / 2;
    if (n_samples <= pad) return fail(TVB_ERR_INVALID, "signal of %lld samples is too short for reflect padding %d", (long long)n_samples, pad);
    const int64_t want = (n_samples + 2 * pad - n_fft) / hop + 1;
    if (n_frames != want) return fail(TVB_ERR_INVALID, "n_frames must be %lld", (long long)want);
    CK(cudaSetDevice(device));
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const unsigned grid = static_cast<unsigned>(n_frames);
    switch (n_fft) {
        case 256: k_mel_spectrogram<256><<<grid, 512, 0, st>>>(wav_dev, n_samples, hop, n_mels, mel_basis_dev, mean, 1.0f / std, out_dev, n_frames); break;
        case 512: k_mel_spectrogram<512><<<grid, 512, 0, st>>>(wav_dev, n_samples, hop, n_mels, mel_basis_dev, mean, 1.0f / std, out_dev, n_frames); break;
        case 1024: k_mel_spectrogram<1024><<<grid, 512, 0, st>>>(wav_dev, n_samples, hop, n_mels, mel_basis_dev, mean, 1.0f / std, out_dev, n_frames); break;
        case 2048: k_mel_spectrogram<2048><<<grid, 512, 0, st>>>(wav_dev, n_samples, hop, n_mels, mel_basis_dev, mean, 1.0f / std, out_dev, n_frames); break;
        default: return fail(TVB_ERR_INVALID, "unsupported n_fft %d (256, 512, 1024, 2048)", n_fft);
    }
    CKL();
    return TVB_OK;
}

// ----------------------------------------------------------------------------------------------------------------
// BigVGAN vocoder path (include/tvb_vocoder.h)
// ----------------------------------------------------------------------------------------------------------------
#include "vocoder_host.inl"
