// ps_cuda.cu — host side of libps_cuda.so: context, weight registry, workspace, KV cache, the operator table and the
// whole-model forward behind the C ABI declared in include/ps_cuda.h.  No torch, no CPU fallback.
#include "../../include/ps_cuda.h"
#include "ps_decode.cuh"
#include "ps_rw.cuh"
#include "ps_mv32.cuh"
#include "ps_tc.cuh"
#include "ps_step.cuh"

#include <dlfcn.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <string>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#define PS_TL_SLOTS 512
enum { PS_TP_SLOT_ATT = 0, PS_TP_SLOT_X1 = 1, PS_TP_SLOT_H = 2, PS_TP_SLOT_X2 = 3, PS_TP_SLOT_PART = 4, PS_TP_SLOT_LOGITS = 5, PS_TP_SLOTS = 6 };

// NCCL is bound at run time (dlopen) and only when a tensor-parallel context is initialised: single-GPU users need no
// NCCL, and inside a torch process the already-loaded libnccl.so.2 is reused.  Just the five entry points we call.
struct PsNcclId { char internal[128]; }; // ncclUniqueId (passed by value to ncclCommInitRank)
namespace {
struct NcclApi {
    void *lib = nullptr;
    int (*GetUniqueId)(void *) = nullptr;
    int (*CommInitRank)(void **, int, PsNcclId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
NcclApi g_nccl;
bool nccl_load() {
    if (g_nccl.lib) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
    g_nccl.GetUniqueId = (int (*)(void *))dlsym(h, "ncclGetUniqueId");
    g_nccl.CommInitRank = (int (*)(void **, int, PsNcclId, int))dlsym(h, "ncclCommInitRank");
    g_nccl.CommDestroy = (int (*)(void *))dlsym(h, "ncclCommDestroy");
    g_nccl.AllGather = (int (*)(const void *, void *, size_t, int, void *, cudaStream_t))dlsym(h, "ncclAllGather");
    g_nccl.GetErrorString = (const char *(*)(int))dlsym(h, "ncclGetErrorString");
    if (!g_nccl.GetUniqueId || !g_nccl.CommInitRank || !g_nccl.CommDestroy || !g_nccl.AllGather || !g_nccl.GetErrorString) return false;
    g_nccl.lib = h;
    return true;
}
} // namespace

namespace {

thread_local std::string g_create_error;

struct DevWeight {
    void *dev = nullptr;
    int type = 0;
    int64_t ne0 = 0, ne1 = 0;
};

struct LayerDev {
    const float *attn_norm, *ffn_norm, *q_bias, *k_bias, *v_bias;
    const uint8_t *wq, *wk, *wv, *wo, *wgate, *wup, *wdown;
    int tq, tk, tv, to, tgate, tup, tdown;
    // octet-interleaved copies for the row-walker mat-vec (ps_rw.cuh): q|k|v rows, o, gate|up slots, down
    uint8_t *rw_qkv = nullptr, *rw_o = nullptr, *rw_gu = nullptr, *rw_down = nullptr;
    // fp16-expanded tensor-core operands for the prefill GEMM (ps_tc.cuh): q|k|v rows, o, gate, up, down
    uint8_t *tc_qkv = nullptr, *tc_o = nullptr, *tc_gate = nullptr, *tc_up = nullptr, *tc_down = nullptr;
    // octet copies for the 32-block mat-vec (ps_mv32.cuh; all-Q4_0 / all-Q8_0 models): q|k|v rows, o, gate|up rows, down
    uint8_t *mv_qkv = nullptr, *mv_o = nullptr, *mv_gu = nullptr, *mv_down = nullptr;
};

} // namespace

struct ps_cuda_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    ps_cuda_model_desc d{};
    std::string err;
    std::mutex mu; // one in-flight generation per context (SURVEY 8b "Threading")
    std::unordered_map<const void *, DevWeight> weights;
    std::vector<void *> owned; // every cudaMalloc we must free
    std::unordered_set<void *> pooled; // ps_cuda_malloc blocks: stream-ordered pool (CUDABuffer intermediates come and go every forward pass)
    // model binding
    bool bound = false;
    const uint8_t *w_embd = nullptr, *w_out = nullptr;
    int t_embd = 0, t_out = 0;
    const float *w_out_norm = nullptr;
    std::vector<LayerDev> layers;
    // kv cache
    std::vector<float *> kc, vct;
    int position = 0;
    // server-side batching: independent KV sets over one set of weights; the ACTIVE session's state lives in kc / vct /
    // position / slot_mask above, the others are parked here
    struct KvSet { std::vector<float *> kc, vct; int position = 0; std::vector<uint8_t> slot_mask; };
    std::unordered_map<int, KvSet> sessions;
    int cur_session = 0, next_session = 1;
    float **sess_ptrs_dev = nullptr, **h_sess_ptrs = nullptr; // [n_layers][2][max_batch] cache pointers of the current session batch
    // speculative decode (kv_cache.hpp:97-276): per-slot mask, last batch kept aside for KVCacheInterface::copy
    std::vector<uint8_t> slot_mask;      // host copy, 1 = masked (not attended)
    bool slot_mask_dirty = true;
    uint8_t *slot_mask_dev = nullptr, *tree_dev = nullptr, *h_tree = nullptr;
    float *tree_bias = nullptr;          // [max_batch][n_ctx] attention bias of the current tree batch
    float *k_stage = nullptr, *v_stage = nullptr; // [n_layers][max_batch][kvd_l]
    float **kv_ptrs_dev = nullptr;       // [2 * n_layers]: K caches, then V^T caches
    int last_batch = 0;                  // tokens held by the staging rows
    // rope table
    float *rope_table = nullptr;
    // workspace
    int64_t maxK = 0;
    float *x = nullptr, *xn = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *qr = nullptr, *kr = nullptr, *att = nullptr;
    float *g = nullptr, *u = nullptr, *kq = nullptr, *logits = nullptr;
    const float *logits_last = nullptr; // where the last forward left its [bs][vocab] logits (ps_cuda_logits_dev)
    int logits_rows = 0;                // rows of that buffer
    float *topk_val = nullptr;          // device top-k scratch: candidates [slices][k], then the result [k]
    int *topk_idx = nullptr;
    uint32_t *aqs = nullptr, *absp = nullptr;
    float *ad = nullptr;
    int32_t *tokens_dev = nullptr, *pos_dev = nullptr, *ids_dev = nullptr;
    int32_t *h_tokens = nullptr, *h_pos = nullptr, *h_ids = nullptr; // pinned staging
    float *h_logits = nullptr;                                          // pinned staging for logits
    size_t h_logits_cap = 0;
    // options / counters
    int opt_graph = 1, opt_fused = 1, opt_pdl = 1, opt_ktime = 0, opt_tc = 1, opt_kb = 0, opt_cta_trace = 1, opt_attn_fused = 0, opt_unroll2 = 1, opt_ksplit = 0, opt_defer = 0, opt_rwm_tile = 1, opt_pv_batch_min = 3, opt_mv_kpar = 1, opt_attn_tile = 1, opt_tc_min = 6, opt_scores_batch_min = 2;
    bool tc_ok = false;        // tensor-core prefill operands are resident
    uint8_t *tc_b = nullptr;   // B operand blocks of the current activation batch
    size_t tc_b_bytes = 0;
    int *err_dev = nullptr;    // device-detected failures, [0] tcgen05 pipeline time-out, [1] peer wait gave up, [2] step-kernel barrier time-out
    int err_seen[3] = {};      // sticky host copy of the flags (counters "tc_error" / "tp_error" / "step_error")
    int *h_err = nullptr;      // pinned mirror, read after every forward / decode (a set flag becomes PS_CUDA_ERR_CUDA)
    int *tc_err_dev = nullptr; // = err_dev + 0 (counter "tc_error")
    int64_t n_tc = 0;          // tcgen05 GEMM launches (counter "tc_gemm_launches")
    std::vector<cudaEvent_t> kt_events; // option "ktime": event pairs around every row-walker mat-vec launch
    size_t kt_used = 0;
    double kt_ms = 0.0;                 // summed mat-vec kernel time of the last decode call
    int64_t kt_launches = 0;
    uint8_t *rw_out = nullptr; // lm_head, octet-interleaved
    uint8_t *tc_out = nullptr; // lm_head, fp16-expanded tensor-core operand (batches of 16+ columns that want logits: session batches, verify batches)
    bool fused_ok = false;   // every matmul weight (and the embedding) is Q4_K: the fused decode path applies
    int opt_attn_group = 0;  // opt-in: decode attention as ONE group-synchronised kernel per layer (ps_k_attn_group); bit-exact, but 23 us vs 11.7 us per layer at ctx 2048
    unsigned long long *ag_ctr = nullptr; // [n_layers][n_kv_heads] arrival counters of the kv-head groups
    float *ag_max = nullptr;              // [n_kv_heads][hs / 8][8] chunk maxima
    double *ag_sum = nullptr;             // [n_kv_heads][hs / 8][8] chunk sums
    bool mv_ok = false;      // every matmul weight is Q4_0 (or every one Q8_0): the fused 32-block decode path applies (ps_mv32.cuh)
    int mv_type = 0;
    uint8_t *mv_out = nullptr;
    bool ops_graph_ok = false; // any supported mix of weight types: the graph-replayed operator-table decode step applies (decode_step_ops)
    int opt_ops_graph = 1;
    int n_sm = 148;
    int32_t *ctr_dev = nullptr;
    // tensor parallelism (row sharding of every matrix + all-gather, bit-exact): local sizes of this rank
    int tp = 1, rank = 0;
    int nh_l = 0, nkv_l = 0, ffn_l = 0, vocab_l = 0, dim_l = 0;
    void *nccl_comm = nullptr;
    float *att_full = nullptr, *h_full = nullptr, *x_part = nullptr, *g_part = nullptr, *logits_part = nullptr;
    float *all_val = nullptr;  // gathered arg-max partials [tp][n_sm]
    int *all_idx = nullptr;
    // peer-memory exchange (fused compute + all-gather): one IPC-exported heap per rank
    //   [att_full qdim][x dim][h_full ffn][all_val tp*1024][all_idx tp*1024][flags PS_TP_SLOTS x PS_TP_MAX u32]
    uint8_t *heap = nullptr;
    size_t heap_bytes = 0, off_att = 0, off_x = 0, off_h = 0, off_val = 0, off_idx = 0, off_logits = 0, off_flags = 0;
    size_t off_ll[4] = {}; // in-band-flag mirrors of the four per-layer exchanges (ATT, X1, H, X2): 8 bytes per element
    int opt_tp_batch = 1;  // option "tp_batch": tensor-parallel batches run as batched row-sharded GEMMs (0: token by token through the fused step)
    int opt_ll = 1;        // option "tp_ll": per-layer exchanges carry their flag in-band (no fences); 0 = fence + epoch flags
    uint8_t *peer_heap[PS_TP_MAX] = {};
    bool p2p = false;          // peers imported: all-gathers run as peer stores inside the producing kernels
    uint32_t *epoch_dev = nullptr; // [PS_TP_SLOTS] local epoch counters
    int *done_dev = nullptr;       // [PS_TP_SLOTS] local CTA arrival counters
    int *tp_err_dev = nullptr;
    PsTpOut *tpo_dev = nullptr; // [2][PS_TP_SLOTS] link tables of the peer-store exchange ([1]: with the in-band-flag pointers)
    PsTpIn *tpi_dev = nullptr;
    float *tp_rows = nullptr;  // logits of a tensor-parallel batch, [max_batch][vocab] (allocated on first use)
    float *tp_tmp = nullptr, *tp_xb = nullptr, *tp_xl = nullptr, *tp_attb = nullptr, *tp_hb = nullptr, *tp_logl = nullptr; // batch forward (forward_ops_tp)
    int64_t n_gather = 0;      // all-gathers enqueued (counter "tp_allgathers")
    uint8_t *ximg = nullptr;   // Q8_K images of up to max_batch activation columns (multi-column row-walker)
    uint8_t *hq = nullptr;     // Q8_K image of the FFN hidden vector, written by the gate/up epilogue for the down mat-vec
    int *blk_cnt = nullptr;    // per-256-block arrival counters of that hand-off (rest state: zero)
    float *part_val = nullptr; // per-CTA partial arg-max of the lm_head kernel (greedy pick, stage 1)
    int *part_idx = nullptr;
    cudaGraphExec_t g_step = nullptr, g_fwd = nullptr; // one decode step (with / without the greedy pick)
    int64_t g_step_kernels = 0, g_fwd_kernels = 0;     // kernels captured in each
    int attn_clusters = -1;                            // cudaOccupancyMaxActiveClusters of the fused attention kernel (counter "attn_clusters")
    // persistent per-step kernel (ps_step.cuh)
    int opt_persist = 0;            // 1: the whole decode step as ONE persistent kernel (ps_step.cuh); bit-exact and tested, but 2x slower than the per-phase kernels so far
    int opt_l2_ahead = 96;          // stages (4736 B) per CTA the L2 look-ahead warp stays ahead of the shared-memory rings (0 = off)
    int opt_attn_chunk = 0;         // testing: cap on the soft-max positions resident in shared memory (forces the chunked attention path)
    bool step_ok = false;           // the model / geometry qualifies for ps_k_step
    PsStLayer *st_layers = nullptr; // device table
    PsStPeers *st_peers = nullptr;  // device table of the exchanged vectors on every rank
    unsigned long long *st_ll[5] = {}; // this rank's (value, epoch) vectors: x, x1, att, hq image, best
    size_t st_off[5] = {};          // tensor parallel: their offsets inside the exchange heap
    unsigned *st_sync = nullptr;    // [0] barrier counter, [1] finish counter, [2] step serial
    int smem_optin = 0, st_static = -1;
    int64_t n_step = 0;             // step-kernel launches (counter "step_launches")
    long long *trace_dev = nullptr; // debug: per-launch timeline of the fused decode step (option "trace"), PS_TL_SLOTS x 4
    int trace_launch = 0;
    int64_t n_launch = 0, n_graph = 0, h2d = 0, d2h = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr; // device timing of the last forward / decode call (stream events)
    double last_ms = 0.0;
};

namespace {

int fail(ps_cuda_ctx *c, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (c) c->err = buf;
    else g_create_error = buf;
    return code;
}

#define PS_CK(call)                                                                                          \
    do {                                                                                                     \
        cudaError_t e_ = (call);                                                                             \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, PS_CUDA_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

#define PS_LAUNCH_CK()                                                                                       \
    do {                                                                                                     \
        ctx->n_launch++;                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                                 \
        if (e_ != cudaSuccess)                                                                               \
            return fail(ctx, PS_CUDA_ERR_CUDA, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

int dev_alloc(ps_cuda_ctx *ctx, void **p, size_t bytes) {
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) return fail(ctx, PS_CUDA_ERR_OOM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    ctx->owned.push_back(*p);
    return 0;
}

// Failures detected on the device (tcgen05 pipeline time-out, a peer wait or a grid barrier that gave up) become a C-ABI
// status: the flags ride along with the call's final device-to-host copy.  Call with work enqueued; synchronises.
int sync_and_check(ps_cuda_ctx *ctx) {
    PS_CK(cudaMemcpyAsync(ctx->h_err, ctx->err_dev, 16, cudaMemcpyDeviceToHost, ctx->stream));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_err[0] | ctx->h_err[1] | ctx->h_err[2]) {
        const int tc = ctx->h_err[0], tp = ctx->h_err[1], st = ctx->h_err[2];
        for (int k = 0; k < 3; k++) ctx->err_seen[k] |= ctx->h_err[k];
        PS_CK(cudaMemsetAsync(ctx->err_dev, 0, 16, ctx->stream));
        if (ctx->ag_ctr) PS_CK(cudaMemsetAsync(ctx->ag_ctr, 0, (size_t)ctx->d.n_layers * ctx->nkv_l * 8, ctx->stream)); // a group that gave up leaves its counter mid-instance
        return fail(ctx, PS_CUDA_ERR_CUDA, "device-side failure:%s%s%s (wait site %d; results of this call are invalid)", tc ? " tcgen05 pipeline time-out" : "",
                    tp ? " tensor-parallel peer wait gave up" : "", st ? " decode-step wait gave up" : "", st);
    }
    return 0;
}

bool type_ok(int t) { return t == 0 || t == 1 || t == 2 || t == 8 || t == 12 || t == 13 || t == 14; } // F16 (1): embedding tables only
int blk_elems(int t) { return (t == 12 || t == 13 || t == 14) ? 256 : (t <= 1 ? 1 : 32); }

// ggml_rope_cache_init (libs/ggml/src/ggml.c:15342-15356) + rope_yarn (:15319-15336) at ext_factor == 0,
// freq_factors == NULL (src/backend/ggml/ggml_wrapper.cpp:104-106), evaluated with the platform libm for
// every position once.
// `ff` (optional, n_dims / 2 entries): the freq_factors of ggml_rope_cache_init - theta / ff[i0 / 2] enters rope_yarn (the
// "llama3" rope scaling of Llama-3.1 / 3.2 GGUFs, rope_freqs.weight).  The reference never passes them, so they
// are off unless ps_cuda_set_rope_freq_factors is called.
void build_rope_table(const ps_cuda_model_desc &d, std::vector<float> &t, const float *ff = nullptr) {
    const int hs = d.head_size;
    t.resize((size_t)d.n_ctx * hs);
    const float theta_scale = powf(d.rope_freq_base, -2.0f / d.rope_n_dims);
    for (int p = 0; p < d.n_ctx; p++) {
        volatile float theta = (float)(int64_t)p;
        float *cache = t.data() + (size_t)p * hs;
        for (int i0 = 0; i0 < hs; i0 += 2) {
            volatile float tx = ff ? theta / ff[i0 / 2] : (float)theta;
            volatile float th = d.rope_freq_scale * tx;
            volatile float c = cosf(th) * d.rope_attn_factor;
            volatile float s = sinf(th) * d.rope_attn_factor;
            s = s * 1.0f;
            cache[i0] = c;
            cache[i0 + 1] = s;
            theta = theta * theta_scale;
        }
    }
}

template <int TYPE, int C> size_t mm_smem(int64_t K) {
    const int64_t nb = K / PsBlk<TYPE>::ELEMS;
    return (size_t)C * (size_t)(K + nb * 4 + (PsBlk<TYPE>::ELEMS == 256 ? nb * 16 : 0));
}

template <int TYPE, int C>
int launch_mm(ps_cuda_ctx *ctx, float *dst, const uint8_t *w, int64_t K, int64_t N, int64_t bs, const float *bias, const float *residual) {
    static bool attr_set[64] = {};
    const size_t smem = mm_smem<TYPE, C>(K);
    if (smem > 48 * 1024 && !attr_set[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_matmul_q<TYPE, C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set[ctx->device] = true;
    }
    dim3 grid((unsigned)((N + 7) / 8), (unsigned)((bs + C - 1) / C));
    ps_k_matmul_q<TYPE, C><<<grid, 256, smem, ctx->stream>>>(w, K, N, bs, ctx->aqs, ctx->ad, ctx->absp, dst, bias, residual);
    PS_LAUNCH_CK();
    return 0;
}

template <int TYPE>
int launch_mm_t(ps_cuda_ctx *ctx, float *dst, const uint8_t *w, int64_t K, int64_t N, int64_t bs, const float *bias, const float *residual) {
    if (bs == 1) return launch_mm<TYPE, 1>(ctx, dst, w, K, N, bs, bias, residual);
    if (bs == 2) return launch_mm<TYPE, 2>(ctx, dst, w, K, N, bs, bias, residual);
    if (bs <= 4) return launch_mm<TYPE, 4>(ctx, dst, w, K, N, bs, bias, residual);
    return launch_mm<TYPE, 8>(ctx, dst, w, K, N, bs, bias, residual);
}

// quantise bs activation columns of K elements into the context's scratch (type of the WEIGHT decides the format)
int quantize_act(ps_cuda_ctx *ctx, int wtype, const float *x, int64_t K, int64_t bs) {
    if (K > ctx->maxK || bs > ctx->d.max_batch) return fail(ctx, PS_CUDA_ERR_INVALID, "activation %lldx%lld exceeds workspace", (long long)K, (long long)bs);
    if (wtype == 12 || wtype == 13 || wtype == 14) {
        if (K % 256) return fail(ctx, PS_CUDA_ERR_INVALID, "K=%lld is not a multiple of 256", (long long)K);
        dim3 grid((unsigned)((K / 256 + 3) / 4), (unsigned)bs);
        ps_k_quantize_q8k<<<grid, 128, 0, ctx->stream>>>(x, K, ctx->aqs, ctx->ad, ctx->absp);
    } else {
        if (K % 32) return fail(ctx, PS_CUDA_ERR_INVALID, "K=%lld is not a multiple of 32", (long long)K);
        dim3 grid((unsigned)((K / 32 + 15) / 16), (unsigned)bs);
        ps_k_quantize_q80<<<grid, 128, 0, ctx->stream>>>(x, K, ctx->aqs, ctx->ad);
    }
    PS_LAUNCH_CK();
    return 0;
}

// matmul on already-quantised activations (the reference re-quantises the same input for q, k, v and for gate, up;
// the bytes are identical, so once is enough)
int matmul_q(ps_cuda_ctx *ctx, float *dst, const uint8_t *w, int wtype, int64_t K, int64_t N, int64_t bs, const float *bias, const float *residual) {
    switch (wtype) {
    case 12: return launch_mm_t<12>(ctx, dst, w, K, N, bs, bias, residual);
    case 13: return launch_mm_t<13>(ctx, dst, w, K, N, bs, bias, residual);
    case 14: return launch_mm_t<14>(ctx, dst, w, K, N, bs, bias, residual);
    case 2: return launch_mm_t<2>(ctx, dst, w, K, N, bs, bias, residual);
    case 8: return launch_mm_t<8>(ctx, dst, w, K, N, bs, bias, residual);
    default: return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "matmul: unsupported weight type %d", wtype);
    }
}

int grid1d(int64_t n, int block = 256) { return (int)std::min<int64_t>((n + block - 1) / block, 148 * 8); }

} // namespace

// ====================================================================================================================
// Fused decode step (bs = 1): EMBED, 32 x {QKV, ATTN1, ATTN2, WO, GATE/UP, DOWN}, LM_HEAD, (ARGMAX) — 195 launches for
// Llama-3.1-8B instead of ~650, chained with programmatic dependent launch and replayed as one CUDA graph.
// ====================================================================================================================
namespace {

template <typename... KArgs, typename... Args>
int launch_k(ps_cuda_ctx *ctx, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (ctx->opt_pdl && !ctx->opt_ktime) ? 1 : 0;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, KArgs(args)...);
    ctx->n_launch++;
    if (e != cudaSuccess) return fail(ctx, PS_CUDA_ERR_CUDA, "kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// timeline slot of the next fused-path launch (nullptr when option "trace" is off)
long long *tl_slot(ps_cuda_ctx *ctx) {
    if (!ctx->trace_dev) return nullptr;
    return ctx->trace_dev + (size_t)(ctx->trace_launch++ % PS_TL_SLOTS) * 8;
}

// ---- row-walker mat-vec (ps_rw.cuh)
int rw_repack(ps_cuda_ctx *ctx, uint8_t *dst, const uint8_t *src, int64_t n_rows, int64_t K, int64_t oct0, int slot, int n_slots) {
    const int64_t chunks = ((n_rows + 7) / 8) * (K / 256) * 72;
    ps_k_rw_repack<<<(unsigned)std::min<int64_t>((chunks + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(dst, src, n_rows, K / 256, oct0, slot, n_slots);
    PS_LAUNCH_CK();
    return 0;
}

int launch_rw_impl(ps_cuda_ctx *ctx, PsRwArgs a, int epi);
int launch_rw(ps_cuda_ctx *ctx, PsRwArgs a, int epi) {
    if (!ctx->opt_ktime) return launch_rw_impl(ctx, a, epi);
    // kernel timing pass: CUDA events on the launching stream around this launch (no graph, no PDL overlap)
    while (ctx->kt_events.size() < ctx->kt_used + 2) {
        cudaEvent_t e;
        PS_CK(cudaEventCreate(&e));
        ctx->kt_events.push_back(e);
    }
    PS_CK(cudaEventRecord(ctx->kt_events[ctx->kt_used], ctx->stream));
    int rc = launch_rw_impl(ctx, a, epi);
    PS_CK(cudaEventRecord(ctx->kt_events[ctx->kt_used + 1], ctx->stream));
    ctx->kt_used += 2;
    return rc;
}
int launch_rw_impl(ps_cuda_ctx *ctx, PsRwArgs a, int epi) {
    a.tl = tl_slot(ctx);
    // per-CTA trace of one kind of launch (1 Wdown, 2 gate|up, 3 QKV, 4 Wo, 5 lm_head): slots 256 .. 256 + grid of the trace buffer
    const int kind = epi == PS_EPI_SILU ? 2 : epi == PS_EPI_RESIDUAL ? (a.xq_in ? 1 : 4) : (a.n_seg == 3 ? 3 : 5);
    a.cta_tl = (a.tl && kind == ctx->opt_cta_trace) ? ctx->trace_dev + (size_t)256 * 8 : nullptr;
    a.inv_k = (a.K & (a.K - 1)) == 0 ? 1.0 / (double)a.K : 0.0;
    const int nb = a.K / 256, rpt = (epi == PS_EPI_SILU) ? 2 : 1;
    if (nb > 4 * PS_RW_WARPS) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "row-walker matvec: K=%d too large", a.K);
    const int grid = std::min(ctx->n_sm, a.n_oct);
    const int per_cta = (a.n_oct + grid - 1) / grid;
    a.n_act = std::min(PS_RW_WARPS, per_cta);
    a.unroll2 = (ctx->opt_unroll2 && per_cta <= 8) ? 1 : 0;
    a.defer = (ctx->opt_defer >> kind) & 1;
    // K-split (ps_rw.cuh): when a CTA owns fewer octets than it has warps, ks warps share an octet - the largest ks in
    // {8, 4, 2} that still gives every octet of the CTA its own group; lm_head and gate|up have plenty of octets per CTA.
    int ksplit = 0;
    if (ctx->opt_ksplit > 1 && rpt == 1 && !a.part_val)
        for (int ks = 8; ks >= 2 && !ksplit; ks >>= 1)
            if (ks <= ctx->opt_ksplit && per_cta <= PS_RW_WARPS / ks) ksplit = ks;
    // stage = kb octet-blocks: the largest divisor of the row's blocks (<= 16 blocks, option "rw_kb") that leaves every warp
    // a ring of >= 2 stages (one stage is enough when it holds the warp's whole stream); a stage boundary (mbarrier wait +
    // re-arm) costs the consuming warp ~150-250 cycles, so fewer is better - but never at the price of resident blocks.
    // K-split: the row's stages are dealt round-robin to ks warps (stage count divisible by ks), every stage is one token hop.
    const size_t unit = (size_t)rpt * PS_RW_OCTET_BLOCK, budget = 208 * 1024;
    const int kb_cap0 = std::max(1, (ctx->opt_kb > 0 ? ctx->opt_kb : 16) / rpt);
    size_t fixed = 0;
    int kb = 0, ns = 0;
    for (; !kb; ksplit = 0) { // second trip: without K-split
        a.ksplit = ksplit;
        const int n_grp = ksplit ? std::min(PS_RW_WARPS / ksplit, per_cta) : std::min(PS_RW_WARPS, per_cta);
        const int kdiv = std::max(1, ksplit);
        a.n_act = n_grp * kdiv;
        const int rounds = (per_cta + n_grp - 1) / n_grp; // octets per warp (group)
        int best_res = 0;
        const int kb_cap = ksplit ? std::min(kb_cap0, PS_RW_KS_KB) : kb_cap0;
        for (int c = std::min(nb, kb_cap); c >= 1; c--) {
            if (nb % c || (nb / c) % kdiv) continue;
            const size_t fx = (size_t)a.K + (size_t)nb * 32;
            if (fx >= budget) continue;
            const int total = rounds * (nb / c / kdiv); // stages of the busiest warp
            const int n = std::min(std::min((int)((budget - fx) / ((size_t)a.n_act * ((size_t)c * unit + 8))), PS_RW_MAX_NS), total);
            if (n < 1 || (n < 2 && n < total)) continue;
            if (!ksplit) { kb = c; ns = n; fixed = fx; break; }
            if (std::min(n * c, total * c) > best_res) { best_res = std::min(n * c, total * c); kb = c; ns = n; fixed = fx; } // K-split: most resident blocks first, then the larger stage
        }
        if (!ksplit) break;
    }
    if (!kb) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "row-walker matvec: no room for a 2-stage ring (K=%d)", a.K);
    const size_t stage = (size_t)kb * unit;
    a.kb = kb;
    a.ns = ns;
    const size_t smem = fixed + (size_t)a.n_act * ns * (stage + 8);
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_rw_matvec<PS_EPI_STORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_rw_matvec<PS_EPI_RESIDUAL>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_rw_matvec<PS_EPI_SILU>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_rw_matvec<PS_EPI_STORE, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_rw_matvec<PS_EPI_RESIDUAL, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
        attr[ctx->device] = true;
    }
    if (a.ksplit) {
        if (epi == PS_EPI_RESIDUAL) return launch_k(ctx, ps_k_rw_matvec<PS_EPI_RESIDUAL, true>, dim3(grid), dim3(PS_RW_THREADS + 32), smem, a);
        return launch_k(ctx, ps_k_rw_matvec<PS_EPI_STORE, true>, dim3(grid), dim3(PS_RW_THREADS + 32), smem, a);
    }
    if (epi == PS_EPI_SILU) return launch_k(ctx, ps_k_rw_matvec<PS_EPI_SILU>, dim3(grid), dim3(PS_RW_THREADS + 32), smem, a);
    if (epi == PS_EPI_RESIDUAL) return launch_k(ctx, ps_k_rw_matvec<PS_EPI_RESIDUAL>, dim3(grid), dim3(PS_RW_THREADS + 32), smem, a);
    return launch_k(ctx, ps_k_rw_matvec<PS_EPI_STORE>, dim3(grid), dim3(PS_RW_THREADS + 32), smem, a);
}

// the four mat-vecs of a layer + lm_head on the row-walker kernel
// peer-store all-gather links of one phase (`slot`), resident in device memory (built once by ps_cuda_tp_import): where
// the producer's rows land on every rank, and what the consumer waits for.  Null when the exchange is not peer-to-peer.
bool tp_ll(ps_cuda_ctx *ctx, int slot) { return ctx->p2p && ctx->opt_ll && slot <= PS_TP_SLOT_X2; }
const PsTpOut *tp_out(ps_cuda_ctx *ctx, int slot) { return ctx->p2p ? ctx->tpo_dev + (tp_ll(ctx, slot) ? PS_TP_SLOTS : 0) + slot : nullptr; }
const PsTpIn *tp_in(ps_cuda_ctx *ctx, int slot) { return (ctx->p2p && !tp_ll(ctx, slot)) ? ctx->tpi_dev + slot : nullptr; }
// consumer side of an in-band-flag exchange: the (value, epoch) mirror of the gathered vector replaces `x` and the flag wait
void tp_ll_in(ps_cuda_ctx *ctx, int slot, PsRwArgs &a) {
    if (!tp_ll(ctx, slot)) return;
    a.x_ll = reinterpret_cast<const unsigned long long *>(ctx->heap + ctx->off_ll[slot]);
    a.x_epoch = ctx->epoch_dev + slot;
    a.tp_err = ctx->tp_err_dev;
}

// all-gather over the tensor-parallel group (NCCL on the context stream; capturable into the decode graph)
int tp_all_gather(ps_cuda_ctx *ctx, const void *send, void *recv, size_t count, bool is_int = false, bool always = false) {
    if (ctx->tp == 1 || (ctx->p2p && !always)) return 0; // p2p: the producing kernel already stored into every rank
    if (!ctx->nccl_comm) return fail(ctx, PS_CUDA_ERR_INVALID, "tensor-parallel context without ps_cuda_tp_init");
    const int rc = g_nccl.AllGather(send, recv, count, is_int ? 2 /* ncclInt32 */ : 7 /* ncclFloat32 */, ctx->nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, PS_CUDA_ERR_CUDA, "ncclAllGather failed: %s", g_nccl.GetErrorString(rc));
    ctx->n_gather++;
    return 0;
}

// q, k, v in one launch; the epilogue applies ROPE to q and k and writes k / v straight into the KV cache at pos_dev[0]
// (tensor parallel: this rank's heads only)
int rw_qkv(ps_cuda_ctx *ctx, const LayerDev &ld, int L) {
    const ps_cuda_model_desc &d = ctx->d;
    const int qdim = ctx->nh_l * d.head_size, kvd = ctx->nkv_l * d.head_size;
    PsRwArgs a{};
    a.tpi = L > 0 ? tp_in(ctx, PS_TP_SLOT_X2) : nullptr;
    if (L > 0) tp_ll_in(ctx, PS_TP_SLOT_X2, a);
    a.w = ld.rw_qkv; a.n_oct = (qdim + 2 * kvd) / 8; a.K = d.dim; a.n_seg = 3;
    a.seg[0] = {ctx->q, d.qkv_bias ? ld.q_bias : nullptr, 0, qdim, PS_RW_OUT_ROPE};
    a.seg[1] = {ctx->kc[L], d.qkv_bias ? ld.k_bias : nullptr, qdim, qdim + kvd, PS_RW_OUT_ROPE_KCACHE};
    a.seg[2] = {ctx->vct[L], d.qkv_bias ? ld.v_bias : nullptr, qdim + kvd, qdim + 2 * kvd, PS_RW_OUT_VCACHE_T};
    a.x = ctx->x; a.norm_w = ld.attn_norm; a.eps = d.norm_eps;
    a.pos_dev = ctx->pos_dev; a.rope_table = ctx->rope_table; a.hs = d.head_size; a.kvd = kvd; a.n_ctx = d.n_ctx;
    return launch_rw(ctx, a, PS_EPI_STORE);
}

// one fused attention kernel per layer: clusters of 8 CTAs, one per (kv head, 64 output dims)
template <int R2, int STEPS> int launch_attn_fused_s(ps_cuda_ctx *ctx, int L, bool *done) {
    const ps_cuda_model_desc &d = ctx->d;
    const int hs = d.head_size, nkv = ctx->nkv_l, cl = 8;
    const size_t row = (size_t)((d.n_ctx + 31) & ~31) * 4;
    const int s_cap = (((d.n_ctx + cl - 1) / cl + 7) & ~7) + 8;
    const size_t smem = (size_t)(R2 + 8) * row + (size_t)R2 * s_cap * 4;
    if (smem > 212 * 1024 || d.max_batch < 2) return 0; // rows too long for shared memory: the two-kernel path streams them
    const dim3 grid((unsigned)cl, (unsigned)(hs / 64), (unsigned)nkv);
    static int ok[64] = {}; // per device: 0 unknown, 1 usable, -1 not
    if (ok[ctx->device] == 0) {
        cudaError_t e = cudaFuncSetAttribute(ps_k_attn_fused<R2, STEPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 212 * 1024);
        int n_clusters = 0;
        if (e == cudaSuccess) {
            cudaLaunchConfig_t qc{};
            qc.gridDim = grid;
            qc.blockDim = dim3(PS_AF_THREADS);
            qc.dynamicSmemBytes = 212 * 1024;
            cudaLaunchAttribute qa[1];
            qa[0].id = cudaLaunchAttributeClusterDimension;
            qa[0].val.clusterDim.x = (unsigned)cl; qa[0].val.clusterDim.y = 1; qa[0].val.clusterDim.z = 1;
            qc.attrs = qa; qc.numAttrs = 1;
            e = cudaOccupancyMaxActiveClusters(&n_clusters, ps_k_attn_fused<R2, STEPS>, &qc);
        }
        if (e != cudaSuccess) cudaGetLastError();
        ok[ctx->device] = (e == cudaSuccess && n_clusters >= 8) ? 1 : -1; // fewer co-resident clusters than that would serialise the layer
        ctx->attn_clusters = n_clusters;
    }
    if (ok[ctx->device] < 0) return 0;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(PS_AF_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)cl; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = (ctx->opt_pdl && !ctx->opt_ktime) ? 2 : 1;
    const float kq_scale = 1.0f / sqrtf((float)hs);
    // ctx->kq ([n_heads][max_batch][n_ctx]) doubles as the clusters' exponential rows: [hs / 64][kv heads][R2][n_ctx]
    cudaError_t e = cudaLaunchKernelEx(&cfg, ps_k_attn_fused<R2, STEPS>, ctx->att, (const float *)ctx->kc[L], (const float *)ctx->vct[L], (const float *)ctx->q, ctx->kq,
                                       (const int32_t *)ctx->pos_dev, nkv, d.n_ctx, kq_scale, s_cap, tl_slot(ctx), tp_out(ctx, PS_TP_SLOT_ATT));
    ctx->n_launch++;
    if (e != cudaSuccess) return fail(ctx, PS_CUDA_ERR_CUDA, "fused attention launch failed: %s", cudaGetErrorString(e));
    if (ctx->trace_dev) tl_slot(ctx); // keep the timeline's six slots per layer (the second attention slot stays empty)
    *done = true;
    return 0;
}
template <int R2> int launch_attn_fused(ps_cuda_ctx *ctx, int L, bool *done) {
    *done = false;
    if (!ctx->opt_attn_fused || ctx->d.n_ctx % 8) return 0;
    if (ctx->d.head_size == 64) return launch_attn_fused_s<R2, 2>(ctx, L, done);
    if (ctx->d.head_size == 128) return launch_attn_fused_s<R2, 4>(ctx, L, done);
    return 0;
}

// R2 = the template's query heads per kv head, r2 <= R2 the model's (Qwen2-0.5B has 7); q_rot = the rotated query vector
template <int R2> int launch_attn(ps_cuda_ctx *ctx, int L, int r2 = R2, const float *q_rot = nullptr) {
    const ps_cuda_model_desc &d = ctx->d;
    const int hs = d.head_size, nkv = ctx->nkv_l;
    const float kq_scale = 1.0f / sqrtf((float)hs);
    if (!q_rot) q_rot = ctx->q;
    if (r2 == R2 && q_rot == ctx->q) {
        bool done = false;
        int rc0 = launch_attn_fused<R2>(ctx, L, &done);
        if (rc0 || done) return rc0;
    }
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_attn2<R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_attn_group<R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 215 * 1024));
        attr[ctx->device] = true;
    }
    int rc;
    {   // one group-synchronised kernel per layer when every CTA of every kv-head group is resident at once
        const int NC = hs / 8;
        const size_t row_b = (size_t)((d.n_ctx + 31) & ~31) * 4;
        const int chunk_cap = (((d.n_ctx + NC - 1) / NC) + 7) & ~7;
        const size_t smem_g = (size_t)(R2 + 8) * row_b + (size_t)R2 * chunk_cap * 4;
        if (ctx->opt_attn_group && ctx->ag_ctr && ctx->tp == 1 && hs % 8 == 0 && NC >= 1 && NC * nkv <= ctx->n_sm && d.n_ctx % 4 == 0 && smem_g <= 212 * 1024) {
            if ((rc = launch_k(ctx, ps_k_attn_group<R2>, dim3((unsigned)NC, (unsigned)nkv), dim3(PS_AG_THREADS), smem_g, ctx->att, (const float *)ctx->kc[L],
                               (const float *)ctx->vct[L], q_rot, ctx->kq, (const int32_t *)ctx->pos_dev, hs, nkv, d.n_ctx, kq_scale, r2, ctx->ag_ctr + (size_t)L * nkv,
                               ctx->ag_max, ctx->ag_sum, ctx->err_dev + 2, chunk_cap, tl_slot(ctx)))) return rc;
            if (ctx->trace_dev) tl_slot(ctx); // keep the timeline's six slots per layer
            return 0;
        }
    }
    {
        auto *k1 = (hs == 128) ? ps_k_attn1<R2, 4> : (hs == 64) ? ps_k_attn1<R2, 2> : ps_k_attn1<R2, 8>;
        if ((rc = launch_k(ctx, k1, dim3((unsigned)(ctx->n_sm * (R2 <= 4 ? 4 : 2))), dim3(128), 0, ctx->kq, (const float *)ctx->kc[L], q_rot,
                           (const int32_t *)ctx->pos_dev, hs, nkv, d.n_ctx, kq_scale, tl_slot(ctx), r2))) return rc;
    }
    // probabilities of the group + (when they fit) the CTA's eight V^T rows, all sized for a full context
    const size_t row = (size_t)((d.n_ctx + 31) & ~31) * 4;
    const int v_smem = (R2 + 8) * row <= 200 * 1024 && d.n_ctx % 4 == 0;
    const size_t a2smem = (size_t)(R2 + (v_smem ? 8 : 0)) * row;
    if ((size_t)R2 * row > 200 * 1024 || d.n_ctx % 4) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "fused decode attention: n_ctx = %d does not fit shared memory", d.n_ctx);
    return launch_k(ctx, ps_k_attn2<R2>, dim3((unsigned)((hs + 7) / 8), (unsigned)nkv), dim3(PS_A2_THREADS), a2smem, ctx->att, (const float *)ctx->kq,
                    (const float *)ctx->vct[L], (const int32_t *)ctx->pos_dev, hs, d.n_ctx, tl_slot(ctx),
                    tp_out(ctx, PS_TP_SLOT_ATT), v_smem, r2);
}
int launch_attn_any(ps_cuda_ctx *ctx, int L, const float *q_rot = nullptr) {
    const int r2 = ctx->d.n_heads / ctx->d.n_kv_heads;
    if (r2 == 1) return launch_attn<1>(ctx, L, r2, q_rot);
    if (r2 == 2) return launch_attn<2>(ctx, L, r2, q_rot);
    if (r2 <= 4) return launch_attn<4>(ctx, L, r2, q_rot);
    return launch_attn<8>(ctx, L, r2, q_rot);
}
int rw_single(ps_cuda_ctx *ctx, const uint8_t *w, int n_rows, int K, float *dst, const float *x, const float *norm_w, const float *residual,
              bool partial_argmax = false, const uint8_t *xq_in = nullptr, const float *next_norm_w = nullptr, int idx_offset = 0,
              const PsTpIn *tpi = nullptr, const PsTpOut *tpo = nullptr, int in_slot = -1) {
    PsRwArgs a{};
    a.tpi = tpi; a.tpo = tpo;
    if (in_slot >= 0) tp_ll_in(ctx, in_slot, a);
    a.xq_in = xq_in;
    a.next_norm_w = next_norm_w; a.next_norm_n = ctx->d.dim;
    if (partial_argmax) { a.part_val = ctx->part_val; a.part_idx = ctx->part_idx; a.idx_offset = idx_offset; }
    a.w = w; a.n_oct = (n_rows + 7) / 8; a.K = K; a.n_seg = 1;
    a.seg[0] = {dst, nullptr, 0, n_rows, 0};
    a.x = x; a.norm_w = norm_w; a.eps = ctx->d.norm_eps; a.residual = residual;
    return launch_rw(ctx, a, residual ? PS_EPI_RESIDUAL : PS_EPI_STORE);
}
int rw_gate_up(ps_cuda_ctx *ctx, const LayerDev &ld) {
    const ps_cuda_model_desc &d = ctx->d;
    PsRwArgs a{};
    a.w = ld.rw_gu; a.n_oct = (ctx->ffn_l + 7) / 8; a.K = d.dim; a.n_seg = 1;
    a.seg[0] = {ctx->g_part, nullptr, 0, ctx->ffn_l, 0};
    a.x = ctx->x; a.norm_w = ld.ffn_norm; a.eps = d.norm_eps;
    a.tpi = tp_in(ctx, PS_TP_SLOT_X1);
    tp_ll_in(ctx, PS_TP_SLOT_X1, a);
    a.tpo = tp_out(ctx, PS_TP_SLOT_H);
    if (ctx->tp == 1) { a.xq_out = ctx->hq; a.blk_cnt = ctx->blk_cnt; } // the Q8_K hand-off needs the whole vector on one GPU
    return launch_rw(ctx, a, PS_EPI_SILU);
}

// ---- multi-column row-walker (prefill chunks / verify batches)
template <int C> int launch_rwm_c(ps_cuda_ctx *ctx, PsRwmArgs a) {
    const int nb = a.K / 256;
    int kb = 2;
    while (nb % kb) kb >>= 1;
    a.kb = kb;
    a.n_act = PS_RW_WARPS;
    const size_t stage = (size_t)kb * PS_RW_OCTET_BLOCK, fixed = (size_t)C * ((size_t)a.K + (size_t)nb * 32);
    const size_t budget = 212 * 1024;
    if (fixed + (size_t)PS_RW_WARPS * 2 * (stage + 8) > budget) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "row-walker matmul: K=%d does not fit", a.K);
    a.ns = (int)std::min<size_t>(PS_RW_MAX_NS, (budget - fixed) / ((size_t)PS_RW_WARPS * (stage + 8)));
    const size_t smem = fixed + (size_t)PS_RW_WARPS * a.ns * (stage + 8);
    // the smallest tile (octets per unit) whose units still fit into one wave of CTAs: fewer walking warps per scheduler
    const int n_cg = (a.bs + C - 1) / C;
    int tile = PS_RW_WARPS;
    while (ctx->opt_rwm_tile && tile > 2 && ((a.n_oct + tile / 2 - 1) / (tile / 2)) * n_cg <= ctx->n_sm) tile >>= 1;
    a.tile = tile;
    const int n_units = n_cg * ((a.n_oct + tile - 1) / tile);
    const int grid = std::min(ctx->n_sm, n_units);
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_rw_matmul<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
        attr[ctx->device] = true;
    }
    ps_k_rw_matmul<C><<<grid, PS_RW_THREADS + 32, smem, ctx->stream>>>(a);
    PS_LAUNCH_CK();
    return 0;
}

// quantise bs columns of K floats into Q8_K images (ctx->ximg), then dst = W . x on the octet-interleaved weights
int rwm_quantize(ps_cuda_ctx *ctx, const float *x, int K, int bs) {
    ps_k_rw_quant_img<<<dim3((unsigned)((K / 256 + 3) / 4), (unsigned)bs), 128, 0, ctx->stream>>>(x, K, ctx->ximg);
    PS_LAUNCH_CK();
    return 0;
}
int launch_rwm(ps_cuda_ctx *ctx, PsRwmArgs a) {
    a.x_img = ctx->ximg;
    const size_t img = (size_t)a.K + (size_t)(a.K / 256) * 32;
    if (16 * img <= 100 * 1024 && a.bs > 8) return launch_rwm_c<16>(ctx, a);
    if (8 * img <= 140 * 1024 && a.bs > 4) return launch_rwm_c<8>(ctx, a);
    return launch_rwm_c<4>(ctx, a);
}
int rwm_single(ps_cuda_ctx *ctx, const uint8_t *w, int n_rows, int K, int slot, int n_slots, float *dst, int bs, const float *bias, const float *residual) {
    PsRwmArgs a{};
    a.w = w; a.n_oct = (n_rows + 7) / 8; a.K = K; a.slot = slot; a.n_slots = n_slots; a.n_seg = 1; a.bs = bs;
    a.seg[0] = {dst, bias, 0, n_rows, 0};
    a.residual = residual;
    return launch_rwm(ctx, a);
}

// one decode step on the token in tokens_dev[0] at position pos_dev[0]; `pick`: run
// the greedy pick + bookkeeping.  Tensor parallel (ctx->tp > 1): every matrix is ROW-sharded, so each dot product keeps
// its full K and the arithmetic stays bit-identical to one GPU; the sharded outputs are exchanged with all-gathers
// (attention output, the two residual updates, the FFN hidden vector: 4 per layer) instead of K-split all-reduces.
int decode_step_fused(ps_cuda_ctx *ctx, bool lm_head, bool pick) {
    const ps_cuda_model_desc &d = ctx->d;
    const int dim = d.dim, hs = d.head_size, nh = d.n_heads, nkv = d.n_kv_heads, qdim = nh * hs, ffn = d.ffn_dim;
    const int tp = ctx->tp, rank = ctx->rank, dim_l = ctx->dim_l;
    int rc;
    ctx->trace_launch = 0;
    if ((rc = launch_k(ctx, ps_k_embed_dev, dim3(std::max(1, dim / 256)), dim3(256), 0, ctx->x, ctx->w_embd, ctx->t_embd, (int64_t)dim, ctx->tokens_dev, tl_slot(ctx)))) return rc;
    for (int L = 0; L < d.n_layers; L++) {
        const LayerDev &ld = ctx->layers[L];
        if ((rc = rw_qkv(ctx, ld, L))) return rc;
        if ((rc = launch_attn_any(ctx, L))) return rc;
        if ((rc = tp_all_gather(ctx, ctx->att, ctx->att_full, (size_t)qdim / tp))) return rc;
        const float *norm_after = (L + 1 < d.n_layers) ? ctx->layers[L + 1].attn_norm : ctx->w_out_norm;
        // x[rows of this rank] += Wo[rows] . att
        // in-band-flag exchange: the gathered x only exists as (value, epoch) words, so the residual input is this rank's own
        // plain copy - x_part, updated in place (layer 0: the embedding row every rank computed for itself)
        const float *res_o = (tp_ll(ctx, PS_TP_SLOT_X2) && L > 0) ? ctx->x_part : ctx->x + (size_t)rank * dim_l;
        const float *res_d = tp_ll(ctx, PS_TP_SLOT_X1) ? ctx->x_part : ctx->x + (size_t)rank * dim_l;
        if ((rc = rw_single(ctx, ld.rw_o, dim_l, qdim, ctx->x_part, ctx->att_full, nullptr, res_o, false, nullptr, ld.ffn_norm, 0,
                            tp_in(ctx, PS_TP_SLOT_ATT), tp_out(ctx, PS_TP_SLOT_X1), PS_TP_SLOT_ATT))) return rc;
        if ((rc = tp_all_gather(ctx, ctx->x_part, ctx->x, (size_t)dim_l))) return rc;
        if ((rc = rw_gate_up(ctx, ld))) return rc;                                                        // g = silu(Wg.xn) * (Wu.xn)
        if ((rc = tp_all_gather(ctx, ctx->g_part, ctx->h_full, (size_t)ctx->ffn_l))) return rc;
        if ((rc = rw_single(ctx, ld.rw_down, dim_l, ffn, ctx->x_part, ctx->h_full, nullptr, res_d, false,
                            tp == 1 ? ctx->hq : nullptr, norm_after, 0, tp_in(ctx, PS_TP_SLOT_H),
                            tp_out(ctx, PS_TP_SLOT_X2), PS_TP_SLOT_H))) return rc;     // x[rows] += Wdown[rows] . g
        if ((rc = tp_all_gather(ctx, ctx->x_part, ctx->x, (size_t)dim_l))) return rc;
    }
    if (lm_head) {
        const int n_part = std::min(ctx->n_sm, (ctx->vocab_l + 7) / 8);
        if ((rc = rw_single(ctx, ctx->rw_out, ctx->vocab_l, dim, ctx->logits_part, ctx->x, ctx->w_out_norm, nullptr, pick, nullptr, nullptr,
                            rank * ctx->vocab_l, tp_in(ctx, PS_TP_SLOT_X2),
                            pick ? tp_out(ctx, PS_TP_SLOT_PART) : tp_out(ctx, PS_TP_SLOT_LOGITS), PS_TP_SLOT_X2))) return rc;
        if (pick) {
            const float *pv = ctx->part_val;
            const int *pi = ctx->part_idx;
            if (tp > 1) {
                if ((rc = tp_all_gather(ctx, ctx->part_val, ctx->all_val, (size_t)n_part))) return rc;
                if ((rc = tp_all_gather(ctx, ctx->part_idx, ctx->all_idx, (size_t)n_part, true))) return rc;
                pv = ctx->all_val; pi = ctx->all_idx;
            }
            if ((rc = launch_k(ctx, ps_k_argmax_step, dim3(1), dim3(256), 0, pv, pi, n_part * tp, ctx->ids_dev, ctx->ctr_dev, ctx->tokens_dev,
                               ctx->pos_dev, tl_slot(ctx), tp_in(ctx, PS_TP_SLOT_PART)))) return rc;
        } else if (tp > 1) { // host-visible logits
            if (ctx->p2p) rc = launch_k(ctx, ps_k_tp_wait, dim3(1), dim3(32), 0, tp_in(ctx, PS_TP_SLOT_LOGITS));
            else rc = tp_all_gather(ctx, ctx->logits_part, ctx->logits, (size_t)ctx->vocab_l);
            if (rc) return rc;
        }
    }
    return 0;
}

// ---- fused decode step for the 32-element block formats (all-Q4_0 or all-Q8_0 models; ps_mv32.cuh): 7 launches per layer
// (q|k|v, rope + cache store, scores, soft-max / P.V, Wo, gate|up, down), PDL-chained and graph-replayed like the Q4_K step
template <int TYPE> int mv_repack(ps_cuda_ctx *ctx, uint8_t *dst, const uint8_t *src, int64_t n_rows, int64_t K, int64_t oct0) {
    ps_k_mv32_repack<TYPE><<<(unsigned)std::min<int64_t>((n_rows * (K / 32) + 255) / 256, 148 * 16), 256, 0, ctx->stream>>>(dst, src, n_rows, K / 32, oct0);
    PS_LAUNCH_CK();
    return 0;
}
int mv_repack_t(ps_cuda_ctx *ctx, uint8_t *dst, const uint8_t *src, int64_t n_rows, int64_t K, int64_t oct0) {
    return ctx->mv_type == 2 ? mv_repack<2>(ctx, dst, src, n_rows, K, oct0) : mv_repack<8>(ctx, dst, src, n_rows, K, oct0);
}
size_t mv_bytes(int type, int64_t rows, int64_t K) { return (size_t)((rows + 7) / 8) * (size_t)(K / 32) * (type == 2 ? PsMv32<2>::BLK : PsMv32<8>::BLK); }

int launch_mv(ps_cuda_ctx *ctx, PsMvArgs a) {
    const int blk = ctx->mv_type == 2 ? PsMv32<2>::BLK : PsMv32<8>::BLK;
    const int nb = a.K / 32;
    const int grid = std::max(1, std::min(ctx->n_sm, a.n_oct));
    const int per_cta = (a.n_oct + grid - 1) / grid, rounds = (per_cta + PS_MV_WARPS - 1) / PS_MV_WARPS;
    // stage = the largest divisor of the row's blocks that fits PS_MV_STAGE_CAP bytes; ring = up to PS_MV_MAX_NS stages per warp
    int sb = 1;
    for (int c = std::min(nb, PS_MV_STAGE_CAP / blk); c >= 1; c--)
        if (nb % c == 0) { sb = c; break; }
    a.sb = sb;
    a.ns = std::max(1, std::min(PS_MV_MAX_NS, rounds * (nb / sb)));
    a.inv_k = 1.0 / (double)a.K;
    if ((double)a.K * a.inv_k != 1.0) a.inv_k = 0.0; // only exact reciprocals replace the division (power-of-two K)
    const size_t act = ((size_t)a.K + (size_t)nb * 4 + 127) & ~(size_t)127;
    size_t smem = act + (size_t)PS_MV_WARPS * a.ns * ((size_t)sb * blk + 8);
    // block-parallel mode (ps_mv32.cuh): at most one octet per CTA and a long row - the CTA's eight warps share the row's blocks
    a.kpar = 0;
    if (ctx->opt_mv_kpar && per_cta == 1 && nb >= 16 && !a.part_val) {
        const size_t slice = ((size_t)((nb + PS_MV_WARPS - 1) / PS_MV_WARPS) * blk + 127) & ~(size_t)127;
        const size_t need = act + (size_t)PS_MV_WARPS * slice + (size_t)nb * 96 * 4 + PS_MV_WARPS * 8;
        if (need <= 200 * 1024) { a.kpar = (int)slice; smem = need; }
    }
    if (smem > 200 * 1024) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "32-block mat-vec: K = %d does not fit shared memory", a.K);
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_mv32<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_mv32<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr[ctx->device] = true;
    }
    if (ctx->mv_type == 2) return launch_k(ctx, ps_k_mv32<2>, dim3(grid), dim3(PS_MV_THREADS), smem, a);
    return launch_k(ctx, ps_k_mv32<8>, dim3(grid), dim3(PS_MV_THREADS), smem, a);
}

int decode_step_mv32(ps_cuda_ctx *ctx, bool lm_head, bool pick) {
    const ps_cuda_model_desc &d = ctx->d;
    const int dim = d.dim, hs = d.head_size, nh = d.n_heads, nkv = d.n_kv_heads, kvd = hs * nkv, qdim = nh * hs, ffn = d.ffn_dim;
    int rc;
    ctx->trace_launch = 0;
    ps_k_get_embedding<<<1, 256, 0, ctx->stream>>>(ctx->x, ctx->w_embd, ctx->t_embd, (int64_t)dim, ctx->tokens_dev);
    PS_LAUNCH_CK();
    for (int L = 0; L < d.n_layers; L++) {
        const LayerDev &ld = ctx->layers[L];
        {   // q | k | v = W . rmsnorm(x) (+ bias)
            PsMvArgs a{};
            a.w = ld.mv_qkv; a.n_oct = (qdim + 2 * kvd) / 8; a.K = dim; a.x = ctx->x; a.norm_w = ld.attn_norm; a.eps = d.norm_eps; a.pro = PS_MV_PRO_RMSNORM;
            a.seg[0] = {ctx->q, d.qkv_bias ? ld.q_bias : nullptr, 0, qdim};
            a.seg[1] = {ctx->k, d.qkv_bias ? ld.k_bias : nullptr, qdim, qdim + kvd};
            a.seg[2] = {ctx->v, d.qkv_bias ? ld.v_bias : nullptr, qdim + kvd, qdim + 2 * kvd};
            a.n_seg = 3;
            if ((rc = launch_mv(ctx, a))) return rc;
        }
        if ((rc = launch_k(ctx, ps_k_rope_kv, dim3((unsigned)(nh + 2 * nkv)), dim3(64), 0, ctx->qr, (const float *)ctx->q, (const float *)ctx->k, (const float *)ctx->v,
                           ctx->kc[L], ctx->vct[L], hs, nh, nkv, d.rope_n_dims, d.rope_type & 2, (const int32_t *)ctx->pos_dev, (const float *)ctx->rope_table,
                           (int64_t)d.n_ctx))) return rc;
        if ((rc = launch_attn_any(ctx, L, ctx->qr))) return rc;
        {   // x += Wo . att
            PsMvArgs a{};
            a.w = ld.mv_o; a.n_oct = dim / 8; a.K = qdim; a.x = ctx->att; a.pro = PS_MV_PRO_PLAIN;
            a.seg[0] = {ctx->x, nullptr, 0, dim}; a.n_seg = 1; a.residual = ctx->x;
            if ((rc = launch_mv(ctx, a))) return rc;
        }
        {   // g | u = Wgate | Wup . rmsnorm(x)
            PsMvArgs a{};
            a.w = ld.mv_gu; a.n_oct = 2 * ffn / 8; a.K = dim; a.x = ctx->x; a.norm_w = ld.ffn_norm; a.eps = d.norm_eps; a.pro = PS_MV_PRO_RMSNORM;
            a.seg[0] = {ctx->g, nullptr, 0, ffn};
            a.seg[1] = {ctx->u, nullptr, ffn, 2 * ffn};
            a.n_seg = 2;
            if ((rc = launch_mv(ctx, a))) return rc;
        }
        {   // x += Wdown . (silu(g) * u)
            PsMvArgs a{};
            a.w = ld.mv_down; a.n_oct = dim / 8; a.K = ffn; a.x = ctx->g; a.x2 = ctx->u; a.pro = PS_MV_PRO_SILU;
            a.seg[0] = {ctx->x, nullptr, 0, dim}; a.n_seg = 1; a.residual = ctx->x;
            if ((rc = launch_mv(ctx, a))) return rc;
        }
    }
    if (lm_head) {
        PsMvArgs a{};
        a.w = ctx->mv_out; a.n_oct = d.vocab_size / 8; a.K = dim; a.x = ctx->x; a.norm_w = ctx->w_out_norm; a.eps = d.norm_eps; a.pro = PS_MV_PRO_RMSNORM;
        a.seg[0] = {ctx->logits, nullptr, 0, d.vocab_size}; a.n_seg = 1;
        if (pick) { a.part_val = ctx->part_val; a.part_idx = ctx->part_idx; }
        if ((rc = launch_mv(ctx, a))) return rc;
        if (pick) {
            const int n_part = std::max(1, std::min(ctx->n_sm, a.n_oct));
            if ((rc = launch_k(ctx, ps_k_argmax_step, dim3(1), dim3(256), 0, (const float *)ctx->part_val, (const int *)ctx->part_idx, n_part, ctx->ids_dev, ctx->ctr_dev,
                               ctx->tokens_dev, ctx->pos_dev, (long long *)nullptr, (const PsTpIn *)nullptr))) return rc;
        }
    }
    return 0;
}
bool mv_usable(ps_cuda_ctx *ctx) { return ctx->mv_ok && ctx->opt_fused && ctx->tp == 1; }

// ---- graph-replayed decode step for ANY supported mix of weight types (Q4_0 / Q8_0 / Q6_K / Q4_K matrices, NEOX or NORM rope,
// q|k|v biases, 1..8 query heads per kv head): the operator-table kernels for the weight products - they hold the block
// arithmetic of every type - plus the type-agnostic decode attention kernels of the fused path (ps_k_attn1 / ps_k_attn2),
// which read the position from device memory.  Nothing in the step depends on a host value, so it is captured once and
// replayed per token like the Q4_K step; the pick runs on the device.  LlamaModel::forward / Qwen2Model::forward at bs = 1.
__global__ void __launch_bounds__(256) ps_k_argmax_parts(const float *__restrict__ logits, int n, float *__restrict__ part_val, int *__restrict__ part_idx) {
    __shared__ float sv[8];
    __shared__ int si[8];
    const int per = (n + gridDim.x - 1) / gridDim.x, lo = blockIdx.x * per, hi = min(n, lo + per);
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = lo + threadIdx.x; t < hi; t += blockDim.x) {
        const float v = logits[t];
        if (v > best) { best = v; bi = t; } // ascending t per thread: the first maximum wins
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(PS_FULL, best, o);
        const int oi = __shfl_xor_sync(PS_FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < 8; t++)
            if (sv[t] > best || (sv[t] == best && si[t] < bi)) { best = sv[t]; bi = si[t]; }
        part_val[blockIdx.x] = best;
        part_idx[blockIdx.x] = bi;
    }
}

int decode_step_ops(ps_cuda_ctx *ctx, bool lm_head, bool pick) {
    const ps_cuda_model_desc &d = ctx->d;
    const int64_t dim = d.dim, hs = d.head_size, nh = d.n_heads, nkv = d.n_kv_heads, kvd = hs * nkv, qdim = nh * hs, ffn = d.ffn_dim;
    int rc;
    ctx->trace_launch = 0;
    ps_k_get_embedding<<<1, 256, 0, ctx->stream>>>(ctx->x, ctx->w_embd, ctx->t_embd, dim, ctx->tokens_dev); // every embedding type, token from device memory
    PS_LAUNCH_CK();
    for (int L = 0; L < d.n_layers; L++) {
        const LayerDev &ld = ctx->layers[L];
        ps_k_rmsnorm<<<1, 256, 0, ctx->stream>>>(ctx->xn, ctx->x, ld.attn_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if ((rc = quantize_act(ctx, ld.tq, ctx->xn, dim, 1))) return rc;
        if ((rc = matmul_q(ctx, ctx->q, ld.wq, ld.tq, dim, qdim, 1, d.qkv_bias ? ld.q_bias : nullptr, nullptr))) return rc;
        if ((rc = matmul_q(ctx, ctx->k, ld.wk, ld.tk, dim, kvd, 1, d.qkv_bias ? ld.k_bias : nullptr, nullptr))) return rc;
        if ((rc = matmul_q(ctx, ctx->v, ld.wv, ld.tv, dim, kvd, 1, d.qkv_bias ? ld.v_bias : nullptr, nullptr))) return rc;
        ps_k_rope<<<dim3((unsigned)nh, 1), 64, 0, ctx->stream>>>(ctx->qr, ctx->q, (int)hs, d.rope_n_dims, d.rope_type & 2, ctx->pos_dev, ctx->rope_table);
        PS_LAUNCH_CK();
        ps_k_rope<<<dim3((unsigned)nkv, 1), 64, 0, ctx->stream>>>(ctx->kr, ctx->k, (int)hs, d.rope_n_dims, d.rope_type & 2, ctx->pos_dev, ctx->rope_table);
        PS_LAUNCH_CK();
        ps_k_kv_store<<<grid1d(kvd), 256, 0, ctx->stream>>>(ctx->kc[L], ctx->vct[L], ctx->kr, ctx->v, kvd, d.n_ctx, ctx->pos_dev, 1);
        PS_LAUNCH_CK();
        if ((rc = launch_attn_any(ctx, L, ctx->qr))) return rc;
        if ((rc = quantize_act(ctx, ld.to, ctx->att, qdim, 1))) return rc;
        if ((rc = matmul_q(ctx, ctx->x, ld.wo, ld.to, qdim, dim, 1, nullptr, ctx->x))) return rc;
        ps_k_rmsnorm<<<1, 256, 0, ctx->stream>>>(ctx->xn, ctx->x, ld.ffn_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if ((rc = quantize_act(ctx, ld.tgate, ctx->xn, dim, 1))) return rc;
        if ((rc = matmul_q(ctx, ctx->g, ld.wgate, ld.tgate, dim, ffn, 1, nullptr, nullptr))) return rc;
        if ((rc = matmul_q(ctx, ctx->u, ld.wup, ld.tup, dim, ffn, 1, nullptr, nullptr))) return rc;
        ps_k_silu_hadamard<<<grid1d(ffn), 256, 0, ctx->stream>>>(ctx->g, ctx->g, ctx->u, ffn);
        PS_LAUNCH_CK();
        if ((rc = quantize_act(ctx, ld.tdown, ctx->g, ffn, 1))) return rc;
        if ((rc = matmul_q(ctx, ctx->x, ld.wdown, ld.tdown, ffn, dim, 1, nullptr, ctx->x))) return rc;
    }
    if (lm_head) {
        ps_k_rmsnorm<<<1, 256, 0, ctx->stream>>>(ctx->xn, ctx->x, ctx->w_out_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if ((rc = quantize_act(ctx, ctx->t_out, ctx->xn, dim, 1))) return rc;
        if ((rc = matmul_q(ctx, ctx->logits, ctx->w_out, ctx->t_out, dim, d.vocab_size, 1, nullptr, nullptr))) return rc;
        if (pick) {
            const int n_part = std::min(ctx->n_sm, (d.vocab_size + 255) / 256);
            ps_k_argmax_parts<<<n_part, 256, 0, ctx->stream>>>(ctx->logits, d.vocab_size, ctx->part_val, ctx->part_idx);
            PS_LAUNCH_CK();
            if ((rc = launch_k(ctx, ps_k_argmax_step, dim3(1), dim3(256), 0, (const float *)ctx->part_val, (const int *)ctx->part_idx, n_part, ctx->ids_dev, ctx->ctr_dev,
                               ctx->tokens_dev, ctx->pos_dev, (long long *)nullptr, (const PsTpIn *)nullptr))) return rc;
        }
    }
    return 0;
}
bool ops_graph_usable(ps_cuda_ctx *ctx) { return ctx->ops_graph_ok && ctx->opt_ops_graph && ctx->tp == 1 && !(ctx->fused_ok && ctx->opt_fused) && !mv_usable(ctx); }

// ---- persistent per-step kernel (ps_step.cuh): the whole decode step in ONE cooperative launch
bool step_usable(ps_cuda_ctx *ctx) { return ctx->opt_persist && ctx->opt_fused && ctx->fused_ok && ctx->step_ok && (ctx->tp == 1 || ctx->p2p); }

template <int R2> int step_static_smem(ps_cuda_ctx *ctx, int *out) {
    cudaFuncAttributes fa;
    PS_CK(cudaFuncGetAttributes(&fa, ps_k_step<R2>));
    *out = (int)fa.sharedSizeBytes;
    return 0;
}
template <int R2> int launch_step_r(ps_cuda_ctx *ctx, const PsStArgs &a, int img_bytes, size_t smem) {
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_step<R2>, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin - ctx->st_static));
        attr[ctx->device] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)ctx->n_sm);
    cfg.blockDim = dim3(PS_ST_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative; // all CTAs resident together (the kernel's device-wide barriers depend on it)
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, ps_k_step<R2>, a, img_bytes);
    ctx->n_launch++;
    ctx->n_step++;
    if (e != cudaSuccess) return fail(ctx, PS_CUDA_ERR_CUDA, "step kernel launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// one decode step on the token in tokens_dev[0] at position pos_dev[0]; n_kv_max = the longest soft-max row this launch may see
int launch_step(ps_cuda_ctx *ctx, int mode, int n_kv_max) {
    const ps_cuda_model_desc &d = ctx->d;
    const int r2 = d.n_heads / d.n_kv_heads;
    PsStArgs a{};
    a.layers = ctx->st_layers; a.peers = ctx->st_peers;
    a.n_layers = d.n_layers; a.dim = d.dim; a.qdim = d.n_heads * d.head_size; a.ffn = d.ffn_dim; a.hs = d.head_size; a.n_ctx = d.n_ctx;
    a.qdim_l = ctx->nh_l * d.head_size; a.kvd_l = ctx->nkv_l * d.head_size; a.ffn_l = ctx->ffn_l; a.vocab_l = ctx->vocab_l; a.nkv_l = ctx->nkv_l;
    a.tp = ctx->tp; a.rank = ctx->rank;
    a.eps = d.norm_eps; a.kq_scale = 1.0f / sqrtf((float)d.head_size);
    a.w_embd = ctx->w_embd; a.w_out = ctx->rw_out; a.out_norm = ctx->w_out_norm; a.rope_table = ctx->rope_table;
    a.tokens_dev = ctx->tokens_dev; a.pos_dev = ctx->pos_dev; a.ids_dev = ctx->ids_dev; a.ctr_dev = ctx->ctr_dev;
    a.x_ll = ctx->st_ll[0]; a.x1_ll = ctx->st_ll[1]; a.att_ll = ctx->st_ll[2]; a.hq_ll = ctx->st_ll[3]; a.best_ll = ctx->st_ll[4];
    a.q = ctx->q; a.sc = ctx->kq; a.h = ctx->g_part; a.logits = ctx->logits_part;
    a.blk_cnt = ctx->blk_cnt; a.part_val = ctx->part_val; a.part_idx = ctx->part_idx;
    a.bar_ctr = ctx->st_sync; a.done_ctr = ctx->st_sync + 1; a.serial = ctx->st_sync + 2;
    a.err = ctx->err_dev + 2;
    a.tpo_logits = (ctx->tp > 1 && (mode & PS_ST_MODE_LMHEAD) && !(mode & PS_ST_MODE_PICK)) ? ctx->tpo_dev + PS_TP_SLOT_LOGITS : nullptr;
    auto kb_of = [](int nb, int cap) { int kb = cap; while (nb % kb) kb >>= 1; return kb; };
    a.kb_dim = kb_of(a.dim / 256, 4); a.kb_gu = kb_of(a.dim / 256, 2); a.kb_q = kb_of(a.qdim / 256, 4); a.kb_ffn = kb_of(a.ffn / 256, 4);
    a.mode = mode;
    a.l2_ahead = ctx->opt_l2_ahead;
    a.timeout_ns = ctx->tp > 1 ? 30000000000LL : 2000000000LL;
    a.tl = ctx->trace_dev;
    // P.V work split: halve the dims per CTA while fewer than half of the SMs would have an item
    a.dpc = 8;
    while (a.dpc > 1 && a.nkv_l * (a.hs / a.dpc) * 2 <= ctx->n_sm) a.dpc >>= 1;
    // shared memory: [activation image | soft-max rows][two rings]
    if (ctx->st_static < 0) { // the kernel's static shared memory comes out of the same opt-in limit
        int rc0, v = 0;
        switch (r2) {
        case 1: rc0 = step_static_smem<1>(ctx, &v); break;
        case 2: rc0 = step_static_smem<2>(ctx, &v); break;
        case 4: rc0 = step_static_smem<4>(ctx, &v); break;
        default: rc0 = step_static_smem<8>(ctx, &v); break;
        }
        if (rc0) return rc0;
        ctx->st_static = (v + 127) & ~127;
    }
    const int budget = ctx->smem_optin - ctx->st_static;
    const int img_need = std::max(std::max(a.dim + a.dim / 8, a.qdim + a.qdim / 8), a.ffn + a.ffn / 8);
    int ch = std::min((n_kv_max + 255) & ~255, std::max(256, (64 * 1024 / (4 * r2)) & ~255));
    if (ctx->opt_attn_chunk > 0) ch = std::min(ch, std::max(256, ctx->opt_attn_chunk & ~255));
    int img_bytes = 0, nsp = 0;
    for (;;) {
        img_bytes = (std::max(img_need, r2 * ch * 4) + 127) & ~127;
        nsp = (budget - img_bytes) / (PS_ST_PROD * (PS_ST_SLOT + 16));
        if (nsp >= 14 || ch <= 256) break;
        ch = std::max(256, (ch / 2 + 255) & ~255); // long contexts: keep the weight rings deep, rebuild the soft-max rows in chunks
    }
    if (nsp < 4) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "step kernel: no room for the weight rings (image %d bytes)", img_bytes);
    nsp = std::min(nsp, 64);
    a.nsp = nsp; a.attn_chunk = ch;
    const size_t smem = (size_t)img_bytes + (size_t)PS_ST_PROD * nsp * (PS_ST_SLOT + 16);
    int rc;
    switch (r2) {
    case 1: rc = launch_step_r<1>(ctx, a, img_bytes, smem); break;
    case 2: rc = launch_step_r<2>(ctx, a, img_bytes, smem); break;
    case 4: rc = launch_step_r<4>(ctx, a, img_bytes, smem); break;
    default: rc = launch_step_r<8>(ctx, a, img_bytes, smem); break;
    }
    if (rc) return rc;
    if (a.tpo_logits) return launch_k(ctx, ps_k_tp_wait, dim3(1), dim3(32), 0, tp_in(ctx, PS_TP_SLOT_LOGITS)); // every rank's logits rows have landed
    return 0;
}

// capture one step into a graph (lazily), then replay it
int run_step(ps_cuda_ctx *ctx, bool pick, int n_kv_max) {
    if (step_usable(ctx)) {
        const int mode = PS_ST_MODE_LMHEAD | (pick ? PS_ST_MODE_PICK : 0);
        if (!ctx->opt_ktime) return launch_step(ctx, mode, n_kv_max);
        // kernel timing pass: CUDA events on the launching stream around every step launch
        while (ctx->kt_events.size() < ctx->kt_used + 2) {
            cudaEvent_t e;
            PS_CK(cudaEventCreate(&e));
            ctx->kt_events.push_back(e);
        }
        PS_CK(cudaEventRecord(ctx->kt_events[ctx->kt_used], ctx->stream));
        const int rc = launch_step(ctx, mode, n_kv_max);
        PS_CK(cudaEventRecord(ctx->kt_events[ctx->kt_used + 1], ctx->stream));
        ctx->kt_used += 2;
        return rc;
    }
    const bool ops = ops_graph_usable(ctx);
    const bool mv = mv_usable(ctx);
    auto build = [&]() { return mv ? decode_step_mv32(ctx, true, pick) : ops ? decode_step_ops(ctx, true, pick) : decode_step_fused(ctx, true, pick); };
    if (!ctx->opt_graph || ctx->opt_ktime) return build();
    cudaGraphExec_t &ge = pick ? ctx->g_step : ctx->g_fwd;
    if (!ge) {
        cudaGraph_t graph = nullptr;
        const int64_t n0 = ctx->n_launch;
        PS_CK(cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal));
        int rc = build();
        cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
        (pick ? ctx->g_step_kernels : ctx->g_fwd_kernels) = ctx->n_launch - n0;
        ctx->n_launch = n0; // capture enqueued nothing
        if (rc) { if (graph) cudaGraphDestroy(graph); return rc; }
        if (e != cudaSuccess) return fail(ctx, PS_CUDA_ERR_CUDA, "graph capture failed: %s", cudaGetErrorString(e));
        e = cudaGraphInstantiate(&ge, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) { ge = nullptr; return fail(ctx, PS_CUDA_ERR_CUDA, "graph instantiate failed: %s", cudaGetErrorString(e)); }
    }
    PS_CK(cudaGraphLaunch(ge, ctx->stream));
    ctx->n_graph++;
    ctx->n_launch += pick ? ctx->g_step_kernels : ctx->g_fwd_kernels; // OUR kernels inside the replayed graph (NCCL's are not counted)
    return 0;
}

} // namespace


// ====================================================================================================================
extern "C" {

int ps_cuda_abi_version(void) { return PS_CUDA_ABI_VERSION; }

int ps_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

const char *ps_cuda_last_error(const ps_cuda_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int ps_cuda_create(ps_cuda_ctx **out, int device, const ps_cuda_model_desc *desc) {
    ps_cuda_ctx *ctx = nullptr;
    if (!out || !desc) return fail(nullptr, PS_CUDA_ERR_INVALID, "null argument");
    *out = nullptr;
    const int ndev = ps_cuda_device_count();
    if (ndev == 0) return fail(nullptr, PS_CUDA_ERR_NO_DEVICE, "no CUDA device: the PowerServe CUDA backend has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, PS_CUDA_ERR_INVALID, "device %d out of range [0,%d)", device, ndev);
    const ps_cuda_model_desc &d = *desc;
    if (d.dim <= 0 || d.n_layers <= 0 || d.n_heads <= 0 || d.n_kv_heads <= 0 || d.head_size <= 0 || d.n_ctx <= 0 || d.vocab_size <= 0 ||
        d.ffn_dim <= 0 || d.max_batch <= 0)
        return fail(nullptr, PS_CUDA_ERR_INVALID, "model descriptor has non-positive fields");
    if (d.n_heads % d.n_kv_heads || d.head_size % 32 || d.head_size > 256 || d.rope_n_dims != d.head_size)
        return fail(nullptr, PS_CUDA_ERR_UNSUPPORTED, "unsupported head geometry (heads %d/%d, head_size %d, rope dims %d)", d.n_heads,
                    d.n_kv_heads, d.head_size, d.rope_n_dims);
    const int tp = d.tp_size > 0 ? d.tp_size : 1;
    if (tp > 1 && (d.tp_rank < 0 || d.tp_rank >= tp || d.n_heads % tp || d.n_kv_heads % tp || d.ffn_dim % (8 * tp) || d.vocab_size % (8 * tp) ||
                   d.dim % (8 * tp)))
        return fail(nullptr, PS_CUDA_ERR_UNSUPPORTED, "tensor parallel size %d does not divide heads %d/%d, ffn %d, vocab %d or dim %d into octets", tp,
                    d.n_heads, d.n_kv_heads, d.ffn_dim, d.vocab_size, d.dim);
    ctx = new ps_cuda_ctx();
    ctx->device = device;
    ctx->d = d;
    ctx->tp = tp;
    ctx->rank = tp > 1 ? d.tp_rank : 0;
    ctx->nh_l = d.n_heads / tp; ctx->nkv_l = d.n_kv_heads / tp; ctx->ffn_l = d.ffn_dim / tp; ctx->vocab_l = d.vocab_size / tp; ctx->dim_l = d.dim / tp;
    auto bail = [&](int rc) {
        g_create_error = ctx->err;
        ps_cuda_destroy(ctx);
        return rc;
    };
#define PS_CKC(call)                                                                      \
    do {                                                                                  \
        cudaError_t e_ = (call);                                                          \
        if (e_ != cudaSuccess) {                                                          \
            fail(ctx, PS_CUDA_ERR_CUDA, "%s failed: %s", #call, cudaGetErrorString(e_)); \
            return bail(PS_CUDA_ERR_CUDA);                                                \
        }                                                                                 \
    } while (0)
#define PS_AL(ptr, bytes)                                         \
    do {                                                          \
        int rc_ = dev_alloc(ctx, (void **)&(ptr), (size_t)(bytes)); \
        if (rc_) return bail(rc_);                                \
    } while (0)
    PS_CKC(cudaSetDevice(device));
    PS_CKC(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    const int64_t B = d.max_batch, dim = d.dim, qdim = (int64_t)d.n_heads * d.head_size, kvd = (int64_t)d.n_kv_heads * d.head_size;
    ctx->maxK = std::max<int64_t>(std::max<int64_t>(dim, qdim), d.ffn_dim);
    PS_AL(ctx->x, 4 * dim * B);
    PS_AL(ctx->xn, 4 * dim * B);
    PS_AL(ctx->q, 4 * qdim * B);
    PS_AL(ctx->k, 4 * kvd * B);
    PS_AL(ctx->v, 4 * kvd * B);
    PS_AL(ctx->qr, 4 * qdim * B);
    PS_AL(ctx->kr, 4 * kvd * B);
    PS_AL(ctx->att, 4 * qdim * B);
    PS_AL(ctx->g, 4 * (int64_t)d.ffn_dim * B);
    PS_AL(ctx->u, 4 * (int64_t)d.ffn_dim * B);
    PS_AL(ctx->kq, 4 * (int64_t)d.n_heads * B * d.n_ctx);
    PS_AL(ctx->logits, 4 * (int64_t)d.vocab_size * B);
    PS_AL(ctx->aqs, ctx->maxK * B);
    PS_AL(ctx->ad, 4 * (ctx->maxK / 32) * B);
    PS_AL(ctx->absp, 4 * (ctx->maxK / 256 + 1) * 4 * B);
    PS_AL(ctx->tokens_dev, 4 * B);
    PS_AL(ctx->pos_dev, 4 * B);
    PS_AL(ctx->ids_dev, 4 * 4096);
    PS_AL(ctx->ctr_dev, 16);
    PS_AL(ctx->err_dev, 128); // [0..2] flags, [3..] debug record of the step kernel (PS_ST_DEBUG builds)
    PS_CKC(cudaMemsetAsync(ctx->err_dev, 0, 128, ctx->stream));
    ctx->tc_err_dev = ctx->err_dev;
    ctx->tp_err_dev = ctx->err_dev + 1;
    PS_CKC(cudaMallocHost(&ctx->h_err, 16));
    if (tp > 1) {
        auto up = [](size_t v) { return (v + 255) & ~(size_t)255; };
        ctx->off_att = 0;
        ctx->off_x = up(ctx->off_att + 4 * (size_t)qdim);
        ctx->off_h = up(ctx->off_x + 4 * (size_t)dim);
        ctx->off_val = up(ctx->off_h + 4 * (size_t)d.ffn_dim);
        ctx->off_idx = up(ctx->off_val + 4 * 1024 * (size_t)tp);
        ctx->off_logits = up(ctx->off_idx + 4 * 1024 * (size_t)tp);
        ctx->off_flags = up(ctx->off_logits + 4 * (size_t)d.vocab_size);
        ctx->off_ll[PS_TP_SLOT_ATT] = up(ctx->off_flags + 4 * PS_TP_SLOTS * PS_TP_MAX);
        ctx->off_ll[PS_TP_SLOT_X1] = up(ctx->off_ll[PS_TP_SLOT_ATT] + 8 * (size_t)qdim);
        ctx->off_ll[PS_TP_SLOT_H] = up(ctx->off_ll[PS_TP_SLOT_X1] + 8 * (size_t)dim);
        ctx->off_ll[PS_TP_SLOT_X2] = up(ctx->off_ll[PS_TP_SLOT_H] + 8 * (size_t)d.ffn_dim);
        // (value, epoch) vectors of the persistent step kernel (ps_step.cuh): x, x1, att, Q8_K image of the FFN hidden vector, best
        const size_t st_words[5] = {(size_t)dim, (size_t)dim, (size_t)qdim, (size_t)(d.ffn_dim / 256 + 1) * 72, 2 * PS_TP_MAX};
        ctx->st_off[0] = up(ctx->off_ll[PS_TP_SLOT_X2] + 8 * (size_t)dim);
        for (int k = 1; k < 5; k++) ctx->st_off[k] = up(ctx->st_off[k - 1] + 8 * st_words[k - 1]);
        ctx->heap_bytes = up(ctx->st_off[4] + 8 * st_words[4]);
        PS_AL(ctx->heap, ctx->heap_bytes);
        for (int k = 0; k < 5; k++) ctx->st_ll[k] = reinterpret_cast<unsigned long long *>(ctx->heap + ctx->st_off[k]);
        PS_CKC(cudaMemsetAsync(ctx->heap, 0, ctx->heap_bytes, ctx->stream));
        PS_AL(ctx->epoch_dev, 4 * PS_TP_SLOTS);
        PS_AL(ctx->done_dev, 4 * PS_TP_SLOTS);
        PS_CKC(cudaMemsetAsync(ctx->epoch_dev, 0, 4 * PS_TP_SLOTS, ctx->stream));
        PS_CKC(cudaMemsetAsync(ctx->done_dev, 0, 4 * PS_TP_SLOTS, ctx->stream));
        // the gathered vectors live in the heap (x replaces the workspace x allocated above for the decode path)
        ctx->att_full = reinterpret_cast<float *>(ctx->heap + ctx->off_att);
        ctx->x = reinterpret_cast<float *>(ctx->heap + ctx->off_x);
        ctx->h_full = reinterpret_cast<float *>(ctx->heap + ctx->off_h);
        ctx->all_val = reinterpret_cast<float *>(ctx->heap + ctx->off_val);
        ctx->all_idx = reinterpret_cast<int *>(ctx->heap + ctx->off_idx);
        ctx->logits = reinterpret_cast<float *>(ctx->heap + ctx->off_logits); // single-token logits (batches are staged through tp_rows)
        PS_AL(ctx->x_part, 4 * (int64_t)ctx->dim_l);
        PS_AL(ctx->g_part, 4 * (int64_t)ctx->ffn_l);
        PS_AL(ctx->logits_part, 4 * (int64_t)ctx->vocab_l);
    } else {
        ctx->att_full = ctx->att; ctx->h_full = ctx->g; ctx->x_part = ctx->x; ctx->g_part = ctx->g; ctx->logits_part = ctx->logits;
        const size_t st_words[5] = {(size_t)dim, (size_t)dim, (size_t)qdim, (size_t)(d.ffn_dim / 256 + 1) * 72, 2 * PS_TP_MAX};
        for (int k = 0; k < 5; k++) {
            PS_AL(ctx->st_ll[k], 8 * st_words[k]);
            PS_CKC(cudaMemsetAsync(ctx->st_ll[k], 0, 8 * st_words[k], ctx->stream));
        }
    }
    PS_AL(ctx->st_sync, 16);
    PS_CKC(cudaMemsetAsync(ctx->st_sync, 0, 16, ctx->stream));
    PS_AL(ctx->st_peers, sizeof(PsStPeers));
    {   // until ps_cuda_tp_import: only this rank's own copies (a single-GPU context never needs more)
        PsStPeers pe;
        memset(&pe, 0, sizeof pe);
        pe.x[ctx->rank] = ctx->st_ll[0]; pe.x1[ctx->rank] = ctx->st_ll[1]; pe.att[ctx->rank] = ctx->st_ll[2]; pe.hq[ctx->rank] = ctx->st_ll[3]; pe.best[ctx->rank] = ctx->st_ll[4];
        PS_CKC(cudaMemcpy(ctx->st_peers, &pe, sizeof pe, cudaMemcpyHostToDevice));
    }
    { int v = 0; if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) == cudaSuccess) ctx->smem_optin = v; }
    PS_AL(ctx->ximg, (size_t)B * ((size_t)ctx->maxK + (size_t)(ctx->maxK / 256 + 1) * 32));
    PS_AL(ctx->hq, (size_t)d.ffn_dim + (size_t)(d.ffn_dim / 256 + 1) * 32);
    PS_AL(ctx->blk_cnt, 4 * (size_t)(d.ffn_dim / 256 + 1));
    PS_CKC(cudaMemsetAsync(ctx->blk_cnt, 0, 4 * (size_t)(d.ffn_dim / 256 + 1), ctx->stream));
    PS_AL(ctx->part_val, 4 * 1024);
    PS_AL(ctx->part_idx, 4 * 1024);
    { int v = 148; if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) == cudaSuccess) ctx->n_sm = v; }
    if (const char *e = getenv("PS_CUDA_TP_LL")) ctx->opt_ll = atoi(e) != 0; // A-B runs of the two peer-store protocols
    ctx->kc.resize(d.n_layers);
    ctx->vct.resize(d.n_layers);
    for (int L = 0; L < d.n_layers; L++) {
        const int64_t kvd_l = kvd / tp; // a tensor-parallel rank caches only its own kv heads
        PS_AL(ctx->kc[L], 4 * kvd_l * d.n_ctx);
        PS_AL(ctx->vct[L], 4 * kvd_l * d.n_ctx);
        PS_CKC(cudaMemsetAsync(ctx->kc[L], 0, 4 * kvd_l * d.n_ctx, ctx->stream));
        PS_CKC(cudaMemsetAsync(ctx->vct[L], 0, 4 * kvd_l * d.n_ctx, ctx->stream));
    }
    if (tp == 1 && d.head_size % 8 == 0) { // group-synchronised decode attention: arrival counters and the chunk maxima / sums of every kv-head group
        const size_t nc = (size_t)d.n_layers * d.n_kv_heads, np = (size_t)d.n_kv_heads * (d.head_size / 8) * 8;
        PS_AL(ctx->ag_ctr, nc * 8);
        PS_AL(ctx->ag_max, np * 4);
        PS_AL(ctx->ag_sum, np * 8);
        PS_CKC(cudaMemsetAsync(ctx->ag_ctr, 0, nc * 8, ctx->stream));
    }
    {   // speculative decode: slot mask, staging rows of the last batch, cache pointer table, tree bias
        const int64_t kvd_l = kvd / tp, tb = std::min<int64_t>(B, 32);
        ctx->slot_mask.assign((size_t)d.n_ctx, 1); // nothing committed yet: every slot is masked (rollback state)
        PS_AL(ctx->slot_mask_dev, (size_t)d.n_ctx);
        PS_AL(ctx->tree_dev, (size_t)(tb * tb));
        PS_AL(ctx->tree_bias, 4 * (size_t)tb * (size_t)d.n_ctx);
        PS_AL(ctx->k_stage, 4 * (size_t)d.n_layers * (size_t)tb * (size_t)kvd_l);
        PS_AL(ctx->v_stage, 4 * (size_t)d.n_layers * (size_t)tb * (size_t)kvd_l);
        PS_AL(ctx->kv_ptrs_dev, sizeof(float *) * 2 * (size_t)d.n_layers);
        std::vector<float *> ptrs(2 * (size_t)d.n_layers);
        for (int L = 0; L < d.n_layers; L++) { ptrs[L] = ctx->kc[L]; ptrs[d.n_layers + L] = ctx->vct[L]; }
        PS_CKC(cudaMemcpy(ctx->kv_ptrs_dev, ptrs.data(), sizeof(float *) * ptrs.size(), cudaMemcpyHostToDevice));
        PS_CKC(cudaMallocHost(&ctx->h_tree, (size_t)(tb * tb)));
    }
    {
        std::vector<float> t;
        build_rope_table(d, t);
        PS_AL(ctx->rope_table, t.size() * 4);
        PS_CKC(cudaMemcpy(ctx->rope_table, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    }
    PS_CKC(cudaMallocHost(&ctx->h_tokens, 4 * B));
    PS_CKC(cudaMallocHost(&ctx->h_pos, 4 * B));
    PS_CKC(cudaMallocHost(&ctx->h_ids, 4 * 8192));
    PS_CKC(cudaEventCreate(&ctx->ev0));
    PS_CKC(cudaEventCreate(&ctx->ev1));
    PS_CKC(cudaStreamSynchronize(ctx->stream));
#undef PS_CKC
#undef PS_AL
    *out = ctx;
    return 0;
}

void ps_cuda_destroy(ps_cuda_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->pooled) cudaFreeAsync(p, ctx->stream);
    cudaStreamSynchronize(ctx->stream);
    for (void *p : ctx->owned) cudaFree(p);
    if (ctx->h_tokens) cudaFreeHost(ctx->h_tokens);
    if (ctx->h_pos) cudaFreeHost(ctx->h_pos);
    if (ctx->h_ids) cudaFreeHost(ctx->h_ids);
    if (ctx->h_err) cudaFreeHost(ctx->h_err);
    if (ctx->h_tree) cudaFreeHost(ctx->h_tree);
    if (ctx->h_logits) cudaFreeHost(ctx->h_logits);
    if (ctx->g_step) cudaGraphExecDestroy(ctx->g_step);
    if (ctx->g_fwd) cudaGraphExecDestroy(ctx->g_fwd);
    for (int p = 0; p < PS_TP_MAX; p++)
        if (ctx->peer_heap[p] && p != ctx->rank) cudaIpcCloseMemHandle(ctx->peer_heap[p]);
    if (ctx->nccl_comm && g_nccl.lib) g_nccl.CommDestroy(ctx->nccl_comm);
    for (cudaEvent_t e : ctx->kt_events) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

int ps_cuda_sync(ps_cuda_ctx *ctx) {
    PS_CK(cudaStreamSynchronize(ctx->stream));
    return 0;
}

void *ps_cuda_stream(ps_cuda_ctx *ctx) { return (void *)ctx->stream; }

// ---------------------------------------------------------------------------------------------- memory
// CUDABuffer backing for graph intermediates: the executor allocates and frees ~30 tensors per layer on every forward pass
// (executor.cpp:23-45, cpu_buffer.hpp:41-51), so these come from the device's stream-ordered memory pool - freed blocks are
// cached by the pool (release threshold: never) and handed out again in stream order, without a device synchronisation.
int ps_cuda_malloc(ps_cuda_ctx *ctx, size_t bytes, void **dev) {
    PS_CK(cudaSetDevice(ctx->device));
    static bool pool_set[64] = {};
    if (!pool_set[ctx->device]) {
        cudaMemPool_t pool;
        PS_CK(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
        unsigned long long keep = ~0ull;
        PS_CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
        pool_set[ctx->device] = true;
    }
    *dev = nullptr;
    cudaError_t e = cudaMallocAsync(dev, std::max<size_t>(bytes, 256), ctx->stream);
    if (e != cudaSuccess) return fail(ctx, PS_CUDA_ERR_OOM, "ps_cuda_malloc(%zu): %s", bytes, cudaGetErrorString(e));
    ctx->pooled.insert(*dev);
    return 0;
}

int ps_cuda_free(ps_cuda_ctx *ctx, void *dev) {
    if (ctx->pooled.erase(dev)) {
        PS_CK(cudaFreeAsync(dev, ctx->stream));
        return 0;
    }
    for (size_t i = 0; i < ctx->owned.size(); i++)
        if (ctx->owned[i] == dev) {
            PS_CK(cudaStreamSynchronize(ctx->stream));
            PS_CK(cudaFree(dev));
            ctx->owned.erase(ctx->owned.begin() + i);
            return 0;
        }
    return fail(ctx, PS_CUDA_ERR_INVALID, "ps_cuda_free: pointer not owned by this context");
}

int ps_cuda_memcpy_h2d(ps_cuda_ctx *ctx, void *dev, const void *host, size_t bytes) {
    PS_CK(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PS_CK(cudaStreamSynchronize(ctx->stream)); // pageable source: make the call safe to return from
    ctx->h2d += (int64_t)bytes;
    return 0;
}

int ps_cuda_memcpy_d2h(ps_cuda_ctx *ctx, void *host, const void *dev, size_t bytes) {
    PS_CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    ctx->d2h += (int64_t)bytes;
    return 0;
}

int ps_cuda_register_weight(ps_cuda_ctx *ctx, const void *host, int type, int64_t ne0, int64_t ne1, void **dev) {
    if (!host || !type_ok(type) || ne0 <= 0 || ne1 <= 0 || ne0 % blk_elems(type))
        return fail(ctx, PS_CUDA_ERR_INVALID, "register_weight: bad tensor (type %d, %lld x %lld)", type, (long long)ne0, (long long)ne1);
    auto it = ctx->weights.find(host);
    if (it != ctx->weights.end()) {
        // a row prefix of a registered tensor aliases it (tied lm_head shard of tensor-parallel rank 0 == first rows of token_embd)
        if (it->second.type != type || it->second.ne0 != ne0 || it->second.ne1 < ne1)
            return fail(ctx, PS_CUDA_ERR_INVALID, "register_weight: host pointer already registered with another shape");
        if (dev) *dev = it->second.dev;
        return 0;
    }
    PS_CK(cudaSetDevice(ctx->device));
    const size_t bytes = (size_t)ps_row_bytes(type, ne0) * (size_t)ne1;
    DevWeight w;
    w.type = type; w.ne0 = ne0; w.ne1 = ne1;
    int rc = dev_alloc(ctx, &w.dev, bytes + 16);
    if (rc) return rc;
    PS_CK(cudaMemcpy(w.dev, host, bytes, cudaMemcpyHostToDevice));
    ctx->h2d += (int64_t)bytes;
    ctx->weights[host] = w;
    if (dev) *dev = w.dev;
    return 0;
}

void *ps_cuda_lookup_weight(ps_cuda_ctx *ctx, const void *host) {
    auto it = ctx->weights.find(host);
    return it == ctx->weights.end() ? nullptr : it->second.dev;
}

int ps_cuda_unregister_weight(ps_cuda_ctx *ctx, const void *host) {
    auto it = ctx->weights.find(host);
    if (it == ctx->weights.end()) return fail(ctx, PS_CUDA_ERR_INVALID, "unregister_weight: unknown host pointer");
    void *dev = it->second.dev;
    ctx->weights.erase(it);
    return ps_cuda_free(ctx, dev);
}

// ---------------------------------------------------------------------------------------------- operator table
static int stage_ints(ps_cuda_ctx *ctx, int32_t *dev, int32_t *pinned, const int32_t *host, int64_t n) {
    if (n > ctx->d.max_batch) return fail(ctx, PS_CUDA_ERR_INVALID, "batch %lld exceeds max_batch %d", (long long)n, ctx->d.max_batch);
    PS_CK(cudaStreamSynchronize(ctx->stream)); // the pinned staging buffer may still be in flight
    memcpy(pinned, host, (size_t)n * 4);
    PS_CK(cudaMemcpyAsync(dev, pinned, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += n * 4;
    return 0;
}

int ps_cuda_get_embedding(ps_cuda_ctx *ctx, float *dst, const void *w, int wtype, int64_t dim, const int32_t *tokens, int64_t bs) {
    if (!type_ok(wtype) || dim % blk_elems(wtype)) return fail(ctx, PS_CUDA_ERR_INVALID, "get_embedding: bad type/dim");
    int rc = stage_ints(ctx, ctx->tokens_dev, ctx->h_tokens, tokens, bs);
    if (rc) return rc;
    ps_k_get_embedding<<<(unsigned)bs, 256, 0, ctx->stream>>>(dst, (const uint8_t *)w, wtype, dim, ctx->tokens_dev);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_rmsnorm(ps_cuda_ctx *ctx, float *dst, const float *x, const float *w, int64_t dim, int64_t bs, float eps) {
    if (!(eps > 0.0f)) return fail(ctx, PS_CUDA_ERR_INVALID, "rmsnorm: eps must be > 0 (GGML_ASSERT, ggml.c:12684)");
    ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(dst, x, w, dim, eps);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_matmul(ps_cuda_ctx *ctx, float *dst, const void *w, int wtype, int64_t K, int64_t N, const float *x, int64_t bs) {
    if (bs <= 0 || N <= 0) return fail(ctx, PS_CUDA_ERR_INVALID, "matmul: empty shape");
    int rc = quantize_act(ctx, wtype, x, K, bs);
    if (rc) return rc;
    return matmul_q(ctx, dst, (const uint8_t *)w, wtype, K, N, bs, nullptr, nullptr);
}

int ps_cuda_rope(ps_cuda_ctx *ctx, float *dst, const float *src, int64_t head_size, int64_t n_heads, int64_t bs, const int32_t *pos) {
    if (head_size != ctx->d.head_size) return fail(ctx, PS_CUDA_ERR_INVALID, "rope: head_size %lld != model head_size %d", (long long)head_size, ctx->d.head_size);
    for (int64_t i = 0; i < bs; i++)
        if (pos[i] < 0 || pos[i] >= ctx->d.n_ctx) return fail(ctx, PS_CUDA_ERR_INVALID, "rope: position %d outside [0,%d)", pos[i], ctx->d.n_ctx);
    int rc = stage_ints(ctx, ctx->pos_dev, ctx->h_pos, pos, bs);
    if (rc) return rc;
    ps_k_rope<<<dim3((unsigned)n_heads, (unsigned)bs), 64, 0, ctx->stream>>>(dst, src, (int)head_size, ctx->d.rope_n_dims, ctx->d.rope_type & 2,
                                                                              ctx->pos_dev, ctx->rope_table);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_set_rope_freq_factors(ps_cuda_ctx *ctx, const float *factors, int n) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (factors && n != ctx->d.rope_n_dims / 2) return fail(ctx, PS_CUDA_ERR_INVALID, "rope_freq_factors: %d entries, expected rope_n_dims / 2 = %d", n, ctx->d.rope_n_dims / 2);
    for (int i = 0; factors && i < n; i++)
        if (!(factors[i] > 0.f)) return fail(ctx, PS_CUDA_ERR_INVALID, "rope_freq_factors: entry %d is not positive", i);
    PS_CK(cudaSetDevice(ctx->device));
    std::vector<float> t;
    build_rope_table(ctx->d, t, factors);
    PS_CK(cudaStreamSynchronize(ctx->stream));
    PS_CK(cudaMemcpy(ctx->rope_table, t.data(), t.size() * 4, cudaMemcpyHostToDevice));
    return 0;
}

int ps_cuda_add(ps_cuda_ctx *ctx, float *dst, const float *a, const float *b, int64_t n, int64_t nb) {
    if (nb <= 0 || n % nb) return fail(ctx, PS_CUDA_ERR_INVALID, "add: broadcast size %lld does not divide %lld", (long long)nb, (long long)n);
    ps_k_add<<<grid1d(n), 256, 0, ctx->stream>>>(dst, a, b, n, nb);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_silu_hadamard(ps_cuda_ctx *ctx, float *dst, const float *gate, const float *up, int64_t n) {
    ps_k_silu_hadamard<<<grid1d(n), 256, 0, ctx->stream>>>(dst, gate, up, n);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_get_mask(ps_cuda_ctx *ctx, float *mask, int64_t n_kv, int64_t bs, const int32_t *pos) {
    int rc = stage_ints(ctx, ctx->pos_dev, ctx->h_pos, pos, bs);
    if (rc) return rc;
    ps_k_get_mask<<<dim3((unsigned)std::min<int64_t>((n_kv + 255) / 256, 64), (unsigned)bs), 256, 0, ctx->stream>>>(mask, n_kv, ctx->pos_dev);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_softmax_ext(ps_cuda_ctx *ctx, float *dst, const float *x, const float *mask, int64_t ne0, int64_t ne1, int64_t ne2, float scale) {
    if (!mask) return fail(ctx, PS_CUDA_ERR_INVALID, "softmax_ext: mask is required (use get_mask)");
    if (ne0 * 4 > 160 * 1024) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "softmax_ext: row of %lld exceeds shared memory", (long long)ne0);
    static bool attr[64] = {};
    if (!attr[ctx->device]) { PS_CK(cudaFuncSetAttribute(ps_k_softmax_ext, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr[ctx->device] = true; }
    ps_k_softmax_ext<<<(unsigned)(ne1 * ne2), 256, (size_t)ne0 * 4, ctx->stream>>>(dst, x, mask, nullptr, ne0, ne1, scale);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_attn_scores(ps_cuda_ctx *ctx, float *kq, const float *k_cache, const float *q, int64_t hs, int64_t n_heads, int64_t n_kv_heads,
                        int64_t n_kv, int64_t bs) {
    if (hs % 32 || hs > 256 || n_heads % n_kv_heads) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "attn_scores: head geometry");
    ps_k_attn_scores<<<dim3((unsigned)((n_kv + 3) / 4), (unsigned)n_kv_heads), 128, 0, ctx->stream>>>(kq, k_cache, q, (int)hs, (int)n_heads,
                                                                                                     (int)n_kv_heads, n_kv, (int)bs);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_attn_pv(ps_cuda_ctx *ctx, float *out, const float *v_cache_t, const float *p, int64_t hs, int64_t n_heads, int64_t n_kv_heads,
                    int64_t n_kv, int64_t n_ctx, int64_t bs) {
    if (hs % 32 || n_heads % n_kv_heads) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "attn_pv: head geometry");
    ps_k_attn_pv<<<dim3((unsigned)((hs + 3) / 4), (unsigned)n_kv_heads), 128, 0, ctx->stream>>>(out, v_cache_t, p, (int)hs, (int)n_heads,
                                                                                               (int)n_kv_heads, n_kv, n_ctx, (int)bs);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_copy_2d(ps_cuda_ctx *ctx, void *dst, int64_t ds0, int64_t ds1, const void *src, int64_t ss0, int64_t ss1, int64_t ne0, int64_t ne1) {
    ps_k_copy_2d<<<grid1d(ne0 * ne1), 256, 0, ctx->stream>>>((uint8_t *)dst, ds0, ds1, (const uint8_t *)src, ss0, ss1, ne0, ne1);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_copy_4d(ps_cuda_ctx *ctx, void *dst, const int64_t dst_ne[4], const int64_t dst_nb[4], const void *src, const int64_t src_ne[4],
                    const int64_t src_nb[4]) {
    PsNd d, s;
    int64_t nd = 1, ns = 1;
    for (int k = 0; k < 4; k++) {
        d.ne[k] = dst_ne[k]; d.nb[k] = dst_nb[k]; s.ne[k] = src_ne[k]; s.nb[k] = src_nb[k];
        if (dst_ne[k] <= 0 || src_ne[k] <= 0) return fail(ctx, PS_CUDA_ERR_INVALID, "copy_4d: empty dimension");
        nd *= dst_ne[k]; ns *= src_ne[k];
    }
    if (nd != ns) return fail(ctx, PS_CUDA_ERR_INVALID, "copy_4d: %lld elements into %lld", (long long)ns, (long long)nd);
    ps_k_copy_4d<<<grid1d(ns), 256, 0, ctx->stream>>>((uint8_t *)dst, d, (const uint8_t *)src, s, ns);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_matmul_f32(ps_cuda_ctx *ctx, float *dst, const void *src0, int64_t ne00, int64_t ne01, int64_t ne02, int64_t nb01, int64_t nb02,
                       const void *src1, int64_t ne11, int64_t ne12, int64_t nb11, int64_t nb12) {
    if (ne00 <= 0 || ne01 <= 0 || ne02 <= 0 || ne11 <= 0 || ne12 <= 0 || ne12 % ne02)
        return fail(ctx, PS_CUDA_ERR_INVALID, "matmul_f32: src1 heads (%lld) must be a multiple of src0 heads (%lld)", (long long)ne12, (long long)ne02);
    const int64_t n_out = ne01 * ne11 * ne12;
    ps_k_matmul_f32<<<(unsigned)std::min<int64_t>((n_out + 3) / 4, 1 << 20), 128, 0, ctx->stream>>>(dst, (const uint8_t *)src0, ne00, ne01, ne02, nb01, nb02,
                                                                                                    (const uint8_t *)src1, ne11, ne12, nb11, nb12);
    PS_LAUNCH_CK();
    return 0;
}

int ps_cuda_softmax(ps_cuda_ctx *ctx, float *dst, const float *x, int64_t ne0, int64_t n_rows) {
    if (ne0 * 4 > 160 * 1024) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "softmax: row of %lld exceeds shared memory", (long long)ne0);
    static bool attr[64] = {};
    if (!attr[ctx->device]) { PS_CK(cudaFuncSetAttribute(ps_k_softmax_ext, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr[ctx->device] = true; }
    ps_k_softmax_ext<<<(unsigned)n_rows, 256, (size_t)ne0 * 4, ctx->stream>>>(dst, x, nullptr, nullptr, ne0, 1, 1.0f);
    PS_LAUNCH_CK();
    return 0;
}

// ---------------------------------------------------------------------------------------------- KV cache
// mask bookkeeping shared by every path that moves the position (kv_cache.hpp:243-271: advance unmasks, rollback masks)
static void kv_set_mask(ps_cuda_ctx *ctx, int from, int to, uint8_t v) {
    from = std::max(from, 0);
    to = std::min(to, ctx->d.n_ctx);
    for (int i = from; i < to; i++)
        if (ctx->slot_mask[i] != v) { ctx->slot_mask[i] = v; ctx->slot_mask_dirty = true; }
}
int ps_cuda_kv_position(ps_cuda_ctx *ctx) { return ctx->position; }
int ps_cuda_kv_reset(ps_cuda_ctx *ctx) {
    kv_set_mask(ctx, 0, ctx->position, 1);
    ctx->position = 0;
    return 0;
}
int ps_cuda_kv_rollback(ps_cuda_ctx *ctx, int n) {
    if (n < 0 || n > ctx->position) return fail(ctx, PS_CUDA_ERR_INVALID, "kv_rollback: %d > position %d (POWERSERVE_ASSERT_KVCACHE)", n, ctx->position);
    ctx->position -= n;
    kv_set_mask(ctx, ctx->position, ctx->position + n, 1); // rollback_tokens, kv_cache.hpp:255-263
    return 0;
}
int ps_cuda_kv_truncate(ps_cuda_ctx *ctx, int n) {
    if (n < 0) return fail(ctx, PS_CUDA_ERR_INVALID, "kv_truncate: negative size");
    if (n < ctx->position) return ps_cuda_kv_rollback(ctx, ctx->position - n); // truncate_tokens, kv_cache.hpp:265-271
    return 0;
}
int ps_cuda_kv_advance(ps_cuda_ctx *ctx, int n) {
    if (n < 0 || ctx->position + n > ctx->d.n_ctx) return fail(ctx, PS_CUDA_ERR_KV_FULL, "the length of kvcache is up to the preset threshold: %d", ctx->d.n_ctx);
    kv_set_mask(ctx, ctx->position, ctx->position + n, 0); // advance_tokens, kv_cache.hpp:243-253
    ctx->position += n;
    return 0;
}
// ---- slot operations of KVCacheInterface (speculative decode, kv_cache.hpp:120-143, 188-231)
static int kv_slot_ck(ps_cuda_ctx *ctx, const char *what, int idx, int limit) {
    if (idx < 0 || idx >= limit) return fail(ctx, PS_CUDA_ERR_INVALID, "%s: index %d outside [0,%d) (POWERSERVE_ASSERT_KVCACHE)", what, idx, limit);
    return 0;
}
int ps_cuda_kv_copy_slot(ps_cuda_ctx *ctx, int dst_cache_index, int src_token_index) {
    int rc;
    if ((rc = kv_slot_ck(ctx, "kv_copy_slot (cache index)", dst_cache_index, ctx->d.n_ctx))) return rc;
    if ((rc = kv_slot_ck(ctx, "kv_copy_slot (token of the last batch)", src_token_index, ctx->last_batch))) return rc;
    if (ctx->tp > 1) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "kv slot operations are single-GPU");
    const int64_t kvd = (int64_t)ctx->d.n_kv_heads * ctx->d.head_size, lstride = std::min<int64_t>(ctx->d.max_batch, 32) * kvd;
    ps_k_kv_copy_slot<<<dim3((unsigned)((kvd + 255) / 256), (unsigned)ctx->d.n_layers), 256, 0, ctx->stream>>>(ctx->kv_ptrs_dev, ctx->d.n_layers, ctx->k_stage, ctx->v_stage, lstride,
                                                                                                             kvd, ctx->d.n_ctx, dst_cache_index, src_token_index);
    PS_LAUNCH_CK();
    return 0;
}
int ps_cuda_kv_move_slot(ps_cuda_ctx *ctx, int dst_cache_index, int src_cache_index) {
    int rc;
    if ((rc = kv_slot_ck(ctx, "kv_move_slot (dst)", dst_cache_index, ctx->d.n_ctx))) return rc;
    if ((rc = kv_slot_ck(ctx, "kv_move_slot (src)", src_cache_index, ctx->d.n_ctx))) return rc;
    if (ctx->tp > 1) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "kv slot operations are single-GPU");
    if (dst_cache_index == src_cache_index) return 0;
    const int64_t kvd = (int64_t)ctx->d.n_kv_heads * ctx->d.head_size;
    ps_k_kv_move_slot<<<dim3((unsigned)((kvd + 255) / 256), (unsigned)ctx->d.n_layers), 256, 0, ctx->stream>>>(ctx->kv_ptrs_dev, ctx->d.n_layers, kvd, ctx->d.n_ctx, dst_cache_index,
                                                                                                             src_cache_index);
    PS_LAUNCH_CK();
    return 0;
}
int ps_cuda_kv_mask_slot(ps_cuda_ctx *ctx, int cache_index) {
    int rc = kv_slot_ck(ctx, "kv_mask_slot", cache_index, ctx->position); // POWERSERVE_ASSERT_KVCACHE(cache_index < position), kv_cache.hpp:224
    if (rc) return rc;
    kv_set_mask(ctx, cache_index, cache_index + 1, 1);
    return 0;
}
int ps_cuda_kv_unmask_slot(ps_cuda_ctx *ctx, int cache_index) {
    int rc = kv_slot_ck(ctx, "kv_unmask_slot", cache_index, ctx->position);
    if (rc) return rc;
    kv_set_mask(ctx, cache_index, cache_index + 1, 0);
    return 0;
}
float *ps_cuda_kv_k(ps_cuda_ctx *ctx, int layer) { return (layer >= 0 && layer < ctx->d.n_layers) ? ctx->kc[layer] : nullptr; }
float *ps_cuda_kv_v(ps_cuda_ctx *ctx, int layer) { return (layer >= 0 && layer < ctx->d.n_layers) ? ctx->vct[layer] : nullptr; }

// ---------------------------------------------------------------------------------------------- whole model
// ---- tcgen05 prefill GEMM (ps_tc.cuh)
static int tc_expand(ps_cuda_ctx *ctx, uint8_t *dst, const uint8_t *w, int64_t n_rows, int64_t K, int64_t tile0) {
    ps_k_tc_expand_a<<<dim3((unsigned)(K / 256), (unsigned)((n_rows + PS_TC_M - 1) / PS_TC_M)), 256, 0, ctx->stream>>>(dst, w, n_rows, K / 256, tile0);
    PS_LAUNCH_CK();
    return 0;
}
static int tc_prep_b(ps_cuda_ctx *ctx, const float *x, int K, int bs) {
    const int n_cg = (bs + PS_TC_N - 1) / PS_TC_N;
    ps_k_tc_prep_b<<<dim3((unsigned)((K / 256 + 3) / 4), (unsigned)(n_cg * PS_TC_N)), 128, 0, ctx->stream>>>(ctx->tc_b, x, K, bs);
    PS_LAUNCH_CK();
    return 0;
}
static int tc_gemm(ps_cuda_ctx *ctx, const uint8_t *a_blocks, int n_rows, int K, int bs, const PsRwSeg *segs, int n_seg, const float *residual) {
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_tc_gemm, cudaFuncAttributeMaxDynamicSharedMemorySize, PS_TC_STAGES * PS_TC_STAGE));
        attr[ctx->device] = true;
    }
    PsTcArgs a{};
    a.err = ctx->tc_err_dev;
    a.a = a_blocks; a.b = ctx->tc_b; a.nb = K / 256; a.n_cg = (bs + PS_TC_N - 1) / PS_TC_N; a.n_seg = n_seg; a.bs = bs; a.residual = residual;
    for (int t = 0; t < n_seg; t++) a.seg[t] = segs[t];
    const int n_tiles = (n_rows + PS_TC_M - 1) / PS_TC_M;
    a.n_units = n_tiles * a.n_cg;
    ps_k_tc_gemm<<<(unsigned)std::min(a.n_units, ctx->n_sm), PS_TC_THREADS, PS_TC_STAGES * PS_TC_STAGE, ctx->stream>>>(a);
    PS_LAUNCH_CK();
    ctx->n_tc++;
    return 0;
}
static int tc_single(ps_cuda_ctx *ctx, const uint8_t *a_blocks, int n_rows, int K, float *dst, int bs, const float *residual) {
    PsRwSeg sg = {dst, nullptr, 0, n_rows, 0};
    return tc_gemm(ctx, a_blocks, n_rows, K, bs, &sg, 1, residual);
}

int ps_cuda_bind_model(ps_cuda_ctx *ctx, const ps_cuda_model_weights *w) {
    const ps_cuda_model_desc &d = ctx->d;
    const int64_t qdim = (int64_t)d.n_heads * d.head_size, kvd = (int64_t)d.n_kv_heads * d.head_size;
    auto reg = [&](const ps_cuda_tensor &t, int64_t ne0, int64_t ne1, const void **dev, int *type) -> int {
        if (!t.host) return fail(ctx, PS_CUDA_ERR_INVALID, "bind_model: missing tensor");
        void *p = nullptr;
        int rc = ps_cuda_register_weight(ctx, t.host, t.type, ne0, ne1, &p);
        if (rc) return rc;
        *dev = p;
        if (type) *type = t.type;
        return 0;
    };
#define REG(t, ne0, ne1, dev, type)                                   \
    do {                                                              \
        int rc_ = reg(t, ne0, ne1, (const void **)&(dev), type);      \
        if (rc_) return rc_;                                          \
    } while (0)
    // tensor parallel rank r owns rows [r * rows / tp, (r + 1) * rows / tp) of EVERY matrix (rows are contiguous in GGUF)
    const int tp = ctx->tp, rank = ctx->rank;
    const int64_t qdim_l = qdim / tp, kvd_l = kvd / tp, ffn_l = ctx->ffn_l, vocab_l = ctx->vocab_l, dim_l = ctx->dim_l;
    auto rows_of = [&](ps_cuda_tensor t, int64_t ne0, int64_t rows_l) {
        if (t.host && tp > 1) t.host = (const uint8_t *)t.host + (size_t)rank * (size_t)rows_l * (size_t)ps_row_bytes(t.type, ne0);
        return t;
    };
    REG(w->token_embd, d.dim, d.vocab_size, ctx->w_embd, &ctx->t_embd);
    REG(rows_of(w->output, d.dim, vocab_l), d.dim, vocab_l, ctx->w_out, &ctx->t_out);
    REG(w->output_norm, d.dim, 1, ctx->w_out_norm, nullptr);
    if (w->output_norm.type != 0) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "norm weights must be F32");
    ctx->layers.assign(d.n_layers, LayerDev{});
    for (int L = 0; L < d.n_layers; L++) {
        const ps_cuda_layer_weights &lw = w->layers[L];
        LayerDev &ld = ctx->layers[L];
        if (lw.attn_norm.type != 0 || lw.ffn_norm.type != 0) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "norm weights must be F32");
        REG(lw.attn_norm, d.dim, 1, ld.attn_norm, nullptr);
        REG(lw.ffn_norm, d.dim, 1, ld.ffn_norm, nullptr);
        REG(rows_of(lw.attn_q, d.dim, qdim_l), d.dim, qdim_l, ld.wq, &ld.tq);
        REG(rows_of(lw.attn_k, d.dim, kvd_l), d.dim, kvd_l, ld.wk, &ld.tk);
        REG(rows_of(lw.attn_v, d.dim, kvd_l), d.dim, kvd_l, ld.wv, &ld.tv);
        REG(rows_of(lw.attn_output, qdim, dim_l), qdim, dim_l, ld.wo, &ld.to);
        REG(rows_of(lw.ffn_gate, d.dim, ffn_l), d.dim, ffn_l, ld.wgate, &ld.tgate);
        REG(rows_of(lw.ffn_up, d.dim, ffn_l), d.dim, ffn_l, ld.wup, &ld.tup);
        REG(rows_of(lw.ffn_down, d.ffn_dim, dim_l), d.ffn_dim, dim_l, ld.wdown, &ld.tdown);
        if (d.qkv_bias) {
            if (lw.attn_q_bias.type != 0 || lw.attn_k_bias.type != 0 || lw.attn_v_bias.type != 0)
                return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "bias tensors must be F32");
            REG(rows_of(lw.attn_q_bias, 1, qdim_l), qdim_l, 1, ld.q_bias, nullptr);
            REG(rows_of(lw.attn_k_bias, 1, kvd_l), kvd_l, 1, ld.k_bias, nullptr);
            REG(rows_of(lw.attn_v_bias, 1, kvd_l), kvd_l, 1, ld.v_bias, nullptr);
        }
        // q/k/v (and gate/up) share one quantised activation: their vec_dot_type must agree
        auto kq = [](int t) { return t == 12 || t == 13 || t == 14; };
        if (kq(ld.tq) != kq(ld.tk) || kq(ld.tq) != kq(ld.tv) || kq(ld.tgate) != kq(ld.tup))
            return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "layer %d mixes K-quant and 32-block weights on one input", L);
    }
#undef REG
    ctx->fused_ok = (ctx->t_embd == 12 && ctx->t_out == 12);
    for (const LayerDev &ld : ctx->layers)
        if (ld.tq != 12 || ld.tk != 12 || ld.tv != 12 || ld.to != 12 || ld.tgate != 12 || ld.tup != 12 || ld.tdown != 12) ctx->fused_ok = false;
    if (d.dim % 256 || d.ffn_dim % 256 || (int64_t)d.n_heads * d.head_size % 256 || d.dim / 256 > 64 || d.ffn_dim / 256 > 64 || d.n_heads / d.n_kv_heads > 8 ||
        (qdim_l % 8) || (kvd_l % 8) || (d.rope_type & 2) || d.n_ctx % 4 ||
        (size_t)(d.n_heads / d.n_kv_heads) * ((d.n_ctx + 31) & ~31) * 4 > 200 * 1024 || // the soft-max rows of a kv group live in shared memory
        !(d.n_heads / d.n_kv_heads == 1 || d.n_heads / d.n_kv_heads == 2 || d.n_heads / d.n_kv_heads == 4 || d.n_heads / d.n_kv_heads == 8))
        ctx->fused_ok = false;
    {   // the graph-replayed operator-table step only needs the decode attention kernels to apply
        const int r2 = d.n_heads / d.n_kv_heads, r2t = r2 <= 1 ? 1 : r2 <= 2 ? 2 : r2 <= 4 ? 4 : 8;
        ctx->ops_graph_ok = tp == 1 && d.n_heads % d.n_kv_heads == 0 && r2 <= 8 && d.head_size % 32 == 0 && d.head_size <= 256 && d.n_ctx % 4 == 0 &&
                            (size_t)r2t * ((d.n_ctx + 31) & ~31) * 4 <= 200 * 1024 && d.vocab_size >= 256;
    }
    {   // all-Q4_0 / all-Q8_0 model: octet copies for the fused 32-block decode path (ps_mv32.cuh)
        const int t0 = ctx->t_out;
        bool same = (t0 == 2 || t0 == 8) && ctx->ops_graph_ok;
        for (const LayerDev &ld : ctx->layers)
            if (ld.tq != t0 || ld.tk != t0 || ld.tv != t0 || ld.to != t0 || ld.tgate != t0 || ld.tup != t0 || ld.tdown != t0) same = false;
        if (d.dim % 32 || d.ffn_dim % 32 || qdim % 32 || qdim % 8 || kvd % 8 || d.dim % 8 || d.ffn_dim % 8 || d.vocab_size % 8 || d.head_size % 2) same = false;
        ctx->mv_ok = false;
        if (same) {
            ctx->mv_type = t0;
            int rc;
            auto mk = [&](uint8_t **p, int64_t rows, int64_t K) -> int {
                const size_t bytes = mv_bytes(t0, rows, K);
                int rc2 = dev_alloc(ctx, (void **)p, bytes);
                if (rc2) return rc2;
                PS_CK(cudaMemsetAsync(*p, 0, bytes, ctx->stream));
                return 0;
            };
            const int64_t dim = d.dim, ffn = d.ffn_dim;
            for (LayerDev &ld : ctx->layers) {
                if ((rc = mk(&ld.mv_qkv, qdim + 2 * kvd, dim)) || (rc = mk(&ld.mv_o, dim, qdim)) || (rc = mk(&ld.mv_gu, 2 * ffn, dim)) || (rc = mk(&ld.mv_down, dim, ffn))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_qkv, ld.wq, qdim, dim, 0))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_qkv, ld.wk, kvd, dim, qdim / 8))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_qkv, ld.wv, kvd, dim, (qdim + kvd) / 8))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_o, ld.wo, dim, qdim, 0))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_gu, ld.wgate, ffn, dim, 0))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_gu, ld.wup, ffn, dim, ffn / 8))) return rc;
                if ((rc = mv_repack_t(ctx, ld.mv_down, ld.wdown, dim, ffn, 0))) return rc;
            }
            if ((rc = mk(&ctx->mv_out, d.vocab_size, dim))) return rc;
            if ((rc = mv_repack_t(ctx, ctx->mv_out, ctx->w_out, d.vocab_size, dim, 0))) return rc;
            PS_CK(cudaStreamSynchronize(ctx->stream));
            ctx->mv_ok = true;
        }
    }
    if (tp > 1 && !ctx->fused_ok) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "tensor parallelism needs the fused Q4_K decode path (all-Q4_K llama-style model)");
    if (ctx->fused_ok) {
        // octet-interleaved copies for the row-walker mat-vec (a permutation of the same bytes; see ps_rw.cuh)
        const int64_t dim = d.dim, ffn = d.ffn_dim;
        auto oct_bytes = [](int64_t rows, int64_t K, int slots) { return (size_t)((rows + 7) / 8) * (size_t)(K / 256) * slots * PS_RW_OCTET_BLOCK; };
        int rc;
        for (LayerDev &ld : ctx->layers) {
            if ((rc = dev_alloc(ctx, (void **)&ld.rw_qkv, oct_bytes(qdim_l + 2 * kvd_l, dim, 1)))) return rc;
            if ((rc = dev_alloc(ctx, (void **)&ld.rw_o, oct_bytes(dim_l, qdim, 1)))) return rc;
            if ((rc = dev_alloc(ctx, (void **)&ld.rw_gu, oct_bytes(ffn_l, dim, 2)))) return rc;
            if ((rc = dev_alloc(ctx, (void **)&ld.rw_down, oct_bytes(dim_l, ffn, 1)))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_qkv, ld.wq, qdim_l, dim, 0, 0, 1))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_qkv, ld.wk, kvd_l, dim, qdim_l / 8, 0, 1))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_qkv, ld.wv, kvd_l, dim, (qdim_l + kvd_l) / 8, 0, 1))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_o, ld.wo, dim_l, qdim, 0, 0, 1))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_gu, ld.wgate, ffn_l, dim, 0, 0, 2))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_gu, ld.wup, ffn_l, dim, 0, 1, 2))) return rc;
            if ((rc = rw_repack(ctx, ld.rw_down, ld.wdown, dim_l, ffn, 0, 0, 1))) return rc;
        }
        // fp16-expanded tensor-core operands for the prefill GEMM (2 B / weight of this rank's rows; only if HBM has room)
        ctx->tc_ok = false;
        if (ctx->opt_tc && qdim_l % PS_TC_M == 0 && kvd_l % PS_TC_M == 0) {
            auto a_bytes = [](int64_t rows, int64_t K) { return (size_t)((rows + PS_TC_M - 1) / PS_TC_M) * (size_t)(K / 256) * PS_TC_A_BLOCK; };
            const size_t per_layer = a_bytes(qdim_l + 2 * kvd_l, dim) + a_bytes(dim_l, qdim) + 2 * a_bytes(ffn_l, dim) + a_bytes(dim_l, ffn);
            size_t free_b = 0, total_b = 0;
            PS_CK(cudaMemGetInfo(&free_b, &total_b));
            if (per_layer * (size_t)d.n_layers + ((size_t)8 << 30) < free_b) {
                for (LayerDev &ld : ctx->layers) {
                    if ((rc = dev_alloc(ctx, (void **)&ld.tc_qkv, a_bytes(qdim_l + 2 * kvd_l, dim)))) return rc;
                    if ((rc = dev_alloc(ctx, (void **)&ld.tc_o, a_bytes(dim_l, qdim)))) return rc;
                    if ((rc = dev_alloc(ctx, (void **)&ld.tc_gate, a_bytes(ffn_l, dim)))) return rc;
                    if ((rc = dev_alloc(ctx, (void **)&ld.tc_up, a_bytes(ffn_l, dim)))) return rc;
                    if ((rc = dev_alloc(ctx, (void **)&ld.tc_down, a_bytes(dim_l, ffn)))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_qkv, ld.wq, qdim_l, dim, 0))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_qkv, ld.wk, kvd_l, dim, qdim_l / PS_TC_M))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_qkv, ld.wv, kvd_l, dim, (qdim_l + kvd_l) / PS_TC_M))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_o, ld.wo, dim_l, qdim, 0))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_gate, ld.wgate, ffn_l, dim, 0))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_up, ld.wup, ffn_l, dim, 0))) return rc;
                    if ((rc = tc_expand(ctx, ld.tc_down, ld.wdown, dim_l, ffn, 0))) return rc;
                }
                const size_t bb = (size_t)((d.max_batch + PS_TC_N - 1) / PS_TC_N) * (size_t)(ctx->maxK / 256) * PS_TC_B_BLOCK;
                if ((rc = dev_alloc(ctx, (void **)&ctx->tc_b, bb))) return rc;
                PS_CK(cudaMemsetAsync(ctx->tc_b, 0, bb, ctx->stream)); // the mins tiles are zero outside each lane's four K positions
                ctx->tc_b_bytes = bb;
                ctx->tc_ok = true;
                if (tp == 1 && a_bytes(vocab_l, dim) + ((size_t)8 << 30) < free_b - per_layer * (size_t)d.n_layers) { // lm_head too (1 GB for the 8B model)
                    if ((rc = dev_alloc(ctx, (void **)&ctx->tc_out, a_bytes(vocab_l, dim)))) return rc;
                    if ((rc = tc_expand(ctx, ctx->tc_out, ctx->w_out, vocab_l, dim, 0))) return rc;
                }
            }
        }
        if ((rc = dev_alloc(ctx, (void **)&ctx->rw_out, oct_bytes(vocab_l, dim, 1)))) return rc;
        if ((rc = rw_repack(ctx, ctx->rw_out, ctx->w_out, vocab_l, dim, 0, 0, 1))) return rc;
        {   // the persistent step kernel's per-layer table
            std::vector<PsStLayer> tab(d.n_layers);
            for (int L = 0; L < d.n_layers; L++) {
                const LayerDev &ld = ctx->layers[L];
                tab[L] = PsStLayer{ld.rw_qkv, ld.rw_o, ld.rw_gu, ld.rw_down, ld.attn_norm, ld.ffn_norm, d.qkv_bias ? ld.q_bias : nullptr,
                                   d.qkv_bias ? ld.k_bias : nullptr, d.qkv_bias ? ld.v_bias : nullptr, ctx->kc[L], ctx->vct[L]};
            }
            if (!ctx->st_layers && (rc = dev_alloc(ctx, (void **)&ctx->st_layers, sizeof(PsStLayer) * (size_t)d.n_layers))) return rc;
            PS_CK(cudaMemcpyAsync(ctx->st_layers, tab.data(), sizeof(PsStLayer) * (size_t)d.n_layers, cudaMemcpyHostToDevice, ctx->stream));
            PS_CK(cudaStreamSynchronize(ctx->stream)); // `tab` dies at the end of this scope
            ctx->step_ok = (d.head_size == 64 || d.head_size == 128) && ffn_l % 256 == 0 && d.vocab_size / 8 >= 1 && ctx->smem_optin >= 160 * 1024;
        }
        PS_CK(cudaStreamSynchronize(ctx->stream));
    }
    if (ctx->g_step) { cudaGraphExecDestroy(ctx->g_step); ctx->g_step = nullptr; }
    if (ctx->g_fwd) { cudaGraphExecDestroy(ctx->g_fwd); ctx->g_fwd = nullptr; }
    ctx->bound = true;
    return 0;
}

// LlamaModel::forward / Qwen2Model::forward on the device, one kernel per table op (the fused / graph-replayed decode
// path lives in ps_decode.cuh and is bit-identical).
// `tree_base` >= 0: a tree batch (ps_cuda_forward_tree) - K / V rows go to cache SLOTS tree_base .. tree_base + bs - 1 whatever the
// token positions are, and the attention bias is ctx->tree_bias (slot mask + in-batch tree mask) instead of the causal `pos` mask.
struct SessRun { int n_kv_max; };
// Scores and P.V of a batch over one cache (prefill chunk, verify batch): register-tiled kernels (option attn_tile, on
// by default), else the round-1 kernels; narrow batches use the per-position / per-(head, dim) warp kernels.  nh / nkv
// are the heads THIS rank owns.  The caller checks the launch (PS_LAUNCH_CK) after launch_batch_scores.
static void launch_batch_scores(ps_cuda_ctx *ctx, int L, int nh, int nkv, int hs, int64_t n_kv, int bs) {
    if (bs >= ctx->opt_scores_batch_min) {
        const int r2 = nh / nkv;
        const int qb = std::max(1, std::min(bs, 8192 / (r2 * hs)));
        const dim3 grid((unsigned)((n_kv + 31) / 32), (unsigned)nkv, (unsigned)((bs + qb - 1) / qb));
        const size_t smem = (size_t)qb * r2 * hs * 4;
        if (ctx->opt_attn_tile && hs == 128) ps_k_attn_scores_tile<4><<<grid, 128, smem, ctx->stream>>>(ctx->kq, ctx->kc[L], ctx->qr, nh, nkv, n_kv, bs, qb);
        else if (ctx->opt_attn_tile && hs == 64) ps_k_attn_scores_tile<2><<<grid, 128, smem, ctx->stream>>>(ctx->kq, ctx->kc[L], ctx->qr, nh, nkv, n_kv, bs, qb);
        else if (ctx->opt_attn_tile && hs == 32) ps_k_attn_scores_tile<1><<<grid, 128, smem, ctx->stream>>>(ctx->kq, ctx->kc[L], ctx->qr, nh, nkv, n_kv, bs, qb);
        else if (ctx->opt_attn_tile && hs == 256) ps_k_attn_scores_tile<8><<<grid, 128, smem, ctx->stream>>>(ctx->kq, ctx->kc[L], ctx->qr, nh, nkv, n_kv, bs, qb);
        else ps_k_attn_scores_batch<<<grid, 128, (size_t)qb * r2 * hs * 4, ctx->stream>>>(ctx->kq, ctx->kc[L], ctx->qr, hs, nh, nkv, n_kv, bs, qb);
    } else {
        ps_k_attn_scores<<<dim3((unsigned)((n_kv + 3) / 4), (unsigned)nkv), 128, 0, ctx->stream>>>(ctx->kq, ctx->kc[L], ctx->qr, hs, nh, nkv, n_kv, bs);
    }
}
static int launch_batch_pv(ps_cuda_ctx *ctx, int L, int nh, int nkv, int hs, int64_t n_kv, int bs, int batch_min) {
    const int64_t n_ctx = ctx->d.n_ctx;
    if (bs >= batch_min && (size_t)PS_PV_QB * n_kv * 4 <= 200 * 1024) {
        static bool pv_attr[64] = {};
        if (!pv_attr[ctx->device]) {
            PS_CK(cudaFuncSetAttribute(ps_k_attn_pv_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            PS_CK(cudaFuncSetAttribute(ps_k_attn_pv_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            pv_attr[ctx->device] = true;
        }
        dim3 grid((unsigned)((bs + PS_PV_QB - 1) / PS_PV_QB), (unsigned)nh);
        const dim3 grid1 = grid;
        const bool tile = ctx->opt_attn_tile && hs % PS_PVT_D == 0 && PS_PVT_RING + (size_t)PS_PVT_Q * n_kv * 4 <= 200 * 1024;
        if (tile)
            while ((int)(grid.x * grid.y * grid.z) * 2 <= ctx->n_sm * 2 && (int)grid.z * 2 * 8 * PS_PVT_D <= hs) grid.z *= 2; // narrow batch: split the dim groups over more CTAs
        static_assert(PS_PV_QB == PS_PVT_Q, "both P.V kernels block the queries by 8");
        if (tile) ps_k_attn_pv_tile<<<grid, 256, PS_PVT_RING + (size_t)PS_PVT_Q * n_kv * 4, ctx->stream>>>(ctx->att, ctx->vct[L], ctx->kq, hs, nh, nkv, n_kv, n_ctx, bs);
        else ps_k_attn_pv_batch<<<grid1, 256, (size_t)PS_PV_QB * n_kv * 4, ctx->stream>>>(ctx->att, ctx->vct[L], ctx->kq, hs, nh, nkv, n_kv, n_ctx, bs);
    } else {
        ps_k_attn_pv<<<dim3((unsigned)((hs + 3) / 4), (unsigned)nkv), 128, 0, ctx->stream>>>(ctx->att, ctx->vct[L], ctx->kq, hs, nh, nkv, n_kv, n_ctx, bs);
    }
    PS_LAUNCH_CK();
    return 0;
}

static int forward_ops(ps_cuda_ctx *ctx, int bs, int lm_head, int pos0, int tree_base = -1, const SessRun *sess = nullptr) {
    const ps_cuda_model_desc &d = ctx->d;
    const int64_t dim = d.dim, hs = d.head_size, nh = d.n_heads, nkv = d.n_kv_heads, kvd = hs * nkv, qdim = nh * hs, ffn = d.ffn_dim;
    const bool tree = tree_base >= 0;
    const int64_t n_kv = sess ? (int64_t)sess->n_kv_max : tree ? (int64_t)tree_base + bs : (int64_t)pos0 + bs; // pos.back() + 1 (session batch: the longest row)
    const float kq_scale = 1.0f / sqrtf((float)hs);
    const bool tc = ctx->tc_ok && ctx->opt_tc && ctx->opt_fused && bs >= ctx->opt_tc_min; // tensor-core GEMM on the fp16-expanded operands
    const bool rw = ctx->fused_ok && ctx->opt_fused && (bs > 1 || tree); // octet-interleaved copies exist: multi-column row-walker
    int rc;
    ps_k_get_embedding<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->x, ctx->w_embd, ctx->t_embd, dim, ctx->tokens_dev);
    PS_LAUNCH_CK();
    static bool attr[64] = {};
    if (!attr[ctx->device]) {
        PS_CK(cudaFuncSetAttribute(ps_k_softmax_ext, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        PS_CK(cudaFuncSetAttribute(ps_k_sess_softmax, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        attr[ctx->device] = true;
    }
    for (int L = 0; L < d.n_layers; L++) {
        const LayerDev &ld = ctx->layers[L];
        ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->xn, ctx->x, ld.attn_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if (tc) {
            if ((rc = tc_prep_b(ctx, ctx->xn, (int)dim, bs))) return rc;
            PsRwSeg sg[3] = {{ctx->q, d.qkv_bias ? ld.q_bias : nullptr, 0, (int)qdim, 0},
                             {ctx->k, d.qkv_bias ? ld.k_bias : nullptr, (int)qdim, (int)(qdim + kvd), 0},
                             {ctx->v, d.qkv_bias ? ld.v_bias : nullptr, (int)(qdim + kvd), (int)(qdim + 2 * kvd), 0}};
            if ((rc = tc_gemm(ctx, ld.tc_qkv, (int)(qdim + 2 * kvd), (int)dim, bs, sg, 3, nullptr))) return rc;
        } else if (rw) { // q | k | v rows in one pass over the concatenated octets
            if ((rc = rwm_quantize(ctx, ctx->xn, (int)dim, bs))) return rc;
            PsRwmArgs a{};
            a.w = ld.rw_qkv; a.n_oct = (int)((qdim + 2 * kvd) / 8); a.K = (int)dim; a.slot = 0; a.n_slots = 1; a.n_seg = 3; a.bs = bs;
            a.seg[0] = {ctx->q, d.qkv_bias ? ld.q_bias : nullptr, 0, (int)qdim, 0};
            a.seg[1] = {ctx->k, d.qkv_bias ? ld.k_bias : nullptr, (int)qdim, (int)(qdim + kvd), 0};
            a.seg[2] = {ctx->v, d.qkv_bias ? ld.v_bias : nullptr, (int)(qdim + kvd), (int)(qdim + 2 * kvd), 0};
            if ((rc = launch_rwm(ctx, a))) return rc;
        } else {
            if ((rc = quantize_act(ctx, ld.tq, ctx->xn, dim, bs))) return rc;
            if ((rc = matmul_q(ctx, ctx->q, ld.wq, ld.tq, dim, qdim, bs, d.qkv_bias ? ld.q_bias : nullptr, nullptr))) return rc;
            if ((rc = matmul_q(ctx, ctx->k, ld.wk, ld.tk, dim, kvd, bs, d.qkv_bias ? ld.k_bias : nullptr, nullptr))) return rc;
            if ((rc = matmul_q(ctx, ctx->v, ld.wv, ld.tv, dim, kvd, bs, d.qkv_bias ? ld.v_bias : nullptr, nullptr))) return rc;
        }
        ps_k_rope<<<dim3((unsigned)nh, (unsigned)bs), 64, 0, ctx->stream>>>(ctx->qr, ctx->q, (int)hs, d.rope_n_dims, d.rope_type & 2, ctx->pos_dev, ctx->rope_table);
        PS_LAUNCH_CK();
        ps_k_rope<<<dim3((unsigned)nkv, (unsigned)bs), 64, 0, ctx->stream>>>(ctx->kr, ctx->k, (int)hs, d.rope_n_dims, d.rope_type & 2, ctx->pos_dev, ctx->rope_table);
        PS_LAUNCH_CK();
        if (sess) {
            float **pl = ctx->sess_ptrs_dev + (size_t)L * 2 * d.max_batch;
            ps_k_sess_kv_store<<<grid1d(kvd * bs), 256, 0, ctx->stream>>>(pl, pl + d.max_batch, ctx->kr, ctx->v, kvd, d.n_ctx, ctx->pos_dev, bs);
        } else if (tree) {
            const size_t lstride = (size_t)std::min<int64_t>(d.max_batch, 32) * (size_t)kvd;
            ps_k_kv_store_at<<<grid1d(kvd * bs), 256, 0, ctx->stream>>>(ctx->kc[L], ctx->vct[L], ctx->k_stage + L * lstride, ctx->v_stage + L * lstride, ctx->kr, ctx->v,
                                                                         kvd, d.n_ctx, tree_base, bs);
        } else {
            ps_k_kv_store<<<grid1d(kvd * bs), 256, 0, ctx->stream>>>(ctx->kc[L], ctx->vct[L], ctx->kr, ctx->v, kvd, d.n_ctx, ctx->pos_dev, bs);
        }
        PS_LAUNCH_CK();
        if (sess) { // every column attends over its own session's cache at its own position
            float **pl = ctx->sess_ptrs_dev + (size_t)L * 2 * d.max_batch;
            const int64_t cs = nh * (int64_t)d.n_ctx;
            ps_k_sess_scores<<<dim3((unsigned)((n_kv + 3) / 4), (unsigned)nkv, (unsigned)bs), 128, 0, ctx->stream>>>(ctx->kq, pl, ctx->qr, (int)hs, (int)nh, (int)nkv, ctx->pos_dev, cs);
            PS_LAUNCH_CK();
            ps_k_sess_softmax<<<dim3((unsigned)nh, (unsigned)bs), 256, (size_t)n_kv * 4, ctx->stream>>>(ctx->kq, ctx->pos_dev, cs, kq_scale);
            PS_LAUNCH_CK();
            ps_k_sess_pv<<<dim3((unsigned)((hs + 3) / 4), (unsigned)nkv, (unsigned)bs), 128, 0, ctx->stream>>>(ctx->att, pl + d.max_batch, ctx->kq, (int)hs, (int)nh, (int)nkv, ctx->pos_dev, d.n_ctx, cs);
            PS_LAUNCH_CK();
        } else {
        launch_batch_scores(ctx, L, (int)nh, (int)nkv, (int)hs, n_kv, bs);
        PS_LAUNCH_CK();
        ps_k_softmax_ext<<<(unsigned)(bs * nh), 256, (size_t)n_kv * 4, ctx->stream>>>(ctx->kq, ctx->kq, tree ? ctx->tree_bias : nullptr, ctx->pos_dev, n_kv, bs, kq_scale);
        PS_LAUNCH_CK();
        if ((rc = launch_batch_pv(ctx, L, (int)nh, (int)nkv, (int)hs, n_kv, bs, ctx->opt_pv_batch_min))) return rc;
        PS_LAUNCH_CK();
        }
        if (tc) {
            if ((rc = tc_prep_b(ctx, ctx->att, (int)qdim, bs))) return rc;
            if ((rc = tc_single(ctx, ld.tc_o, (int)dim, (int)qdim, ctx->x, bs, ctx->x))) return rc;
        } else if (rw) {
            if ((rc = rwm_quantize(ctx, ctx->att, (int)qdim, bs))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_o, (int)dim, (int)qdim, 0, 1, ctx->x, bs, nullptr, ctx->x))) return rc;
        } else {
            if ((rc = quantize_act(ctx, ld.to, ctx->att, qdim, bs))) return rc;
            if ((rc = matmul_q(ctx, ctx->x, ld.wo, ld.to, qdim, dim, bs, nullptr, ctx->x))) return rc; // x = x + Wo.att
        }
        ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->xn, ctx->x, ld.ffn_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if (tc) {
            if ((rc = tc_prep_b(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = tc_single(ctx, ld.tc_gate, (int)ffn, (int)dim, ctx->g, bs, nullptr))) return rc;
            if ((rc = tc_single(ctx, ld.tc_up, (int)ffn, (int)dim, ctx->u, bs, nullptr))) return rc;
        } else if (rw) {
            if ((rc = rwm_quantize(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_gu, (int)ffn, (int)dim, 0, 2, ctx->g, bs, nullptr, nullptr))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_gu, (int)ffn, (int)dim, 1, 2, ctx->u, bs, nullptr, nullptr))) return rc;
        } else {
            if ((rc = quantize_act(ctx, ld.tgate, ctx->xn, dim, bs))) return rc;
            if ((rc = matmul_q(ctx, ctx->g, ld.wgate, ld.tgate, dim, ffn, bs, nullptr, nullptr))) return rc;
            if ((rc = matmul_q(ctx, ctx->u, ld.wup, ld.tup, dim, ffn, bs, nullptr, nullptr))) return rc;
        }
        ps_k_silu_hadamard<<<grid1d(ffn * bs), 256, 0, ctx->stream>>>(ctx->g, ctx->g, ctx->u, ffn * bs);
        PS_LAUNCH_CK();
        if (tc) {
            if ((rc = tc_prep_b(ctx, ctx->g, (int)ffn, bs))) return rc;
            if ((rc = tc_single(ctx, ld.tc_down, (int)dim, (int)ffn, ctx->x, bs, ctx->x))) return rc;
        } else if (rw) {
            if ((rc = rwm_quantize(ctx, ctx->g, (int)ffn, bs))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_down, (int)dim, (int)ffn, 0, 1, ctx->x, bs, nullptr, ctx->x))) return rc;
        } else {
            if ((rc = quantize_act(ctx, ld.tdown, ctx->g, ffn, bs))) return rc;
            if ((rc = matmul_q(ctx, ctx->x, ld.wdown, ld.tdown, ffn, dim, bs, nullptr, ctx->x))) return rc; // x = x + Wdown.h
        }
    }
    if (lm_head) {
        ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->xn, ctx->x, ctx->w_out_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if (tc && ctx->tc_out) {
            if ((rc = tc_prep_b(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = tc_single(ctx, ctx->tc_out, d.vocab_size, (int)dim, ctx->logits, bs, nullptr))) return rc;
        } else if (rw) {
            if ((rc = rwm_quantize(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = rwm_single(ctx, ctx->rw_out, d.vocab_size, (int)dim, 0, 1, ctx->logits, bs, nullptr, nullptr))) return rc;
        } else {
            if ((rc = quantize_act(ctx, ctx->t_out, ctx->xn, dim, bs))) return rc;
            if ((rc = matmul_q(ctx, ctx->logits, ctx->w_out, ctx->t_out, dim, d.vocab_size, bs, nullptr, nullptr))) return rc;
        }
    }
    return 0;
}


// ---- tensor-parallel BATCH forward (prefill chunks, verify batches): the same row sharding as the decode step - every
// rank owns rows [r * rows / tp, ...) of every matrix and its own kv heads, so every dot product keeps its full K and the
// result is bit-identical to one GPU - with ONE all-gather per exchange for the whole chunk (4 per layer; NCCL on the
// context stream) instead of feeding the chunk token by token.  Activations are replicated in full ([bs][dim] etc.);
// `x_part` ([bs][dim_l]) is this rank's slice of the residual stream, the send buffer of its all-gathers.
__global__ void ps_k_tp_slice(float *__restrict__ dst, const float *__restrict__ src, int64_t bs, int64_t n, int64_t n_l, int64_t rank) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < bs * n_l; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = src[(t / n_l) * n + rank * n_l + t % n_l];
}
// all-gather output [tp][bs][n_l] -> [bs][tp * n_l]
__global__ void ps_k_tp_unshard(float *__restrict__ dst, const float *__restrict__ src, int64_t bs, int64_t n_l, int64_t tp) {
    const int64_t n = n_l * tp;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < bs * n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / n, e = t % n, r = e / n_l;
        dst[t] = src[(r * bs + i) * n_l + (e - r * n_l)];
    }
}
static int tp_gather_rows(ps_cuda_ctx *ctx, float *dst_full, const float *src_part, int64_t bs, int64_t n_l) {
    int rc = tp_all_gather(ctx, src_part, ctx->tp_tmp, (size_t)(bs * n_l), false, true);
    if (rc) return rc;
    ps_k_tp_unshard<<<grid1d(bs * n_l * ctx->tp), 256, 0, ctx->stream>>>(dst_full, ctx->tp_tmp, bs, n_l, ctx->tp);
    PS_LAUNCH_CK();
    return 0;
}
static int forward_ops_tp(ps_cuda_ctx *ctx, int bs, int lm_head, int pos0) {
    const ps_cuda_model_desc &d = ctx->d;
    const int64_t dim = d.dim, hs = d.head_size, qdim = (int64_t)d.n_heads * hs, ffn = d.ffn_dim;
    const int64_t nh = ctx->nh_l, nkv = ctx->nkv_l, kvd = hs * nkv, qdim_l = nh * hs, ffn_l = ctx->ffn_l, dim_l = ctx->dim_l, vocab_l = ctx->vocab_l;
    const int tp = ctx->tp, rank = ctx->rank;
    const int64_t n_kv = (int64_t)pos0 + bs;
    const float kq_scale = 1.0f / sqrtf((float)hs);
    const bool tc = ctx->tc_ok && ctx->opt_tc && bs >= ctx->opt_tc_min;
    int rc;
    if (!ctx->nccl_comm) return fail(ctx, PS_CUDA_ERR_INVALID, "tensor-parallel batch forward needs ps_cuda_tp_init");
    if (!ctx->tp_tmp) {
        const size_t widest = (size_t)std::max<int64_t>(std::max<int64_t>(qdim, dim), std::max<int64_t>(ffn, lm_head ? d.vocab_size : 0));
        if ((rc = dev_alloc(ctx, (void **)&ctx->tp_tmp, 4 * (size_t)d.max_batch * std::max<size_t>(widest, (size_t)d.vocab_size)))) return rc;
        if ((rc = dev_alloc(ctx, (void **)&ctx->tp_xb, 4 * (size_t)d.max_batch * (size_t)dim))) return rc;
        if ((rc = dev_alloc(ctx, (void **)&ctx->tp_xl, 4 * (size_t)d.max_batch * (size_t)dim_l))) return rc;
        if ((rc = dev_alloc(ctx, (void **)&ctx->tp_attb, 4 * (size_t)d.max_batch * (size_t)qdim))) return rc;
        if ((rc = dev_alloc(ctx, (void **)&ctx->tp_hb, 4 * (size_t)d.max_batch * (size_t)ffn))) return rc;
    }
    float *x = ctx->tp_xb, *xl = ctx->tp_xl, *attb = ctx->tp_attb, *hb = ctx->tp_hb; // batch-sized twins of the single-token exchange buffers
    static bool attr[64] = {};
    if (!attr[ctx->device]) { PS_CK(cudaFuncSetAttribute(ps_k_softmax_ext, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024)); attr[ctx->device] = true; }
    ps_k_get_embedding<<<(unsigned)bs, 256, 0, ctx->stream>>>(x, ctx->w_embd, ctx->t_embd, dim, ctx->tokens_dev);
    PS_LAUNCH_CK();
    ps_k_tp_slice<<<grid1d(bs * dim_l), 256, 0, ctx->stream>>>(xl, x, bs, dim, dim_l, rank);
    PS_LAUNCH_CK();
    for (int L = 0; L < d.n_layers; L++) {
        const LayerDev &ld = ctx->layers[L];
        ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->xn, x, ld.attn_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        PsRwSeg sg[3] = {{ctx->q, d.qkv_bias ? ld.q_bias : nullptr, 0, (int)qdim_l, 0},
                         {ctx->k, d.qkv_bias ? ld.k_bias : nullptr, (int)qdim_l, (int)(qdim_l + kvd), 0},
                         {ctx->v, d.qkv_bias ? ld.v_bias : nullptr, (int)(qdim_l + kvd), (int)(qdim_l + 2 * kvd), 0}};
        if (tc) {
            if ((rc = tc_prep_b(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = tc_gemm(ctx, ld.tc_qkv, (int)(qdim_l + 2 * kvd), (int)dim, bs, sg, 3, nullptr))) return rc;
        } else {
            if ((rc = rwm_quantize(ctx, ctx->xn, (int)dim, bs))) return rc;
            PsRwmArgs a{};
            a.w = ld.rw_qkv; a.n_oct = (int)((qdim_l + 2 * kvd) / 8); a.K = (int)dim; a.slot = 0; a.n_slots = 1; a.n_seg = 3; a.bs = bs;
            for (int t = 0; t < 3; t++) a.seg[t] = sg[t];
            if ((rc = launch_rwm(ctx, a))) return rc;
        }
        ps_k_rope<<<dim3((unsigned)nh, (unsigned)bs), 64, 0, ctx->stream>>>(ctx->qr, ctx->q, (int)hs, d.rope_n_dims, d.rope_type & 2, ctx->pos_dev, ctx->rope_table);
        PS_LAUNCH_CK();
        ps_k_rope<<<dim3((unsigned)nkv, (unsigned)bs), 64, 0, ctx->stream>>>(ctx->kr, ctx->k, (int)hs, d.rope_n_dims, d.rope_type & 2, ctx->pos_dev, ctx->rope_table);
        PS_LAUNCH_CK();
        ps_k_kv_store<<<grid1d(kvd * bs), 256, 0, ctx->stream>>>(ctx->kc[L], ctx->vct[L], ctx->kr, ctx->v, kvd, d.n_ctx, ctx->pos_dev, bs);
        PS_LAUNCH_CK();
        launch_batch_scores(ctx, L, (int)nh, (int)nkv, (int)hs, n_kv, bs);
        PS_LAUNCH_CK();
        ps_k_softmax_ext<<<(unsigned)(bs * nh), 256, (size_t)n_kv * 4, ctx->stream>>>(ctx->kq, ctx->kq, nullptr, ctx->pos_dev, n_kv, bs, kq_scale);
        PS_LAUNCH_CK();
        if ((rc = launch_batch_pv(ctx, L, (int)nh, (int)nkv, (int)hs, n_kv, bs, 0))) return rc;
        PS_LAUNCH_CK();
        if ((rc = tp_gather_rows(ctx, attb, ctx->att, bs, qdim_l))) return rc;                      // exchange 1: attention output
        if (tc) {
            if ((rc = tc_prep_b(ctx, attb, (int)qdim, bs))) return rc;
            if ((rc = tc_single(ctx, ld.tc_o, (int)dim_l, (int)qdim, xl, bs, xl))) return rc;       // x[rows of this rank] += Wo[rows] . att
        } else {
            if ((rc = rwm_quantize(ctx, attb, (int)qdim, bs))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_o, (int)dim_l, (int)qdim, 0, 1, xl, bs, nullptr, xl))) return rc;
        }
        if ((rc = tp_gather_rows(ctx, x, xl, bs, dim_l))) return rc;                                   // exchange 2: x after Wo
        ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->xn, x, ld.ffn_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if (tc) {
            if ((rc = tc_prep_b(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = tc_single(ctx, ld.tc_gate, (int)ffn_l, (int)dim, ctx->g, bs, nullptr))) return rc;
            if ((rc = tc_single(ctx, ld.tc_up, (int)ffn_l, (int)dim, ctx->u, bs, nullptr))) return rc;
        } else {
            if ((rc = rwm_quantize(ctx, ctx->xn, (int)dim, bs))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_gu, (int)ffn_l, (int)dim, 0, 2, ctx->g, bs, nullptr, nullptr))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_gu, (int)ffn_l, (int)dim, 1, 2, ctx->u, bs, nullptr, nullptr))) return rc;
        }
        ps_k_silu_hadamard<<<grid1d(ffn_l * bs), 256, 0, ctx->stream>>>(ctx->g, ctx->g, ctx->u, ffn_l * bs);
        PS_LAUNCH_CK();
        if ((rc = tp_gather_rows(ctx, hb, ctx->g, bs, ffn_l))) return rc;                             // exchange 3: FFN hidden vector
        if (tc) {
            if ((rc = tc_prep_b(ctx, hb, (int)ffn, bs))) return rc;
            if ((rc = tc_single(ctx, ld.tc_down, (int)dim_l, (int)ffn, xl, bs, xl))) return rc;     // x[rows] += Wdown[rows] . h
        } else {
            if ((rc = rwm_quantize(ctx, hb, (int)ffn, bs))) return rc;
            if ((rc = rwm_single(ctx, ld.rw_down, (int)dim_l, (int)ffn, 0, 1, xl, bs, nullptr, xl))) return rc;
        }
        if ((rc = tp_gather_rows(ctx, x, xl, bs, dim_l))) return rc;                                   // exchange 4: x after Wdown
    }
    if (lm_head) {
        if (!ctx->tp_rows && (rc = dev_alloc(ctx, (void **)&ctx->tp_rows, (size_t)d.max_batch * d.vocab_size * 4))) return rc;
        ps_k_rmsnorm<<<(unsigned)bs, 256, 0, ctx->stream>>>(ctx->xn, x, ctx->w_out_norm, dim, d.norm_eps);
        PS_LAUNCH_CK();
        if ((rc = rwm_quantize(ctx, ctx->xn, (int)dim, bs))) return rc;
        if ((rc = rwm_single(ctx, ctx->rw_out, (int)vocab_l, (int)dim, 0, 1, ctx->tp_logl, bs, nullptr, nullptr))) return rc;
        if ((rc = tp_gather_rows(ctx, ctx->tp_rows, ctx->tp_logl, bs, vocab_l))) return rc;
    }
    return 0;
}


static int check_forward_args(ps_cuda_ctx *ctx, const int32_t *tokens, const int32_t *pos, int bs) {
    if (!ctx->bound) return fail(ctx, PS_CUDA_ERR_INVALID, "forward: no model bound");
    if (bs <= 0 || bs > ctx->d.max_batch) return fail(ctx, PS_CUDA_ERR_INVALID, "forward: batch %d outside [1,%d]", bs, ctx->d.max_batch);
    for (int i = 0; i < bs; i++) {
        if (tokens[i] < 0 || tokens[i] >= ctx->d.vocab_size) return fail(ctx, PS_CUDA_ERR_INVALID, "forward: token %d outside the vocabulary", tokens[i]);
        if (pos[i] != pos[0] + i) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "forward: positions must be consecutive (the CPU path ignores tree masks, SURVEY F7)");
    }
    if (pos[0] < 0 || pos[0] + bs > ctx->d.n_ctx) return fail(ctx, PS_CUDA_ERR_KV_FULL, "the length of kvcache is up to the preset threshold: %d", ctx->d.n_ctx);
    return 0;
}

int ps_cuda_forward(ps_cuda_ctx *ctx, const int32_t *tokens, const int32_t *pos, int bs, int lm_head, float *logits_host) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    int rc = check_forward_args(ctx, tokens, pos, bs);
    if (rc) return rc;
    // lm_head with logits_host == NULL: the logits stay on the device (lazy read-back: ps_cuda_logits_dev, ps_cuda_sample_topk)
    PS_CK(cudaSetDevice(ctx->device));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->h_tokens, tokens, (size_t)bs * 4);
    memcpy(ctx->h_pos, pos, (size_t)bs * 4);
    PS_CK(cudaMemcpyAsync(ctx->tokens_dev, ctx->h_tokens, (size_t)bs * 4, cudaMemcpyHostToDevice, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->pos_dev, ctx->h_pos, (size_t)bs * 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += (int64_t)bs * 8;
    PS_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    ctx->logits_last = ctx->logits;
    if (ctx->tp > 1 && bs > 1 && ctx->nccl_comm && ctx->opt_tp_batch) {
        // tensor parallel batch (prefill chunk / verify batch): row-sharded GEMMs, one all-gather per exchange for the whole chunk
        if (lm_head && !ctx->tp_logl) { int rc2 = dev_alloc(ctx, (void **)&ctx->tp_logl, (size_t)ctx->d.max_batch * ctx->vocab_l * 4); if (rc2) return rc2; }
        rc = forward_ops_tp(ctx, bs, lm_head, pos[0]);
        if (lm_head) ctx->logits_last = ctx->tp_rows;
    } else if (ctx->tp > 1 && bs > 1) {
        // tensor parallel without NCCL (or option tp_batch = 0): the sharded single-token step is fed token by token, i.e.
        // it equals the reference run with batch_size = 1 (NOT a batched pass: the reference's soft-max row length and its
        // SIMD / libm exp split depend on the chunking)
        for (int i = 0; i < bs && !rc; i++) {
            PS_CK(cudaMemcpyAsync(ctx->tokens_dev, ctx->h_tokens + i, 4, cudaMemcpyHostToDevice, ctx->stream));
            PS_CK(cudaMemcpyAsync(ctx->pos_dev, ctx->h_pos + i, 4, cudaMemcpyHostToDevice, ctx->stream));
            rc = step_usable(ctx) ? launch_step(ctx, lm_head ? PS_ST_MODE_LMHEAD : 0, pos[i] + 1) : decode_step_fused(ctx, lm_head != 0, false);
            if (!rc && lm_head) {
                if (!ctx->tp_rows) { int rc2 = dev_alloc(ctx, (void **)&ctx->tp_rows, (size_t)ctx->d.max_batch * ctx->d.vocab_size * 4); if (rc2) return rc2; }
                PS_CK(cudaMemcpyAsync(ctx->tp_rows + (size_t)i * ctx->d.vocab_size, ctx->logits, (size_t)ctx->d.vocab_size * 4, cudaMemcpyDeviceToDevice, ctx->stream));
            }
        }
        if (lm_head) ctx->logits_last = ctx->tp_rows; // the single-token logits slot of the exchange heap holds ONE row: batches are read from tp_rows
    } else if (bs == 1 && lm_head && (ops_graph_usable(ctx) || mv_usable(ctx))) {
        rc = run_step(ctx, false, pos[0] + 1);
    } else if (bs == 1 && ctx->opt_fused && ctx->fused_ok) {
        if (lm_head) rc = run_step(ctx, false, pos[0] + 1);
        else rc = step_usable(ctx) ? launch_step(ctx, 0, pos[0] + 1) : decode_step_fused(ctx, false, false);
    } else {
        rc = forward_ops(ctx, bs, lm_head, pos[0]);
    }
    if (rc) return rc;
    PS_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->logits_rows = lm_head ? bs : 0;
    if (lm_head && logits_host) {
        const size_t bytes = (size_t)bs * ctx->d.vocab_size * 4;
        if (bytes > ctx->h_logits_cap) {
            if (ctx->h_logits) cudaFreeHost(ctx->h_logits);
            ctx->h_logits = nullptr;
            ctx->h_logits_cap = 0;
            PS_CK(cudaMallocHost(&ctx->h_logits, bytes));
            ctx->h_logits_cap = bytes;
        }
        PS_CK(cudaMemcpyAsync(ctx->h_logits, ctx->logits_last, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = sync_and_check(ctx))) return rc;
        memcpy(logits_host, ctx->h_logits, bytes);
        ctx->d2h += (int64_t)bytes;
    } else {
        if ((rc = sync_and_check(ctx))) return rc;
    }
    { float ms = 0.f; PS_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); ctx->last_ms = ms; }
    ctx->position = pos[0] + bs; // m_kv->advance(batch_size), llama_model.cpp:109
    kv_set_mask(ctx, 0, ctx->position, 0);
    return 0;
}

// ---------------------------------------------------------------------------------------------- sessions (server-side batching)
// SURVEY section 8 f4 (app/server/server_handler.hpp:512-720): the reference serves one generation per model at a time - its
// KV position is shared state.  Here a context can hold several independent KV sets ("sessions") over ONE set of weights:
// ps_cuda_session_select makes one of them the target of every single-sequence call (forward / decode_greedy / kv_*: prefill
// and bookkeeping work unchanged), and ps_cuda_forward_sessions advances n sessions by one token each in ONE forward pass -
// the weight stream is read once for the n columns (multi-column row-walker for n < 16, tcgen05 GEMM from 16 up), attention
// runs per column over its own session's cache.  A session's logits are bit-identical to decoding it alone with batch 1.
static int refresh_kv_tables(ps_cuda_ctx *ctx) {
    const ps_cuda_model_desc &d = ctx->d;
    std::vector<float *> ptrs(2 * (size_t)d.n_layers);
    for (int L = 0; L < d.n_layers; L++) { ptrs[L] = ctx->kc[L]; ptrs[d.n_layers + L] = ctx->vct[L]; }
    PS_CK(cudaMemcpyAsync(ctx->kv_ptrs_dev, ptrs.data(), sizeof(float *) * ptrs.size(), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->st_layers && ctx->bound && ctx->fused_ok) {
        std::vector<PsStLayer> tab(d.n_layers);
        for (int L = 0; L < d.n_layers; L++) {
            const LayerDev &ld = ctx->layers[L];
            tab[L] = PsStLayer{ld.rw_qkv, ld.rw_o, ld.rw_gu, ld.rw_down, ld.attn_norm, ld.ffn_norm, d.qkv_bias ? ld.q_bias : nullptr,
                               d.qkv_bias ? ld.k_bias : nullptr, d.qkv_bias ? ld.v_bias : nullptr, ctx->kc[L], ctx->vct[L]};
        }
        PS_CK(cudaMemcpyAsync(ctx->st_layers, tab.data(), sizeof(PsStLayer) * (size_t)d.n_layers, cudaMemcpyHostToDevice, ctx->stream));
    }
    PS_CK(cudaStreamSynchronize(ctx->stream)); // the host vectors die here
    if (ctx->g_step) { cudaGraphExecDestroy(ctx->g_step); ctx->g_step = nullptr; } // the captured steps hold the old cache pointers
    if (ctx->g_fwd) { cudaGraphExecDestroy(ctx->g_fwd); ctx->g_fwd = nullptr; }
    return 0;
}

int ps_cuda_session_create(ps_cuda_ctx *ctx, int *session_id) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!session_id) return fail(ctx, PS_CUDA_ERR_INVALID, "session_create: null argument");
    if (ctx->tp > 1) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "sessions are single-GPU");
    PS_CK(cudaSetDevice(ctx->device));
    const ps_cuda_model_desc &d = ctx->d;
    const size_t bytes = 4 * (size_t)d.n_kv_heads * d.head_size * d.n_ctx;
    ps_cuda_ctx::KvSet ks;
    ks.kc.resize(d.n_layers); ks.vct.resize(d.n_layers);
    for (int L = 0; L < d.n_layers; L++) {
        int rc;
        if ((rc = dev_alloc(ctx, (void **)&ks.kc[L], bytes)) || (rc = dev_alloc(ctx, (void **)&ks.vct[L], bytes))) return rc;
        PS_CK(cudaMemsetAsync(ks.kc[L], 0, bytes, ctx->stream));
        PS_CK(cudaMemsetAsync(ks.vct[L], 0, bytes, ctx->stream));
    }
    ks.slot_mask.assign((size_t)d.n_ctx, 1);
    const int id = ctx->next_session++;
    ctx->sessions.emplace(id, std::move(ks));
    *session_id = id;
    return 0;
}

int ps_cuda_session_destroy(ps_cuda_ctx *ctx, int session_id) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (session_id == ctx->cur_session) return fail(ctx, PS_CUDA_ERR_INVALID, "session_destroy: session %d is selected", session_id);
    auto it = ctx->sessions.find(session_id);
    if (it == ctx->sessions.end()) return fail(ctx, PS_CUDA_ERR_INVALID, "session_destroy: unknown session %d", session_id);
    PS_CK(cudaSetDevice(ctx->device));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    for (auto *vec : {&it->second.kc, &it->second.vct})
        for (float *p : *vec) {
            for (size_t i = 0; i < ctx->owned.size(); i++)
                if (ctx->owned[i] == p) { ctx->owned.erase(ctx->owned.begin() + i); break; }
            PS_CK(cudaFree(p));
        }
    ctx->sessions.erase(it);
    return 0;
}

int ps_cuda_session_current(ps_cuda_ctx *ctx) { return ctx->cur_session; }

int ps_cuda_session_select(ps_cuda_ctx *ctx, int session_id) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (session_id == ctx->cur_session) return 0;
    auto it = ctx->sessions.find(session_id);
    if (it == ctx->sessions.end() && session_id != 0) return fail(ctx, PS_CUDA_ERR_INVALID, "session_select: unknown session %d", session_id);
    PS_CK(cudaSetDevice(ctx->device));
    // park the active state under its id, then load the requested one (session 0 = the context's own cache)
    ps_cuda_ctx::KvSet cur;
    cur.kc = ctx->kc; cur.vct = ctx->vct; cur.position = ctx->position; cur.slot_mask = ctx->slot_mask;
    ps_cuda_ctx::KvSet next = std::move(ctx->sessions.at(session_id));
    ctx->sessions.erase(session_id);
    ctx->sessions.emplace(ctx->cur_session, std::move(cur));
    ctx->kc = next.kc; ctx->vct = next.vct; ctx->position = next.position; ctx->slot_mask = next.slot_mask;
    ctx->slot_mask_dirty = true;
    ctx->last_batch = 0;
    ctx->cur_session = session_id;
    return refresh_kv_tables(ctx);
}

int ps_cuda_session_position(ps_cuda_ctx *ctx, int session_id) {
    if (session_id == ctx->cur_session) return ctx->position;
    auto it = ctx->sessions.find(session_id);
    return it == ctx->sessions.end() ? -1 : it->second.position;
}

int ps_cuda_forward_sessions(ps_cuda_ctx *ctx, const int32_t *session_ids, const int32_t *tokens, int n, int lm_head, float *logits_host, int32_t *greedy_ids) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    const ps_cuda_model_desc &d = ctx->d;
    if (!ctx->bound) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_sessions: no model bound");
    if (ctx->tp > 1) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "forward_sessions: sessions are single-GPU");
    if (!session_ids || !tokens || n <= 0 || n > d.max_batch) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_sessions: batch %d outside [1,%d]", n, d.max_batch);
    if ((size_t)d.n_ctx * 4 > 160 * 1024) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "forward_sessions: n_ctx = %d exceeds the soft-max kernel's shared memory", d.n_ctx);
    PS_CK(cudaSetDevice(ctx->device));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    if (!ctx->sess_ptrs_dev) {
        const size_t bytes = sizeof(float *) * 2 * (size_t)d.n_layers * (size_t)d.max_batch;
        int rc = dev_alloc(ctx, (void **)&ctx->sess_ptrs_dev, bytes);
        if (rc) return rc;
        PS_CK(cudaMallocHost(&ctx->h_sess_ptrs, bytes));
    }
    std::vector<ps_cuda_ctx::KvSet *> sets(n);
    int n_kv_max = 0;
    for (int i = 0; i < n; i++) {
        if (tokens[i] < 0 || tokens[i] >= d.vocab_size) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_sessions: token %d outside the vocabulary", tokens[i]);
        for (int j = 0; j < i; j++)
            if (session_ids[j] == session_ids[i]) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_sessions: session %d appears twice", session_ids[i]);
        int pos;
        const std::vector<float *> *kc, *vct;
        if (session_ids[i] == ctx->cur_session) { pos = ctx->position; kc = &ctx->kc; vct = &ctx->vct; sets[i] = nullptr; }
        else {
            auto it = ctx->sessions.find(session_ids[i]);
            if (it == ctx->sessions.end()) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_sessions: unknown session %d", session_ids[i]);
            pos = it->second.position; kc = &it->second.kc; vct = &it->second.vct; sets[i] = &it->second;
        }
        if (pos >= d.n_ctx) return fail(ctx, PS_CUDA_ERR_KV_FULL, "the length of kvcache is up to the preset threshold: %d", d.n_ctx);
        ctx->h_tokens[i] = tokens[i];
        ctx->h_pos[i] = pos;
        n_kv_max = std::max(n_kv_max, pos + 1);
        for (int L = 0; L < d.n_layers; L++) {
            ctx->h_sess_ptrs[((size_t)L * 2 + 0) * d.max_batch + i] = (*kc)[L];
            ctx->h_sess_ptrs[((size_t)L * 2 + 1) * d.max_batch + i] = (*vct)[L];
        }
    }
    PS_CK(cudaMemcpyAsync(ctx->tokens_dev, ctx->h_tokens, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->pos_dev, ctx->h_pos, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->sess_ptrs_dev, ctx->h_sess_ptrs, sizeof(float *) * 2 * (size_t)d.n_layers * (size_t)d.max_batch, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += (int64_t)n * 8 + (int64_t)sizeof(float *) * 2 * d.n_layers * d.max_batch;
    PS_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    SessRun sr{n_kv_max};
    int rc = forward_ops(ctx, n, lm_head, 0, -1, &sr);
    if (rc) return rc;
    ctx->logits_last = ctx->logits;
    ctx->logits_rows = lm_head ? n : 0;
    if (lm_head && greedy_ids) {
        ps_k_argmax<<<n, 1024, 0, ctx->stream>>>(ctx->logits, d.vocab_size, ctx->ids_dev, ctx->ids_dev + d.max_batch); // one CTA per column
        PS_LAUNCH_CK();
        PS_CK(cudaMemcpyAsync(ctx->h_ids, ctx->ids_dev, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    }
    PS_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    if (lm_head && logits_host) {
        const size_t bytes = (size_t)n * d.vocab_size * 4;
        if (bytes > ctx->h_logits_cap) {
            if (ctx->h_logits) cudaFreeHost(ctx->h_logits);
            ctx->h_logits = nullptr;
            ctx->h_logits_cap = 0;
            PS_CK(cudaMallocHost(&ctx->h_logits, bytes));
            ctx->h_logits_cap = bytes;
        }
        PS_CK(cudaMemcpyAsync(ctx->h_logits, ctx->logits, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = sync_and_check(ctx))) return rc;
        memcpy(logits_host, ctx->h_logits, bytes);
        ctx->d2h += (int64_t)bytes;
    } else if ((rc = sync_and_check(ctx))) return rc;
    if (lm_head && greedy_ids) { memcpy(greedy_ids, ctx->h_ids, (size_t)n * 4); ctx->d2h += (int64_t)n * 4; }
    { float ms = 0.f; PS_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); ctx->last_ms = ms; }
    for (int i = 0; i < n; i++) { // m_kv->advance(1) of every session in the batch
        if (!sets[i]) { kv_set_mask(ctx, ctx->position, ctx->position + 1, 0); ctx->position += 1; }
        else { sets[i]->slot_mask[sets[i]->position] = 0; sets[i]->position += 1; }
    }
    return 0;
}

// LlamaModel::forward as the speculative path uses it (src/speculative/spec_model.hpp:96-103, token_tree.cpp:131): arbitrary
// token positions (ROPE), an in-batch tree mask (row i = the batch tokens token i attends to; NULL = causal, i >= j,
// attention_mask.cpp:36-41), attention over the UNMASKED cache slots below the current position, and the KV semantics of
// CausalLM::Batch::save_kv + advance (src/backend/qnn/causal_models.cpp:353-359): the batch's K / V rows go to cache slots
// position .. position + bs - 1 (kept aside as well for ps_cuda_kv_copy_slot), which are unmasked, and the position advances by bs.
int ps_cuda_forward_tree(ps_cuda_ctx *ctx, const int32_t *tokens, const int32_t *pos, int bs, const uint8_t *tree_mask, int lm_head, float *logits_host) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->bound) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_tree: no model bound");
    if (ctx->tp > 1) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "forward_tree: tree batches are single-GPU");
    const int tb = std::min(ctx->d.max_batch, 32);
    if (bs <= 0 || bs > tb) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_tree: batch %d outside [1,%d]", bs, tb);
    if (lm_head && !logits_host) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_tree: lm_head requested without an output buffer");
    const int base = ctx->position;
    if (base + bs > ctx->d.n_ctx) return fail(ctx, PS_CUDA_ERR_KV_FULL, "the length of kvcache is up to the preset threshold: %d", ctx->d.n_ctx);
    for (int i = 0; i < bs; i++) {
        if (tokens[i] < 0 || tokens[i] >= ctx->d.vocab_size) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_tree: token %d outside the vocabulary", tokens[i]);
        if (pos[i] < 0 || pos[i] >= ctx->d.n_ctx) return fail(ctx, PS_CUDA_ERR_INVALID, "forward_tree: position %d outside [0,%d)", pos[i], ctx->d.n_ctx);
    }
    const size_t row = (size_t)(base + bs) * 4;
    if (row > 160 * 1024) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "forward_tree: soft-max row of %d exceeds shared memory", base + bs);
    PS_CK(cudaSetDevice(ctx->device));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    memcpy(ctx->h_tokens, tokens, (size_t)bs * 4);
    memcpy(ctx->h_pos, pos, (size_t)bs * 4);
    for (int i = 0; i < bs; i++)
        for (int j = 0; j < bs; j++) ctx->h_tree[i * bs + j] = tree_mask ? (tree_mask[i * bs + j] != 0) : (i >= j);
    PS_CK(cudaMemcpyAsync(ctx->tokens_dev, ctx->h_tokens, (size_t)bs * 4, cudaMemcpyHostToDevice, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->pos_dev, ctx->h_pos, (size_t)bs * 4, cudaMemcpyHostToDevice, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->tree_dev, ctx->h_tree, (size_t)bs * bs, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += (int64_t)bs * 8 + (int64_t)bs * bs;
    if (ctx->slot_mask_dirty) { // pageable source, small: the synchronous copy keeps the host vector free to change afterwards
        PS_CK(cudaMemcpy(ctx->slot_mask_dev, ctx->slot_mask.data(), (size_t)ctx->d.n_ctx, cudaMemcpyHostToDevice));
        ctx->h2d += ctx->d.n_ctx;
        ctx->slot_mask_dirty = false;
    }
    PS_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    ps_k_tree_mask<<<dim3((unsigned)std::min<int64_t>((base + bs + 255) / 256, 64), (unsigned)bs), 256, 0, ctx->stream>>>(ctx->tree_bias, ctx->slot_mask_dev, ctx->tree_dev, base, bs,
                                                                                                                        base + bs);
    PS_LAUNCH_CK();
    int rc = forward_ops(ctx, bs, lm_head, 0, base);
    if (rc) return rc;
    PS_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    ctx->logits_last = ctx->logits;
    if (lm_head == 2) { // greedy ids only (ProbArray + greedy_sample with top_k = 1 per row, first maximum): bs ints come back instead of bs x vocab floats
        ps_k_argmax<<<bs, 1024, 0, ctx->stream>>>(ctx->logits, ctx->d.vocab_size, ctx->ids_dev, ctx->ids_dev + 2048); // one CTA per tree node
        PS_LAUNCH_CK();
        PS_CK(cudaMemcpyAsync(ctx->h_ids, ctx->ids_dev, (size_t)bs * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = sync_and_check(ctx))) return rc;
        memcpy(logits_host, ctx->h_ids, (size_t)bs * 4);
        ctx->d2h += (int64_t)bs * 4;
    } else if (lm_head) {
        const size_t bytes = (size_t)bs * ctx->d.vocab_size * 4;
        if (bytes > ctx->h_logits_cap) {
            if (ctx->h_logits) cudaFreeHost(ctx->h_logits);
            ctx->h_logits = nullptr;
            ctx->h_logits_cap = 0;
            PS_CK(cudaMallocHost(&ctx->h_logits, bytes));
            ctx->h_logits_cap = bytes;
        }
        PS_CK(cudaMemcpyAsync(ctx->h_logits, ctx->logits, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        if ((rc = sync_and_check(ctx))) return rc;
        memcpy(logits_host, ctx->h_logits, bytes);
        ctx->d2h += (int64_t)bytes;
    } else if ((rc = sync_and_check(ctx))) return rc;
    { float ms = 0.f; PS_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); ctx->last_ms = ms; }
    ctx->last_batch = bs;
    kv_set_mask(ctx, base, base + bs, 0);   // advance_tokens(bs)
    ctx->position = base + bs;
    return 0;
}

int ps_cuda_decode_greedy(ps_cuda_ctx *ctx, int32_t first_token, int n_steps, int32_t *ids_host) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (!ctx->bound) return fail(ctx, PS_CUDA_ERR_INVALID, "decode_greedy: no model bound");
    if (n_steps <= 0 || n_steps > 4096) return fail(ctx, PS_CUDA_ERR_INVALID, "decode_greedy: n_steps %d outside [1,4096]", n_steps);
    if (first_token < 0 || first_token >= ctx->d.vocab_size) return fail(ctx, PS_CUDA_ERR_INVALID, "decode_greedy: bad token");
    if (ctx->position + n_steps > ctx->d.n_ctx) return fail(ctx, PS_CUDA_ERR_KV_FULL, "the length of kvcache is up to the preset threshold: %d", ctx->d.n_ctx);
    PS_CK(cudaSetDevice(ctx->device));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    ctx->h_tokens[0] = first_token;
    PS_CK(cudaMemcpyAsync(ctx->tokens_dev, ctx->h_tokens, 4, cudaMemcpyHostToDevice, ctx->stream));
    ctx->h2d += 4;
    if ((ctx->opt_fused && ctx->fused_ok) || ops_graph_usable(ctx) || mv_usable(ctx)) {
        int32_t *slot = &ctx->h_ids[4096];
        slot[0] = ctx->position;
        slot[1] = 0;
        PS_CK(cudaMemcpyAsync(ctx->pos_dev, &slot[0], 4, cudaMemcpyHostToDevice, ctx->stream));
        PS_CK(cudaMemcpyAsync(ctx->ctr_dev, &slot[1], 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += 8;
        ctx->kt_used = 0;
        PS_CK(cudaEventRecord(ctx->ev0, ctx->stream));
        for (int s = 0; s < n_steps; s++) {
            int rc = run_step(ctx, true, ctx->position + n_steps);
            if (rc) return rc;
        }
        PS_CK(cudaEventRecord(ctx->ev1, ctx->stream));
        PS_CK(cudaMemcpyAsync(ctx->h_ids, ctx->ids_dev, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, ctx->stream));
        { int rc = sync_and_check(ctx); if (rc) return rc; }
        if (ctx->opt_ktime) {
            ctx->kt_ms = 0.0;
            for (size_t i = 0; i + 1 < ctx->kt_used; i += 2) {
                float ms = 0.f;
                PS_CK(cudaEventElapsedTime(&ms, ctx->kt_events[i], ctx->kt_events[i + 1]));
                ctx->kt_ms += ms;
            }
            ctx->kt_launches = (int64_t)(ctx->kt_used / 2);
        }
        { float ms = 0.f; PS_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); ctx->last_ms = ms; }
        memcpy(ids_host, ctx->h_ids, (size_t)n_steps * 4);
        ctx->d2h += (int64_t)n_steps * 4;
        kv_set_mask(ctx, 0, ctx->position + n_steps, 0);
        ctx->position += n_steps;
        return 0;
    }
    PS_CK(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int s = 0; s < n_steps; s++) {
        const int pos = ctx->position + s;
        int32_t *slot = &ctx->h_ids[4096 + s]; // one pinned slot per step: the async copies never race with the host writes
        *slot = pos;
        PS_CK(cudaMemcpyAsync(ctx->pos_dev, slot, 4, cudaMemcpyHostToDevice, ctx->stream));
        ctx->h2d += 4;
        int rc = forward_ops(ctx, 1, 1, pos);
        if (rc) return rc;
        ps_k_argmax<<<1, 1024, 0, ctx->stream>>>(ctx->logits, ctx->d.vocab_size, ctx->ids_dev + s, ctx->tokens_dev);
        PS_LAUNCH_CK();
    }
    PS_CK(cudaEventRecord(ctx->ev1, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->h_ids, ctx->ids_dev, (size_t)n_steps * 4, cudaMemcpyDeviceToHost, ctx->stream));
    { int rc = sync_and_check(ctx); if (rc) return rc; }
    { float ms = 0.f; PS_CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1)); ctx->last_ms = ms; }
    memcpy(ids_host, ctx->h_ids, (size_t)n_steps * 4);
    ctx->d2h += (int64_t)n_steps * 4;
    kv_set_mask(ctx, 0, ctx->position + n_steps, 0);
    ctx->position += n_steps;
    return 0;
}

// TopKSampler on the device (sampler.cpp:39-56): the k largest logits of row `row` of the last forward pass, descending (equal
// logits by ascending token id), as (logit, token) pairs - what ProbArray holds after TopKSampler::apply.
int ps_cuda_sample_topk(ps_cuda_ctx *ctx, int row, int k, float *logits_out, int32_t *tokens_out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    const int vocab = ctx->d.vocab_size;
    if (k <= 0 || k > PS_TOPK_MAX || k > vocab) return fail(ctx, PS_CUDA_ERR_INVALID, "sample_topk: k = %d outside [1,%d]", k, std::min(PS_TOPK_MAX, vocab));
    if (!ctx->logits_last || row < 0 || row >= std::max(ctx->logits_rows, 1)) return fail(ctx, PS_CUDA_ERR_INVALID, "sample_topk: no logits row %d", row);
    if (!logits_out || !tokens_out) return fail(ctx, PS_CUDA_ERR_INVALID, "sample_topk: null output");
    PS_CK(cudaSetDevice(ctx->device));
    const int n_slices = (vocab + PS_TOPK_SLICE - 1) / PS_TOPK_SLICE;
    if (!ctx->topk_val) {
        const size_t n = (size_t)((ctx->d.vocab_size + PS_TOPK_SLICE - 1) / PS_TOPK_SLICE + 1) * PS_TOPK_MAX;
        int rc;
        if ((rc = dev_alloc(ctx, (void **)&ctx->topk_val, n * 4)) || (rc = dev_alloc(ctx, (void **)&ctx->topk_idx, n * 4))) return rc;
    }
    float *res_val = ctx->topk_val + (size_t)n_slices * PS_TOPK_MAX;
    int *res_idx = ctx->topk_idx + (size_t)n_slices * PS_TOPK_MAX;
    ps_k_topk_stage1<<<n_slices, 256, 0, ctx->stream>>>(ctx->logits_last + (size_t)row * vocab, vocab, k, ctx->topk_val, ctx->topk_idx);
    PS_LAUNCH_CK();
    ps_k_topk_stage2<<<1, 256, 0, ctx->stream>>>(ctx->topk_val, ctx->topk_idx, n_slices * k, k, res_val, res_idx);
    PS_LAUNCH_CK();
    PS_CK(cudaMemcpyAsync(ctx->h_ids, res_idx, (size_t)k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    PS_CK(cudaMemcpyAsync(ctx->h_ids + PS_TOPK_MAX, res_val, (size_t)k * 4, cudaMemcpyDeviceToHost, ctx->stream));
    int rc = sync_and_check(ctx);
    if (rc) return rc;
    memcpy(tokens_out, ctx->h_ids, (size_t)k * 4);
    memcpy(logits_out, ctx->h_ids + PS_TOPK_MAX, (size_t)k * 4);
    ctx->d2h += (int64_t)k * 8;
    return 0;
}

const float *ps_cuda_logits_dev(ps_cuda_ctx *ctx) { return ctx->logits_last ? ctx->logits_last : ctx->logits; }

int ps_cuda_tp_unique_id(void *out128) {
    if (!out128 || !nccl_load()) return PS_CUDA_ERR_UNSUPPORTED;
    return g_nccl.GetUniqueId(out128) == 0 ? 0 : PS_CUDA_ERR_CUDA;
}

int ps_cuda_tp_export(ps_cuda_ctx *ctx, void *handle64) {
    if (ctx->tp <= 1 || !ctx->heap) return fail(ctx, PS_CUDA_ERR_INVALID, "tp_export: not a tensor-parallel context");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    PS_CK(cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    PS_CK(cudaIpcGetMemHandle(&h, ctx->heap));
    memcpy(handle64, &h, 64);
    return 0;
}

int ps_cuda_tp_import(ps_cuda_ctx *ctx, const void *handles, int n) {
    if (ctx->tp <= 1 || n != ctx->tp) return fail(ctx, PS_CUDA_ERR_INVALID, "tp_import: expected %d handles", ctx->tp);
    PS_CK(cudaSetDevice(ctx->device));
    PS_CK(cudaStreamSynchronize(ctx->stream));
    for (int p = 0; p < n; p++) {
        if (p == ctx->rank) { ctx->peer_heap[p] = ctx->heap; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const uint8_t *)handles + (size_t)p * 64, 64);
        void *ptr = nullptr;
        PS_CK(cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
        ctx->peer_heap[p] = (uint8_t *)ptr;
    }
    {   // link tables: per slot, where this rank's shard lands on every rank, the flag it publishes, and what it waits for
        const ps_cuda_model_desc &d = ctx->d;
        const int n_part = std::min(ctx->n_sm, (ctx->vocab_l + 7) / 8);
        const size_t off[PS_TP_SLOTS] = {ctx->off_att, ctx->off_x, ctx->off_h, ctx->off_x, ctx->off_val, ctx->off_logits};
        const size_t mine[PS_TP_SLOTS] = {(size_t)ctx->rank * ctx->nh_l * d.head_size, (size_t)ctx->rank * ctx->dim_l, (size_t)ctx->rank * ctx->ffn_l,
                                          (size_t)ctx->rank * ctx->dim_l, (size_t)ctx->rank * n_part, (size_t)ctx->rank * ctx->vocab_l};
        PsTpOut to[2 * PS_TP_SLOTS]; // [0..): fence + epoch flags; [PS_TP_SLOTS..): the same links with the in-band-flag mirrors
        PsTpIn ti[PS_TP_SLOTS];
        memset(to, 0, sizeof to);
        memset(ti, 0, sizeof ti);
        for (int s = 0; s < PS_TP_SLOTS; s++) {
            to[s].n = ti[s].n = ctx->tp;
            for (int p = 0; p < ctx->tp; p++) {
                to[s].peer_dst[p] = reinterpret_cast<float *>(ctx->peer_heap[p] + off[s]) + mine[s];
                to[s].peer_idx[p] = reinterpret_cast<int *>(ctx->peer_heap[p] + ctx->off_idx) + mine[s];
                to[s].peer_flag[p] = reinterpret_cast<uint32_t *>(ctx->peer_heap[p] + ctx->off_flags) + s * PS_TP_MAX + ctx->rank;
            }
            to[s].epoch = ctx->epoch_dev + s;
            to[s].done = ctx->done_dev + s;
            to[PS_TP_SLOTS + s] = to[s];
            if (s <= PS_TP_SLOT_X2)
                for (int p = 0; p < ctx->tp; p++)
                    to[PS_TP_SLOTS + s].peer_ll[p] = reinterpret_cast<unsigned long long *>(ctx->peer_heap[p] + ctx->off_ll[s]) + mine[s];
            ti[s].flags = reinterpret_cast<const uint32_t *>(ctx->heap + ctx->off_flags) + s * PS_TP_MAX;
            ti[s].epoch = ctx->epoch_dev + s;
            ti[s].err = ctx->tp_err_dev;
        }
        if (!ctx->tpo_dev) {
            int rc = dev_alloc(ctx, (void **)&ctx->tpo_dev, sizeof to);
            if (rc) return rc;
            if ((rc = dev_alloc(ctx, (void **)&ctx->tpi_dev, sizeof ti))) return rc;
        }
        PS_CK(cudaMemcpy(ctx->tpo_dev, to, sizeof to, cudaMemcpyHostToDevice));
        PS_CK(cudaMemcpy(ctx->tpi_dev, ti, sizeof ti, cudaMemcpyHostToDevice));
        PsStPeers pe; // the step kernel's exchanged vectors on every rank
        memset(&pe, 0, sizeof pe);
        for (int p = 0; p < ctx->tp; p++) {
            pe.x[p] = reinterpret_cast<unsigned long long *>(ctx->peer_heap[p] + ctx->st_off[0]);
            pe.x1[p] = reinterpret_cast<unsigned long long *>(ctx->peer_heap[p] + ctx->st_off[1]);
            pe.att[p] = reinterpret_cast<unsigned long long *>(ctx->peer_heap[p] + ctx->st_off[2]);
            pe.hq[p] = reinterpret_cast<unsigned long long *>(ctx->peer_heap[p] + ctx->st_off[3]);
            pe.best[p] = reinterpret_cast<unsigned long long *>(ctx->peer_heap[p] + ctx->st_off[4]);
        }
        PS_CK(cudaMemcpy(ctx->st_peers, &pe, sizeof pe, cudaMemcpyHostToDevice));
    }
    ctx->p2p = true;
    if (ctx->g_step) { cudaGraphExecDestroy(ctx->g_step); ctx->g_step = nullptr; }
    if (ctx->g_fwd) { cudaGraphExecDestroy(ctx->g_fwd); ctx->g_fwd = nullptr; }
    return 0;
}

int ps_cuda_tp_init(ps_cuda_ctx *ctx, const void *id128) {
    if (ctx->tp <= 1) return 0;
    if (!nccl_load()) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "tensor parallelism needs libnccl.so.2 (dlopen failed)");
    if (ctx->nccl_comm) return fail(ctx, PS_CUDA_ERR_INVALID, "tp_init: already initialised");
    PS_CK(cudaSetDevice(ctx->device));
    PsNcclId id;
    memcpy(&id, id128, sizeof id);
    const int rc = g_nccl.CommInitRank(&ctx->nccl_comm, ctx->tp, id, ctx->rank);
    if (rc != 0) { ctx->nccl_comm = nullptr; return fail(ctx, PS_CUDA_ERR_CUDA, "ncclCommInitRank failed: %s", g_nccl.GetErrorString(rc)); }
    return 0;
}

int ps_cuda_set_option(ps_cuda_ctx *ctx, const char *name, int value) {
    if (ctx->g_step) { cudaGraphExecDestroy(ctx->g_step); ctx->g_step = nullptr; }
    if (ctx->g_fwd) { cudaGraphExecDestroy(ctx->g_fwd); ctx->g_fwd = nullptr; }
    if (!strcmp(name, "graph")) ctx->opt_graph = value;
    else if (!strcmp(name, "fused")) {
        if (!value && ctx->tp > 1) return fail(ctx, PS_CUDA_ERR_UNSUPPORTED, "option fused = 0: the table-op path is not sharded; tensor-parallel contexts run the fused path only");
        ctx->opt_fused = value;
    }
    else if (!strcmp(name, "ops_graph")) ctx->opt_ops_graph = value; // 0: models off the Q4_K fused path decode through one launch per table op from the host (no graph replay, host-fed positions)
    else if (!strcmp(name, "pdl")) ctx->opt_pdl = value;
    else if (!strcmp(name, "rw_ksplit")) ctx->opt_ksplit = value;   // opt-in (default 0): most warps that may share a row octet in the mat-vec launches with few octets per CTA; bit-exact but measured 3-4 % slower per step (profiles/r02_ab_matvec_ksplit_8b_ctx2048.txt)
    else if (!strcmp(name, "rw_defer")) ctx->opt_defer = value;     // bit k: launch kind k (1 Wdown, 2 gate|up, 3 q|k|v, 4 Wo, 5 lm_head) requests its weight stream after its activation vector
    else if (!strcmp(name, "mv_kpar")) ctx->opt_mv_kpar = value;         // 1 (default): 32-block mat-vec launches with one octet per CTA and long rows share the row's blocks among the CTA's warps
    else if (!strcmp(name, "rwm_tile")) ctx->opt_rwm_tile = value;       // 1 (default): the multi-column row-walker shrinks its row tile so that narrow batches use every SM; 0: 16-octet tiles
    else if (!strcmp(name, "scores_batch_min")) ctx->opt_scores_batch_min = std::max(1, value); // batches at least this wide use the query-blocked scores kernel
    else if (!strcmp(name, "tc_min")) ctx->opt_tc_min = std::max(2, value); // narrowest batch that takes the tcgen05 GEMM (narrower ones: multi-column row-walker)
    else if (!strcmp(name, "attn_tile")) ctx->opt_attn_tile = value; // register-tiled scores / P.V kernels for batches (0: the round-1 kernels, for A/B)
    else if (!strcmp(name, "pv_batch_min")) ctx->opt_pv_batch_min = value; // batches at least this wide use the query-blocked P.V kernel (prefill), narrower ones the per-(head, dim) warp kernel
    else if (!strcmp(name, "rw_unroll2")) ctx->opt_unroll2 = value; // tuning: two blocks per loop trip in the mat-vec launches with <= 8 octets per CTA
    else if (!strcmp(name, "attn_group")) ctx->opt_attn_group = value; // 1: decode attention as one group-synchronised kernel per layer (bit-exact, slower so far: DESIGN.md 5b); 0 (default): scores kernel + soft-max / P.V kernel
    else if (!strcmp(name, "attn_fused")) ctx->opt_attn_fused = value; // 1: decode attention as ONE cluster kernel per layer (bit-exact, but slower so far: DESIGN.md); 0 (default): scores kernel + soft-max / P.V kernel
    else if (!strcmp(name, "l2_ahead")) ctx->opt_l2_ahead = value;     // tuning: L2 look-ahead of the step kernel's weight stream, in 4736-byte stages per CTA
    else if (!strcmp(name, "attn_chunk")) ctx->opt_attn_chunk = value; // testing: soft-max positions resident in shared memory (multiple of 256)
    else if (!strcmp(name, "persist")) ctx->opt_persist = value; // 1: the whole decode step as ONE persistent kernel (ps_step.cuh); 0 (default): one kernel per phase
    else if (!strcmp(name, "ktime")) ctx->opt_ktime = value;
    else if (!strcmp(name, "tc")) ctx->opt_tc = value;
    else if (!strcmp(name, "cta_trace")) ctx->opt_cta_trace = value;
    else if (!strcmp(name, "rw_kb")) ctx->opt_kb = value; // tuning: cap on the blocks per TMA stage of the row-walker mat-vec
    else if (!strcmp(name, "tp_ll")) ctx->opt_ll = value;
    else if (!strcmp(name, "tp_batch")) ctx->opt_tp_batch = value;
    else if (!strcmp(name, "tp_p2p")) ctx->p2p = value && ctx->peer_heap[ctx->tp > 1 ? (ctx->rank + 1) % ctx->tp : 0] != nullptr;
    else if (!strcmp(name, "trace")) {
        if (value && !ctx->trace_dev) {
            int rc = dev_alloc(ctx, (void **)&ctx->trace_dev, sizeof(long long) * PS_TL_SLOTS * 8);
            if (rc) return rc;
        }
        if (!value && ctx->trace_dev) { ps_cuda_free(ctx, ctx->trace_dev); ctx->trace_dev = nullptr; }
        if (ctx->trace_dev) {
            std::vector<long long> init((size_t)PS_TL_SLOTS * 8, 0);
            for (size_t i = 0; i < init.size(); i += 8) init[i] = init[i + 2] = 0x7fffffffffffffffLL; // [0],[2] take minima
            PS_CK(cudaStreamSynchronize(ctx->stream));
            PS_CK(cudaMemcpy(ctx->trace_dev, init.data(), init.size() * sizeof(long long), cudaMemcpyHostToDevice));
        }
        ctx->trace_launch = 0;
    }
    else return fail(ctx, PS_CUDA_ERR_INVALID, "unknown option %s", name);
    return 0;
}

int64_t ps_cuda_get_counter(ps_cuda_ctx *ctx, const char *name) {
    if (!strcmp(name, "kernel_launches")) return ctx->n_launch;
    if (!strcmp(name, "graph_replays")) return ctx->n_graph;
    if (!strcmp(name, "step_launches")) return ctx->n_step;          // persistent step-kernel launches
    if (!strcmp(name, "step_ok")) return ctx->step_ok ? 1 : 0;
    if (!strcmp(name, "mv32_ok")) return ctx->mv_ok ? 1 : 0;           // all-Q4_0 / all-Q8_0 model: the fused 32-block decode path is bound
    if (!strcmp(name, "fused_ok")) return ctx->fused_ok ? 1 : 0;
    if (!strcmp(name, "attn_clusters")) return ctx->attn_clusters;
    if (!strcmp(name, "step_kernel_ns")) return (int64_t)(ctx->kt_ms * 1e6);   // option "ktime": summed CUDA-event time of the step-kernel launches
    if (!strcmp(name, "step_kernel_launches")) return ctx->kt_launches;
    if (!strncmp(name, "step_dbg", 8)) { // PS_ST_DEBUG builds: word k of the step kernel's debug record
        int v[32] = {};
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(v, ctx->err_dev, 128, cudaMemcpyDeviceToHost);
        return v[3 + atoi(name + 8)];
    }
    if (!strcmp(name, "step_error")) {
        int v = 0;
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(&v, ctx->err_dev + 2, 4, cudaMemcpyDeviceToHost);
        return v | ctx->err_seen[2];
    }
    if (!strcmp(name, "h2d_bytes")) return ctx->h2d;
    if (!strcmp(name, "d2h_bytes")) return ctx->d2h;
    if (!strcmp(name, "last_device_ns")) return (int64_t)(ctx->last_ms * 1e6);
    if (!strcmp(name, "matvec_kernel_ns")) return (int64_t)(ctx->kt_ms * 1e6);   // option "ktime": summed CUDA-event time of the mat-vec launches
    if (!strcmp(name, "matvec_kernel_launches")) return ctx->kt_launches;
    if (!strcmp(name, "tp_allgathers")) return ctx->n_gather;
    if (!strcmp(name, "tc_gemm_launches")) return ctx->n_tc;
    if (!strcmp(name, "tc_ok")) return ctx->tc_ok ? 1 : 0;
    if (!strcmp(name, "tc_error")) {
        int v = 0;
        if (ctx->tc_err_dev) { cudaStreamSynchronize(ctx->stream); cudaMemcpy(&v, ctx->tc_err_dev, 4, cudaMemcpyDeviceToHost); }
        return v | ctx->err_seen[0];
    }
    if (!strcmp(name, "tp_p2p")) return ctx->p2p ? 1 : 0;
    if (!strcmp(name, "tp_error")) { // 1 if a peer wait gave up (bounded spin)
        int v = 0;
        if (ctx->tp_err_dev) { cudaStreamSynchronize(ctx->stream); cudaMemcpy(&v, ctx->tp_err_dev, 4, cudaMemcpyDeviceToHost); }
        return v | ctx->err_seen[1];
    } // CUDA-event time of the last forward / decode
    return -1;
}

int ps_cuda_read_trace(ps_cuda_ctx *ctx, long long *host, int n_launches) {
    if (!ctx->trace_dev) return fail(ctx, PS_CUDA_ERR_INVALID, "trace is off");
    if (n_launches < 0 || n_launches > PS_TL_SLOTS) return fail(ctx, PS_CUDA_ERR_INVALID, "read_trace: at most %d slots", PS_TL_SLOTS);
    PS_CK(cudaStreamSynchronize(ctx->stream));
    PS_CK(cudaMemcpy(host, ctx->trace_dev, sizeof(long long) * (size_t)n_launches * 8, cudaMemcpyDeviceToHost));
    return 0;
}

float ps_cuda_host_expf_ref(float x) { return ps_expf_glibc(x); }
float ps_cuda_host_v_expf(float x) { return ps_v_expf(x); }

} // extern "C"
