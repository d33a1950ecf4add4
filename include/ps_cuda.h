/* include/ps_cuda.h — C ABI of the B200-native CUDA backend for PowerServe's decode / prefill hot path.
 *
 * This is the drop-in boundary: one shared library (libps_cuda.so), plain pointers and sizes only, no C++ or torch
 * types.  A PowerServe maintainer binds it from a `CUDABackend` class that mirrors `powerserve::ggml::GGMLBackend`
 * (/root/reference/src/backend/ggml/ggml.hpp:186-250) — see INTEGRATION.md for the stub — exactly the way the QNN
 * backend is bound today (src/backend/platform.hpp:31-33, src/executor/executor.cpp:142-165).
 *
 * Conventions
 *   - ggml dimension order everywhere: ne0 is the contiguous dim; a weight W is {K, N} = N rows of K quantised
 *     elements; activations are {dim, bs} fp32 (src/core/tensor.hpp:27-31, src/graph/graph.cpp:64-76).
 *   - every call enqueues on the context's stream and returns an int status: 0 = ok, otherwise an error whose text
 *     `ps_cuda_last_error` returns.  The C++ side converts non-zero into POWERSERVE_ABORT to keep the reference's
 *     error convention (src/core/logger.hpp:56-82).
 *   - results are BIT-IDENTICAL to the reference's x86 AVX2+FMA build (see DESIGN.md "Numerics contract").
 *   - there is NO CPU fallback: without a CUDA device every entry point fails (PS_CUDA_ERR_NO_DEVICE).
 */
#ifndef PS_CUDA_H
#define PS_CUDA_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PS_CUDA_ABI_VERSION 1

/* ggml type ids as stored in GGUF (libs/ggml/include/ggml.h:386-401) */
enum ps_cuda_type { PS_TYPE_F32 = 0, PS_TYPE_Q4_0 = 2, PS_TYPE_Q8_0 = 8, PS_TYPE_Q4_K = 12, PS_TYPE_Q6_K = 14 };

enum ps_cuda_status {
    PS_CUDA_OK = 0,
    PS_CUDA_ERR_NO_DEVICE = 1,
    PS_CUDA_ERR_CUDA = 2,
    PS_CUDA_ERR_INVALID = 3,
    PS_CUDA_ERR_UNSUPPORTED = 4,
    PS_CUDA_ERR_OOM = 5,
    PS_CUDA_ERR_KV_FULL = 6
};

typedef struct ps_cuda_ctx ps_cuda_ctx;

/* Mirrors ModelConfig::LLMConfig + RopeConfig (src/core/config.hpp:86-109) — the fields the hot path reads. */
typedef struct {
    int32_t dim, ffn_dim, n_layers, n_heads, n_kv_heads, head_size, vocab_size;
    int32_t n_ctx;          /* KV slots to allocate (cap of model.json's n_ctx) */
    float   norm_eps;
    int32_t rope_n_dims;    /* == head_size (asserted by NormAttention::build, norm_attention.cpp:38) */
    int32_t rope_type;      /* mode & 2 -> NEOX pairs, else adjacent pairs (ggml.c:15440) */
    float   rope_freq_base, rope_freq_scale, rope_attn_factor;
    int32_t qkv_bias;       /* Qwen2 (src/model/qwen2/qwen2_model.cpp:89) */
    int32_t max_batch;      /* largest bs a forward() will see (prefill chunk / verify width); workspace is sized once */
    int32_t tp_rank, tp_size; /* tensor-parallel position of this context (1 = single GPU) */
} ps_cuda_model_desc;

typedef struct { const void *host; int32_t type; int32_t _pad; } ps_cuda_tensor; /* host pointer into GGUF memory */

/* Same names as LayerWeights / Weight (src/model/common/weights.hpp:24-74). */
typedef struct {
    ps_cuda_tensor attn_norm, ffn_norm, attn_q, attn_k, attn_v, attn_output, ffn_gate, ffn_up, ffn_down;
    ps_cuda_tensor attn_q_bias, attn_k_bias, attn_v_bias;
} ps_cuda_layer_weights;

typedef struct {
    ps_cuda_tensor token_embd, output_norm, output; /* output.host == token_embd.host when tied (weights.hpp:67) */
    const ps_cuda_layer_weights *layers;            /* n_layers entries */
} ps_cuda_model_weights;

/* ---------------------------------------------------------------------------------------------- lifecycle
 * replaces Platform::init_ggml_backend / destroy_ggml_backend (src/backend/platform.cpp:19-25) */
int  ps_cuda_abi_version(void);
int  ps_cuda_device_count(void);
int  ps_cuda_create(ps_cuda_ctx **out, int device, const ps_cuda_model_desc *desc);
void ps_cuda_destroy(ps_cuda_ctx *ctx);
const char *ps_cuda_last_error(const ps_cuda_ctx *ctx); /* ctx may be NULL: error of a failed create */
int  ps_cuda_sync(ps_cuda_ctx *ctx);
void *ps_cuda_stream(ps_cuda_ctx *ctx);                 /* the cudaStream_t every call enqueues on */

/* ---------------------------------------------------------------------------------------------- memory
 * CUDABuffer backing for graph intermediates (replaces CPUBuffer::create_buffer, src/backend/cpu_buffer.hpp:41-51)
 * and the weight registry: graph leaves arrive as CPUBuffer views of GGUF memory (weights.hpp:45-52), so weights are
 * uploaded once and looked up by host pointer. */
int   ps_cuda_malloc(ps_cuda_ctx *ctx, size_t bytes, void **dev);
int   ps_cuda_free(ps_cuda_ctx *ctx, void *dev);
int   ps_cuda_memcpy_h2d(ps_cuda_ctx *ctx, void *dev, const void *host, size_t bytes);
int   ps_cuda_memcpy_d2h(ps_cuda_ctx *ctx, void *host, const void *dev, size_t bytes);  /* synchronises */
int   ps_cuda_register_weight(ps_cuda_ctx *ctx, const void *host, int type, int64_t ne0, int64_t ne1, void **dev);
void *ps_cuda_lookup_weight(ps_cuda_ctx *ctx, const void *host);
int   ps_cuda_unregister_weight(ps_cuda_ctx *ctx, const void *host); /* frees the device copy (model unload) */

/* ---------------------------------------------------------------------------------------------- operator table
 * One entry per GGMLBackend method on the hot path (ggml.hpp:216-244); all pointers are DEVICE pointers except
 * `tokens` / `pos`, which are host arrays like the reference's std::vector<int> arguments. */
int ps_cuda_get_embedding(ps_cuda_ctx *ctx, float *dst, const void *w, int wtype, int64_t dim, const int32_t *tokens, int64_t bs);
int ps_cuda_rmsnorm(ps_cuda_ctx *ctx, float *dst, const float *x, const float *w, int64_t dim, int64_t bs, float eps);
int ps_cuda_matmul(ps_cuda_ctx *ctx, float *dst, const void *w, int wtype, int64_t K, int64_t N, const float *x, int64_t bs);
int ps_cuda_rope(ps_cuda_ctx *ctx, float *dst, const float *src, int64_t head_size, int64_t n_heads, int64_t bs, const int32_t *pos);
/* RoPE frequency factors (rope_freqs.weight of Llama-3.1 / 3.2 GGUFs; ggml_rope_cache_init, ggml.c:15342-15356: theta / ff).
 * The reference never passes them (ggml_wrapper.cpp:104-106), so the default - and parity - is WITHOUT;
 * factors == NULL restores that.  n must be rope_n_dims / 2.  Rebuilds the context's cos / sin table. */
int ps_cuda_set_rope_freq_factors(ps_cuda_ctx *ctx, const float *factors, int n);
int ps_cuda_add(ps_cuda_ctx *ctx, float *dst, const float *a, const float *b, int64_t n, int64_t nb); /* b row-broadcast */
int ps_cuda_silu_hadamard(ps_cuda_ctx *ctx, float *dst, const float *gate, const float *up, int64_t n);
int ps_cuda_get_mask(ps_cuda_ctx *ctx, float *mask, int64_t n_kv, int64_t bs, const int32_t *pos);
int ps_cuda_softmax_ext(ps_cuda_ctx *ctx, float *dst, const float *x, const float *mask, int64_t ne0, int64_t ne1, int64_t ne2, float scale);
/* the two fp32 attention matmuls over the reference's cache layout: K {kv_dim, n_kv} rows, V transposed
 * {n_ctx, kv_dim} (norm_attention.cpp:82-147); q is the rope output {head_size, n_heads, bs}. */
int ps_cuda_attn_scores(ps_cuda_ctx *ctx, float *kq, const float *k_cache, const float *q, int64_t head_size, int64_t n_heads,
                        int64_t n_kv_heads, int64_t n_kv, int64_t bs);
int ps_cuda_attn_pv(ps_cuda_ctx *ctx, float *out, const float *v_cache_t, const float *p, int64_t head_size, int64_t n_heads,
                    int64_t n_kv_heads, int64_t n_kv, int64_t n_ctx, int64_t bs);
/* strided fp32 copy (GGMLBackend::copy / cont -> powerserve_compute_forward_dup): 2-D, byte strides */
int ps_cuda_copy_2d(ps_cuda_ctx *ctx, void *dst, int64_t dst_stride0, int64_t dst_stride1, const void *src,
                    int64_t src_stride0, int64_t src_stride1, int64_t ne0, int64_t ne1);

/* the same for views of up to four dims whose shapes may differ (GGMLBackend::cont of a permuted view, copy into a
 * cache view; ggml_wrapper.cpp:135-161 -> powerserve_compute_forward_dup, ggml.c:9519-9558): element t of the source in
 * its row-major order lands on element t of the destination; shapes in elements, strides in bytes */
int ps_cuda_copy_4d(ps_cuda_ctx *ctx, void *dst, const int64_t dst_ne[4], const int64_t dst_nb[4], const void *src,
                    const int64_t src_ne[4], const int64_t src_nb[4]);
/* GGMLBackend::matmul with an FP32 src0 (ggml_wrapper.cpp:20-40 -> vec_dot_f32; the attention products over strided
 * cache views, norm_attention.cpp:117-147): dst {ne01, ne11, ne12} contiguous; src0 {ne00, ne01, ne02}, src1 {ne00, ne11,
 * ne12}, ne12 % ne02 == 0 (GQA broadcast, ggml.c:13365); innermost dims contiguous, nb* in bytes */
int ps_cuda_matmul_f32(ps_cuda_ctx *ctx, float *dst, const void *src0, int64_t ne00, int64_t ne01, int64_t ne02, int64_t nb01,
                       int64_t nb02, const void *src1, int64_t ne11, int64_t ne12, int64_t nb11, int64_t nb12);
/* GGMLBackend::softmax (ggml_wrapper.cpp:57-69 -> powerserve_compute_forward_soft_max, ggml.c:15060-15089): scale 1, no mask */
int ps_cuda_softmax(ps_cuda_ctx *ctx, float *dst, const float *x, int64_t ne0, int64_t n_rows);

/* ---------------------------------------------------------------------------------------------- KV cache
 * Device implementation of the position bookkeeping of KVCacheInterface (src/core/kv_cache.hpp:97-163) that
 * Platform::get/reset_kv_position (src/backend/platform.cpp:34-50) and LlamaModel::forward (:84-86,109) use. */
int     ps_cuda_kv_position(ps_cuda_ctx *ctx);
int     ps_cuda_kv_reset(ps_cuda_ctx *ctx);                    /* truncate_tokens(0) */
int     ps_cuda_kv_truncate(ps_cuda_ctx *ctx, int n_tokens);   /* truncate_tokens(n) */
int     ps_cuda_kv_rollback(ps_cuda_ctx *ctx, int n_tokens);   /* rollback_tokens(n) */
int     ps_cuda_kv_advance(ps_cuda_ctx *ctx, int n_tokens);    /* advance_tokens(n) */
/* slot operations of KVCacheInterface used by the speculative path (src/core/kv_cache.hpp:120-143, 188-231;
 * src/speculative/token_tree.cpp:195-212, 295-315): cache SLOTS are decoupled from token positions and carry a mask bit
 * (advance unmasks, rollback masks); `copy` takes a token of the LAST ps_cuda_forward_tree batch. */
int     ps_cuda_kv_copy_slot(ps_cuda_ctx *ctx, int dst_cache_index, int src_token_index);   /* copy(dst, src) */
int     ps_cuda_kv_move_slot(ps_cuda_ctx *ctx, int dst_cache_index, int src_cache_index);   /* move(dst, src) */
int     ps_cuda_kv_mask_slot(ps_cuda_ctx *ctx, int cache_index);                            /* mask(idx), idx < position */
int     ps_cuda_kv_unmask_slot(ps_cuda_ctx *ctx, int cache_index);                          /* unmask(idx) */
float  *ps_cuda_kv_k(ps_cuda_ctx *ctx, int layer);             /* device ptr, [n_ctx][kv_dim] */
float  *ps_cuda_kv_v(ps_cuda_ctx *ctx, int layer);             /* device ptr, [kv_dim][n_ctx] (transposed) */

/* ---------------------------------------------------------------------------------------------- whole-model path
 * The QNN precedent (`g.qnn_forward`, llama_model.cpp:66-77): one graph op that runs the whole forward pass inside
 * the backend.  bind uploads every weight; forward == LlamaModel::forward / Qwen2Model::forward (llama_model.cpp:
 * 52-117): tokens/pos are HOST arrays, logits_host (may be NULL when lm_head == 0) receives [bs][vocab] fp32 and the
 * call returns after the copy has landed; the KV position advances by bs. */
int ps_cuda_bind_model(ps_cuda_ctx *ctx, const ps_cuda_model_weights *w);
int ps_cuda_forward(ps_cuda_ctx *ctx, const int32_t *tokens, const int32_t *pos, int bs, int lm_head, float *logits_host);
/* The same forward as the speculative path calls it (src/speculative/spec_model.hpp:96-103, token_tree.cpp:131-133): arbitrary
 * token positions, `tree_mask` = bs x bs bytes, row i = the batch tokens token i attends to (TokenTree::attention_mask; NULL =
 * causal), attention over the unmasked cache slots below the current position; the batch's K / V rows are appended at cache
 * slots position .. position + bs - 1 and the position advances by bs (CausalLM::Batch::save_kv + advance, causal_models.cpp:
 * 353-359) - callers roll back and `copy` the accepted tokens, as TokenTree::verify does.  bs <= min(max_batch, 32).
 * lm_head = 1: logits_host receives [bs][vocab] fp32; lm_head = 2: the greedy pick is made on the device (first maximum per row,
 * llama_model.cpp:124-128 with top_k = 1) and logits_host receives bs int32 token ids instead. */
int ps_cuda_forward_tree(ps_cuda_ctx *ctx, const int32_t *tokens, const int32_t *pos, int bs, const uint8_t *tree_mask, int lm_head,
                         float *logits_host);
/* ---------------------------------------------------------------------------------------------- sessions
 * Server-side batching (app/server/server_handler.hpp:512-720): the reference serves one generation
 * per model at a time because the KV position is shared state (server_handler.hpp:224-240).  A context can hold several
 * independent KV sets over one set of weights.  Session 0 is the context's own cache.  ps_cuda_session_select makes a
 * session the target of every single-sequence call (forward / forward_tree / decode_greedy / kv_*), so prefill and KV
 * bookkeeping work per session unchanged; ps_cuda_forward_sessions advances n DISTINCT sessions by one token each in one
 * forward pass: one weight stream for the n columns, attention per column over its own cache and position.  logits_host
 * ([n][vocab], may be NULL) and greedy_ids ([n], may be NULL: arg-max on the device) are in batch order.  Each column's
 * result is bit-identical to decoding that session alone with batch 1. */
int ps_cuda_session_create(ps_cuda_ctx *ctx, int *session_id);
int ps_cuda_session_destroy(ps_cuda_ctx *ctx, int session_id);   /* not the selected one */
int ps_cuda_session_select(ps_cuda_ctx *ctx, int session_id);
int ps_cuda_session_current(ps_cuda_ctx *ctx);
int ps_cuda_session_position(ps_cuda_ctx *ctx, int session_id);  /* -1: unknown session */
int ps_cuda_forward_sessions(ps_cuda_ctx *ctx, const int32_t *session_ids, const int32_t *tokens, int n, int lm_head,
                             float *logits_host, int32_t *greedy_ids);

/* Model::decode with top_k = 1 (llama_model.cpp:119-132): forward + arg-max on the device; only the token id comes
 * back.  `n_steps` > 1 keeps feeding the produced id back in without a host round trip (ids_host gets n_steps ids). */
int ps_cuda_decode_greedy(ps_cuda_ctx *ctx, int32_t first_token, int n_steps, int32_t *ids_host);
/* device-side logits of the last forward ([bs][vocab]) for callers that sample on the GPU */
const float *ps_cuda_logits_dev(ps_cuda_ctx *ctx);
/* Device-side sampling: ps_cuda_forward / ps_cuda_forward_sessions accept logits_host == NULL with
 * lm_head = 1 - the logits stay on the device - and ps_cuda_sample_topk returns what ProbArray holds after
 * TopKSampler::apply (src/sampler/sampler.cpp:39-56, prob_array.hpp:43-49): the k <= 64 largest logits of row `row` of the
 * last forward pass, descending (equal logits by ascending token id), with their token ids.  2 k words cross PCIe instead of
 * the vocabulary's logits; the rest of the sampler chain (temperature, soft-max, top-p, stochastic pick) runs on k entries. */
int ps_cuda_sample_topk(ps_cuda_ctx *ctx, int row, int k, float *logits_out, int32_t *tokens_out);

/* ---------------------------------------------------------------------------------------------- tensor parallelism
 * One process per GPU.  A context created with desc.tp_size = N > 1 owns rows [rank * rows / N, (rank + 1) * rows / N) of
 * EVERY matrix (bind_model slices the host tensors itself) and only its own kv heads of the cache; the sharded outputs
 * of the fused decode step are exchanged with NCCL all-gathers on the context stream (4 per layer + 1 or 2 per token),
 * so every dot product keeps its full K and the results stay bit-identical to one GPU.  The host side moves the
 * 128-byte NCCL id from rank 0 to the other ranks (any transport: torch.distributed, MPI, a file). */
int ps_cuda_tp_unique_id(void *out128);                        /* rank 0: ncclGetUniqueId */
int ps_cuda_tp_init(ps_cuda_ctx *ctx, const void *id128);      /* every rank, after create, before the first forward */
/* Fused compute + all-gather over NVLink peer memory (optional, on top of tp_init): every rank exports the CUDA-IPC
 * handle of its exchange heap (64 bytes), the host gathers the N handles in rank order and every rank imports them.
 * From then on the producing kernels store their output rows straight into every rank's copy of the gathered vector
 * and publish an epoch flag; the consuming kernels wait on the flags (bounded spin; counter "tp_error").  Option
 * "tp_p2p" = 0 switches back to NCCL all-gathers (same results; tests compare the two).  The four per-layer exchanges
 * carry their flag in band (64-bit {value, epoch} peer stores, polled by the consumer: no fences); option "tp_ll" = 0
 * makes them use the fence + epoch-flag protocol as well. */
int ps_cuda_tp_export(ps_cuda_ctx *ctx, void *handle64);
int ps_cuda_tp_import(ps_cuda_ctx *ctx, const void *handles, int n);

/* execution switches (0/1): "graph" = replay the decode step as a captured CUDA graph, "fused" = fused decode
 * kernels instead of one kernel per table op, "pdl" = programmatic dependent launch between the fused kernels.  All
 * produce bit-identical results; they exist so tests can prove it.  "ktime" = 1 runs the fused step un-graphed with
 * CUDA events around every mat-vec launch (bench.py's roofline leg); "trace" = 1 records a per-kernel device timeline;
 * "cta_trace" = k (with "trace") adds a per-CTA stream trace of one mat-vec launch site (1 Wdown, 2 gate|up, 3 QKV, 4 Wo,
 * 5 lm_head; trace slots 256..); "rw_kb" = n caps the blocks per TMA ring stage of the mat-vec (tuning; default 16).
 * Kernel choice for batches (prefill chunks, verify batches; every choice computes the same bits): "attn_tile" = 0 runs
 * the round-1 batch attention kernels instead of the register-tiled ones; "tc_min" = n is the narrowest batch that takes
 * the tcgen05 GEMM (default 6; narrower ones walk with the multi-column row-walker); "scores_batch_min" (default 2) and
 * "pv_batch_min" (default 3) are the narrowest batches for the query-blocked scores / P.V kernels. */
int ps_cuda_set_option(ps_cuda_ctx *ctx, const char *name, int value);
/* counters: "kernel_launches" (kernels enqueued since create), "graph_replays", "h2d_bytes", "d2h_bytes",
 * "last_device_ns" (CUDA-event time, on the context stream, of the last forward / decode_greedy call),
 * "matvec_kernel_ns" / "matvec_kernel_launches" (option "ktime": summed event time / count of the mat-vec launches of
 * the last decode_greedy call), "tp_allgathers" (NCCL all-gathers enqueued since create) */
int64_t ps_cuda_get_counter(ps_cuda_ctx *ctx, const char *name);

/* debug: with option "trace" = 1 every kernel of the fused decode step records globaltimer stamps into its slot
 * (8 x int64 per launch: first start, last end, first dependency-resolved, slowest prologue, 3 prologue probes, spare;
 * at most 512 launches); read the first n_launches slots back here.  Not used in production. */
int ps_cuda_read_trace(ps_cuda_ctx *ctx, long long *host, int n_launches);

/* host-side restatement of glibc expf used by the device code; exported so CPU-only tests can pin it against libm */
float ps_cuda_host_expf_ref(float x);
float ps_cuda_host_v_expf(float x);

#ifdef __cplusplus
}
#endif
#endif /* PS_CUDA_H */
