/* include/ps_spec.h — speculative decoding over two CUDA-backend contexts (target + draft), C ABI.
 *
 * Host-side mirror of PowerServe's speculative path — TokenTree::draft / verify / switch_parent
 * (/root/reference/src/speculative/token_tree.cpp:96-234, 295-315) and SpecTokenIterator (src/speculative/spec_model.hpp:
 * 31-113) — written against include/ps_cuda.h: tree batches go through ps_cuda_forward_tree, the cache bookkeeping through
 * the ps_cuda_kv_* slot operations (KVCacheInterface, src/core/kv_cache.hpp:97-276).  The reference only offers this path
 * with the QNN backend (app/run/run.cpp:61, 107-113); here it runs on the CUDA backend (BASELINE.json configs[3]).
 * Greedy target sampling (top_k = 1), so the output is the target model's own greedy continuation ("lossless").
 */
#ifndef PS_SPEC_H
#define PS_SPEC_H
#include "ps_cuda.h"

#ifdef __cplusplus
extern "C" {
#endif

/* SpeculativeConfig defaults (src/speculative/speculative_config.hpp:21-36) */
typedef struct {
    int32_t draft_batch_size; /* 12 */
    int32_t top_k;            /* draft_sampler.top_k = 15 */
    float   temperature;      /* draft_sampler.temperature = 1.5 */
    float   p_base;           /* draft_sampler.p_base = 0.9 */
    int32_t max_fan_out;      /* token_tree.max_fan_out = 3 */
    float   min_prob;         /* token_tree.min_prob = 0.2 */
    int32_t early_stop;       /* token_tree.early_stop = true */
    int32_t n_stop;           /* Tokenizer::should_stop: bos / eos / eot / eom ids (none for synthetic models) */
    int32_t stop_tokens[8];
} ps_spec_config;

/* TokenTree::stat (token_tree.hpp:84-90) + wall-clock split */
typedef struct {
    int64_t n_draft_times, n_draft_tokens, n_accepted_tokens, n_iterations, n_generated_tokens;
    double  prefill_s, draft_s, verify_s, total_s;
} ps_spec_stats;

typedef struct ps_spec ps_spec;

void ps_spec_default_config(ps_spec_config *cfg);
/* both contexts must be bound (ps_cuda_bind_model), share the vocabulary, and have max_batch >= draft_batch_size */
int  ps_spec_create(ps_spec **out, ps_cuda_ctx *target, ps_cuda_ctx *draft, const ps_spec_config *cfg);
void ps_spec_destroy(ps_spec *s);
int  ps_spec_set_vocab(ps_spec *s, int vocab_size); /* vocabulary shared by target and draft (sizes the logits rows) */
/* SpecTokenIterator: prefill prompt[:-1] on both models in chunks of prefill_batch, then iterate draft -> tree verify until
 * n_tokens ids are produced (ids_out).  Returns a ps_cuda_status; ps_spec_last_error explains a failure. */
int  ps_spec_generate(ps_spec *s, const int32_t *prompt, int n_prompt, int n_tokens, int prefill_batch, int32_t *ids_out, ps_spec_stats *stats);
const char *ps_spec_last_error(const ps_spec *s);

#ifdef __cplusplus
}
#endif
#endif /* PS_SPEC_H */
