/* oracle/ps_oracle.h — TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the arithmetic PowerServe's ggml CPU backend performs on the decode/prefill hot
 * path, as compiled for x86-64 AVX2+FMA+F16C (the reference's Release build with GGML_NATIVE=OFF).  The goal is
 * BIT-EXACT agreement with that build: the reference quantises activations before every weight matmul
 *  and a 1-ulp difference anywhere upstream of a quantiser is amplified to percent-level logit
 * noise within one forward pass (F13), so "close" is not a usable parity criterion — identical is.
 *
 * Pinned by tests/test_oracle_vs_ref.py against oracle/_ref/libggml_ref.so (the reference's own ggml compiled from
 * /root/reference) and oracle/_ref/ps_ref_run (the reference's own model stack), and by the golden vectors those
 * produced (tests/golden/).  Nothing under powerserve_b200/ or include/ may include, link or load this.
 */
#ifndef PS_ORACLE_H
#define PS_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ggml type ids (libs/ggml/include/ggml.h:386-401) */
enum { PS_OR_F32 = 0, PS_OR_F16 = 1, PS_OR_Q4_0 = 2, PS_OR_Q8_0 = 8, PS_OR_Q4_K = 12, PS_OR_Q5_K = 13, PS_OR_Q6_K = 14, PS_OR_Q8_K = 15 };

float    ps_or_fp16_to_fp32(uint16_t h);
uint16_t ps_or_fp32_to_fp16(float f);
float    ps_or_v_expf(float x);                  /* one lane of ggml_v_expf (AVX2), ggml.c:2685-2722 */

size_t ps_or_row_size(int type, int64_t k);      /* bytes of k elements of `type` */
int    ps_or_vec_dot_type(int wtype);            /* Q4_0/Q8_0 -> Q8_0 ; Q4_K/Q6_K -> Q8_K ; F32 -> F32 (ggml.c:734-900) */

/* activation quantisers (from_float of the vec_dot_type) */
void ps_or_quantize_row_q8_K(const float *x, void *y, int64_t k);   /* ggml-quants.c:3799-3837 */
void ps_or_quantize_row_q8_0(const float *x, void *y, int64_t k);   /* ggml-quants.c:957-1017 (AVX2 path) */
void ps_or_quantize_row(int qtype, const float *x, void *y, int64_t k);

/* weight -> fp32 (get_embedding and the definition of a weight value) */
void ps_or_dequantize_row(int type, const void *x, float *y, int64_t k);

/* one weight row . one quantised activation row, AVX2 lane order */
float ps_or_vec_dot(int wtype, int64_t k, const void *w_row, const void *xq_row);
float ps_or_vec_dot_f32(int64_t n, const float *x, const float *y);  /* ggml.c:2092-2131 */

/* operator table (ggml dim order: shape[0] contiguous; all buffers contiguous fp32 unless said otherwise) */
void ps_or_matmul(int wtype, const void *w, int64_t K, int64_t N, const float *x, int64_t bs, float *dst);
void ps_or_rmsnorm(float *dst, const float *x, const float *w, int64_t dim, int64_t bs, float eps);
void ps_or_rope(float *dst, const float *src, int64_t head_size, int64_t n_heads, int64_t bs, const int32_t *pos,
                int n_dims, int mode, float freq_base, float freq_scale, float attn_factor);
void ps_or_get_mask(float *mask, int64_t n_kv, int64_t bs, const int32_t *pos);
void ps_or_softmax_ext(float *dst, const float *x, const float *mask, int64_t ne0, int64_t ne1, int64_t ne2, float scale);
void ps_or_add(float *dst, const float *a, const float *b, int64_t n, int64_t nb);  /* b broadcast over rows when nb < n */
void ps_or_silu_hadamard(float *dst, const float *gate, const float *up, int64_t n);
void ps_or_get_embedding(float *dst, const void *w, int wtype, int64_t dim, const int32_t *tokens, int64_t bs);
/* fp32 attention matmuls over the reference's KV layout (K: [pos][kv_dim]; V transposed: [kv_dim][n_ctx]) */
void ps_or_attn_scores(float *kq, const float *k_cache, const float *q, int64_t head_size, int64_t n_heads,
                       int64_t n_kv_heads, int64_t n_kv, int64_t bs);
void ps_or_attn_pv(float *out, const float *v_cache_t, const float *p, int64_t head_size, int64_t n_heads,
                   int64_t n_kv_heads, int64_t n_kv, int64_t n_ctx, int64_t bs);

/* ---------------------------------------------------------------- whole-model forward */
typedef struct {
    int32_t dim, ffn_dim, n_layers, n_heads, n_kv_heads, head_size, vocab_size, n_ctx;
    float   norm_eps;
    int32_t rope_n_dims, rope_type;
    float   rope_freq_base, rope_freq_scale, rope_attn_factor;
    int32_t qkv_bias;
} ps_or_config;

typedef struct { const void *data; int32_t type; int32_t _pad; } ps_or_tensor;

typedef struct {
    ps_or_tensor attn_norm, ffn_norm, attn_q, attn_k, attn_v, attn_output, ffn_gate, ffn_up, ffn_down;
    ps_or_tensor q_bias, k_bias, v_bias;
} ps_or_layer;

typedef struct {
    ps_or_tensor token_embd, output_norm, output;   /* output.data == token_embd.data when tied */
    const ps_or_layer *layers;
} ps_or_weights;

typedef struct ps_or_model ps_or_model;

ps_or_model *ps_or_model_create(const ps_or_config *cfg, const ps_or_weights *w);
void         ps_or_model_free(ps_or_model *m);
void         ps_or_model_reset(ps_or_model *m);                 /* kv position = 0 */
int          ps_or_model_position(const ps_or_model *m);
void         ps_or_model_set_position(ps_or_model *m, int pos); /* rollback / truncate */
/* LlamaModel::forward (llama_model.cpp:52-117): logits [bs][vocab] written when lm_head != 0 */
int          ps_or_model_forward(ps_or_model *m, const int32_t *tokens, const int32_t *pos, int bs, int lm_head, float *logits);
/* debug taps: after a forward, copy an intermediate of layer L of the LAST forward into `out` (returns #floats):
 *   which: 0 = layer input x, 1 = attn output (pre-Wo, [bs][dim]), 2 = x after attention residual, 3 = layer output */
int64_t      ps_or_model_tap(ps_or_model *m, int layer, int which, float *out);
const float *ps_or_model_k_cache(ps_or_model *m, int layer);
const float *ps_or_model_v_cache(ps_or_model *m, int layer);

#ifdef __cplusplus
}
#endif
#endif
