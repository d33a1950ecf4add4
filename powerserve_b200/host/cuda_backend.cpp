// cuda_backend.cpp — see cuda_backend.hpp.  Thin: every method is one call into libps_cuda.so.
#include "cuda_backend.hpp"

#include "backend/ggml/ggml.hpp" // convert_datatype_to_ggml

namespace powerserve::cuda {

#define PS_CHECK(call)                                                                  \
    do {                                                                                \
        if ((call) != 0) POWERSERVE_ABORT("cuda backend: {}", ps_cuda_last_error(m_ctx)); \
    } while (0)

CUDABuffer::~CUDABuffer() {
    if (m_owned && m_data) ps_cuda_free(m_ctx, m_data);
}

static int ggml_type_of(const Tensor &t) { return (int)ggml::convert_datatype_to_ggml(t.m_dtype); }
static const void *host_ptr(const Tensor &t) { return t.m_data ? const_cast<Tensor &>(t).get<CPUBuffer>().m_data : nullptr; }

CUDABackend::CUDABackend(const ModelConfig::LLMConfig &c, const HyperParams &hparams, int device, bool qkv_bias) : m_config(c) {
    ps_cuda_model_desc d{};
    d.dim = (int)c.dim; d.ffn_dim = (int)c.hidden_dim; d.n_layers = (int)c.n_layers; d.n_heads = (int)c.n_heads;
    d.n_kv_heads = (int)c.n_kv_heads; d.head_size = (int)c.head_size; d.vocab_size = (int)c.vocab_size;
    d.n_ctx = (int)c.seq_len;                       // GGMLKV sizes its cache for seq_len too (ggml_kv_cache.cpp:35-58)
    d.norm_eps = c.norm_eps;
    d.rope_n_dims = c.rope_config.n_dims; d.rope_type = c.rope_config.rope_type;
    d.rope_freq_base = c.rope_config.freq_base; d.rope_freq_scale = c.rope_config.freq_scale; d.rope_attn_factor = c.rope_config.attn_factor;
    d.qkv_bias = qkv_bias ? 1 : 0;
    d.max_batch = (int)hparams.batch_size;
    d.tp_rank = 0; d.tp_size = 1;
    if (ps_cuda_create(&m_ctx, device, &d) != 0) POWERSERVE_ABORT("cuda backend: {}", ps_cuda_last_error(nullptr));
}

CUDABackend::~CUDABackend() { ps_cuda_destroy(m_ctx); }

void CUDABackend::bind_weights(const Weight &w) {
    auto T = [](const Tensor &t) { return ps_cuda_tensor{host_ptr(t), t.m_data ? ggml_type_of(t) : 0, 0}; };
    m_layers.clear();
    for (const auto &lw : w.lw) {
        ps_cuda_layer_weights l{};
        l.attn_norm = T(lw.attn_norm); l.ffn_norm = T(lw.ffn_norm);
        l.attn_q = T(lw.attn_q); l.attn_k = T(lw.attn_k); l.attn_v = T(lw.attn_v); l.attn_output = T(lw.attn_output);
        l.ffn_gate = T(lw.ffn_gate); l.ffn_up = T(lw.ffn_up); l.ffn_down = T(lw.ffn_down);
        l.attn_q_bias = T(lw.attn_q_bias); l.attn_k_bias = T(lw.attn_k_bias); l.attn_v_bias = T(lw.attn_v_bias);
        m_layers.push_back(l);
    }
    ps_cuda_model_weights mw{T(w.token_embedding_table), T(w.rms_final_weight), T(w.output_weight), m_layers.data()};
    PS_CHECK(ps_cuda_bind_model(m_ctx, &mw));
}

void CUDABackend::forward(const Tensor *out, const std::vector<int> &tokens, const std::vector<int> &pos, bool lm_head) {
    float *logits = lm_head ? static_cast<float *>(const_cast<Tensor *>(out)->get<CPUBuffer>().m_data) : nullptr;
    PS_CHECK(ps_cuda_forward(m_ctx, tokens.data(), pos.data(), (int)tokens.size(), lm_head ? 1 : 0, logits));
}

std::vector<int> CUDABackend::decode_greedy(int first_token, int n_steps) {
    std::vector<int> ids(n_steps);
    PS_CHECK(ps_cuda_decode_greedy(m_ctx, first_token, n_steps, ids.data()));
    return ids;
}

void *CUDABackend::dev(const Tensor *t) { return const_cast<Tensor *>(t)->get<CUDABuffer>().m_data; }

void *CUDABackend::device_weight(const Tensor *w) const {
    const void *h = host_ptr(*w);
    if (void *d = ps_cuda_lookup_weight(m_ctx, h)) return d;
    void *d = nullptr;
    PS_CHECK(ps_cuda_register_weight(m_ctx, h, ggml_type_of(*w), (int64_t)w->m_shape[0], (int64_t)w->nrows(), &d));
    return d;
}

void CUDABackend::add(const Tensor *dst, const Tensor *src0, const Tensor *src1) const {
    // src1 is either the same shape (residual) or a {N} row broadcast bias held in GGUF memory (qwen2_model / norm_attention)
    const bool bias = dynamic_cast<CPUBuffer *>(src1->m_data.get()) != nullptr;
    const float *b = bias ? static_cast<const float *>(device_weight(src1)) : static_cast<const float *>(dev(src1));
    PS_CHECK(ps_cuda_add(m_ctx, (float *)dev(dst), (const float *)dev(src0), b, (int64_t)src0->n_elements(), (int64_t)src1->n_elements()));
}

void CUDABackend::get_embedding(const Tensor *dst, const Tensor *weight, const std::vector<int> &tokens) const {
    PS_CHECK(ps_cuda_get_embedding(m_ctx, (float *)dev(dst), device_weight(weight), ggml_type_of(*weight), (int64_t)weight->m_shape[0],
                                   tokens.data(), (int64_t)tokens.size()));
}

void CUDABackend::matmul(const Tensor *dst, const Tensor *src0, const Tensor *src1) const {
    POWERSERVE_ASSERT(tensor_can_mul_mat(src0, src1));
    PS_CHECK(ps_cuda_matmul(m_ctx, (float *)dev(dst), device_weight(src0), ggml_type_of(*src0), (int64_t)src0->m_shape[0],
                            (int64_t)src0->m_shape[1], (const float *)dev(src1), (int64_t)src1->nrows()));
}

void CUDABackend::rmsnorm(const Tensor *o, const Tensor *x, const Tensor *weight, float eps) const {
    PS_CHECK(ps_cuda_rmsnorm(m_ctx, (float *)dev(o), (const float *)dev(x), (const float *)device_weight(weight), (int64_t)x->m_shape[0],
                             (int64_t)x->nrows(), eps));
}

void CUDABackend::rope(Tensor *out, const Tensor *src, const std::vector<int> &pos, const ModelConfig::LLMConfig::RopeConfig &) const {
    // src is {head_size, n_heads, bs}; the context was created from the same RopeConfig
    PS_CHECK(ps_cuda_rope(m_ctx, (float *)dev(out), (const float *)dev(src), (int64_t)src->m_shape[0], (int64_t)src->m_shape[1],
                          (int64_t)src->m_shape[2], pos.data()));
}

void CUDABackend::softmax_ext(const Tensor *out, const Tensor *x, const Tensor *mask, float scale, float max_bias) const {
    POWERSERVE_ASSERT(max_bias == 0.0f); // the reference only ever passes 0 (norm_attention.cpp:133)
    PS_CHECK(ps_cuda_softmax_ext(m_ctx, (float *)dev(out), (const float *)dev(x), (const float *)dev(mask), (int64_t)x->m_shape[0],
                                 (int64_t)x->m_shape[1], (int64_t)x->m_shape[2], scale));
}

void CUDABackend::silu_hadamard(const Tensor *out, const Tensor *hb, const Tensor *hb2) const {
    PS_CHECK(ps_cuda_silu_hadamard(m_ctx, (float *)dev(out), (const float *)dev(hb), (const float *)dev(hb2), (int64_t)hb->n_elements()));
}

void CUDABackend::copy(const Tensor *dst, const Tensor *src) const {
    auto &d = const_cast<Tensor *>(dst)->get<CUDABuffer>();
    auto &s = const_cast<Tensor *>(src)->get<CUDABuffer>();
    PS_CHECK(ps_cuda_copy_2d(m_ctx, d.m_data, (int64_t)d.m_stride[0], (int64_t)d.m_stride[1], s.m_data, (int64_t)s.m_stride[0],
                             (int64_t)s.m_stride[1], (int64_t)src->m_shape[0], (int64_t)src->nrows()));
}

void CUDABackend::get_mask(const Tensor *out, const std::vector<int> &pos) const {
    PS_CHECK(ps_cuda_get_mask(m_ctx, (float *)dev(out), (int64_t)out->m_shape[0], (int64_t)pos.size(), pos.data()));
}

size_t CUDABackend::kv_position() const { return (size_t)ps_cuda_kv_position(m_ctx); }
void CUDABackend::reset_kv_cache() { PS_CHECK(ps_cuda_kv_reset(m_ctx)); }
void CUDABackend::rollback_tokens(size_t n) { PS_CHECK(ps_cuda_kv_rollback(m_ctx, (int)n)); }

} // namespace powerserve::cuda
