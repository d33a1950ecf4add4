// oracle/ref_ops_shim.cpp — TEST INFRASTRUCTURE. Plain-pointer extern "C" entry points that call the REFERENCE's own
// operator table — powerserve::ggml::GGMLBackend (src/backend/ggml/ggml.hpp:216-244) on powerserve::Tensor /
// CPUBuffer arguments, running on the reference's own ThreadPool — so that tests can put the reference's result next
// to the oracle's and the CUDA backend's for identical inputs.  No reference code is copied: this file only
// constructs the reference's types and calls its methods.  Built into oracle/_ref/libps_ref_ops.so by
// `make -C oracle ref` (build container only).
#include "backend/cpu_buffer.hpp"
#include "backend/ggml/ggml.hpp"
#include "core/config.hpp"
#include "core/tensor.hpp"

#include <cstring>
#include <memory>
#include <vector>

using namespace powerserve;

namespace {

DataType dtype_of(int ggml_type) {
    switch (ggml_type) {
    case 0: return DataType::FP32;
    case 2: return DataType::GGML_Q4_0;
    case 8: return DataType::GGML_Q8_0;
    case 12: return DataType::GGML_Q4_K;
    case 14: return DataType::GGML_Q6_K;
    default: POWERSERVE_ABORT("shim: unsupported type {}", ggml_type);
    }
}

// contiguous tensor over caller memory
Tensor make(int type, std::vector<size_t> shape, const void *data) {
    Shape s = {1, 1, 1, 1};
    for (size_t i = 0; i < shape.size(); i++) s[i] = shape[i];
    Tensor t(dtype_of(type), s);
    Stride st;
    st[0] = get_type_size(t.m_dtype);
    st[1] = st[0] * (s[0] / get_block_size(t.m_dtype));
    st[2] = st[1] * s[1];
    st[3] = st[2] * s[2];
    t.m_data = std::make_shared<CPUBuffer>(st, const_cast<void *>(data));
    return t;
}

Tensor make_strided(std::vector<size_t> shape, std::vector<size_t> stride, const void *data) {
    Shape s = {1, 1, 1, 1};
    Stride st;
    for (size_t i = 0; i < 4; i++) {
        s[i]  = i < shape.size() ? shape[i] : 1;
        st[i] = i < stride.size() ? stride[i] : st[i - 1] * s[i - 1];
    }
    Tensor t(DataType::FP32, s);
    t.m_data = std::make_shared<CPUBuffer>(st, const_cast<void *>(data));
    return t;
}

struct RefBackend {
    ModelConfig::LLMConfig cfg;
    HyperParams hp;
    std::unique_ptr<ggml::GGMLBackend> be;
};

} // namespace

extern "C" {

void *ref_backend_create(int n_threads) {
    // the fp16->fp32 lookup table used by every block kernel is filled by ggml_init (ggml.c:3700-3712)
    struct ggml_init_params ip = {.mem_size = 1 << 20, .mem_buffer = nullptr, .no_alloc = true};
    static ggml_context *ctx = ggml_init(ip);
    (void)ctx;
    auto *r = new RefBackend();
    r->cfg.dim = 64; r->cfg.hidden_dim = 64; r->cfg.n_layers = 1; r->cfg.n_heads = 1; r->cfg.n_kv_heads = 1;
    r->cfg.seq_len = 8; r->cfg.vocab_size = 8; r->cfg.kv_dim = 64; r->cfg.head_size = 64;
    r->hp.n_threads = n_threads;
    r->be = std::make_unique<ggml::GGMLBackend>(r->cfg, r->hp);
    r->be->setup_threadpool();
    r->be->setup_work_data(size_t(64) << 20);
    return r;
}

void ref_backend_destroy(void *h) {
    auto *r = static_cast<RefBackend *>(h);
    r->be->reset_threadpool();
    delete r;
}

// GGMLBackend::matmul — dst{N,bs} = W{K,N} . x{K,bs}
void ref_matmul(void *h, int wtype, const void *w, int64_t K, int64_t N, const float *x, int64_t bs, float *dst) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    be.setup_work_data(size_t(K) * bs * 2 + (1 << 20));
    Tensor W = make(wtype, {size_t(K), size_t(N)}, w), X = make(0, {size_t(K), size_t(bs)}, x), D = make(0, {size_t(N), size_t(bs)}, dst);
    be.matmul(&D, &W, &X);
}

void ref_rmsnorm(void *h, float *dst, const float *x, const float *w, int64_t dim, int64_t bs, float eps) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    Tensor X = make(0, {size_t(dim), size_t(bs)}, x), W = make(0, {size_t(dim)}, w), D = make(0, {size_t(dim), size_t(bs)}, dst);
    be.rmsnorm(&D, &X, &W, eps);
}

void ref_rope(void *h, float *dst, const float *src, int64_t head_size, int64_t n_heads, int64_t bs, const int32_t *pos,
              int n_dims, int mode, float freq_base, float freq_scale, float attn_factor) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    Tensor S = make(0, {size_t(head_size), size_t(n_heads), size_t(bs)}, src), D = make(0, {size_t(head_size), size_t(n_heads), size_t(bs)}, dst);
    ModelConfig::LLMConfig::RopeConfig rc;
    rc.n_dims = n_dims; rc.n_ctx_orig = 4096; rc.freq_base = freq_base; rc.freq_scale = freq_scale; rc.ext_factor = 0.0f;
    rc.attn_factor = attn_factor; rc.beta_fast = 32.0f; rc.beta_slow = 0.0f; rc.rope_type = mode;
    std::vector<int> p(pos, pos + bs);
    be.rope(&D, &S, p, rc);
}

void ref_softmax_ext(void *h, float *dst, const float *x, const float *mask, int64_t ne0, int64_t ne1, int64_t ne2, float scale) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    Tensor X = make(0, {size_t(ne0), size_t(ne1), size_t(ne2)}, x), M = make(0, {size_t(ne0), size_t(ne1)}, mask),
           D = make(0, {size_t(ne0), size_t(ne1), size_t(ne2)}, dst);
    be.softmax_ext(&D, &X, &M, scale, 0.0f);
}

void ref_add(void *h, float *dst, const float *a, const float *b, int64_t ne0, int64_t rows_a, int64_t rows_b) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    Tensor A = make(0, {size_t(ne0), size_t(rows_a)}, a), B = make(0, {size_t(ne0), size_t(rows_b)}, b), D = make(0, {size_t(ne0), size_t(rows_a)}, dst);
    be.add(&D, &A, &B);
}

void ref_silu_hadamard(void *h, float *dst, const float *gate, const float *up, int64_t n) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    Tensor G = make(0, {size_t(n)}, gate), U = make(0, {size_t(n)}, up), D = make(0, {size_t(n)}, dst);
    be.silu_hadamard(&D, &G, &U);
}

void ref_get_embedding(void *h, float *dst, const void *w, int wtype, int64_t dim, int64_t vocab, const int32_t *tokens, int64_t bs) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    Tensor W = make(wtype, {size_t(dim), size_t(vocab)}, w), D = make(0, {size_t(dim), size_t(bs)}, dst);
    std::vector<int> t(tokens, tokens + bs);
    be.get_embedding(&D, &W, t);
}

// The two fp32 attention matmuls exactly as NormAttention::build views them (norm_attention.cpp:115-147).
void ref_attn_scores(void *h, float *kq, const float *k_cache, const float *q, int64_t hs, int64_t n_heads, int64_t n_kv_heads,
                     int64_t n_kv, int64_t bs) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    const size_t f = sizeof(float), kv_dim = hs * n_kv_heads;
    Tensor K = make_strided({size_t(hs), size_t(n_kv), size_t(n_kv_heads)}, {f, f * kv_dim, f * hs, f * hs * n_kv_heads}, k_cache);
    Tensor Q = make_strided({size_t(hs), size_t(bs), size_t(n_heads)}, {f, f * hs * n_heads, f * hs, f * hs * n_heads * bs}, q);
    Tensor D = make(0, {size_t(n_kv), size_t(bs), size_t(n_heads)}, kq);
    be.matmul(&D, &K, &Q);
}

void ref_attn_pv(void *h, float *out, const float *v_cache_t, const float *p, int64_t hs, int64_t n_heads, int64_t n_kv_heads,
                 int64_t n_kv, int64_t n_ctx, int64_t bs) {
    auto &be = *static_cast<RefBackend *>(h)->be;
    const size_t f = sizeof(float);
    Tensor V = make_strided({size_t(n_kv), size_t(hs), size_t(n_kv_heads)}, {f, f * n_ctx, f * n_ctx * hs, f * n_ctx * hs * n_kv_heads}, v_cache_t);
    Tensor P = make(0, {size_t(n_kv), size_t(bs), size_t(n_heads)}, p);
    std::vector<float> kqv(size_t(hs) * bs * n_heads);
    Tensor KQV = make(0, {size_t(hs), size_t(bs), size_t(n_heads)}, kqv.data());
    be.matmul(&KQV, &V, &P);
    // permute {0,2,1,3} + cont -> {hs, n_heads, bs}
    Tensor PM = make_strided({size_t(hs), size_t(n_heads), size_t(bs)}, {f, f * hs * bs, f * hs, f * hs * bs * n_heads}, kqv.data());
    Tensor D  = make(0, {size_t(hs), size_t(n_heads), size_t(bs)}, out);
    be.cont(&D, &PM);
}

} // extern "C"
