// cuda_backend.cpp — see cuda_backend.hpp.  Thin: every method is one call into libps_cuda.so.
#include "cuda_backend.hpp"

#include "backend/ggml/ggml.hpp" // convert_datatype_to_ggml
#include "graph/graph.hpp"

#include <cstdio>
#include <cstdlib>

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
    if (const char *e = getenv("POWERSERVE_CUDA_PER_OP")) m_per_op = atoi(e) != 0;
    // the device caches in the shape / strides GGMLKV gives its tensors (ggml_kv_cache.cpp:35-58)
    const size_t kv_dim = c.n_kv_heads * c.head_size, n_ctx = c.seq_len;
    const Stride stride = {sizeof(float), sizeof(float) * n_ctx, sizeof(float) * kv_dim * n_ctx, sizeof(float) * kv_dim * n_ctx};
    for (size_t L = 0; L < c.n_layers; L++) {
        m_key_tensors.emplace_back(Tensor(DataType::FP32, {n_ctx, kv_dim, 1, 1}));
        m_value_tensors.emplace_back(Tensor(DataType::FP32, {n_ctx, kv_dim, 1, 1}));
        m_key_tensors[L].m_data   = std::make_shared<CUDABuffer>(m_ctx, stride, ps_cuda_kv_k(m_ctx, (int)L), false);
        m_value_tensors[L].m_data = std::make_shared<CUDABuffer>(m_ctx, stride, ps_cuda_kv_v(m_ctx, (int)L), false);
    }
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
    float *logits = (lm_head && !m_lazy_logits) ? static_cast<float *>(const_cast<Tensor *>(out)->get<CPUBuffer>().m_data) : nullptr;
    PS_CHECK(ps_cuda_forward(m_ctx, tokens.data(), pos.data(), (int)tokens.size(), lm_head ? 1 : 0, logits));
}

std::vector<int> CUDABackend::decode_greedy(int first_token, int n_steps) {
    std::vector<int> ids(n_steps);
    PS_CHECK(ps_cuda_decode_greedy(m_ctx, first_token, n_steps, ids.data()));
    return ids;
}

std::vector<std::pair<float, int>> CUDABackend::topk(int k, int row) const {
    std::vector<float> vals(k);
    std::vector<int> ids(k);
    PS_CHECK(ps_cuda_sample_topk(m_ctx, row, k, vals.data(), ids.data()));
    std::vector<std::pair<float, int>> out(k);
    for (int i = 0; i < k; i++) out[i] = {vals[i], ids[i]};
    return out;
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
    if (dynamic_cast<CUDABuffer *>(src0->m_data.get())) {
        // FP32 x FP32 over device views (the attention products: k_view . q and v_view . kq, norm_attention.cpp:117-147)
        POWERSERVE_ASSERT(src0->m_dtype == DataType::FP32 && src1->m_dtype == DataType::FP32);
        POWERSERVE_ASSERT(src0->m_shape[3] == 1 && src1->m_shape[3] == 1);
        const auto &a = src0->get<CUDABuffer>();
        const auto &b = src1->get<CUDABuffer>();
        POWERSERVE_ASSERT(a.m_stride[0] == sizeof(float) && b.m_stride[0] == sizeof(float));
        PS_CHECK(ps_cuda_matmul_f32(m_ctx, (float *)dev(dst), a.m_data, (int64_t)src0->m_shape[0], (int64_t)src0->m_shape[1], (int64_t)src0->m_shape[2],
                                    (int64_t)a.m_stride[1], (int64_t)a.m_stride[2], b.m_data, (int64_t)src1->m_shape[1], (int64_t)src1->m_shape[2],
                                    (int64_t)b.m_stride[1], (int64_t)b.m_stride[2]));
        return;
    }
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

void CUDABackend::softmax(const Tensor *out, const Tensor *x) const {
    PS_CHECK(ps_cuda_softmax(m_ctx, (float *)dev(out), (const float *)dev(x), (int64_t)x->m_shape[0], (int64_t)x->nrows()));
}

// permute / transpose / VIEW rewrite strides and the data pointer of a view buffer (ggml_wrapper.cpp:125-133,
// ggml.cpp:170-177, executor.cpp:194-199); nothing runs on the device
void CUDABackend::permute(const Tensor *out, const Tensor *x, Shape axes) const {
    const Stride &xs = x->get<CUDABuffer>().m_stride;
    Stride stride{};
    for (size_t i = 0; i < 4; i++) stride[axes[i]] = xs[i];
    out->get<CUDABuffer>().m_stride = stride;
}

void CUDABackend::transpose(const Tensor *out, const Tensor *x) const {
    Stride stride{x->get<CUDABuffer>().m_stride};
    std::swap(stride[0], stride[1]);
    out->get<CUDABuffer>().m_data   = x->get<CUDABuffer>().m_data;
    out->get<CUDABuffer>().m_stride = stride;
}

void CUDABackend::view(const Tensor *out, const Stride &stride, size_t offset) const {
    auto &b    = out->get<CUDABuffer>();
    b.m_stride = stride;
    b.m_data   = (char *)b.m_data + offset;
}

// copy / cont: powerserve_compute_forward_dup over two strided views whose shapes may differ (ggml_wrapper.cpp:135-161)
void CUDABackend::copy(const Tensor *dst, const Tensor *src) const {
    POWERSERVE_ASSERT(dst->m_dtype == DataType::FP32 && src->m_dtype == DataType::FP32);
    const auto &d = dst->get<CUDABuffer>();
    const auto &s = src->get<CUDABuffer>();
    int64_t dne[4], dnb[4], sne[4], snb[4];
    for (size_t i = 0; i < 4; i++) {
        dne[i] = (int64_t)dst->m_shape[i]; dnb[i] = (int64_t)d.m_stride[i];
        sne[i] = (int64_t)src->m_shape[i]; snb[i] = (int64_t)s.m_stride[i];
    }
    PS_CHECK(ps_cuda_copy_4d(m_ctx, d.m_data, dne, dnb, s.m_data, sne, snb));
}

void CUDABackend::cont(const Tensor *out, const Tensor *x) const { copy(out, x); }

void CUDABackend::print(const Tensor *x, size_t size) const {
    POWERSERVE_UNUSED(size);
    POWERSERVE_ASSERT(x->m_dtype == DataType::FP32);
    // GGMLBackend::print (ggml.cpp:131-151): shape, strides, every element, then exit
    Tensor host(DataType::FP32, x->m_shape);
    host.m_data = CPUBuffer::create_buffer<float>(x->m_shape);
    {
        Tensor dense(DataType::FP32, x->m_shape);
        dense.m_data = CUDABuffer::create_buffer<float>(m_ctx, x->m_shape);
        copy(&dense, x);
        PS_CHECK(ps_cuda_memcpy_d2h(m_ctx, host.get<CPUBuffer>().m_data, dense.get<CUDABuffer>().m_data, x->n_elements() * sizeof(float)));
    }
    const auto shape  = x->m_shape;
    const auto stride = x->get<CUDABuffer>().m_stride;
    printf("\n{%ld, %ld, %ld, %ld}\n", shape[3], shape[2], shape[1], shape[0]);
    printf("\n{%ld, %ld, %ld, %ld}\n", stride[3], stride[2], stride[1], stride[0]);
    const float *p = static_cast<const float *>(host.get<CPUBuffer>().m_data);
    for (size_t i = 0; i < x->n_elements(); i++) printf("%.6f\n", (double)p[i]);
    exit(0);
}

void CUDABackend::get_mask(const Tensor *out, const std::vector<int> &pos) const {
    PS_CHECK(ps_cuda_get_mask(m_ctx, (float *)dev(out), (int64_t)out->m_shape[0], (int64_t)pos.size(), pos.data()));
}

// ---- op-by-op execution of the unfused graph (the CUDA counterparts of Executor::allocate_buffers / Executor::run)
void CUDABackend::allocate_buffers(Graph &g) const {
    for (auto tensor : g.tensors) {
        if (tensor->m_data) continue; // weights (CPUBuffer views of GGUF memory) and the device caches
        POWERSERVE_ASSERT(tensor->m_dtype == DataType::FP32, "could not allocate a device buffer for data type: {}", static_cast<int>(tensor->m_dtype));
        if (tensor->type == NodeType::TENSOR_VIEW) {
            auto *parent = dynamic_cast<CUDABuffer *>(tensor->tensor_view()->parent->m_data.get());
            POWERSERVE_ASSERT(parent != nullptr, "view of a tensor that is not on the device");
            tensor->m_data = CUDABuffer::create_buffer_view<float>(*parent, tensor->m_shape);
        } else {
            tensor->m_data = CUDABuffer::create_buffer<float>(m_ctx, tensor->m_shape);
        }
    }
}

void CUDABackend::run(Graph &g) const {
    for (auto op : g.ops) {
        switch (op->op) {
        case OpType::GET_EMBEDDING: {
            auto [tokens] = op->get_params<GetEmbeddingParams>();
            get_embedding(op->output(), op->prev[0]->tensor(), tokens);
        } break;
        case OpType::ADD: add(op->output(), op->prev[0]->tensor(), op->prev[1]->tensor()); break;
        case OpType::MAT_MUL: matmul(op->output(), op->prev[0]->tensor(), op->prev[1]->tensor()); break;
        case OpType::RMS_NORM: {
            auto [eps] = op->get_params<RMSNormParams>();
            rmsnorm(op->output(), op->prev[0]->tensor(), op->prev[1]->tensor(), eps);
        } break;
        case OpType::SILU_HADAMARD: silu_hadamard(op->output(), op->prev[0]->tensor(), op->prev[1]->tensor()); break;
        case OpType::ROPE: {
            auto [pos, rope_cfg] = op->get_params<RopeParams>();
            rope(op->next[0]->tensor(), op->prev[0]->tensor(), pos, rope_cfg);
        } break;
        case OpType::SOFTMAX: softmax(op->output(), op->prev[0]->tensor()); break;
        case OpType::COPY: copy(op->prev[0]->tensor(), op->prev[1]->tensor()); break;
        case OpType::PRINT: print(op->prev[0]->tensor(), op->get_params<PrintParams>().size); break;
        case OpType::PERMUTE: {
            auto [axes] = op->get_params<PermuteParams>();
            permute(op->output(), op->prev[0]->tensor(), axes);
        } break;
        case OpType::CONT: cont(op->output(), op->prev[0]->tensor()); break;
        case OpType::VIEW: {
            auto [stride, offset] = op->get_params<ViewParams>();
            view(op->output(), stride, offset);
        } break;
        case OpType::SOFTMAX_EXT: {
            auto [scale, max_bias] = op->get_params<SoftmaxExtParams>();
            softmax_ext(op->output(), op->prev[0]->tensor(), op->prev[1]->tensor(), scale, max_bias);
        } break;
        case OpType::GET_MASK: {
            auto [mask, pos] = op->get_params<GetMaskParams>();
            get_mask(op->output(), pos);
        } break;
        case OpType::TRANSPOSE: transpose(op->output(), op->prev[0]->tensor()); break;
        default: POWERSERVE_ABORT("cuda backend: OpType {} has no device twin", static_cast<int>(op->op));
        }
    }
}

auto CUDABackend::get_cache(size_t L) -> std::pair<Tensor &, Tensor &> { return {m_key_tensors[L], m_value_tensors[L]}; }

void CUDABackend::advance(size_t n) { PS_CHECK(ps_cuda_kv_advance(m_ctx, (int)n)); }

auto CUDABackend::download(const Tensor *t) const -> BufferPtr {
    auto host = CPUBuffer::create_buffer<float>(t->m_shape);
    PS_CHECK(ps_cuda_memcpy_d2h(m_ctx, dynamic_cast<CPUBuffer &>(*host).m_data, dev(t), t->n_elements() * sizeof(float)));
    return host;
}

size_t CUDABackend::kv_position() const { return (size_t)ps_cuda_kv_position(m_ctx); }
void CUDABackend::reset_kv_cache() { PS_CHECK(ps_cuda_kv_reset(m_ctx)); }
void CUDABackend::rollback_tokens(size_t n) { PS_CHECK(ps_cuda_kv_rollback(m_ctx, (int)n)); }

} // namespace powerserve::cuda
