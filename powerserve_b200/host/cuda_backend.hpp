// cuda_backend.hpp — the PowerServe-side host class of the B200 CUDA backend (C++20, compiled INSIDE the PowerServe
// tree next to src/backend/ggml and src/backend/qnn).
//
// It mirrors powerserve::ggml::GGMLBackend (/root/reference/src/backend/ggml/ggml.hpp:186-250): same method names,
// argument meaning and error behaviour (POWERSERVE_ABORT on failure, no return codes), and adds the whole-model entry
// point `forward`, the analogue of qnn::QNNBackend::forward (src/backend/qnn/qnn_backend.cpp:47-98) that
// Executor::run dispatches for OpType::CUDA_FORWARD.  Everything below calls the C ABI of libps_cuda.so
// (include/ps_cuda.h); no CUDA headers are needed to build PowerServe with it.
#pragma once

#include "backend/backend.hpp"
#include "backend/cpu_buffer.hpp"
#include "core/config.hpp"
#include "core/tensor.hpp"
#include "model/common/weights.hpp"
#include "ps_cuda.h"

#include <memory>
#include <utility>
#include <vector>

namespace powerserve {
struct Graph;
}

namespace powerserve::cuda {

// Device memory behind a Tensor (the CUDA counterpart of CPUBuffer, src/backend/cpu_buffer.hpp:23-65).
struct CUDABuffer : BaseBuffer {
    ps_cuda_ctx *m_ctx = nullptr;
    Stride m_stride{}; // in bytes, like CPUBuffer
    void *m_data = nullptr;
    bool m_owned = false;

    CUDABuffer(ps_cuda_ctx *ctx, Stride stride, void *data, bool owned) : m_ctx(ctx), m_stride(stride), m_data(data), m_owned(owned) {}
    ~CUDABuffer() override;

    // a view of a parent buffer's memory with the contiguous strides of `shape` (CPUBuffer::create_buffer_view, cpu_buffer.hpp:53-64)
    template <typename T>
    static auto create_buffer_view(CUDABuffer &parent, Shape shape) -> BufferPtr {
        Stride stride;
        stride[0] = sizeof(T);
        for (size_t i = 1; i < shape.size(); i++) stride[i] = stride[i - 1] * shape[i - 1];
        POWERSERVE_ASSERT(parent.m_data != nullptr, "parent buffer is nullptr");
        return std::make_shared<CUDABuffer>(parent.m_ctx, stride, parent.m_data, false);
    }

    template <typename T>
    static auto create_buffer(ps_cuda_ctx *ctx, Shape shape) -> BufferPtr {
        Stride stride;
        stride[0] = sizeof(T);
        for (size_t i = 1; i < shape.size(); i++) stride[i] = stride[i - 1] * shape[i - 1];
        void *dev = nullptr;
        if (ps_cuda_malloc(ctx, stride.back() * shape.back(), &dev) != 0) POWERSERVE_ABORT("CUDABuffer: {}", ps_cuda_last_error(ctx));
        return std::make_shared<CUDABuffer>(ctx, stride, dev, true);
    }
};

struct CUDABackend : Backend {
public:
    CUDABackend(const ModelConfig::LLMConfig &config, const HyperParams &hparams, int device = 0, bool qkv_bias = false);
    ~CUDABackend() override;

    // ---- whole-model path (Executor::run, OpType::CUDA_FORWARD): LlamaModel::forward / Qwen2Model::forward on the device.
    // `out` is the CPUBuffer tensor {vocab, bs} the executor allocated for the logits (LogitsVector reads a CPUBuffer,
    // src/model/model.hpp:27-40); it is ignored when lm_head is false.  Advances the KV position by tokens.size().
    void bind_weights(const Weight &w);
    void forward(const Tensor *out, const std::vector<int> &tokens, const std::vector<int> &pos, bool lm_head);
    // Model::decode with a greedy sampler and no host round trip per token
    std::vector<int> decode_greedy(int first_token, int n_steps);
    // Device-side sampling: with lazy logits `forward` leaves the logits on the device (the CPUBuffer of the
    // CUDA_FORWARD op stays untouched) and topk(k) returns what ProbArray holds after TopKSampler::apply (sampler.cpp:39-56):
    // the k largest (logit, token) pairs, descending - the sampler chain then runs on k entries instead of the vocabulary.
    void set_lazy_logits(bool v) { m_lazy_logits = v; }
    std::vector<std::pair<float, int>> topk(int k, int row = 0) const;

    // ---- operator table (same names / argument order as GGMLBackend, ggml.hpp:216-244); tensors carry CUDABuffers,
    // weights are looked up by the host pointer of their CPUBuffer view of GGUF memory.
    void add(const Tensor *dst, const Tensor *src0, const Tensor *src1) const;
    void get_embedding(const Tensor *dst, const Tensor *weight, const std::vector<int> &tokens) const;
    void matmul(const Tensor *dst, const Tensor *src0, const Tensor *src1) const;
    void rmsnorm(const Tensor *o, const Tensor *x, const Tensor *weight, float eps) const;
    void rope(Tensor *out, const Tensor *src, const std::vector<int> &pos, const ModelConfig::LLMConfig::RopeConfig &rope_cfg) const;
    void softmax(const Tensor *out, const Tensor *x) const;
    void softmax_ext(const Tensor *out, const Tensor *x, const Tensor *mask, float scale, float max_bias) const;
    void permute(const Tensor *out, const Tensor *x, Shape axes) const;
    void transpose(const Tensor *out, const Tensor *x) const;
    void cont(const Tensor *out, const Tensor *x) const;
    void copy(const Tensor *dst, const Tensor *src) const;
    void silu_hadamard(const Tensor *out, const Tensor *hb, const Tensor *hb2) const;
    void print(const Tensor *x, size_t size) const;
    void get_mask(const Tensor *out, const std::vector<int> &pos) const;                 // executor-inline GET_MASK, executor.cpp:210-224
    void view(const Tensor *out, const Stride &stride, size_t offset) const;             // executor-inline VIEW, executor.cpp:194-199

    // ---- op-by-op execution of the UNFUSED PowerServe graph on the device (Executor::allocate_buffers / Executor::run,
    // executor.cpp:23-45, 77-235): every intermediate is a CUDABuffer, every op one call of the table above.  A bring-up /
    // debug path (about 30 launches per layer); selected with set_per_op(true) or POWERSERVE_CUDA_PER_OP=1.
    bool per_op() const { return m_per_op; }
    void set_per_op(bool v) { m_per_op = v; }
    void allocate_buffers(Graph &g) const;
    void run(Graph &g) const;
    auto get_cache(size_t L) -> std::pair<Tensor &, Tensor &>;   // GGMLKV::get_cache (ggml_kv_cache.hpp:157-159) over the device caches
    void advance(size_t n);                                      // GGMLKV::advance
    auto download(const Tensor *t) const -> BufferPtr;           // device tensor -> CPUBuffer (LogitsVector reads a CPUBuffer)

    // ---- KV position (Platform::get_kv_position / reset_kv_position, src/backend/platform.cpp:34-50)
    size_t kv_position() const;
    void reset_kv_cache();
    void rollback_tokens(size_t n);

    // GGMLBackend parity no-ops: the CUDA context owns its workspace and has no thread pool
    void plan(std::vector<std::shared_ptr<OpNode>> &) {}
    void setup_threadpool() {}
    void reset_threadpool() {}
    void reset_kv_batch_size(size_t) const {}

    ps_cuda_ctx *ctx() const { return m_ctx; }

private:
    void *device_weight(const Tensor *w) const; // registers on first use
    static void *dev(const Tensor *t);

    ps_cuda_ctx *m_ctx = nullptr;
    ModelConfig::LLMConfig m_config;
    std::vector<ps_cuda_layer_weights> m_layers;
    bool m_per_op = false;
    bool m_lazy_logits = false;
    std::vector<Tensor> m_key_tensors, m_value_tensors; // per layer, CUDABuffers over ps_cuda_kv_k / ps_cuda_kv_v
};

} // namespace powerserve::cuda
