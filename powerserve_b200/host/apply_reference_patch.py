"""Apply the PowerServe-side integration of the CUDA backend to a scratch COPY of the reference sources and (optionally)
print it as a unified diff.  This is exactly the patch INTEGRATION.md asks a maintainer to make; it follows the QNN
precedent (one OpType + one executor case + one Platform member, all behind `#if defined(POWERSERVE_WITH_CUDA)`).

    python apply_reference_patch.py <reference_root> <out_dir> [--diff]

Thirteen files are touched (two of them only for the Q4_K / Q6_K enum enablement, which is independent of CUDA); they are written under <out_dir>/src/... and take precedence on the include path.  Nothing
from the reference is stored in this repository.
"""
import difflib
import os
import sys


def sub(s, old, new, count=1):
    assert s.count(old) >= 1, f"anchor not found: {old[:60]!r}"
    return s.replace(old, new, count)


def patch_op_type(s):
    return sub(s, "    QNN_FORWARD_VL,\n#endif\n", "    QNN_FORWARD_VL,\n#endif\n\n#if defined(POWERSERVE_WITH_CUDA)\n    CUDA_FORWARD,\n#endif\n")


def patch_op_params(s):
    return sub(s, "#if defined(POWERSERVE_WITH_QNN)\nstruct QNNForwardParams", """#if defined(POWERSERVE_WITH_CUDA)
struct CUDAForwardParams : OpParams {
    std::vector<int> tokens;
    std::vector<int> pos;
    bool lm_head;

    explicit CUDAForwardParams(std::vector<int> tokens, std::vector<int> pos, bool lm_head) :
        tokens(std::move(tokens)), pos(std::move(pos)), lm_head(lm_head) {}

    ~CUDAForwardParams() override = default;
};
#endif

#if defined(POWERSERVE_WITH_QNN)
struct QNNForwardParams""")


def patch_graph_hpp(s):
    return sub(s, "#if defined(POWERSERVE_WITH_QNN)\n    auto qnn_forward(", """#if defined(POWERSERVE_WITH_CUDA)
    auto cuda_forward(const std::vector<int> &tokens, const std::vector<int> &pos, size_t vocab_size, bool lm_head) -> TensorNode *;
#endif

#if defined(POWERSERVE_WITH_QNN)
    auto qnn_forward(""")


def patch_graph_cpp(s):
    return sub(s, "#if defined(POWERSERVE_WITH_QNN)\nauto Graph::qnn_forward(", """#if defined(POWERSERVE_WITH_CUDA)
auto Graph::cuda_forward(const std::vector<int> &tokens, const std::vector<int> &pos, size_t vocab_size, bool lm_head)
    -> TensorNode * {
    auto op = new_op(OpType::CUDA_FORWARD);
    op->set_params(CUDAForwardParams(tokens, pos, lm_head));
    auto out = lm_head ? new_tensor(DataType::FP32, {vocab_size, tokens.size()}) : new_tensor(DataType::FP32, {0});
    op->set_outputs({out});
    return out;
}
#endif

#if defined(POWERSERVE_WITH_QNN)
auto Graph::qnn_forward(""")


def patch_executor(s):
    s = sub(s, "#if defined(POWERSERVE_WITH_QNN)\n        case OpType::QNN_FORWARD: {", """#if defined(POWERSERVE_WITH_CUDA)
        case OpType::CUDA_FORWARD: {
            auto &params = op->get_params<CUDAForwardParams>();
            m_platform.cuda_backends[model_id]->forward(op->output(), params.tokens, params.pos, params.lm_head);
        } break;
#endif

#if defined(POWERSERVE_WITH_QNN)
        case OpType::QNN_FORWARD: {""")
    # op-by-op mode: the unfused graph runs on the device, one CUDABackend method per op (buffer type decides the backend)
    s = sub(s, "void Executor::allocate_buffers() {\n", """void Executor::allocate_buffers() {
#if defined(POWERSERVE_WITH_CUDA)
    if (auto it = m_platform.cuda_backends.find(m_graph.m_model_id);
        it != m_platform.cuda_backends.end() && it->second->per_op()) {
        it->second->allocate_buffers(m_graph); // CUDABuffers instead of CPUBuffers
        return;
    }
#endif
""")
    return sub(s, "    auto &model_id = m_graph.m_model_id;\n    plan();\n", """    auto &model_id = m_graph.m_model_id;
#if defined(POWERSERVE_WITH_CUDA)
    if (auto it = m_platform.cuda_backends.find(model_id); it != m_platform.cuda_backends.end() && it->second->per_op()) {
        it->second->run(m_graph); // the same switch over OpType, dispatched to the CUDABackend operator table
        return;
    }
#endif
    plan();
""")


def patch_platform_hpp(s):
    s = sub(s, '#include "backend/ggml/ggml.hpp"\n', '#include "backend/ggml/ggml.hpp"\n\n#if defined(POWERSERVE_WITH_CUDA)\n#include "cuda_backend.hpp"\n#endif\n')
    s = sub(s, "    std::map<std::string, std::unique_ptr<ggml::GGMLBackend>> ggml_backends;\n",
            "    std::map<std::string, std::unique_ptr<ggml::GGMLBackend>> ggml_backends;\n\n#if defined(POWERSERVE_WITH_CUDA)\n"
            "    std::map<std::string, std::unique_ptr<cuda::CUDABackend>> cuda_backends;\n#endif\n")
    return sub(s, "    void destroy_ggml_backend(const std::shared_ptr<ModelConfig> &config);\n",
               "    void destroy_ggml_backend(const std::shared_ptr<ModelConfig> &config);\n\n#if defined(POWERSERVE_WITH_CUDA)\n"
               "    void init_cuda_backend(\n        const std::shared_ptr<ModelConfig> &config, const HyperParams &hparams, const Weight &weights, int device = 0\n    );\n#endif\n")


def patch_platform_cpp(s):
    s = sub(s, "#if defined(POWERSERVE_WITH_QNN)\nvoid Platform::init_qnn_backend", """#if defined(POWERSERVE_WITH_CUDA)
void Platform::init_cuda_backend(
    const std::shared_ptr<ModelConfig> &config, const HyperParams &hparams, const Weight &weights, int device
) {
    const bool qkv_bias = !weights.lw.empty() && weights.lw[0].attn_q_bias.m_data != nullptr;
    auto backend        = std::make_unique<cuda::CUDABackend>(config->llm, hparams, device, qkv_bias);
    backend->bind_weights(weights);
    cuda_backends.insert({config->model_id, std::move(backend)});
}
#endif

#if defined(POWERSERVE_WITH_QNN)
void Platform::init_qnn_backend""")
    s = sub(s, "    size_t position = ggml_backends.at(model_id)->m_kv->kv_cache->position;\n",
            "    size_t position = ggml_backends.at(model_id)->m_kv->kv_cache->position;\n#if defined(POWERSERVE_WITH_CUDA)\n"
            "    if (cuda_backends.count(model_id)) {\n        position = cuda_backends.at(model_id)->kv_position();\n    }\n#endif\n")
    return sub(s, "    ggml_backends[model_id]->m_kv->reset_kv_cache();\n",
               "    ggml_backends[model_id]->m_kv->reset_kv_cache();\n#if defined(POWERSERVE_WITH_CUDA)\n"
               "    if (cuda_backends.count(model_id)) {\n        cuda_backends[model_id]->reset_kv_cache();\n    }\n#endif\n")


def patch_model_forward(s):
    # LlamaModel::forward / Qwen2Model::forward: the whole forward pass becomes ONE graph op, like g.qnn_forward - or, in
    # op-by-op mode, the usual graph is built over the DEVICE caches and executed by the CUDABackend operator table
    s = sub(s, "    auto &llm_config = m_config->llm;\n\n#if defined(POWERSERVE_WITH_QNN)\n", """    auto &llm_config = m_config->llm;

#if defined(POWERSERVE_WITH_CUDA)
    const bool have_cuda    = m_platform->cuda_backends.count(m_config->model_id) > 0;
    const bool use_cuda_ops = have_cuda && m_platform->cuda_backends.at(m_config->model_id)->per_op();
    const bool use_cuda     = have_cuda && !use_cuda_ops;
    if (use_cuda) {
        logits = g.cuda_forward(tokens, pos, llm_config.vocab_size, lm_head);
    } else
#endif
#if defined(POWERSERVE_WITH_QNN)
""")
    s = sub(s, "                auto [k_cache, v_cache] = m_platform->ggml_backends[m_config->model_id]->m_kv->get_cache(L);\n", """#if defined(POWERSERVE_WITH_CUDA)
                auto [k_cache, v_cache] = use_cuda_ops ? m_platform->cuda_backends[m_config->model_id]->get_cache(L)
                                                       : m_platform->ggml_backends[m_config->model_id]->m_kv->get_cache(L);
#else
                auto [k_cache, v_cache] = m_platform->ggml_backends[m_config->model_id]->m_kv->get_cache(L);
#endif
""")
    s = sub(s, "#if defined(POWERSERVE_WITH_QNN)\n    if (!m_platform->qnn_backend)\n#endif\n    {", """#if defined(POWERSERVE_WITH_CUDA)
    if (use_cuda_ops) {
        m_platform->cuda_backends[m_config->model_id]->advance(batch_size);
    } else if (!use_cuda)
#endif
#if defined(POWERSERVE_WITH_QNN)
    if (!m_platform->qnn_backend)
#endif
    {""")
    return sub(s, "    return LogitsVector(logits->m_data, m_config->llm.vocab_size, batch_size);\n", """#if defined(POWERSERVE_WITH_CUDA)
    if (use_cuda_ops) { // the logits tensor lives on the device; LogitsVector reads a CPUBuffer (model.hpp:27-40)
        return LogitsVector(m_platform->cuda_backends[m_config->model_id]->download(logits), m_config->llm.vocab_size, batch_size);
    }
#endif
    return LogitsVector(logits->m_data, m_config->llm.vocab_size, batch_size);
""")


# ---- Q4_K / Q6_K enablement: the reference's DataType layer rejects the K-quants its vendored ggml supports;
# three enum / switch sites, no arithmetic (the same edit the oracle build applies with sed).
def patch_data_type(s):
    s = sub(s, "    GGML_Q8_0,\n", "    GGML_Q8_0,\n    GGML_Q4_K,\n    GGML_Q6_K,\n")
    s = sub(s, "        return ggml_type_size(GGML_TYPE_Q8_0);\n", "        return ggml_type_size(GGML_TYPE_Q8_0);\n    case DataType::GGML_Q4_K:\n"
            "        return ggml_type_size(GGML_TYPE_Q4_K);\n    case DataType::GGML_Q6_K:\n        return ggml_type_size(GGML_TYPE_Q6_K);\n")
    return sub(s, "        return ggml_blck_size(GGML_TYPE_Q8_0);\n", "        return ggml_blck_size(GGML_TYPE_Q8_0);\n    case DataType::GGML_Q4_K:\n"
               "        return ggml_blck_size(GGML_TYPE_Q4_K);\n    case DataType::GGML_Q6_K:\n        return ggml_blck_size(GGML_TYPE_Q6_K);\n")


def patch_ggml_hpp(s):
    s = sub(s, "        return GGML_TYPE_Q8_0;\n", "        return GGML_TYPE_Q8_0;\n    case DataType::GGML_Q4_K:\n        return GGML_TYPE_Q4_K;\n"
            "    case DataType::GGML_Q6_K:\n        return GGML_TYPE_Q6_K;\n")
    return sub(s, "        return DataType::GGML_Q8_0;\n", "        return DataType::GGML_Q8_0;\n    case GGML_TYPE_Q4_K:\n        return DataType::GGML_Q4_K;\n"
               "    case GGML_TYPE_Q6_K:\n        return DataType::GGML_Q6_K;\n")


def patch_ggml_wrapper(s):
    s = sub(s, "            dequantize_row_q8_0((block_q8_0 *)src, dst_tb + i * dim, dim);\n",
            "            dequantize_row_q8_0((block_q8_0 *)src, dst_tb + i * dim, dim);\n        } break;\n        case DataType::GGML_Q4_K: {\n"
            "            dequantize_row_q4_K((block_q4_K *)src, dst_tb + i * dim, dim);\n        } break;\n        case DataType::GGML_Q6_K: {\n"
            "            dequantize_row_q6_K((block_q6_K *)src, dst_tb + i * dim, dim);\n")
    # GGMLBackend::get_n_tasks sizes the CPU plan per op; the whole-model op needs no CPU tasks (same as QNN_FORWARD)
    return sub(s, "#if defined(POWERSERVE_WITH_QNN)\n    case OpType::QNN_FORWARD: {\n        n_tasks = 1;", """#if defined(POWERSERVE_WITH_CUDA)
    case OpType::CUDA_FORWARD: {
        n_tasks = 1;
    } break;
#endif

#if defined(POWERSERVE_WITH_QNN)
    case OpType::QNN_FORWARD: {
        n_tasks = 1;""")


def patch_ggml_cpp(s):
    # GGMLBackend::plan sizes the CPU scratch per op: nothing to size for the whole-model op (same as QNN_FORWARD)
    return sub(s, "#if defined(POWERSERVE_WITH_QNN)\n        case OpType::QNN_FORWARD: {\n        } break;", """#if defined(POWERSERVE_WITH_CUDA)
        case OpType::CUDA_FORWARD: {
        } break;
#endif

#if defined(POWERSERVE_WITH_QNN)
        case OpType::QNN_FORWARD: {
        } break;""")


FILES = {
    "src/core/data_type.hpp": patch_data_type,
    "src/backend/ggml/ggml.hpp": patch_ggml_hpp,
    "src/backend/ggml/ggml.cpp": patch_ggml_cpp,
    "src/backend/ggml/ggml_wrapper.cpp": patch_ggml_wrapper,
    "src/graph/op_type.hpp": patch_op_type,
    "src/graph/op_params.hpp": patch_op_params,
    "src/graph/graph.hpp": patch_graph_hpp,
    "src/graph/graph.cpp": patch_graph_cpp,
    "src/executor/executor.cpp": patch_executor,
    "src/backend/platform.hpp": patch_platform_hpp,
    "src/backend/platform.cpp": patch_platform_cpp,
    "src/model/llama/llama_model.cpp": patch_model_forward,
    "src/model/qwen2/qwen2_model.cpp": patch_model_forward,
}


def main():
    ref, out = sys.argv[1], sys.argv[2]
    for rel, fn in FILES.items():
        old = open(os.path.join(ref, rel)).read()
        new = fn(old)
        dst = os.path.join(out, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        open(dst, "w").write(new)
        if "--diff" in sys.argv:
            sys.stdout.writelines(difflib.unified_diff(old.splitlines(True), new.splitlines(True), "a/" + rel, "b/" + rel, n=1))


if __name__ == "__main__":
    main()
