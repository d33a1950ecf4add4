// ps_cuda_run.cpp — PowerServe's own model stack (Model -> Graph -> Executor -> Platform) running on the CUDA backend.
// Same command line and outputs as the test suite's driver of the unmodified CPU path, so the two can be diffed: it
// mirrors app/run/run.cpp:38-154 without CLI11 / tokenizer (the reference's submodules are empty here).
//
// usage: ps_cuda_run <model_dir> <n_threads> <batch_size> <prompt_ids.txt> <n_decode> <out_prefix> [--dump-logits N] [--device-topk K]
// --device-topk K: the logits stay on the device, TopKSampler runs there (CUDABackend::topk) and the host picks from K pairs
#include "backend/platform.hpp"
#include "core/config.hpp"
#include "model/llama/llama_model.hpp"
#include "model/module/norm_attention.hpp"
#include "model/qwen2/qwen2_model.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>

using namespace powerserve;

static std::vector<int> read_ids(const std::string &path) {
    std::ifstream f(path);
    std::vector<int> v;
    for (int x; f >> x;) v.push_back(x);
    return v;
}
static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s model_dir n_threads batch_size prompt_ids n_decode out_prefix [--dump-logits N]\n", argv[0]);
        return 2;
    }
    const std::string model_dir = argv[1];
    HyperParams hp;
    hp.n_threads  = atoi(argv[2]);
    hp.batch_size = (size_t)atoi(argv[3]);
    const auto prompt = read_ids(argv[4]);
    const int n_decode = atoi(argv[5]);
    const std::string out = argv[6];
    int dump_logits = 0, device_topk = 0;
    for (int i = 7; i < argc; i++) {
        if (!strcmp(argv[i], "--dump-logits") && i + 1 < argc) dump_logits = atoi(argv[++i]);
        if (!strcmp(argv[i], "--device-topk") && i + 1 < argc) device_topk = atoi(argv[++i]);
    }

    auto cfg = std::make_shared<ModelConfig>(Path(model_dir) / "model.json");
    std::shared_ptr<Model> model;
    const std::string weights = (Path(model_dir) / "ggml" / "weights.gguf").string();
    if (cfg->arch == "qwen2") model = std::make_shared<Qwen2Model>(weights, cfg);
    else model = std::make_shared<LlamaModel>(weights, cfg);
    model->m_platform = std::make_shared<Platform>();
    auto &platform    = model->m_platform;
    platform->init_ggml_backend(model->m_config, hp);                      // still owns GET_EMBEDDING of the (unused) graph leaf
    platform->init_cuda_backend(model->m_config, hp, *model->m_weights);   // <- the drop-in: everything else is unchanged
    model->m_attn = std::make_shared<NormAttention>(model->m_config->llm, model->m_weights);
    auto &model_id     = model->m_config->model_id;
    const size_t vocab = cfg->llm.vocab_size;
    platform->ggml_backends[model_id]->setup_threadpool();

    if (device_topk > 0) platform->cuda_backends[model_id]->set_lazy_logits(true);
    const double t0 = now_s();
    platform->reset_kv_position(model_id);
    size_t done = 0;
    while (done + 1 < prompt.size()) { // ModelTokenIterator prefill loop (src/model/model.hpp:147-160)
        const size_t bs = std::min(hp.batch_size, prompt.size() - done - 1);
        std::vector<int> tokens(prompt.begin() + done, prompt.begin() + done + bs), pos(bs);
        std::iota(pos.begin(), pos.end(), (int)platform->get_kv_position(model_id));
        model->forward(tokens, pos, CausalAttentionMask(bs), false);
        done += bs;
    }
    const double t1 = now_s();
    std::vector<int> ids;
    FILE *flog = dump_logits > 0 ? fopen((out + ".logits").c_str(), "wb") : nullptr;
    int tok = prompt.back();
    double t_first = t1;
    for (int step = 0; step < n_decode; step++) {
        auto ret = model->forward({tok}, {(int)platform->get_kv_position(model_id)}, CausalAttentionMask(1), true);
        const auto &logits = ret.logits_vector[0];
        int best = 0;
        if (device_topk > 0) { // TopKSampler on the device: K (logit, token) pairs instead of the vocabulary's logits
            best = platform->cuda_backends[model_id]->topk(device_topk)[0].second;
        } else {
            for (size_t i = 1; i < vocab; i++)
                if (logits[i] > logits[best]) best = (int)i;
            if (flog && step < dump_logits) fwrite(logits.data(), sizeof(float), vocab, flog);
        }
        ids.push_back(best);
        tok = best;
        if (step == 0) t_first = now_s();
    }
    const double t2 = now_s();
    platform->ggml_backends[model_id]->reset_threadpool();
    if (flog) fclose(flog);
    std::ofstream f(out + ".ids");
    for (int id : ids) f << id << "\n";
    printf("{\"backend\": \"cuda\", \"n_prompt\": %zu, \"n_decode\": %d, \"prefill_tok_s\": %.3f, \"decode_tok_s\": %.3f}\n", prompt.size(), n_decode,
           prompt.size() > 1 ? (prompt.size() - 1) / (t1 - t0) : 0.0, n_decode > 1 ? (n_decode - 1) / (t2 - t_first) : 0.0);
    return 0;
}
