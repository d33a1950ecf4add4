// oracle/ref_driver.cpp — TEST INFRASTRUCTURE. Drives the UNMODIFIED PowerServe reference stack
// (Model -> Graph -> Executor -> GGMLBackend -> vendored ggml) on a synthetic model directory and dumps
// token ids / logits / timings. It mirrors what app/run does without CLI11 / tokenizer:
//   setup        app/run/run.cpp:38-73   (load model, Platform::init_ggml_backend, NormAttention)
//   prefill loop src/model/model.hpp:141-160 (chunks of batch_size over prompt[:-1], lm_head=false)
//   decode loop  src/model/model.hpp:170-183 + llama_model.cpp:119-132 (forward(lm_head=true) -> greedy)
//   timing       app/run/run.cpp:96-154  (wall clock; prefill tok/s = (n_prompt-1)/t, decode excludes 1st token)
// Built only by `make -C oracle ref` inside the build container; the binary lives in oracle/_ref (git-ignored).
//
// usage: ps_ref_run <model_dir> <n_threads> <batch_size> <prompt_ids.txt> <n_decode> <out_prefix>
//                   [--force forced_ids.txt] [--dump-logits N] [--quiet]
//   prompt_ids.txt : whitespace separated token ids
//   --force        : teacher forcing — feed these ids as the decode inputs instead of the greedy ones
//   outputs        : <out_prefix>.ids (text, one greedy id per decode step), <out_prefix>.logits (fp32, N rows of vocab)
//                    and one JSON line on stdout with timings.
#include "backend/platform.hpp"
#include "core/config.hpp"
#include "core/timer.hpp"
#include "model/llama/llama_model.hpp"
#include "model/module/norm_attention.hpp"
#include "model/qwen2/qwen2_model.hpp"

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <string>
#include <vector>

using namespace powerserve;

static std::vector<int> read_ids(const std::string &path) {
    std::ifstream f(path);
    std::vector<int> v;
    int x;
    while (f >> x) v.push_back(x);
    return v;
}

static double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char **argv) {
    if (argc < 7) {
        fprintf(stderr, "usage: %s model_dir n_threads batch_size prompt_ids n_decode out_prefix [--force f] [--dump-logits N]\n", argv[0]);
        return 2;
    }
    const std::string model_dir = argv[1];
    const int n_threads         = atoi(argv[2]);
    const size_t batch_size     = (size_t)atoi(argv[3]);
    const auto prompt           = read_ids(argv[4]);
    const int n_decode          = atoi(argv[5]);
    const std::string out       = argv[6];
    std::vector<int> forced;
    int dump_logits = 0;
    for (int i = 7; i < argc; i++) {
        if (!strcmp(argv[i], "--force") && i + 1 < argc) forced = read_ids(argv[++i]);
        else if (!strcmp(argv[i], "--dump-logits") && i + 1 < argc) dump_logits = atoi(argv[++i]);
    }
    if (prompt.empty()) {
        fprintf(stderr, "empty prompt\n");
        return 2;
    }

    auto cfg = std::make_shared<ModelConfig>(Path(model_dir) / "model.json");
    std::shared_ptr<Model> model;
    const std::string weights = (Path(model_dir) / "ggml" / "weights.gguf").string();
    if (cfg->arch == "qwen2") model = std::make_shared<Qwen2Model>(weights, cfg);
    else model = std::make_shared<LlamaModel>(weights, cfg);

    HyperParams hp;
    hp.n_threads  = n_threads;
    hp.batch_size = batch_size;
    model->m_platform = std::make_shared<Platform>();
    auto &platform    = model->m_platform;
    platform->init_ggml_backend(model->m_config, hp);
    model->m_attn = std::make_shared<NormAttention>(model->m_config->llm, model->m_weights);

    auto &model_id = model->m_config->model_id;
    const size_t vocab = cfg->llm.vocab_size;

    const double t0 = now_s();
    platform->reset_kv_position(model_id);
    platform->ggml_backends[model_id]->setup_threadpool();
    // Work around a latent heap overflow in the reference: GGMLBackend::setup_work_data (src/backend/ggml/ggml.cpp:
    // 98-109) adds the per-thread cache-line padding only when it grows the buffer, and returns early when the new
    // request fits the PADDED size — so as n_kv grows by one per decode step, softmax's per-thread scratch
    // (ggml.c:14891, offset (nc+16)*ith) runs up to 64*n_threads bytes past the allocation (ASan: heap-buffer-
    // overflow in ggml_vec_cpy_f32 <- ggml_compute_forward_soft_max_f32; "double free or corruption" without ASan).
    // Pre-sizing the scratch once through the backend's own public method avoids it without touching arithmetic.
    platform->ggml_backends[model_id]->setup_work_data(
        size_t(64) << 20 | (size_t(batch_size) * cfg->llm.hidden_dim * 2 + size_t(cfg->llm.seq_len + 64) * 4 * (n_threads + 1))
    );
    size_t n_prefilled = 0;
    while (n_prefilled + 1 < prompt.size()) {
        size_t bs = std::min(batch_size, prompt.size() - n_prefilled - 1);
        std::vector<int> tokens(prompt.begin() + n_prefilled, prompt.begin() + n_prefilled + bs);
        std::vector<int> pos(bs);
        std::iota(pos.begin(), pos.end(), (int)platform->get_kv_position(model_id));
        model->forward(tokens, pos, CausalAttentionMask(bs), false);
        n_prefilled += bs;
    }
    const double t1 = now_s();

    std::vector<int> ids;
    FILE *flog = dump_logits > 0 ? fopen((out + ".logits").c_str(), "wb") : nullptr;
    int tok = prompt.back();
    double t_first = t1;
    for (int step = 0; step < n_decode; step++) {
        std::vector<int> tokens(1, tok);
        std::vector<int> pos(1, (int)platform->get_kv_position(model_id));
        auto ret           = model->forward(tokens, pos, CausalAttentionMask(1), true);
        const auto &logits = ret.logits_vector[0];
        int best           = 0;
        for (size_t i = 1; i < vocab; i++)
            if (logits[i] > logits[best]) best = (int)i; // first max wins, like ProbArray/greedy (top_k=1)
        if (flog && step < dump_logits) fwrite(logits.data(), sizeof(float), vocab, flog);
        ids.push_back(best);
        tok = (step < (int)forced.size()) ? forced[step] : best;
        if (step == 0) t_first = now_s();
    }
    const double t2 = now_s();
    platform->ggml_backends[model_id]->reset_threadpool();
    if (flog) fclose(flog);

    {
        std::ofstream f(out + ".ids");
        for (int id : ids) f << id << "\n";
    }
    const double prefill_s = t1 - t0;
    const double decode_s  = t2 - t_first;
    printf(
        "{\"n_threads\": %d, \"batch_size\": %zu, \"n_prompt\": %zu, \"n_decode\": %d, \"prefill_s\": %.6f, "
        "\"decode_s_excl_first\": %.6f, \"prefill_tok_s\": %.4f, \"decode_tok_s\": %.4f}\n",
        n_threads, batch_size, prompt.size(), n_decode, prefill_s, decode_s,
        prompt.size() > 1 ? (prompt.size() - 1) / prefill_s : 0.0, n_decode > 1 ? (n_decode - 1) / decode_s : 0.0
    );
    return 0;
}
