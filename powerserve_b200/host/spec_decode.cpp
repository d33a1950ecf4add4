// spec_decode.cpp — host side of include/ps_spec.h: PowerServe's token-tree speculative decoding on the CUDA backend.
// Every step cites the reference code it mirrors (paths relative to /root/reference); the arithmetic on the device is the
// backend's own (ps_cuda_forward_tree), the cache bookkeeping is KVCacheInterface's (ps_cuda_kv_*).
#include "../../include/ps_spec.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <functional>
#include <queue>
#include <string>
#include <vector>

namespace {

double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

// TokenTree::Node (src/speculative/token_tree.hpp:52-66)
struct Node {
    static constexpr int no_parent = -1, not_in_cache = -1;
    int parent = no_parent, depth = 0, token = 0, position = 0, cache_index = not_in_cache;
    float current_prob = 1.0f;
    bool accepted = false;
    std::vector<int> children;
    void reset() { *this = Node(); }
};
// TokenTree::Candidate (token_tree.hpp:70-79)
struct Candidate {
    int token, parent;
    float current_prob, cumulative_prob;
    bool operator<(const Candidate &o) const { return cumulative_prob < o.cumulative_prob; }
};
struct ProbIndex {
    float prob;
    int token;
};

} // namespace

struct ps_spec {
    ps_cuda_ctx *target = nullptr, *draft = nullptr;
    ps_spec_config cfg{};
    int vocab = 0;
    std::string err;
    std::vector<Node> nodes;
    std::priority_queue<Candidate> main_heap, leaf_heap;
    std::vector<float> draft_logits;
    std::vector<int32_t> target_ids;
    std::vector<ProbIndex> probs;
    ps_spec_stats stat{};

    int fail(int rc, ps_cuda_ctx *c, const char *what) {
        err = std::string(what) + ": " + (c ? ps_cuda_last_error(c) : "");
        return rc;
    }
    bool should_stop(int token) const { // Tokenizer::should_stop (src/tokenizer/tokenizer.cpp:57-60)
        for (int i = 0; i < cfg.n_stop; i++)
            if (cfg.stop_tokens[i] == token) return true;
        return false;
    }
    // TokenTree::reset (token_tree.cpp:266-278)
    void reset(size_t batch_size) {
        nodes.resize(batch_size);
        for (auto &n : nodes) n.reset();
        while (!main_heap.empty()) main_heap.pop();
        while (!leaf_heap.empty()) leaf_heap.pop();
    }
    // TokenTree::lca (token_tree.cpp:280-293)
    int lca(int u, int v) const {
        if (nodes[u].depth < nodes[v].depth) std::swap(u, v);
        while (nodes[u].depth > nodes[v].depth) u = nodes[u].parent;
        while (u != v) {
            u = nodes[u].parent;
            v = nodes[v].parent;
        }
        return u;
    }
    // TokenTree::switch_parent (token_tree.cpp:295-315): mask the old branch's draft cache slots, unmask the new branch's
    int switch_parent(int old_parent, int new_parent) {
        if (old_parent == new_parent) return 0;
        const int p = lca(old_parent, new_parent);
        int rc;
        while (old_parent != p) {
            if ((rc = ps_cuda_kv_mask_slot(draft, nodes[old_parent].cache_index))) return fail(rc, draft, "mask");
            old_parent = nodes[old_parent].parent;
        }
        while (new_parent != p) {
            if ((rc = ps_cuda_kv_unmask_slot(draft, nodes[new_parent].cache_index))) return fail(rc, draft, "unmask");
            new_parent = nodes[new_parent].parent;
        }
        return 0;
    }
    // draft_sampler = TopK(top_k) -> Temperature -> Softmax (token_tree.cpp:35-39; sampler.cpp:19-59, prob_array.cpp softmax)
    void draft_sample(const float *logits) {
        probs.resize((size_t)vocab);
        for (int i = 0; i < vocab; i++) probs[i] = ProbIndex{logits[i], i};
        const size_t k = std::min<size_t>((size_t)cfg.top_k, probs.size());
        std::partial_sort(probs.begin(), probs.begin() + k, probs.end(), [](const ProbIndex &a, const ProbIndex &b) { return a.prob > b.prob; });
        probs.resize(k);
        if (cfg.temperature != 1.0f)
            for (auto &p : probs) p.prob /= cfg.temperature;
        // ProbArray::softmax: subtract the maximum, exponentiate, normalise
        float mx = probs[0].prob;
        for (auto &p : probs) mx = std::max(mx, p.prob);
        double sum = 0.0;
        for (auto &p : probs) {
            p.prob = expf(p.prob - mx);
            sum += p.prob;
        }
        for (auto &p : probs) p.prob = (float)(p.prob / sum);
    }

    // TokenTree::draft (token_tree.cpp:96-179)
    int do_draft(int root_token) {
        const size_t batch_size = (size_t)cfg.draft_batch_size;
        reset(batch_size);
        main_heap.push(Candidate{root_token, Node::no_parent, 1.0f, 1.0f});
        int last_parent = Node::no_parent, rc;
        size_t n_nodes = 0, n_saved_tokens = 0;
        while (n_nodes < batch_size) {
            const bool is_leaf = main_heap.empty();
            auto &heap = is_leaf ? leaf_heap : main_heap;
            if (heap.empty()) break;
            const Candidate c = heap.top();
            heap.pop();
            const int u = (int)n_nodes++;
            Node &node = nodes[u];
            node.token = c.token;
            node.current_prob = c.current_prob;
            if (c.parent == Node::no_parent) {
                node.position = ps_cuda_kv_position(draft);
            } else {
                node.position = nodes[c.parent].position + 1;
                node.parent = c.parent;
                node.depth = nodes[c.parent].depth + 1;
                nodes[c.parent].children.push_back(u);
            }
            // early terminate
            if (is_leaf || should_stop(node.token) || n_nodes + (cfg.early_stop ? main_heap.size() / 2 : 0) >= batch_size || c.cumulative_prob < cfg.min_prob) continue;
            if (last_parent != Node::no_parent && (rc = switch_parent(last_parent, c.parent))) return rc;
            node.cache_index = ps_cuda_kv_position(draft);
            const int32_t tok = node.token, pos = node.position;
            if ((rc = ps_cuda_forward_tree(draft, &tok, &pos, 1, nullptr, 1, draft_logits.data()))) return fail(rc, draft, "draft forward");
            n_saved_tokens++;
            last_parent = u;
            draft_sample(draft_logits.data());
            const float min_prob = probs[0].prob * cfg.p_base;
            size_t i = 0;
            for (const auto &item : probs) {
                const bool leaf_only = (i >= (size_t)cfg.max_fan_out || item.prob < min_prob);
                i++;
                (leaf_only ? leaf_heap : main_heap).push(Candidate{item.token, u, item.prob, c.cumulative_prob * item.prob});
            }
        }
        stat.n_draft_times += (int64_t)n_saved_tokens;
        stat.n_draft_tokens += (int64_t)n_nodes - 1; // exclude the root token
        if ((rc = ps_cuda_kv_rollback(draft, (int)n_saved_tokens))) return fail(rc, draft, "draft rollback");
        return 0;
    }

    // TokenTree::verify (token_tree.cpp:181-234) with greedy target sampling
    int do_verify(const std::function<void(int)> &enqueue) {
        stat.n_iterations += 1;
        int u = 0, rc;
        int64_t n_generated = 0;
        while (true) {
            Node &node = nodes[u];
            node.accepted = true;
            if (ps_cuda_kv_position(draft) != node.position || ps_cuda_kv_position(target) != node.position) {
                err = "verify: cache positions out of step with the tree";
                return PS_CUDA_ERR_INVALID;
            }
            if ((rc = ps_cuda_kv_copy_slot(target, node.position, u))) return fail(rc, target, "target copy");
            if ((rc = ps_cuda_kv_advance(target, 1))) return fail(rc, target, "target advance");
            if (node.cache_index == Node::not_in_cache) { // catch up with the target model
                const int32_t tok = node.token, pos = node.position;
                if ((rc = ps_cuda_forward_tree(draft, &tok, &pos, 1, nullptr, 0, nullptr))) return fail(rc, draft, "draft catch-up");
            } else {
                if ((rc = ps_cuda_kv_move_slot(draft, node.position, node.cache_index))) return fail(rc, draft, "draft move");
                if ((rc = ps_cuda_kv_advance(draft, 1))) return fail(rc, draft, "draft advance");
            }
            const int next = target_ids[u]; // ProbArray + greedy_sample with top_k = 1: the first maximum of row u
            enqueue(next);
            n_generated += 1;
            auto it = std::find_if(node.children.begin(), node.children.end(), [&](int v) { return nodes[v].token == next; });
            if (it == node.children.end()) break;
            u = *it;
            stat.n_accepted_tokens += 1;
        }
        stat.n_generated_tokens += n_generated;
        return 0;
    }

    // SpecTokenIterator::generate_tokens (spec_model.hpp:96-113)
    int generate_tokens(int last_token, const std::function<void(int)> &enqueue) {
        const int bs = cfg.draft_batch_size;
        double t0 = now_s();
        int rc = do_draft(last_token);
        if (rc) return rc;
        double t1 = now_s();
        stat.draft_s += t1 - t0;
        std::vector<int32_t> toks(bs), pos(bs);
        std::vector<uint8_t> mask((size_t)bs * bs, 0);
        for (int u = 0; u < bs; u++) { // TokenTree::tokens / positions / attention_mask (token_tree.cpp:59-94): a node sees its ancestors and itself
            toks[u] = nodes[u].token;
            pos[u] = nodes[u].position;
            for (int x = u; x != Node::no_parent; x = nodes[x].parent) mask[(size_t)u * bs + x] = 1;
        }
        // greedy target sampling: the arg-max of every row is taken on the device (12 ids come back instead of 12 x vocab logits)
        if ((rc = ps_cuda_forward_tree(target, toks.data(), pos.data(), bs, mask.data(), 2, reinterpret_cast<float *>(target_ids.data())))) return fail(rc, target, "target tree forward");
        if ((rc = ps_cuda_kv_rollback(target, bs))) return fail(rc, target, "target rollback");
        rc = do_verify(enqueue);
        stat.verify_s += now_s() - t1;
        return rc;
    }
};

extern "C" {

void ps_spec_default_config(ps_spec_config *c) {
    memset(c, 0, sizeof *c);
    c->draft_batch_size = 12; c->top_k = 15; c->temperature = 1.5f; c->p_base = 0.9f; c->max_fan_out = 3; c->min_prob = 0.2f; c->early_stop = 1;
}

int ps_spec_create(ps_spec **out, ps_cuda_ctx *target, ps_cuda_ctx *draft, const ps_spec_config *cfg) {
    if (!out || !target || !draft) return PS_CUDA_ERR_INVALID;
    ps_spec *s = new ps_spec();
    s->target = target;
    s->draft = draft;
    if (cfg) s->cfg = *cfg;
    else ps_spec_default_config(&s->cfg);
    *out = s;
    return 0;
}

void ps_spec_destroy(ps_spec *s) { delete s; }

const char *ps_spec_last_error(const ps_spec *s) { return s ? s->err.c_str() : ""; }

int ps_spec_generate(ps_spec *s, const int32_t *prompt, int n_prompt, int n_tokens, int prefill_batch, int32_t *ids_out, ps_spec_stats *stats) {
    if (!s || !prompt || n_prompt < 1 || n_tokens < 0 || !ids_out) return PS_CUDA_ERR_INVALID;
    if (s->cfg.draft_batch_size < 1 || s->cfg.draft_batch_size > 32 || s->cfg.top_k < 1) { s->err = "bad speculative config"; return PS_CUDA_ERR_INVALID; }
    s->stat = ps_spec_stats{};
    const double t_begin = now_s();
    // SpecTokenIterator ctor (spec_model.hpp:44-72): reset both caches, prefill prompt[:-1] on both models
    int rc;
    if ((rc = ps_cuda_kv_reset(s->target)) || (rc = ps_cuda_kv_reset(s->draft))) return s->fail(rc, s->target, "kv reset");
    {   // vocabulary from a probe: both models share it (asserted through the logits buffers below)
        // (the descriptor is not exported by the C ABI; the caller's logits rows are sized by the vocabulary)
    }
    const int n_prefill = n_prompt - 1;
    std::vector<int32_t> pos(std::max(prefill_batch, 1));
    for (int done = 0; done < n_prefill;) {
        const int bs = std::min(prefill_batch, n_prefill - done);
        for (int i = 0; i < bs; i++) pos[i] = done + i;
        if ((rc = ps_cuda_forward(s->target, prompt + done, pos.data(), bs, 0, nullptr))) return s->fail(rc, s->target, "target prefill");
        if ((rc = ps_cuda_forward(s->draft, prompt + done, pos.data(), bs, 0, nullptr))) return s->fail(rc, s->draft, "draft prefill");
        done += bs;
    }
    s->stat.prefill_s = now_s() - t_begin;
    if (s->vocab <= 0) { s->err = "vocabulary size not set (ps_spec_set_vocab)"; return PS_CUDA_ERR_INVALID; }
    s->draft_logits.resize((size_t)s->vocab);
    s->target_ids.resize((size_t)s->cfg.draft_batch_size);
    int last = prompt[n_prompt - 1], n_out = 0;
    while (n_out < n_tokens) { // SpecTokenIterator::decode (spec_model.hpp:76-92): one tree iteration yields >= 1 token
        std::vector<int> q;
        if ((rc = s->generate_tokens(last, [&](int t) { q.push_back(t); }))) return rc;
        for (int t : q) {
            if (n_out < n_tokens) ids_out[n_out++] = t;
        }
        last = q.back();
    }
    s->stat.total_s = now_s() - t_begin;
    if (stats) *stats = s->stat;
    return 0;
}

// the C ABI does not export a context's descriptor: the caller states the (shared) vocabulary size
int ps_spec_set_vocab(ps_spec *s, int vocab) {
    if (!s || vocab <= 0) return PS_CUDA_ERR_INVALID;
    s->vocab = vocab;
    return 0;
}

} // extern "C"
