// ps_step.cuh — ONE persistent kernel per decode step (bs = 1): the whole forward pass of LlamaModel::forward
// (src/model/llama/llama_model.cpp:52-117) as the executor would run it op by op (src/executor/executor.cpp:77-235), on
// 148 resident CTAs that walk the phases of every layer together.  Same arithmetic as the per-phase kernels in ps_rw.cuh /
// ps_decode.cuh (block math, FMA chains, quantiser, soft-max: all shared device functions), bit-identical to the table ops.
//
// Why (round-1 measurements): the per-phase kernels stream their weights at 5.6-6.0 TB/s once they
// run, but a decode step was 62 % fixed cost - six kernel boundaries per layer, each with ~1 us of dependency latency
// and a 2-5 us prologue during which DRAM idles, because a CTA that holds 200 KB of shared memory cannot overlap with
// its successor.  Here the weight stream never stops:
//   * two PRODUCER warps per CTA (one issuing lane each) walk a static program of TMA bulk copies (cp.async.bulk + mbarrier) over ALL mat-vec
//     phases of ALL layers, bounded only by their shared-memory rings (38 slots x 4736 B per SM = 26 MB chip-wide), so
//     the weights of the next phases arrive while the compute warps sit in a prologue, an attention phase or a barrier;
//   * sixteen COMPUTE warps run the phases.  Activation vectors travel between phases as (value, epoch) words - every
//     store is one 64-bit word whose upper half is the epoch of this (step, layer), and the consumer's prologue polls
//     the very words it is about to read (the in-band-flag exchange of the tensor-parallel path, used inside one GPU as
//     well): no barrier and no fence between Wo -> gate|up -> down -> next layer's QKV.  Only the attention needs two
//     device-wide barriers per layer (q/k/v complete; scores complete);
//   * tensor parallel: the same (value, epoch) stores go to every rank's copy over NVLink peer mappings - the
//     all-gathers are still fused into the producing phase and there is still no NCCL on the data path.
// Every wait is bounded: a dead peer / a sequencing bug raises `step_error` (-> PS_CUDA_ERR_CUDA) instead of hanging.
#pragma once
#include "ps_rw.cuh"

#define PS_ST_WARPS 16                      // compute warps
#define PS_ST_PROD 4                        // producer warps; producer p feeds the compute warps w with w % PS_ST_PROD == p
#define PS_ST_WPP (PS_ST_WARPS / PS_ST_PROD)
#define PS_ST_CT (PS_ST_WARPS * 32)         // compute threads
#define PS_ST_THREADS ((PS_ST_WARPS + PS_ST_PROD) * 32)      // 20 warps: one more would cost 16 registers per thread (warp allocation granularity)
#define PS_ST_SLOT 4736                     // ring slot: four 1184-byte octet blocks
#ifndef PS_ST_DEBUG
#define PS_ST_DEBUG 0
#endif
#define PS_ST_MODE_LMHEAD 1
#define PS_ST_MODE_PICK 2

struct PsStLayer {
    const uint8_t *w_qkv, *w_o, *w_gu, *w_down;  // octet-interleaved (ps_k_rw_repack), rows of this rank
    const float *attn_norm, *ffn_norm, *q_bias, *k_bias, *v_bias;
    float *kc, *vct;                             // K cache [n_ctx][kvd_l], V cache transposed [kvd_l][n_ctx]
};

// where the (value, epoch) words of every exchanged vector live on every rank (own rank included); tp == 1: entry 0 only
struct PsStPeers {
    unsigned long long *x[PS_TP_MAX], *x1[PS_TP_MAX], *att[PS_TP_MAX], *hq[PS_TP_MAX], *best[PS_TP_MAX];
};

struct PsStArgs {
    const PsStLayer *layers;
    const PsStPeers *peers;
    int n_layers, dim, qdim, ffn, hs, n_ctx;     // full sizes (K of Wo = qdim, K of Wdown = ffn)
    int qdim_l, kvd_l, ffn_l, vocab_l, nkv_l;    // this rank's rows / kv heads
    int tp, rank;
    float eps, kq_scale;
    const uint8_t *w_embd;                       // token_embd rows as stored in the GGUF (Q4_K)
    const uint8_t *w_out;                        // lm_head rows of this rank, octet-interleaved
    const float *out_norm;
    const float *rope_table;
    int32_t *tokens_dev, *pos_dev, *ids_dev, *ctr_dev;
    unsigned long long *x_ll, *x1_ll, *att_ll, *hq_ll, *best_ll; // this rank's copies
    float *q, *sc, *h, *logits;                  // plain scratch: rotated q [qdim_l], scores [n_heads_l][n_ctx], FFN hidden [ffn_l], logits [vocab_l]
    int *blk_cnt;                                // per-256-block arrival counters of the hidden vector (rest state 0)
    float *part_val;                             // per-CTA arg-max partials
    int *part_idx;
    unsigned *bar_ctr, *done_ctr, *serial;       // device-wide barrier counter, finish counter, step serial number
    int *err;                                    // set to 1 if any wait gave up
    const PsTpOut *tpo_logits;                   // tensor parallel, host-visible logits: peer stores + epoch flag (ps_tp_signal)
    int kb_dim, kb_gu, kb_q, kb_ffn;             // octet blocks per ring stage of the four weight shapes
    int nsp;                                     // ring slots per producer
    int attn_chunk;                              // positions of a soft-max row resident in shared memory at a time (multiple of 32)
    int dpc;                                     // output dims per CTA in the P.V phase (8, 4, 2 or 1)
    int mode;                                    // PS_ST_MODE_*
    int l2_ahead;                                // producer: stages of look-ahead prefetched into L2 (0 = off)
    long long timeout_ns;
    long long *tl;                               // optional timeline, 8 int64 per phase: [0] first CTA in, [1] last CTA out, [3] last CTA's inputs arrived, [4] last CTA's image ready
};

// ---------------------------------------------------------------------------------------------------- bounded waits
struct PsStCtl {
    int *abort;          // shared-memory flag: once set every wait returns at once
    int *err;            // global flag
    long long timeout_ns;
    volatile uint32_t *dbg_rel; // PS_ST_DEBUG: [slot] index of the stage whose consumer last released the slot, [nsp + slot] that warp
};
// every 256 failed attempts: has anybody given up, or is it time to give up?  Out of line on purpose: the step kernel is one
// large body of code shared by twenty warps in different phases, and its instruction-cache footprint is a first-order cost.
__device__ __noinline__ bool ps_st_spin_check(int *abort, int *err, long long timeout_ns, long long *t0, int site) {
    if (*reinterpret_cast<volatile int *>(abort)) return true;
    const long long now = ps_globaltimer();
    if (!*t0) *t0 = now;
    else if (now - *t0 > timeout_ns) {
        *reinterpret_cast<volatile int *>(abort) = 1;
        *reinterpret_cast<volatile int *>(err) = site;
        return true;
    }
    return false;
}
#define PS_ST_SPIN_UNTIL(ctl, cond, site) PS_ST_SPIN_UNTIL_S(ctl, cond, site, 0)
#define PS_ST_SPIN_UNTIL_S(ctl, cond, site, sleep_ns)                                                 \
    do {                                                                                              \
        long long t0_ = 0;                                                                            \
        uint32_t it_ = 0;                                                                             \
        while (!(cond)) {                                                                             \
            if ((sleep_ns) > 0) __nanosleep(sleep_ns); /* polls back off: spinning warps take issue slots and L2 bandwidth */ \
            if ((++it_ & 255u) == 0 && ps_st_spin_check((ctl).abort, (ctl).err, (ctl).timeout_ns, &t0_, (site))) break; \
        }                                                                                             \
    } while (0)

PS_D bool ps_st_mbar_try(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ps_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
PS_D bool ps_st_mbar_test(uint64_t *bar, uint32_t parity) { // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(ps_smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
PS_D void ps_st_mbar_wait(const PsStCtl &ctl, uint64_t *bar, uint32_t parity, int site) { PS_ST_SPIN_UNTIL(ctl, ps_st_mbar_try(bar, parity), site); }
// the same with a back-off between attempts: producer warps wait for a slot far more often than not, and must not take issue slots from the compute warps
PS_D void ps_st_mbar_wait_idle(const PsStCtl &ctl, uint64_t *bar, uint32_t parity, int site) { PS_ST_SPIN_UNTIL_S(ctl, ps_st_mbar_try(bar, parity), site, 64); }

PS_D unsigned ps_st_ld_acquire_gpu(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// device-wide barrier of the compute warps of all CTAs (all CTAs are resident: cooperative launch, one CTA per SM).
// `target` = n_ctas x (number of barriers passed so far in this step, this one included); the counter is reset by the
// step's finishing CTA.
PS_D void ps_st_grid_barrier(const PsStCtl &ctl, unsigned *ctr, unsigned target) {
    ps_bar_sync(2, PS_ST_CT);
    if (threadIdx.x == 0) {
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        PS_ST_SPIN_UNTIL_S(ctl, ps_st_ld_acquire_gpu(ctr) >= target, 1, 20);
        __threadfence();
    }
    ps_bar_sync(2, PS_ST_CT);
}

// ---------------------------------------------------------------------------------------------------- (value, epoch) words
PS_D void ps_st_ll_store(unsigned long long *const *peer, int n_peer, int64_t idx, uint32_t bits, uint32_t ep) {
    const unsigned long long w = ((unsigned long long)ep << 32) | (unsigned long long)bits;
    for (int p = 0; p < n_peer; p++) asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(peer[p] + idx), "l"(w) : "memory");
}
PS_D bool ps_st_ll_load1(const unsigned long long *p, uint32_t ep, uint32_t &bits) {
    unsigned long long a;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    bits = (uint32_t)a;
    return (uint32_t)(a >> 32) == ep;
}
PS_D bool ps_st_ll_load2(const unsigned long long *p, uint32_t ep, uint32_t &b0, uint32_t &b1) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    b0 = (uint32_t)a;
    b1 = (uint32_t)b;
    return (uint32_t)(a >> 32) == ep && (uint32_t)(b >> 32) == ep;
}
// eight elements of a 256-block (4*lane .. +3, 128 + 4*lane .. +3), polled until they carry epoch `ep`.  While the block
// is not there yet only ONE lane polls ONE 16-byte pair (with a back-off): 2000 warps re-reading their 2 KB every 100 ns
// saturated the L2 (10 TB/s of poll traffic) and slowed every other memory access of the step several-fold.
PS_D void ps_st_load8_ll(const PsStCtl &ctl, const unsigned long long *p, int lane, float e[8], uint32_t ep) {
    const unsigned long long *p0 = p + 4 * lane, *p1 = p + 128 + 4 * lane;
    uint32_t b[8];
    if (lane == 0) {
        uint32_t t0, t1;
        PS_ST_SPIN_UNTIL_S(ctl, ps_st_ll_load2(p + 254, ep, t0, t1), 2, 100); // the block's last pair
    }
    __syncwarp();
    PS_ST_SPIN_UNTIL_S(ctl, __all_sync(PS_FULL, ps_st_ll_load2(p0, ep, b[0], b[1]) & ps_st_ll_load2(p0 + 2, ep, b[2], b[3]) & ps_st_ll_load2(p1, ep, b[4], b[5]) &
                                               ps_st_ll_load2(p1 + 2, ep, b[6], b[7])), 2, 200);
#pragma unroll
    for (int t = 0; t < 8; t++) e[t] = __uint_as_float(b[t]);
}
// the value half of one word that this CTA has already seen valid (residual inputs)
PS_D float ps_st_ll_value(const unsigned long long *p) {
    unsigned long long a;
    asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    return __uint_as_float((uint32_t)a);
}

// ---------------------------------------------------------------------------------------------------- weight stream
// A mat-vec phase as both sides of the ring see it.  Octet j of this CTA (j = 0 .. n_mine-1) belongs to compute warp
// j % 16 in round j / 16; producer p serves the warps w with w % PS_ST_PROD == p, in the order (round, stage, warp), and
// both sides number the stages of a producer's ring with one running counter across all phases of the step.
struct PsStMv {
    const uint8_t *w;
    int n_oct;   // row octets (octet pairs for gate|up)
    int nb;      // super-blocks per row
    int kb;      // super-blocks per stage
    int rpt;     // matrices interleaved per octet block (gate|up: 2)
};
PS_D int ps_st_cnt(int n_mine, int p) { // octets of this CTA served by producer p
    const int rag = n_mine % PS_ST_WARPS;
    return (n_mine / PS_ST_WARPS) * PS_ST_WPP + (rag > p ? (rag - p + PS_ST_PROD - 1) / PS_ST_PROD : 0);
}

PS_D void ps_st_produce(const PsStCtl &ctl, const PsStMv mv, int p, int lane, uint32_t &slot, uint32_t &par, uint32_t &issued, volatile uint32_t *s_issued, int nsp,
                        uint8_t *ring, uint64_t *full, uint64_t *empty) {
    // executed by the WHOLE producer warp, convergently (every lane waits for the slot, one lane issues): the shape of
    // CUTLASS's TMA producer warps.  A first version that let lanes 1..31 exit and ran this loop on a lone lane raised
    // sporadic "illegal instruction" faults on the arrive.expect_tx below (B200, CUDA 12.9).
    const int G = gridDim.x, bid = blockIdx.x;
    const int o0 = (int)(((long long)bid * mv.n_oct) / G), o1 = (int)(((long long)(bid + 1) * mv.n_oct) / G);
    const int n_mine = o1 - o0, spo = mv.nb / mv.kb;
    const uint32_t stage_bytes = (uint32_t)mv.kb * mv.rpt * PS_RW_OCTET_BLOCK;
    const size_t oct_bytes = (size_t)mv.nb * mv.rpt * PS_RW_OCTET_BLOCK;
    for (int m0 = 0; m0 < n_mine; m0 += PS_ST_WARPS) {
        const int n_m = min(PS_ST_WARPS, n_mine - m0);
        for (int sb = 0; sb < spo; sb++) {
            const uint8_t *src = mv.w + (size_t)(o0 + m0 + p) * oct_bytes + (size_t)sb * stage_bytes;
            for (int w = p; w < n_m; w += PS_ST_PROD, src += oct_bytes * PS_ST_PROD) {
                // ONE lane polls, without parking in the barrier unit (threads parked in try_wait slowed every shared-memory
                // operation of the compute warps several-fold), the others wait at the warp barrier
                if (lane == 0) PS_ST_SPIN_UNTIL_S(ctl, ps_st_mbar_test(empty + slot, par ^ 1), 4, 100);
                __syncwarp();
                if (*reinterpret_cast<volatile int *>(ctl.abort)) return; // a wait gave up somewhere: stop feeding (an un-waited expect_tx would over-arrive and trap)
#if PS_ST_DEBUG
                if (lane == 0 && issued >= (uint32_t)nsp && !ps_st_mbar_try(full + slot, par ^ 1)) { // the slot's previous round has not even landed, yet its empty barrier let us through
                    int *dbg = ctl.err + 1;
                    if (atomicCAS(dbg, 0, 1) == 0) {
                        dbg[1] = blockIdx.x; dbg[2] = p; dbg[3] = (int)issued; dbg[4] = (int)slot; dbg[5] = (int)par; dbg[6] = (int)ctl.dbg_rel[slot];
                        dbg[7] = m0; dbg[8] = sb; dbg[9] = w; dbg[10] = n_mine; dbg[11] = spo; dbg[12] = (int)ctl.dbg_rel[nsp + slot];
                    }
                    *reinterpret_cast<volatile int *>(ctl.abort) = 1;
                    *reinterpret_cast<volatile int *>(ctl.err) = 9;
                }
#endif
                ++issued;
                if (lane == 0) {
                    ps_mbar_expect_tx(full + slot, stage_bytes);
                    ps_bulk_g2s(ring + (size_t)slot * PS_ST_SLOT, src, stage_bytes, full + slot);
                    *s_issued = issued; // a compute warp that sees the count knows the slot's previous round has been released
                }
                __syncwarp();
                if (++slot == (uint32_t)nsp) { slot = 0; par ^= 1; }
            }
        }
    }
}

// compute side: walk this warp's octets of the phase; `epi(oct_index_in_matrix, acc[])` consumes a finished octet.
// `base` = stages this warp's producer has issued before the phase (advanced here).
struct PsStNoIdle {
    PS_D void operator()(int, int) const {}
};
struct PsStRing { // one compute warp's view of its producer's ring
    const volatile uint32_t *s_issued;
    uint8_t *ring;
    uint64_t *full, *empty;
    const uint4 *s_qa;
    const uint2 *s_meta;
    int nsp;
};
// all stages of ONE row octet (stage indices g, g + n_lw, ...), out of line: one copy of the block loop per RPT in the
// whole kernel.  Returns the cycles spent waiting for weight stages (trace only).
template <int RPT>
__device__ __noinline__ long long ps_st_octet(int *abort, int *err, long long timeout_ns, const PsStRing rg, uint32_t g, int n_lw, int spo, int kb, int lane, bool timed,
                                              PsRwAcc *out) {
    const PsStCtl ctl{abort, err, timeout_ns, nullptr};
    const int r = lane >> 2, q = lane & 3, moff = ps_rw_mins_off(r, q);
    uint32_t slot = g % (uint32_t)rg.nsp, par = (g / (uint32_t)rg.nsp) & 1;
    long long wcyc = 0;
    PsRwAcc acc[RPT];
#pragma unroll
    for (int t = 0; t < RPT; t++) acc[t].a0 = acc[t].a1 = acc[t].am = 0.f;
#pragma unroll 1
    for (int sb = 0; sb < spo; sb++) {
        const long long c0_ = timed ? clock64() : 0;
        // A parity wait is only unambiguous once the stage has been ISSUED (a warp may run more than a ring revolution ahead
        // of the slowest one): first the producer's issue count, then the barrier.  One lane polls, the warp reconverges.
        if (lane == 0) {
            PS_ST_SPIN_UNTIL_S(ctl, *rg.s_issued > g, 5, 32);
            PS_ST_SPIN_UNTIL_S(ctl, ps_st_mbar_test(rg.full + slot, par), 6, 20);
        }
        __syncwarp(); // lane 0 observed the phase completion (acquire); the warp barrier orders the other lanes' reads after it
        if (timed) wcyc += clock64() - c0_;
        const uint8_t *st = rg.ring + (size_t)slot * PS_ST_SLOT;
#pragma unroll 1
        for (int b = 0; b < kb; b++) {
            const int i = sb * kb + b;
            const uint4 *qa = rg.s_qa + (size_t)i * 16;
            const uint2 meta = rg.s_meta[i * 4 + q];
#pragma unroll
            for (int t = 0; t < RPT; t++) ps_rw_block(st + (size_t)(b * RPT + t) * PS_RW_OCTET_BLOCK, r, q, moff, qa, meta, acc[t]);
        }
        __syncwarp();
        if (lane == 0) ps_mbar_arrive(rg.empty + slot);
        g += n_lw;
        slot += n_lw;
        if (slot >= (uint32_t)rg.nsp) { slot -= rg.nsp; par ^= 1; }
    }
#pragma unroll
    for (int t = 0; t < RPT; t++) out[t] = acc[t];
    return wcyc;
}

// compute side: walk this warp's octets of the phase; `epi(oct_index_in_matrix, acc[])` consumes a finished octet;
// `idle(k, n)` runs on the warps that own no octet of this phase (k = 0 .. n-1).
// `base` = stages this warp's producer has issued before the phase (advanced here).
template <int RPT, class Epi, class Idle = PsStNoIdle>
PS_D void ps_st_walk(const PsStCtl &ctl, const PsStMv mv, uint32_t &base, const PsStRing &rg, int warp, int lane, Epi epi, Idle idle = Idle(), long long *tl = nullptr) {
    const int G = gridDim.x, bid = blockIdx.x;
    const int o0 = (int)(((long long)bid * mv.n_oct) / G), o1 = (int)(((long long)(bid + 1) * mv.n_oct) / G);
    const int n_mine = o1 - o0, spo = mv.nb / mv.kb;
    const int p = warp % PS_ST_PROD, lw = warp / PS_ST_PROD;
    if (warp >= n_mine) idle(warp - n_mine, PS_ST_WARPS - n_mine);
    long long wcyc = 0, ecyc = 0;
    const long long c_begin = tl ? clock64() : 0;
#pragma unroll 1
    for (int m0 = 0, m = 0; m0 + warp < n_mine; m0 += PS_ST_WARPS, m++) {
        const int n_m = min(PS_ST_WARPS, n_mine - m0);
        const int n_lw = (n_m - p + PS_ST_PROD - 1) / PS_ST_PROD; // warps of my producer active in this round
        PsRwAcc acc[RPT];
        wcyc += ps_st_octet<RPT>(ctl.abort, ctl.err, ctl.timeout_ns, rg, base + (uint32_t)m * spo * PS_ST_WPP + lw, n_lw, spo, mv.kb, lane, tl != nullptr, acc);
        const long long ce_ = tl ? clock64() : 0;
        epi(o0 + m0 + warp, acc);
        if (tl) ecyc += clock64() - ce_;
    }
    if (tl && lane == 0 && warp < n_mine) { // trace: slowest warp's walk, the most any warp waited for weight stages / spent in epilogues (cycles)
        atomicMax(reinterpret_cast<unsigned long long *>(tl + 5), (unsigned long long)(clock64() - c_begin));
        atomicMax(reinterpret_cast<unsigned long long *>(tl + 6), (unsigned long long)wcyc);
        atomicMax(reinterpret_cast<unsigned long long *>(tl + 7), (unsigned long long)ecyc);
    }
    base += (uint32_t)ps_st_cnt(n_mine, p) * spo;
}

// ---------------------------------------------------------------------------------------------------- prologues
// Q8_K image of an exchanged fp32 vector (K elements as (value, epoch) words), optionally through RMSNorm
// (ggml.c:12667-12721 + ggml-quants.c:3799-3837): one warp per 256-block, exactly the prologue of ps_k_rw_matvec.
__device__ __noinline__ void ps_st_prologue_vec(const PsStCtl &ctl, const unsigned long long *v_ll, uint32_t ep, int K, const float *norm_w, float eps, uint4 *s_qa, uint2 *s_meta,
                             double *sh_red, int warp, int lane, long long *tl) {
    const int nb = K / 256;
    const double inv_k = (K & (K - 1)) == 0 ? 1.0 / (double)K : 0.0;
    float e0[8], wv0[8];
    const bool have0 = warp < nb;
    if (have0 && norm_w) ps_rw_load8(norm_w + warp * 256, lane, wv0);
    if (have0) ps_st_load8_ll(ctl, v_ll + (size_t)warp * 256, lane, e0, ep);
    float nscale = 1.f;
    if (norm_w) {
        double ss = 0.0;
        if (have0) {
#pragma unroll
            for (int t = 0; t < 8; t++) ss += (double)__fmul_rn(e0[t], e0[t]);
        }
#pragma unroll 1
        for (int i = warp + PS_ST_WARPS; i < nb; i += PS_ST_WARPS) {
            float e[8];
            ps_st_load8_ll(ctl, v_ll + (size_t)i * 256, lane, e, ep);
#pragma unroll
            for (int t = 0; t < 8; t++) ss += (double)__fmul_rn(e[t], e[t]);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(PS_FULL, ss, o);
        ps_bar_sync(2, PS_ST_CT); // everybody is done with the previous phase's image and sh_red
        if (lane == 0) sh_red[warp] = ss;
        ps_bar_sync(2, PS_ST_CT);
        double t = (lane < PS_ST_WARPS) ? sh_red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(PS_FULL, t, o);
        const float mean = (float)(inv_k != 0.0 ? t * inv_k : t / (double)K);
        nscale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    } else {
        ps_bar_sync(2, PS_ST_CT); // everybody is done with the previous phase's image
    }
    if (tl && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + 3), (unsigned long long)ps_globaltimer()); // inputs arrived
    if (have0) {
        if (norm_w) {
#pragma unroll
            for (int t = 0; t < 8; t++) e0[t] = __fmul_rn(e0[t], __fmul_rn(wv0[t], nscale)); // y = x * (w * scale)
        }
        ps_rw_quant_store(e0, lane, reinterpret_cast<uint32_t *>(s_qa) + (size_t)warp * 64, s_meta + warp * 4);
    }
#pragma unroll 1
    for (int i = warp + PS_ST_WARPS; i < nb; i += PS_ST_WARPS) {
        float e[8];
        ps_st_load8_ll(ctl, v_ll + (size_t)i * 256, lane, e, ep);
        if (norm_w) {
            float wv[8];
            ps_rw_load8(norm_w + i * 256, lane, wv);
#pragma unroll
            for (int t = 0; t < 8; t++) e[t] = __fmul_rn(e[t], __fmul_rn(wv[t], nscale));
        }
        ps_rw_quant_store(e, lane, reinterpret_cast<uint32_t *>(s_qa) + (size_t)i * 64, s_meta + i * 4);
    }
    ps_bar_sync(2, PS_ST_CT);
    if (tl && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + 4), (unsigned long long)ps_globaltimer()); // image ready
}

// the Q8_K image the gate|up epilogue produced (words [nb x 64 quants][nb x 8 meta] as (value, epoch) pairs) -> shared memory
__device__ __noinline__ void ps_st_prologue_img(const PsStCtl &ctl, const unsigned long long *img_ll, uint32_t ep, int K, uint32_t *s_img, int tid, long long *tl) {
    const int n_pairs = (K / 256) * 36;
    ps_bar_sync(2, PS_ST_CT); // everybody is done with the previous phase's image
    for (int t0 = (tid & ~31); t0 < n_pairs; t0 += PS_ST_CT) { // warp-uniform trip count
        const int t = t0 + (tid & 31);
        if ((tid & 31) == 0) { // while the words are not there only one lane polls one pair (see ps_st_load8_ll)
            uint32_t u0, u1;
            PS_ST_SPIN_UNTIL_S(ctl, ps_st_ll_load2(img_ll + 2 * min(t0 + 31, n_pairs - 1), ep, u0, u1), 3, 100);
        }
        __syncwarp();
        uint32_t b0 = 0, b1 = 0;
        PS_ST_SPIN_UNTIL_S(ctl, __all_sync(PS_FULL, t >= n_pairs || ps_st_ll_load2(img_ll + 2 * t, ep, b0, b1)), 3, 200);
        if (t < n_pairs) *reinterpret_cast<uint2 *>(s_img + 2 * t) = make_uint2(b0, b1);
    }
    if (tl && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + 3), (unsigned long long)ps_globaltimer());
    ps_bar_sync(2, PS_ST_CT);
    if (tl && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + 4), (unsigned long long)ps_globaltimer());
}

// ---------------------------------------------------------------------------------------------------- attention
// scores = mat_mul(k_view, q) * scale + mask (norm_attention.cpp:115-134; ggml_vec_dot_f32, ggml.c:2092-2131): every
// compute warp of every CTA takes (8 cache positions, kv head) items; arithmetic of ps_k_attn1.
template <int R2, int STEPS>
__device__ __noinline__ void ps_st_scores(const float *__restrict__ kc, const float *__restrict__ q, float *__restrict__ sc, int n_kv, int nkv, int n_ctx,
                                          float scale, int warp, int lane) {
    constexpr int hs = 32 * STEPS;
    const int kvd = hs * nkv;
    const int n_items = ((n_kv + 7) >> 3) * nkv;
    for (int it = (int)blockIdx.x * PS_ST_WARPS + warp; it < n_items; it += (int)gridDim.x * PS_ST_WARPS) {
        const int chunk = it / nkv, g = it - chunk * nkv;
        const int j0 = chunk * 8;
        float kv[8][STEPS], qv[R2][STEPS];
#pragma unroll
        for (int t = 0; t < 8; t++)
#pragma unroll
            for (int s = 0; s < STEPS; s++) kv[t][s] = (j0 + t < n_kv) ? __ldcg(kc + (size_t)(j0 + t) * kvd + g * hs + 32 * s + lane) : 0.f;
#pragma unroll
        for (int hh = 0; hh < R2; hh++)
#pragma unroll
            for (int s = 0; s < STEPS; s++) qv[hh][s] = __ldcg(q + (size_t)(g * R2 + hh) * hs + 32 * s + lane);
#pragma unroll
        for (int t = 0; t < 8; t++) {
            float sum[R2];
#pragma unroll
            for (int hh = 0; hh < R2; hh++) {
                sum[hh] = 0.f;
#pragma unroll
                for (int s = 0; s < STEPS; s++) sum[hh] = __fmaf_rn(kv[t][s], qv[hh][s], sum[hh]);
            }
            ps_f32x8_reduce_n<R2>(sum);
            if (lane < R2 && j0 + t < n_kv) {
                float v = sum[0];
#pragma unroll
                for (int hh = 1; hh < R2; hh++)
                    if (lane == hh) v = sum[hh];
                sc[(size_t)(g * R2 + lane) * n_ctx + j0 + t] = __fadd_rn(__fmul_rn(v, scale), 0.0f);
            }
        }
    }
}

// soft-max (ggml.c:14846-14940, 2814-2868) + mat_mul(v_view, kq) + permute/cont (norm_attention.cpp:133-151) for the
// (kv head, block of `dpc` output dims) items; the R2 probability rows of the group are rebuilt in shared memory, CH
// positions at a time (rows longer than CH: the maximum and the double-precision sum take their own passes over the
// scores and the exponentials are recomputed chunk by chunk - the FMA chains of the P.V product stay in position order).
template <int R2>
__device__ __noinline__ void ps_st_pv(const PsStCtl &ctl, const PsStArgs &a, const float *__restrict__ vct, int n_kv, float *s_p, double *shd, float *shf,
                                      uint32_t ep, int tid, long long *tl) {
    const long long t_in = tl ? ps_globaltimer() : 0;
#define PS_PV_PROBE(k)                                                                                                     \
    do {                                                                                                                   \
        if (tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + (k)), (unsigned long long)(ps_globaltimer() - t_in)); \
    } while (0)
    const int warp = tid >> 5, lane = tid & 31;
    const int hs = a.hs, dpc = a.dpc, CH = a.attn_chunk;
    const int n_items = a.nkv_l * (hs / dpc);
    constexpr int TPH = PS_ST_CT / R2, WPH = PS_ST_WARPS / R2; // threads / warps per head
    const int hh = tid / TPH, ht = tid % TPH;
    const int n_chunks = (n_kv + CH - 1) / CH;
    const int n8 = n_kv & ~7, np = n_kv & ~31;
    const int wpd = PS_ST_WARPS / dpc;                         // warps per output dim
    const int hpw = (R2 + wpd - 1) / wpd;                      // heads per warp
    const int dd = warp % dpc, h_lo = (warp / dpc) * hpw;
    for (int it = blockIdx.x; it < n_items; it += gridDim.x) {
        const int g = it / (hs / dpc), d = (it % (hs / dpc)) * dpc + dd;
        const float *srow = a.sc + (size_t)(g * R2 + hh) * a.n_ctx;
        float *pp = s_p + (size_t)hh * CH;
        // the V^T row of this warp: 16 loads per lane in flight before the soft-max is rebuilt (they do not depend on it)
        const float *vrow = vct + ((size_t)g * hs + d) * a.n_ctx;
        const bool pv_warp = h_lo < R2;
        const int ntail = n_kv - np;
        auto load_v = [&](float (&v)[16], int s0, int c_end) {
#pragma unroll
            for (int u = 0; u < 16; u++) v[u] = (s0 + 32 * u < c_end) ? __ldcs(vrow + s0 + 32 * u + lane) : 0.f;
        };
        float cur[16];
        if (pv_warp) load_v(cur, 0, min(min(CH, n_kv), np));
        const float vtail = (pv_warp && lane < ntail) ? __ldcs(vrow + np + lane) : 0.f;
        // ---- pass 1: row maximum (a single chunk also lands in shared memory)
        float mx = -INFINITY;
        {   // rows are 16-byte aligned (n_ctx % 4 == 0): four scores per load, several loads in flight
            const int n4 = n_kv >> 2;
#pragma unroll 4
            for (int j4 = ht; j4 < n4; j4 += TPH) {
                const float4 v = __ldcg(reinterpret_cast<const float4 *>(srow) + j4);
                if (n_chunks == 1) reinterpret_cast<float4 *>(pp)[j4] = v;
                mx = fmaxf(fmaxf(mx, fmaxf(v.x, v.y)), fmaxf(v.z, v.w));
            }
            const int j = (n4 << 2) + ht;
            if (j < n_kv) {
                const float v = __ldcg(srow + j);
                if (n_chunks == 1) pp[j] = v;
                mx = fmaxf(mx, v);
            }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PS_FULL, mx, o));
        if (lane == 0) shf[warp] = mx;
        ps_bar_sync(2, PS_ST_CT);
        mx = shf[hh * WPH];
#pragma unroll
        for (int t = 1; t < WPH; t++) mx = fmaxf(mx, shf[hh * WPH + t]);
        PS_PV_PROBE(3);
        // exponentials of chunk c into shared memory (ggml_v_expf on full 8-groups of the ROW, libm expf on its tail); returns this thread's partial sum
        auto exp_chunk = [&](int c0, int cn, bool load) -> double {
            if (load) { // c0 % 4 == 0
                const int n4 = cn >> 2;
#pragma unroll 4
                for (int j4 = ht; j4 < n4; j4 += TPH) reinterpret_cast<float4 *>(pp)[j4] = __ldcg(reinterpret_cast<const float4 *>(srow + c0) + j4);
                const int j = (n4 << 2) + ht;
                if (j < cn) pp[j] = __ldcg(srow + c0 + j);
                ps_bar_sync(2, PS_ST_CT);
            }
            double s = 0.0;
            const int g_end = (min(c0 + cn, n8) - c0) >> 3; // full 8-groups of this chunk (c0 % 8 == 0)
            int it_no = 0;
            for (int gi = ht; gi < g_end; gi += TPH) {
                const long long cc0 = tl ? clock64() : 0;
                float4 *p4 = reinterpret_cast<float4 *>(pp + gi * 8);
                const float4 xa = p4[0], xb = p4[1];
                float vv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
                for (int l = 0; l < 8; l++) vv[l] = ps_v_expf(__fadd_rn(vv[l], -mx));
                p4[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
                p4[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
                const float r0 = __fadd_rn(vv[4], vv[0]), r1 = __fadd_rn(vv[5], vv[1]), r2_ = __fadd_rn(vv[6], vv[2]), r3 = __fadd_rn(vv[7], vv[3]);
                s += (double)__fadd_rn(__fadd_rn(r0, r2_), __fadd_rn(r1, r3));
                if (tl && (tid & 31) == 0 && it_no < 2) atomicMax(reinterpret_cast<unsigned long long *>(tl + 6 + it_no), (unsigned long long)(clock64() - cc0)); // trace: 1st / 2nd trip through the same code
                it_no++;
            }
            for (int j = max(n8 - c0, 0) + ht; j < cn; j += TPH) { // scalar tail of the row: libm expf
                const float vv = ps_expf_glibc(__fadd_rn(pp[j], -mx));
                pp[j] = vv;
                s += (double)vv;
            }
            return s;
        };
        // ---- pass 2: sum of the exponentials in double
        double s = 0.0;
        for (int c = 0; c < n_chunks; c++) {
            const int c0 = c * CH, cn = min(CH, n_kv - c0);
            if (n_chunks > 1) ps_bar_sync(2, PS_ST_CT); // the previous chunk's values are no longer needed
            s += exp_chunk(c0, cn, n_chunks > 1);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(PS_FULL, s, o);
        if (lane == 0) shd[warp] = s;
        ps_bar_sync(2, PS_ST_CT);
        double sum = shd[hh * WPH];
#pragma unroll
        for (int t = 1; t < WPH; t++) sum += shd[hh * WPH + t];
        const float inv = (float)(1.0 / sum);
        PS_PV_PROBE(4);
        // ---- pass 3: P.V  (the first 16 V loads of the row were issued before the soft-max; batches are double-buffered in registers)
        float acc[R2];
#pragma unroll
        for (int h2 = 0; h2 < R2; h2++) acc[h2] = 0.f;
        int c0_last = 0;
        for (int c = 0; c < n_chunks; c++) {
            const int c0 = c * CH, cn = min(CH, n_kv - c0);
            c0_last = c0;
            if (n_chunks > 1) {
                ps_bar_sync(2, PS_ST_CT);
                exp_chunk(c0, cn, true);
                ps_bar_sync(2, PS_ST_CT); // the scaling below reads exponentials other threads wrote
            }
            for (int j = ht; j < cn; j += TPH) pp[j] = __fmul_rn(pp[j], inv);
            ps_bar_sync(2, PS_ST_CT);
            if (pv_warp) {
                const int c_end = min(c0 + cn, np);
                if (c > 0) load_v(cur, c0, c_end);
                for (int s0 = c0; s0 < c_end; s0 += 512) { // the FMA chains stay in position order
                    float nxt[16];
                    if (s0 + 512 < c_end) load_v(nxt, s0 + 512, c_end);
#pragma unroll
                    for (int u = 0; u < 16; u++)
                        if (s0 + 32 * u < c_end) {
#pragma unroll
                            for (int h2 = 0; h2 < R2; h2++)
                                if (h2 < hpw && h_lo + h2 < R2) acc[h2] = __fmaf_rn(cur[u], s_p[(size_t)(h_lo + h2) * CH + (s0 - c0) + 32 * u + lane], acc[h2]);
                        }
#pragma unroll
                    for (int u = 0; u < 16; u++) cur[u] = nxt[u];
                }
            }
        }
        PS_PV_PROBE(5);
        if (pv_warp) {
            ps_f32x8_reduce_n<R2>(acc);
            for (int t = 0; t < ntail; t++) { // leftovers: mul, then add, in order (every lane computes the same chain)
                const float v = __shfl_sync(PS_FULL, vtail, t);
#pragma unroll
                for (int h2 = 0; h2 < R2; h2++)
                    if (h2 < hpw && h_lo + h2 < R2) acc[h2] = __fadd_rn(acc[h2], __fmul_rn(v, s_p[(size_t)(h_lo + h2) * CH + (np - c0_last) + t]));
            }
            if (lane < hpw && h_lo + lane < R2) {
                float v = acc[0];
#pragma unroll
                for (int h2 = 1; h2 < R2; h2++)
                    if (lane == h2) v = acc[h2];
                const int64_t idx = (int64_t)a.rank * a.qdim_l + (int64_t)(g * R2 + h_lo + lane) * hs + d;
                ps_st_ll_store(a.peers->att, a.tp, idx, __float_as_uint(v), ep);
            }
        }
        ps_bar_sync(2, PS_ST_CT); // the probabilities are dead: the next item (or the next phase's image) may overwrite them
    }
}

template <int R2> PS_D void ps_st_scores_dispatch(const PsStArgs &a, const float *kc, int n_kv, int warp, int lane) {
    if (a.hs == 64) ps_st_scores<R2, 2>(kc, a.q, a.sc, n_kv, a.nkv_l, a.n_ctx, a.kq_scale, warp, lane);
    else ps_st_scores<R2, 4>(kc, a.q, a.sc, n_kv, a.nkv_l, a.n_ctx, a.kq_scale, warp, lane);
}

// pull the first n_kv positions of a layer's K cache and transposed V cache into L2 (126 MB) ahead of its attention phases:
// thread `t` of `n_t` chip-wide takes every n_t-th 128-byte line
PS_D void ps_st_prefetch_kv(const float *kc, const float *vct, int n_kv, int kvd, int n_ctx, int t, int n_t) {
    const int k_lines = (int)(((size_t)n_kv * kvd * 4 + 127) >> 7);
    for (int i = t; i < k_lines; i += n_t) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(kc) + ((size_t)i << 7)));
    const int row_lines = (n_kv * 4 + 127) >> 7, v_lines = kvd * row_lines;
    for (int i = t; i < v_lines; i += n_t) {
        const int row = i / row_lines, l = i - row * row_lines;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char *>(vct + (size_t)row * n_ctx) + ((size_t)l << 7)));
    }
}

// ---------------------------------------------------------------------------------------------------- timeline
PS_D void ps_st_tl_enter(long long *tl) {
    if (tl && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long *>(tl), (unsigned long long)ps_globaltimer());
}
PS_D void ps_st_tl_exit(long long *tl) {
    if (tl && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + 1), (unsigned long long)ps_globaltimer());
}

// ---------------------------------------------------------------------------------------------------- the kernel
// Dynamic shared memory: [image / soft-max scratch: img_bytes][rings: PS_ST_PROD x nsp x 4736][full / empty barriers]
template <int R2>
__global__ void __launch_bounds__(PS_ST_THREADS, 1) ps_k_step(const __grid_constant__ PsStArgs a, const int img_bytes) {
    extern __shared__ __align__(128) uint8_t ps_st_smem[];
    __shared__ double sh_red[PS_ST_WARPS];
    __shared__ float sh_f[PS_ST_WARPS];
    __shared__ int sh_i[PS_ST_WARPS];
    __shared__ int s_abort, s_last;
    __shared__ uint32_t s_issued[PS_ST_PROD];   // stages issued by each producer so far
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nsp = a.nsp;
    uint8_t *s_ring = ps_st_smem + img_bytes;
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_ring + (size_t)PS_ST_PROD * nsp * PS_ST_SLOT);
    uint64_t *s_empty = s_full + PS_ST_PROD * nsp;
#if PS_ST_DEBUG
    __shared__ uint32_t s_dbg_rel[2][2 * 64];
#endif
    PsStCtl ctl{&s_abort, a.err, a.timeout_ns, nullptr};
    if (tid == 0) {
        s_abort = 0;
        s_last = 0;
        for (int k = 0; k < PS_ST_PROD; k++) s_issued[k] = 0;
        for (int s = 0; s < PS_ST_PROD * nsp; s++) {
            ps_mbar_init(s_full + s, 1);
            ps_mbar_init(s_empty + s, 1);
        }
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    const int n_layers = a.n_layers, dim = a.dim;
    const int n_qkv = a.qdim_l + 2 * a.kvd_l;
    const bool lm_head = a.mode & PS_ST_MODE_LMHEAD;

    if (warp >= PS_ST_WARPS) {
        // =================================================================== producer warps: the weight stream of the whole step
        const int p = warp - PS_ST_WARPS;
        uint8_t *ring = s_ring + (size_t)p * nsp * PS_ST_SLOT;
        uint64_t *full = s_full + p * nsp, *empty = s_empty + p * nsp;
        uint32_t slot = 0, par = 0, issued = 0;
        volatile uint32_t *iss = s_issued + p;
#if PS_ST_DEBUG
        ctl.dbg_rel = s_dbg_rel[p];
#endif
        const int n_phases = 4 * n_layers + (lm_head ? 1 : 0);
#pragma unroll 1
        for (int k = 0; k < n_phases; k++) { // the phases in execution order: (q|k|v, o, gate|up, down) per layer, lm_head
            const int L = k >> 2, t = k & 3;
            PsStMv mv;
            if (L == n_layers) mv = PsStMv{a.w_out, (a.vocab_l + 7) / 8, dim / 256, a.kb_dim, 1};
            else if (t == 0) mv = PsStMv{a.layers[L].w_qkv, n_qkv / 8, dim / 256, a.kb_dim, 1};
            else if (t == 1) mv = PsStMv{a.layers[L].w_o, dim / 8 / a.tp, a.qdim / 256, a.kb_q, 1};
            else if (t == 2) mv = PsStMv{a.layers[L].w_gu, (a.ffn_l + 7) / 8, dim / 256, a.kb_gu, 2};
            else mv = PsStMv{a.layers[L].w_down, dim / 8 / a.tp, a.ffn / 256, a.kb_ffn, 1};
            ps_st_produce(ctl, mv, p, lane, slot, par, issued, iss, nsp, ring, full, empty);
        }
        return;
    }

    // ======================================================================= compute warps
    uint4 *s_qa = reinterpret_cast<uint4 *>(ps_st_smem);
    const int p = warp % PS_ST_PROD;
    uint8_t *ring = s_ring + (size_t)p * nsp * PS_ST_SLOT;
    uint64_t *full = s_full + p * nsp, *empty = s_empty + p * nsp;
    uint32_t base = 0;                 // stages my producer has issued before the current phase
#if PS_ST_DEBUG
    ctl.dbg_rel = s_dbg_rel[p];
#endif
    const int r = lane >> 2, q = lane & 3;
    const int G = gridDim.x;
    const uint32_t serial = *reinterpret_cast<volatile unsigned *>(a.serial);
    const uint32_t ep0 = serial * (uint32_t)(n_layers + 2) + 1; // epoch of (this step, layer L) = ep0 + L; never reused within 2^32 / (n_layers + 2) steps
    const int pos = a.pos_dev[0];
    const int n_kv = pos + 1;
    unsigned bar_target = 0;
    const int dim_l = dim / a.tp;
    long long *tl = a.tl;
#define PS_ST_TL(k) (tl ? tl + (size_t)(k) * 8 : nullptr)

    // ---- phase 0: GGMLBackend::get_embedding (ggml_wrapper.cpp:181-211, dequantize_row_q4_K) of the token in device memory:
    // 256 elements per CTA, published like the output of a down projection
    ps_st_tl_enter(PS_ST_TL(0));
    {
        const int64_t tok = a.tokens_dev[0];
        const uint8_t *row = a.w_embd + tok * (int64_t)(dim / 256) * PS_Q4_K_BYTES;
        for (int e = (int)blockIdx.x * 256 + tid; e < dim; e += G * 256) {
            if (tid < 256) {
                const uint8_t *blk = row + (e / 256) * PS_Q4_K_BYTES;
                const int rr = e % 256, j = rr / 32, el = rr % 32;
                const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
                const float mn = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 2));
                const uint8_t *scp = blk + 4;
                int sv, mv;
                if (j < 4) { sv = scp[j] & 63; mv = scp[j + 4] & 63; }
                else { sv = (scp[j + 4] & 0xF) | ((scp[j - 4] >> 6) << 4); mv = (scp[j + 4] >> 4) | ((scp[j] >> 6) << 4); }
                const uint8_t qb = blk[16 + 32 * (j / 2) + el];
                const int qv = (j & 1) ? (qb >> 4) : (qb & 0xF);
                const float v = __fmaf_rn(__fmul_rn(d, (float)sv), (float)qv, -__fmul_rn(mn, (float)mv));
                ps_st_ll_store(&a.peers->x[a.rank], 1, e, __float_as_uint(v), ep0); // every rank computes the whole row for itself
            }
        }
    }
    ps_st_tl_exit(PS_ST_TL(0));

    for (int L = 0; L < n_layers; L++) {
        const PsStLayer &ly = a.layers[L];
        const uint32_t ep = ep0 + L;
        int nb = dim / 256;
        uint2 *s_meta = reinterpret_cast<uint2 *>(ps_st_smem + dim);
        // ---------------------------------------------------------------- QKV: rmsnorm + quantise, q|k|v rows, ROPE, KV store
        long long *t_ = PS_ST_TL(1 + 6 * L);
        ps_st_tl_enter(t_);
        ps_st_prologue_vec(ctl, a.x_ll, ep, dim, ly.attn_norm, a.eps, s_qa, s_meta, sh_red, warp, lane, t_);
        ps_st_walk<1>(ctl, PsStMv{ly.w_qkv, n_qkv / 8, nb, a.kb_dim, 1}, base, PsStRing{s_issued + p, ring, full, empty, s_qa, s_meta, nsp}, warp, lane, [&](int oct, PsRwAcc *acc) {
            float res = ps_rw_row_result(acc[0]);
            const int row = oct * 8 + r;
            // segments are octet-aligned: the branch is warp-uniform
            if (row < a.qdim_l + a.kvd_l) {
                const bool is_k = row >= a.qdim_l;
                const int n = is_k ? row - a.qdim_l : row;
                const float *bias = is_k ? ly.k_bias : ly.q_bias;
                if (bias) res = __fadd_rn(res, bias[n]);
                // ggml_compute_forward_rope_f32, adjacent pairs (ggml.c:15455-15486): rows (2p, 2p+1) sit in neighbouring quads
                const float other = __shfl_xor_sync(PS_FULL, res, 4);
                const int i0 = (n % a.hs) & ~1;
                const float c = a.rope_table[(size_t)pos * a.hs + i0], sn = a.rope_table[(size_t)pos * a.hs + i0 + 1];
                const float x0 = (r & 1) ? other : res, x1 = (r & 1) ? res : other;
                const float out = (r & 1) ? __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c)) : __fadd_rn(__fmul_rn(x0, c), -__fmul_rn(x1, sn));
                if (q == 0) {
                    if (is_k) ly.kc[(size_t)pos * a.kvd_l + n] = out;
                    else a.q[n] = out;
                }
            } else {
                const int n = row - a.qdim_l - a.kvd_l;
                if (ly.v_bias) res = __fadd_rn(res, ly.v_bias[n]);
                if (q == 0) ly.vct[(size_t)n * a.n_ctx + pos] = res; // V cache is stored transposed
            }
        }, [&](int k, int) { // the warps without a row octet pull this layer's K / V cache into L2: the attention phases come next
            const int n_fix = PS_ST_WARPS - (n_qkv / 8 + G - 1) / G; // idle warps EVERY CTA has in this phase
            if (k < n_fix) ps_st_prefetch_kv(ly.kc, ly.vct, pos, a.kvd_l, a.n_ctx, ((int)blockIdx.x * n_fix + k) * 32 + lane, G * n_fix * 32);
        }, t_);
        ps_st_tl_exit(t_);
        bar_target += G;
        ps_st_grid_barrier(ctl, a.bar_ctr, bar_target);
        // ---------------------------------------------------------------- attention scores
        t_ = PS_ST_TL(2 + 6 * L);
        ps_st_tl_enter(t_);
        ps_st_scores_dispatch<R2>(a, ly.kc, n_kv, warp, lane);
        ps_st_tl_exit(t_);
        bar_target += G;
        ps_st_grid_barrier(ctl, a.bar_ctr, bar_target);
        // ---------------------------------------------------------------- soft-max + P.V -> attention output (exchanged)
        t_ = PS_ST_TL(3 + 6 * L);
        ps_st_tl_enter(t_);
        ps_st_pv<R2>(ctl, a, ly.vct, n_kv, reinterpret_cast<float *>(ps_st_smem), sh_red, sh_f, ep, tid, t_);
        ps_st_tl_exit(t_);
        // ---------------------------------------------------------------- Wo + residual: x1 = x + Wo . att
        t_ = PS_ST_TL(4 + 6 * L);
        ps_st_tl_enter(t_);
        nb = a.qdim / 256;
        s_meta = reinterpret_cast<uint2 *>(ps_st_smem + a.qdim);
        ps_st_prologue_vec(ctl, a.att_ll, ep, a.qdim, nullptr, 0.f, s_qa, s_meta, sh_red, warp, lane, t_);
        ps_st_walk<1>(ctl, PsStMv{ly.w_o, dim_l / 8, nb, a.kb_q, 1}, base, PsStRing{s_issued + p, ring, full, empty, s_qa, s_meta, nsp}, warp, lane, [&](int oct, PsRwAcc *acc) {
            float res = ps_rw_row_result(acc[0]);
            if (q == 0) {
                const int64_t n = (int64_t)a.rank * dim_l + oct * 8 + r;
                res = __fadd_rn(ps_st_ll_value(a.x_ll + n), res);
                ps_st_ll_store(a.peers->x1, a.tp, n, __float_as_uint(res), ep);
            }
        }, PsStNoIdle(), t_);
        ps_st_tl_exit(t_);
        // ---------------------------------------------------------------- gate | up + SiLU: h = silu(Wg . xn) * (Wu . xn), quantised by its producers
        t_ = PS_ST_TL(5 + 6 * L);
        ps_st_tl_enter(t_);
        nb = dim / 256;
        s_meta = reinterpret_cast<uint2 *>(ps_st_smem + dim);
        ps_st_prologue_vec(ctl, a.x1_ll, ep, dim, ly.ffn_norm, a.eps, s_qa, s_meta, sh_red, warp, lane, t_);
        {
            const int n_oct_gu = (a.ffn_l + 7) / 8;
            const int nbf = a.ffn / 256;
            ps_st_walk<2>(ctl, PsStMv{ly.w_gu, n_oct_gu, nb, a.kb_gu, 2}, base, PsStRing{s_issued + p, ring, full, empty, s_qa, s_meta, nsp}, warp, lane, [&](int oct, PsRwAcc *acc) {
                const float gv = ps_rw_row_result(acc[0]);
                const float uv = ps_rw_row_result(acc[1]);
                const int row = oct * 8 + r;
                if (q == 0 && row < a.ffn_l) a.h[row] = ps_silu_mul(gv, uv);
                // the last-arriving warp of every 256-row block quantises it (quantize_row_q8_K_ref, one warp) and publishes the image words
                const int i = oct >> 5; // 32 octets per 256-row block
                __syncwarp();
                int old = 0;
                if (lane == 0) // release: the warp's h stores above; acquire: the other warps' before we read the block
                    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(a.blk_cnt + i) : "memory");
                old = __shfl_sync(PS_FULL, old, 0);
                const int expect = min(32, n_oct_gu - i * 32);
                if (old == expect - 1) {
                    const float4 v0 = __ldcg(reinterpret_cast<const float4 *>(a.h + i * 256 + 4 * lane));
                    const float4 v1 = __ldcg(reinterpret_cast<const float4 *>(a.h + i * 256 + 128 + 4 * lane));
                    const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t words[2], bsp4;
                    float yd;
                    ps_quant_block_q8k_regs(e, lane, words, yd, bsp4);
                    const int ig = a.rank * (a.ffn_l / 256) + i; // block index in the full hidden vector
                    const int l = lane & 7, jA = lane >> 3, jB = 4 + (lane >> 3);
                    ps_st_ll_store(a.peers->hq, a.tp, (int64_t)ig * 64 + ((jA >> 1) * 4 + (l >> 1)) * 4 + (jA & 1) * 2 + (l & 1), words[0], ep);
                    ps_st_ll_store(a.peers->hq, a.tp, (int64_t)ig * 64 + ((jB >> 1) * 4 + (l >> 1)) * 4 + (jB & 1) * 2 + (l & 1), words[1], ep);
                    if (lane < 4) {
                        ps_st_ll_store(a.peers->hq, a.tp, (int64_t)nbf * 64 + (int64_t)ig * 8 + 2 * lane, __float_as_uint(yd), ep);
                        ps_st_ll_store(a.peers->hq, a.tp, (int64_t)nbf * 64 + (int64_t)ig * 8 + 2 * lane + 1, bsp4, ep);
                    }
                    if (lane == 0) a.blk_cnt[i] = 0;
                }
            }, PsStNoIdle(), t_);
        }
        ps_st_tl_exit(t_);
        // ---------------------------------------------------------------- down + residual: x = x1 + Wdown . h
        t_ = PS_ST_TL(6 + 6 * L);
        ps_st_tl_enter(t_);
        nb = a.ffn / 256;
        s_meta = reinterpret_cast<uint2 *>(ps_st_smem + a.ffn);
        ps_st_prologue_img(ctl, a.hq_ll, ep, a.ffn, reinterpret_cast<uint32_t *>(ps_st_smem), tid, t_);
        ps_st_walk<1>(ctl, PsStMv{ly.w_down, dim_l / 8, nb, a.kb_ffn, 1}, base, PsStRing{s_issued + p, ring, full, empty, s_qa, s_meta, nsp}, warp, lane, [&](int oct, PsRwAcc *acc) {
            float res = ps_rw_row_result(acc[0]);
            if (q == 0) {
                const int64_t n = (int64_t)a.rank * dim_l + oct * 8 + r;
                res = __fadd_rn(ps_st_ll_value(a.x1_ll + n), res);
                ps_st_ll_store(a.peers->x, a.tp, n, __float_as_uint(res), ep + 1);
            }
        }, PsStNoIdle(), t_);
        ps_st_tl_exit(t_);
    }

    // -------------------------------------------------------------------- lm_head (+ greedy pick, stage 1)
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    if (lm_head) {
        long long *t_ = PS_ST_TL(1 + 6 * n_layers);
        ps_st_tl_enter(t_);
        const int nb = dim / 256;
        uint2 *s_meta = reinterpret_cast<uint2 *>(ps_st_smem + dim);
        ps_st_prologue_vec(ctl, a.x_ll, ep0 + n_layers, dim, a.out_norm, a.eps, s_qa, s_meta, sh_red, warp, lane, t_);
        ps_st_walk<1>(ctl, PsStMv{a.w_out, (a.vocab_l + 7) / 8, nb, a.kb_dim, 1}, base, PsStRing{s_issued + p, ring, full, empty, s_qa, s_meta, nsp}, warp, lane, [&](int oct, PsRwAcc *acc) {
            const float res = ps_rw_row_result(acc[0]);
            const int n = oct * 8 + r;
            if (q == 0 && n < a.vocab_l) {
                a.logits[n] = res;
                if (a.tpo_logits)
                    for (int pr = 0; pr < a.tpo_logits->n; pr++) a.tpo_logits->peer_dst[pr][n] = res;
                if (res > best_v || (res == best_v && n < best_i)) { best_v = res; best_i = n; } // first maximum wins
            }
        }, PsStNoIdle(), t_);
        ps_st_tl_exit(t_);
    }
    // -------------------------------------------------------------------- finish: per-CTA partial, last CTA picks and does the step bookkeeping
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(PS_FULL, best_v, o);
        const int oi = __shfl_xor_sync(PS_FULL, best_i, o);
        if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
    }
    if (lane == 0) { sh_f[warp] = best_v; sh_i[warp] = best_i; }
    ps_bar_sync(2, PS_ST_CT);
    if (tid == 0) {
        for (int t = 1; t < PS_ST_WARPS; t++)
            if (sh_f[t] > best_v || (sh_f[t] == best_v && sh_i[t] < best_i)) { best_v = sh_f[t]; best_i = sh_i[t]; }
        a.part_val[blockIdx.x] = best_v;
        a.part_idx[blockIdx.x] = (best_i == 0x7fffffff) ? best_i : best_i + a.rank * a.vocab_l;
        if (a.tpo_logits && lm_head && !(a.mode & PS_ST_MODE_PICK)) ps_tp_signal(a.tpo_logits, G); // host-visible logits: fence + epoch flag on every rank
        __threadfence();
        s_last = (atomicAdd(a.done_ctr, 1u) == (unsigned)(G - 1));
        __threadfence();
    }
    ps_bar_sync(2, PS_ST_CT);
    if (s_last && warp == 0) {
        // ps_k_argmax_step: reduce the partial maxima (first maximum wins), then ids[*ctr] = argmax, token feedback, position / counter advance
        float bv = -INFINITY;
        int bi = 0x7fffffff;
        for (int t = lane; t < G; t += 32) {
            const float v = __ldcg(a.part_val + t);
            const int i = __ldcg(a.part_idx + t);
            if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(PS_FULL, bv, o);
            const int oi = __shfl_xor_sync(PS_FULL, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (lane == 0) {
            if ((a.mode & PS_ST_MODE_PICK) && lm_head) {
                if (a.tp > 1) { // every rank publishes its best (value, index) to every rank and reduces the tp pairs in rank order
                    const uint32_t epb = ep0 + n_layers + 1;
                    ps_st_ll_store(a.peers->best, a.tp, 2 * a.rank, __float_as_uint(bv), epb);
                    ps_st_ll_store(a.peers->best, a.tp, 2 * a.rank + 1, (uint32_t)bi, epb);
                    bv = -INFINITY;
                    bi = 0x7fffffff;
                    for (int rk = 0; rk < a.tp; rk++) {
                        uint32_t b0 = 0, b1 = 0;
                        PS_ST_SPIN_UNTIL_S(ctl, ps_st_ll_load2(a.best_ll + 2 * rk, epb, b0, b1), 7, 100);
                        const float v = __uint_as_float(b0);
                        const int i = (int)b1;
                        if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; }
                    }
                }
                if (bi == 0x7fffffff) bi = 0;
                a.ids_dev[*a.ctr_dev] = bi;
                *a.ctr_dev += 1;
                *a.tokens_dev = bi;
                *a.pos_dev += 1;
            }
            *a.bar_ctr = 0;
            *a.done_ctr = 0;
            *a.serial = serial + 1;
        }
    }
#undef PS_ST_TL
}
