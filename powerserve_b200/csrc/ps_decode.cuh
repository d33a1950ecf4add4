// ps_decode.cuh — building blocks of the fused decode path (bs = 1): PTX helpers (mbarrier, 1-D TMA, PDL), the warp-level
// Q8_K activation quantiser, the two-kernel decode attention, and the device-side token feedback (embedding gather /
// greedy pick).  The weight-streaming mat-vec itself lives in ps_rw.cuh.  Everything here is bit-identical to the
// table-op kernels in ps_kernels.cuh (tests/test_gpu_decode.py proves it op by op and end to end).
#pragma once
#include "ps_kernels.cuh"

enum { PS_EPI_STORE = 0, PS_EPI_RESIDUAL = 1, PS_EPI_SILU = 2 };

// ---------------------------------------------------------------------------------------------------- PTX helpers
PS_D uint32_t ps_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
PS_D void ps_mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ps_smem_u32(bar)), "r"(count));
}
PS_D void ps_mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ps_smem_u32(bar)), "r"(bytes) : "memory");
}
PS_D void ps_mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ps_smem_u32(bar)) : "memory");
}
PS_D void ps_mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(ps_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D TMA: global -> shared, completion counted in bytes on an mbarrier
PS_D void ps_bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ps_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(ps_smem_u32(bar))
                 : "memory");
}
PS_D void ps_fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PS_D void ps_grid_dep_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
PS_D void ps_grid_dep_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
PS_D void ps_bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
PS_D long long ps_globaltimer() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
// timeline probe (option "trace"): slot[0] = min over CTAs of the start stamp, [1] = max of the end stamp,
// [2] = min of "dependencies resolved", [3] = max of "prologue done"; globaltimer ticks (ns, ~0.25 us granularity)
PS_D void ps_tl_min(long long *slot, int k) {
    if (slot && threadIdx.x == 0) atomicMin(reinterpret_cast<unsigned long long *>(slot + k), (unsigned long long)ps_globaltimer());
}
PS_D void ps_tl_max(long long *slot, int k) {
    if (slot && threadIdx.x == 0) atomicMax(reinterpret_cast<unsigned long long *>(slot + k), (unsigned long long)ps_globaltimer());
}
PS_D int ps_dp4a_us(uint32_t a, int b, int c) { // unsigned bytes of a  x  signed bytes of b
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---------------------------------------------------------------------------------------------------- tensor parallel
// Fused compute + all-gather over NVLink peer memory (one process per GPU, buffers mapped with CUDA IPC).  A producer
// kernel stores its output rows straight into EVERY rank's copy of the gathered vector; when its last CTA is done it
// publishes a monotonically increasing epoch into flags[slot][my_rank] on every rank (release, system scope).  The
// consumer kernel of the same phase waits (acquire, system scope) until all ranks' flags have reached the epoch its own
// rank published in that phase.  Gathers are in place: the barrier chain of the decode step makes every buffer's next
// write happen after all of its readers are done ("tensor parallelism").
#define PS_TP_MAX 8
struct PsTpOut {
    unsigned long long *peer_ll[PS_TP_MAX]; // in-band-flag exchange (else null): (value, epoch) words on every rank, already offset
    float *peer_dst[PS_TP_MAX];      // where this kernel's output goes on every rank (own rank included), already offset
    int *peer_idx[PS_TP_MAX];        // lm_head partial arg-max only: the index array next to the value array
    uint32_t *peer_flag[PS_TP_MAX];  // &flags[slot][my_rank] on every rank
    uint32_t *epoch;                 // local: epoch counter of the slot (single writer: the last CTA)
    int *done;                       // local: CTA arrival counter (rest state 0)
    int n;                           // ranks; 0 = not tensor parallel
};
struct PsTpIn {
    const uint32_t *flags;           // local flags[slot][0..n)
    const uint32_t *epoch;           // local epoch counter of the slot
    int *err;                        // set to 1 if the wait gave up (a peer died): bounded spin, never a hang
    int n;
};
PS_D uint32_t ps_ld_acquire_sys(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
PS_D void ps_st_release_sys(uint32_t *p, uint32_t v) { asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// consumer: ONE thread of the CTA calls this after griddepcontrol.wait and before the CTA touches the gathered vector
PS_D void ps_tp_wait(const PsTpIn *inp) {
    if (!inp) return;
    const PsTpIn in = *inp;
    if (in.n == 0) return;
    const uint32_t e = *reinterpret_cast<const volatile uint32_t *>(in.epoch);
    for (int s = 0; s < in.n; s++) {
        long long spins = 0;
        while ((int32_t)(ps_ld_acquire_sys(in.flags + s) - e) < 0) {
            if (++spins > (1ll << 26)) { *in.err = 1; break; }   // ~seconds; results are garbage but the GPU is not hung
        }
    }
}
// producer: every thread's stores are done (CTA-wide barrier before the call); ONE thread of the CTA calls this
PS_D void ps_tp_signal(const PsTpOut *outp, int n_ctas) {
    if (!outp) return;
    const PsTpOut &out = *outp;
    if (out.n == 0) return;
    __threadfence_system();
    if (atomicAdd(out.done, 1) == n_ctas - 1) {
        *out.done = 0;
        const uint32_t e = *out.epoch + 1;
        *out.epoch = e;
        __threadfence_system();
        for (int p = 0; p < out.n; p++) ps_st_release_sys(out.peer_flag[p], e);
    }
}

// ---- in-band flags (the per-layer exchanges): every peer store is ONE naturally aligned 64-bit word {value bits, epoch},
// single-copy atomic, so the producer needs no fence, no arrival flag and no system-scope release - the consumer polls
// the very words it is about to read until they carry the epoch of this phase.  The epoch counter of the slot is local
// and advances once per kernel instance (last CTA, plain gpu-scope bookkeeping), in lock step on every rank.
PS_D uint32_t ps_tp_ll_epoch(const PsTpOut *outp) { return *reinterpret_cast<const volatile uint32_t *>(outp->epoch) + 1; }
PS_D void ps_tp_ll_store(const PsTpOut *outp, int64_t idx, float v, uint32_t e) {
    const unsigned long long w = ((unsigned long long)e << 32) | (unsigned long long)__float_as_uint(v);
    for (int p = 0; p < outp->n; p++) asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(outp->peer_ll[p] + idx), "l"(w) : "memory");
}
// producer: ONE thread of every CTA, after the CTA's stores were issued (no ordering needed: the data carries the flag)
PS_D void ps_tp_ll_done(const PsTpOut *outp, int n_ctas, uint32_t e) {
    if (atomicAdd(outp->done, 1) == n_ctas - 1) {
        *outp->done = 0;
        *outp->epoch = e; // read by the consumer kernel after its grid-dependency wait
    }
}
// consumer: two adjacent words (16 bytes, each half single-copy atomic); returns false while either is stale
PS_D bool ps_tp_ll_load2(const unsigned long long *p, uint32_t e, float &v0, float &v1) {
    unsigned long long a, b;
    asm volatile("ld.relaxed.sys.global.v2.b64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
    v0 = __uint_as_float((uint32_t)a);
    v1 = __uint_as_float((uint32_t)b);
    return (uint32_t)(a >> 32) == e && (uint32_t)(b >> 32) == e;
}
#define PS_TP_LL_SPINS (1 << 25) // bounded (~a minute; ranks may enter their first step seconds apart): a dead peer yields garbage + the tp_error counter, never a hung GPU

// stand-alone consumer wait (before a device-to-host copy of a gathered vector)
__global__ void ps_k_tp_wait(const PsTpIn *tpi) {
    ps_grid_dep_wait();
    if (threadIdx.x == 0) ps_tp_wait(tpi);
}

// ---------------------------------------------------------------------------------------------------- prologue
// quantize_row_q8_K_ref (ggml-quants.c:3799-3837) of one 256-block by one warp; e[0..3] = elements 4*lane..+3,
// e[4..7] = elements 128+4*lane..+3.  Writes the natural-order words (word w = elements 4w..4w+3), d and the four
// int16 pairs of sub-block sums.
// Register form: words[0] = natural word `lane` (elements 4*lane..), words[1] = natural word 32 + lane; `d` is the block
// scale (valid in every lane); `bsp` is the int16 pair (s_{2k}, s_{2k+1}) for k = lane (valid in lanes 0..3).
PS_D void ps_quant_block_q8k_regs(const float e[8], int lane, uint32_t words[2], float &d_out, uint32_t &bsp_out) {
    const int idx0 = 4 * lane, idx1 = 128 + 4 * lane;
    float amax = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++) amax = fmaxf(amax, fabsf(e[t]));
    // warp-wide reductions with redux.sync (one instruction each): |x| orders like its bit pattern
    amax = __uint_as_float(__reduce_max_sync(PS_FULL, __float_as_uint(amax)));
    // `if (ax > amax) { amax = ax; max = x[j]; }` keeps the FIRST element that attains the maximum magnitude
    int first = 1 << 20;
#pragma unroll
    for (int t = 7; t >= 0; t--)
        if (fabsf(e[t]) == amax) first = (t < 4 ? idx0 + t : idx1 + t - 4);
    const int wfirst = __reduce_min_sync(PS_FULL, first);
    float mx = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++)
        if ((t < 4 ? idx0 + t : idx1 + t - 4) == wfirst) mx = e[t];
    mx = __uint_as_float(__reduce_or_sync(PS_FULL, (first == wfirst) ? __float_as_uint(mx) : 0u)); // exactly one owner
    if (amax == 0.f) { // warp-uniform
        words[0] = words[1] = 0;
        d_out = 0.f;
        bsp_out = 0;
        return;
    }
    const float iscale = __fdiv_rn(-127.f, mx);
    int q[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const float val = __fadd_rn(__fmul_rn(iscale, e[t]), 12582912.f); // nearest_int, ggml-quants.c:1653-1658
        q[t] = min(127, (int)(ps_f2u(val) & 0x007fffffu) - 0x00400000);
    }
    words[0] = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    words[1] = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    // sub-block sums: lanes 8j..8j+7 hold sub-blocks j (first word) and 4 + j (second word)
    int s0 = q[0] + q[1] + q[2] + q[3], s1 = q[4] + q[5] + q[6] + q[7];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const int t0 = __shfl_xor_sync(PS_FULL, s0, o), t1 = __shfl_xor_sync(PS_FULL, s1, o);
        s0 += t0;
        s1 += t1;
    }
    // lane k < 4 packs (s_{2k}, s_{2k+1}): k = 0,1 read the first-word sums of groups 2k, 2k+1; k = 2,3 the second-word sums
    const int src = 8 * ((2 * lane) & 3);
    const int a0 = __shfl_sync(PS_FULL, s0, src), a1 = __shfl_sync(PS_FULL, s0, src + 8);
    const int b0 = __shfl_sync(PS_FULL, s1, src), b1 = __shfl_sync(PS_FULL, s1, src + 8);
    const int lo = (lane < 2) ? a0 : b0, hi = (lane < 2) ? a1 : b1;
    bsp_out = ((uint32_t)lo & 0xffffu) | ((uint32_t)hi << 16);
    d_out = __fdiv_rn(1.f, iscale);
}

PS_D void ps_quant_block_q8k_warp(const float e[8], int lane, uint32_t *qs_words, float *d_out, uint32_t *bsp_out) {
    uint32_t words[2], bsp;
    float d;
    ps_quant_block_q8k_regs(e, lane, words, d, bsp);
    qs_words[lane] = words[0];
    qs_words[32 + lane] = words[1];
    if (lane == 0) *d_out = d;
    if (lane < 4) bsp_out[lane] = bsp;
}

// ====================================================================================================================
// Decode attention (bs = 1), two kernels, positions read from device memory so the step can be replayed as a graph.
// ROPE(q), ROPE(k) and the two KV-cache COPY ops are fused into the QKV mat-vec epilogue (ps_rw.cuh), so by the time
// these run q is rotated and row `pos` of the K cache / column `pos` of the transposed V cache are in place.
// ====================================================================================================================
// the five butterfly steps of GGML_F32x8_REDUCE (see ps_f32x8_reduce) on N independent values, interleaved so the
// shuffle latencies overlap
template <int N> PS_D void ps_f32x8_reduce_n(float (&v)[N]) {
    const int steps[5] = {16, 8, 4, 1, 2};
#pragma unroll
    for (int k = 0; k < 5; k++) {
        float o[N];
#pragma unroll
        for (int t = 0; t < N; t++) o[t] = __shfl_xor_sync(PS_FULL, v[t], steps[k]);
#pragma unroll
        for (int t = 0; t < N; t++) v[t] = __fadd_rn(v[t], o[t]);
    }
}

// ATTN1 = mat_mul(k_view, q) + scale + mask                                   (norm_attention.cpp:115-134, first half)
//   wp[h][j] = fl(fl(dot_f32(K[j], q_h) * scale) + 0.0f)   written to `sc` ({n_ctx} per head), j < n_kv = pos + 1.
// Persistent grid; a work item is (32-position chunk, kv head); a warp scores 8 cache rows against the R2 query heads
// of the group, with all K loads of the item in flight before anything is consumed.
// ST = head size / 32 when that is 1, 2 or 4 (no run-time predicates in the unrolled loops); ST = 8: any head size up
// to 256.  The 8 x R2 dot products of a warp are reduced as ONE reduce-scatter over the 32 lanes: butterfly stage 16 / 8 /
// 4 halves the cache rows a lane keeps, stages 1 / 2 halve the heads, and a + b == b + a makes the value a lane ends up
// with the one the full GGML_F32x8_REDUCE butterfly would leave there - 8 R2 - 1 (+ plain stages) shuffles instead of
// 40 R2.  Register slot (t, hh) of a lane holds cache row t ^ (lane >> 2) and head hh ^ ch (ch from lane bits 0, 1), so the
// half a lane keeps is always "the lower slots" and no selects are needed.
template <int R2, int ST>
__global__ void __launch_bounds__(128) ps_k_attn1(float *__restrict__ sc, const float *__restrict__ kc, const float *__restrict__ q,
                                                  const int32_t *__restrict__ pos_dev, int hs, int n_kv_heads, int n_ctx, float scale, long long *tl, int r2) {
    __shared__ float s_q[R2][256];
    constexpr int HB = (R2 == 1) ? 0 : (R2 == 2) ? 1 : (R2 == 4) ? 2 : 3;
    static_assert(R2 == 1 || R2 == 2 || R2 == 4 || R2 == 8, "R2 is a power of two");
    ps_tl_min(tl, 0);
    // Everything this kernel reads except the query vector and cache row `pos` is older than the kernel before the
    // previous one (pos_dev: the last kernel of the previous step; cache rows < pos: earlier steps / the prefill), and a
    // grid can only start once its predecessor is past ITS dependency wait - so those loads are issued BEFORE the wait
    // and their DRAM latency overlaps the tail of the q|k|v kernel.  Row `pos` (written by that kernel) follows the wait.
    const int pos = pos_dev[0];
    const int64_t n_kv = (int64_t)pos + 1;
    const int n_items = (int)((n_kv + 31) / 32) * n_kv_heads;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, steps = (ST == 8) ? hs / 32 : ST;
    const int kvd = hs * n_kv_heads;
    const int ct = lane >> 2;
    const int ch = (HB >= 1 ? (lane & 1) << (HB - 1) : 0) | (HB >= 2 ? ((lane >> 1) & 1) << (HB - 2) : 0);
    bool waited = false;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int chunk = item / n_kv_heads, g = item % n_kv_heads;
        const int64_t j0 = (int64_t)chunk * 32 + warp * 8;
        float kv[8][ST];
#pragma unroll
        for (int t = 0; t < 8; t++)
#pragma unroll
            for (int s = 0; s < ST; s++) {
                kv[t][s] = 0.f;
                if (s < steps && j0 + (t ^ ct) < pos) kv[t][s] = kc[(j0 + (t ^ ct)) * (int64_t)kvd + g * hs + 32 * s + lane];
            }
        if (!waited) {
            ps_grid_dep_wait();
            ps_grid_dep_launch();
            ps_tl_min(tl, 2);
            waited = true;
        }
        if (j0 <= pos && pos < j0 + 8) { // warp-uniform: the row the q|k|v kernel has just written
#pragma unroll
            for (int t = 0; t < 8; t++)
#pragma unroll
                for (int s = 0; s < ST; s++)
                    if (s < steps && j0 + (t ^ ct) == pos) kv[t][s] = kc[(j0 + (t ^ ct)) * (int64_t)kvd + g * hs + 32 * s + lane];
        }
        __syncthreads(); // the previous item's queries are no longer needed
        for (int idx = tid; idx < R2 * hs; idx += 128) s_q[idx / hs][idx % hs] = (idx < r2 * hs) ? q[(int64_t)g * r2 * hs + idx] : 0.f; // r2 <= R2 query heads per kv head
        __syncthreads();
        float sum[8][R2];
#pragma unroll
        for (int hh = 0; hh < R2; hh++) {
            float qv[ST];
#pragma unroll
            for (int s = 0; s < ST; s++) qv[s] = (s < steps) ? s_q[hh ^ ch][32 * s + lane] : 0.f;
#pragma unroll
            for (int t = 0; t < 8; t++) {
                sum[t][hh] = 0.f;
#pragma unroll
                for (int s = 0; s < ST; s++)
                    if (s < steps) sum[t][hh] = __fmaf_rn(kv[t][s], qv[s], sum[t][hh]); // ggml_vec_dot_f32 lane chain (ggml.c:2092-2131)
            }
        }
#pragma unroll
        for (int t = 0; t < 4; t++)
#pragma unroll
            for (int hh = 0; hh < R2; hh++) sum[t][hh] = __fadd_rn(sum[t][hh], __shfl_xor_sync(PS_FULL, sum[t + 4][hh], 16));
#pragma unroll
        for (int t = 0; t < 2; t++)
#pragma unroll
            for (int hh = 0; hh < R2; hh++) sum[t][hh] = __fadd_rn(sum[t][hh], __shfl_xor_sync(PS_FULL, sum[t + 2][hh], 8));
#pragma unroll
        for (int hh = 0; hh < R2; hh++) sum[0][hh] = __fadd_rn(sum[0][hh], __shfl_xor_sync(PS_FULL, sum[1][hh], 4));
        constexpr int N4 = (HB >= 1) ? R2 / 2 : 1; // values a lane keeps after stage xor 1
#pragma unroll
        for (int hh = 0; hh < N4; hh++) sum[0][hh] = __fadd_rn(sum[0][hh], __shfl_xor_sync(PS_FULL, sum[0][HB >= 1 ? hh + N4 : hh], 1));
        constexpr int N5 = (HB >= 2) ? R2 / 4 : 1; // ... and after stage xor 2
#pragma unroll
        for (int hh = 0; hh < N5; hh++) sum[0][hh] = __fadd_rn(sum[0][hh], __shfl_xor_sync(PS_FULL, sum[0][HB >= 2 ? hh + N5 : hh], 2));
        // slot hh < N5 now holds cache row j0 + ct, head hh ^ ch; lanes that differ only in unused code bits hold copies
        const bool owner = (HB >= 2) || (HB == 1 ? (lane & 2) == 0 : (lane & 3) == 0);
#pragma unroll
        for (int hh = 0; hh < N5; hh++) {
            const int head = hh ^ ch;
            if (owner && head < r2 && j0 + ct < n_kv) sc[(int64_t)(g * r2 + head) * n_ctx + j0 + ct] = __fadd_rn(__fmul_rn(sum[0][hh], scale), 0.0f);
        }
    }
    if (!waited) { // no item for this CTA: it still takes part in the dependency chain
        ps_grid_dep_wait();
        ps_grid_dep_launch();
        ps_tl_min(tl, 2);
    }
    ps_tl_max(tl, 1);
}

// ATTN2 = softmax_ext + mat_mul(v_view, kq) + permute/cont                    (norm_attention.cpp:133-151)
// One CTA (8 warps) per (8 output dims d, kv head).  Right after the dependency wait one thread asks the TMA for the R2
// score rows of the group and (v_smem) the CTA's eight V^T rows, so the whole kernel pays ONE memory latency.  The
// soft-max rows are rebuilt in shared memory by all 256 threads (256 / R2 threads per head: max, ggml_v_expf on 8-groups
// / expf tail, double sum, scale: ggml.c:14846-14940, 2814-2868); then each warp walks one V^T row once for all R2 heads
// (ggml_vec_dot_f32 lane order, leftovers in order).  v_smem = 0 (contexts too long for shared memory) streams the V^T
// row from global memory with 32 loads in flight per lane instead.
#define PS_A2_THREADS 512 // 16 warps rebuild the soft-max rows (four per scheduler hide the exp / shared-memory latencies: 7.3 -> 6.8 us per layer at ctx 2048 vs 8 warps), then eight of them walk one V^T row each (two rows per warp on four warps measured the same)
template <int R2>
__global__ void __launch_bounds__(PS_A2_THREADS) ps_k_attn2(float *__restrict__ att, const float *__restrict__ sc, const float *__restrict__ vct,
                                                  const int32_t *__restrict__ pos_dev, int hs, int n_ctx, long long *tl, const PsTpOut *tpo,
                                                  int v_smem, int r2) {
    extern __shared__ __align__(128) float s_p[]; // [R2][stride] probabilities, then (v_smem) [8][stride] V^T rows
    __shared__ double shd[PS_A2_THREADS / 32];
    __shared__ float shf[PS_A2_THREADS / 32];
    __shared__ __align__(8) uint64_t bar_s, bar_v;
    const int g = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        ps_mbar_init(&bar_s, 1);
        ps_mbar_init(&bar_v, 1);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    ps_tl_min(tl, 0);
    // The position (last kernel of the previous step) and the V^T rows (column `pos` comes from the q|k|v kernel, which was
    // complete before the scores kernel let this grid start) are older than this kernel's predecessor: the V^T copies are
    // requested BEFORE the dependency wait; only the score rows follow it.
    const int64_t n_kv_all = (int64_t)pos_dev[0] + 1, n_kv = n_kv_all;
    constexpr int TPH = PS_A2_THREADS / R2, WPH = TPH / 32;          // threads / warps per head
    const int hh = tid / TPH, ht = tid % TPH;
    const int64_t stride = (n_kv + 31) & ~(int64_t)31;
    const int64_t n8_all = n_kv & ~(int64_t)7;
    float *s_v = s_p + R2 * stride;
    const int n_rows = min(8, hs - (int)blockIdx.x * 8);
    // rows are 16-byte aligned (n_ctx % 4 == 0); a copy may run up to 3 floats past n_kv, still inside its row
    const uint32_t bytes = (uint32_t)(((n_kv + 3) & ~(int64_t)3) * 4);
    if (tid == 0 && v_smem) {
        ps_mbar_expect_tx(&bar_v, bytes * n_rows);
        for (int w = 0; w < n_rows; w++) ps_bulk_g2s(s_v + w * stride, vct + ((int64_t)g * hs + blockIdx.x * 8 + w) * n_ctx, bytes, &bar_v);
    }
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    ps_tl_min(tl, 2);
    const uint32_t ll_epoch = (tpo && tpo->peer_ll[0]) ? ps_tp_ll_epoch(tpo) : 0;
    const long long t_dep = (tl && tid == 0) ? ps_globaltimer() : 0;
#define PS_A2_PROBE(k)                                                                                                   \
    do {                                                                                                                 \
        if (tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + (k)), (unsigned long long)(ps_globaltimer() - t_dep)); \
    } while (0)
    if (tid == 0) {
        ps_mbar_expect_tx(&bar_s, bytes * r2);
#pragma unroll
        for (int h2 = 0; h2 < R2; h2++)
            if (h2 < r2) ps_bulk_g2s(s_p + h2 * stride, sc + (int64_t)(g * r2 + h2) * n_ctx, bytes, &bar_s);
    }
    __syncthreads(); // the barrier words are initialised for everybody
    ps_mbar_wait(&bar_s, 0);
    {
        float *pp = s_p + hh * stride;
        // 32-bit loop counters and 16-byte shared-memory accesses: this kernel is instruction-bound (two warps per scheduler)
        const int n_kv = (hh < r2) ? (int)n_kv_all : 0, n8 = (hh < r2) ? (int)n8_all : 0; // template heads beyond r2 (r2 <= R2 query heads per kv head) idle
        const int n4 = n_kv & ~3;
        float mx = -INFINITY;
        for (int j = 4 * ht; j < n4; j += 4 * TPH) {
            const float4 x = *reinterpret_cast<const float4 *>(pp + j);
            mx = fmaxf(fmaxf(mx, fmaxf(x.x, x.y)), fmaxf(x.z, x.w));
        }
        for (int j = n4 + ht; j < n_kv; j += TPH) mx = fmaxf(mx, pp[j]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PS_FULL, mx, o));
        if (lane == 0) shf[warp] = mx;
        __syncthreads();
        PS_A2_PROBE(4);
        mx = shf[hh * WPH];
#pragma unroll
        for (int t = 1; t < WPH; t++) mx = fmaxf(mx, shf[hh * WPH + t]);
        double s = 0.0;
        for (int gi = ht; gi < n8 / 8; gi += TPH) {
            float4 *p4 = reinterpret_cast<float4 *>(pp + gi * 8);
            const float4 xa = p4[0], xb = p4[1];
            float vv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int l = 0; l < 8; l++) vv[l] = ps_v_expf(__fadd_rn(vv[l], -mx));
            p4[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            p4[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
            const float r0 = __fadd_rn(vv[4], vv[0]), r1 = __fadd_rn(vv[5], vv[1]), r2_ = __fadd_rn(vv[6], vv[2]), r3 = __fadd_rn(vv[7], vv[3]);
            s += (double)__fadd_rn(__fadd_rn(r0, r2_), __fadd_rn(r1, r3));
        }
        for (int j = n8 + ht; j < n_kv; j += TPH) { // scalar tail: libm expf
            const float vv = ps_expf_glibc(__fadd_rn(pp[j], -mx));
            pp[j] = vv;
            s += (double)vv;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(PS_FULL, s, o);
        if (lane == 0) shd[warp] = s;
        __syncthreads();
        PS_A2_PROBE(5);
        double sum = shd[hh * WPH];
#pragma unroll
        for (int t = 1; t < WPH; t++) sum += shd[hh * WPH + t];
        const float inv = (float)(1.0 / sum);
        for (int j = 4 * ht; j < n4; j += 4 * TPH) {
            float4 x = *reinterpret_cast<const float4 *>(pp + j);
            x.x = __fmul_rn(x.x, inv); x.y = __fmul_rn(x.y, inv); x.z = __fmul_rn(x.z, inv); x.w = __fmul_rn(x.w, inv);
            *reinterpret_cast<float4 *>(pp + j) = x;
        }
        for (int j = n4 + ht; j < n_kv; j += TPH) pp[j] = __fmul_rn(pp[j], inv);
    }
    __syncthreads();
    ps_tl_max(tl, 3);
    const int d = blockIdx.x * 8 + warp;
    if (warp < 8 && d < hs) { // warps 8.. only help with the soft-max
        const float *vrow = vct + ((int64_t)g * hs + d) * n_ctx;
        const int64_t np = n_kv & ~(int64_t)31;
        float sum[R2];
#pragma unroll
        for (int h2 = 0; h2 < R2; h2++) sum[h2] = 0.f;
        const int ntail = (int)(n_kv - np);
        float vtail;
        if (v_smem) {
            ps_mbar_wait(&bar_v, 0);
            PS_A2_PROBE(6);
            const float *vs = s_v + warp * stride;
            vtail = (lane < ntail) ? vs[np + lane] : 0.f;
#pragma unroll 8
            for (int s0 = 0; s0 < (int)np; s0 += 32) { // the FMA chains stay in position order
                const float v = vs[s0 + lane];
#pragma unroll
                for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fmaf_rn(v, s_p[h2 * stride + s0 + lane], sum[h2]);
            }
        } else {
            vtail = (lane < ntail) ? vrow[np + lane] : 0.f;
            for (int64_t s0 = 0; s0 < np; s0 += 1024) { // 32 independent V loads in flight per lane
                float vv[32];
#pragma unroll
                for (int u = 0; u < 32; u++) vv[u] = (s0 + 32 * u < np) ? __ldcs(vrow + s0 + 32 * u + lane) : 0.f;
#pragma unroll
                for (int u = 0; u < 32; u++)
                    if (s0 + 32 * u < np) {
#pragma unroll
                        for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fmaf_rn(vv[u], s_p[h2 * stride + s0 + 32 * u + lane], sum[h2]);
                    }
            }
        }
        PS_A2_PROBE(7);
        ps_f32x8_reduce_n<R2>(sum);
        for (int t = 0; t < ntail; t++) { // leftovers: mul, then add, in order (every lane computes the same chain)
            const float v = __shfl_sync(PS_FULL, vtail, t);
#pragma unroll
            for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fadd_rn(sum[h2], __fmul_rn(v, s_p[h2 * stride + np + t]));
        }
        if (lane < r2) {
            float v = sum[0];
#pragma unroll
            for (int h2 = 1; h2 < R2; h2++)
                if (lane == h2) v = sum[h2];
            att[(int64_t)(g * r2 + lane) * hs + d] = v;
            if (tpo) { // all-gather by peer stores
                if (tpo->peer_ll[0]) ps_tp_ll_store(tpo, (int64_t)(g * r2 + lane) * hs + d, v, ll_epoch);
                else
                    for (int p = 0; p < tpo->n; p++) tpo->peer_dst[p][(int64_t)(g * r2 + lane) * hs + d] = v;
            }
        }
    }
    if (tpo) {
        if (tpo->peer_ll[0]) {
            if (tid == 0) ps_tp_ll_done(tpo, (int)(gridDim.x * gridDim.y), ll_epoch);
        } else {
            __syncthreads();
            if (tid == 0) ps_tp_signal(tpo, (int)(gridDim.x * gridDim.y));
        }
    }
    ps_tl_max(tl, 1);
}

// ====================================================================================================================
// Decode attention in ONE kernel without clusters ("group-synchronised"): the hs / 8 CTAs of a kv head form a GROUP that
// meets twice at a counter in global memory.  CTA (c, g) first scores ITS chunk of the cache positions against the R2
// query heads (K rows requested before the dependency wait, 8 rows per warp in registers, as ps_k_attn1), publishes the
// chunk maxima, meets the group (1), turns its chunk into exponentials with the ROW maximum - ggml_v_expf on the row's full
// 8-groups, libm expf on its tail - publishes them and the chunk sums, meets the group (2), gathers the R2 rows of
// exponentials, scales them and walks its eight V^T rows (staged by TMA right after the dependency wait) as ps_k_attn2 does.
// One kernel boundary per layer disappears and the soft-max row is built ONCE per kv head instead of once per CTA (16x for
// head size 128).  All CTAs of a group are co-resident by construction (grid <= SM count, one CTA per SM; the dependent
// grid starts only after every CTA of this one has started), the waits are bounded all the same (error flag, no hang).
// Arithmetic = ps_k_attn1 + ps_k_attn2: ggml_vec_dot_f32 lane chains, GGML_F32x8_REDUCE, scale, + 0.0f mask, soft-max with
// double sums (tree order), position-ordered FMA chains for P.V.  norm_attention.cpp:115-151.
// ====================================================================================================================
#define PS_AG_THREADS 512
PS_D unsigned long long ps_ag_ld_acquire(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
// every thread of the CTA calls this; returns false if the group did not arrive in time
PS_D bool ps_ag_meet(unsigned long long *ctr, unsigned long long *s_target, int step, int n_cta, int *err) {
    __syncthreads(); // the CTA's global stores of this phase are issued
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned long long old = atomicAdd(ctr, 1ull);
        // a kernel instance adds 2 * n_cta to its group's counter, and instances of one layer never overlap
        if (step == 0) *s_target = old / (2ull * n_cta) * (2ull * n_cta) + n_cta;
        else *s_target += n_cta;
        const unsigned long long target = *s_target;
        int spins = 0;
        while (ps_ag_ld_acquire(ctr) < target)
            if (++spins > (1 << 22)) { *err = 5; break; }
        __threadfence();
    }
    __syncthreads();
    return true;
}

template <int R2>
__global__ void __launch_bounds__(PS_AG_THREADS, 1) ps_k_attn_group(float *__restrict__ att, const float *__restrict__ kc, const float *__restrict__ vct,
                                                                     const float *__restrict__ q, float *__restrict__ ex, const int32_t *__restrict__ pos_dev, int hs,
                                                                     int n_kv_heads, int n_ctx, float scale, int r2, unsigned long long *__restrict__ sync_ctr,
                                                                     float *__restrict__ part_max, double *__restrict__ part_sum, int *__restrict__ err,
                                                                     int chunk_cap, long long *tl) {
    extern __shared__ __align__(128) float s_dyn[]; // [R2][stride] exponentials / probabilities, [8][stride] V^T rows, [R2][chunk_cap] chunk scores
    __shared__ float s_q[R2][256];
    __shared__ double shd[PS_AG_THREADS / 32];
    __shared__ float shf[PS_AG_THREADS / 32];
    __shared__ __align__(8) uint64_t bar_v;
    __shared__ unsigned long long s_target;
    const int c = blockIdx.x, g = blockIdx.y, NC = gridDim.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int steps = hs / 32, kvd = hs * n_kv_heads;
    if (tid == 0) {
        ps_mbar_init(&bar_v, 1);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    ps_tl_min(tl, 0);
    // ---- before the dependency wait: position and the chunk's K rows except row `pos` (see ps_k_attn1)
    const int pos = pos_dev[0];
    const int64_t n_kv = (int64_t)pos + 1, stride = (n_kv + 31) & ~(int64_t)31, n8 = n_kv & ~(int64_t)7;
    const int P = (int)((((n_kv + NC - 1) / NC) + 7) & ~(int64_t)7); // positions per chunk, a multiple of 8
    const int j_lo = c * P, j_hi = (int)min((int64_t)j_lo + P, n_kv), len = max(j_hi - j_lo, 0);
    float *s_p = s_dyn, *s_v = s_dyn + (size_t)R2 * stride, *s_sc = s_v + (size_t)8 * stride;
    const int n_rows = min(8, hs - c * 8);
    const uint32_t row_bytes = (uint32_t)(((n_kv + 3) & ~(int64_t)3) * 4);
    constexpr int RPB = (PS_AG_THREADS / 32) * 8; // cache rows per batch: 8 per warp
    float kv[8][8];
    {
        const int j0 = j_lo + warp * 8;
#pragma unroll
        for (int t = 0; t < 8; t++)
#pragma unroll
            for (int s = 0; s < 8; s++) {
                kv[t][s] = 0.f;
                if (s < steps && j0 + t < j_hi && j0 + t < pos) kv[t][s] = kc[(int64_t)(j0 + t) * kvd + g * hs + 32 * s + lane];
            }
    }
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    ps_tl_min(tl, 2);
    const long long t_dep = (tl && tid == 0) ? ps_globaltimer() : 0;
#define PS_AG_PROBE(k)                                                                                                   \
    do {                                                                                                                 \
        if (tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + (k)), (unsigned long long)(ps_globaltimer() - t_dep)); \
    } while (0)
    // the V^T rows only now: their column `pos` comes from the kernel this one waits for (they are needed last, so the copy
    // still hides behind the two score phases)
    if (tid == 0) {
        ps_mbar_expect_tx(&bar_v, row_bytes * n_rows);
        for (int w = 0; w < n_rows; w++) ps_bulk_g2s(s_v + w * stride, vct + ((int64_t)g * hs + c * 8 + w) * n_ctx, row_bytes, &bar_v);
    }
    for (int idx = tid; idx < R2 * hs; idx += PS_AG_THREADS) s_q[idx / hs][idx % hs] = (idx < r2 * hs) ? q[(int64_t)g * r2 * hs + idx] : 0.f;
    __syncthreads();
    // ---- phase 1: scores of this chunk -> s_sc[h][j - j_lo]
    for (int b0 = 0; b0 < len; b0 += RPB) {
        const int j0 = j_lo + b0 + warp * 8;
        if (b0 > 0) { // chunks longer than one batch (contexts beyond RPB * NC positions): further batches are loaded here
#pragma unroll
            for (int t = 0; t < 8; t++)
#pragma unroll
                for (int s = 0; s < 8; s++) {
                    kv[t][s] = 0.f;
                    if (s < steps && j0 + t < j_hi && j0 + t < pos) kv[t][s] = kc[(int64_t)(j0 + t) * kvd + g * hs + 32 * s + lane];
                }
        }
        if (j0 <= pos && pos < j0 + 8 && pos < j_hi) { // warp-uniform: the row the q|k|v kernel has just written
#pragma unroll
            for (int t = 0; t < 8; t++)
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (s < steps && j0 + t == pos) kv[t][s] = kc[(int64_t)(j0 + t) * kvd + g * hs + 32 * s + lane];
        }
        float qv[R2][8];
#pragma unroll
        for (int hh = 0; hh < R2; hh++)
#pragma unroll
            for (int s = 0; s < 8; s++) qv[hh][s] = (s < steps) ? s_q[hh][32 * s + lane] : 0.f;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            float sum[R2];
#pragma unroll
            for (int hh = 0; hh < R2; hh++) {
                sum[hh] = 0.f;
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (s < steps) sum[hh] = __fmaf_rn(kv[t][s], qv[hh][s], sum[hh]);
            }
            ps_f32x8_reduce_n<R2>(sum);
            if (lane < r2 && j0 + t < j_hi) {
                float v = sum[0];
#pragma unroll
                for (int hh = 1; hh < R2; hh++)
                    if (lane == hh) v = sum[hh];
                s_sc[lane * chunk_cap + (j0 + t - j_lo)] = __fadd_rn(__fmul_rn(v, scale), 0.0f);
            }
        }
    }
    __syncthreads();
    constexpr int TPH = PS_AG_THREADS / R2, WPH = TPH / 32; // threads / warps per head
    const int hh = tid / TPH, ht = tid % TPH;
    const int my_len = (hh < r2) ? len : 0;
    float *sc_row = s_sc + hh * chunk_cap;
    {
        float mx = -INFINITY;
        for (int j = ht; j < my_len; j += TPH) mx = fmaxf(mx, sc_row[j]);
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PS_FULL, mx, o));
        if (lane == 0) shf[warp] = mx;
        __syncthreads();
        if (ht == 0 && hh < r2) {
            mx = shf[hh * WPH];
#pragma unroll
            for (int t = 1; t < WPH; t++) mx = fmaxf(mx, shf[hh * WPH + t]);
            part_max[((size_t)g * NC + c) * R2 + hh] = mx;
        }
    }
    unsigned long long *ctr = sync_ctr + g;
    PS_AG_PROBE(4);
    ps_ag_meet(ctr, &s_target, 0, NC, err);
    PS_AG_PROBE(5);
    // ---- phase 2: exponentials of this chunk with the row maximum
    {
        float mx = -INFINITY;
        if (hh < r2)
            for (int cc = 0; cc < NC; cc++) mx = fmaxf(mx, __ldcg(part_max + ((size_t)g * NC + cc) * R2 + hh));
        float *erow = ex + (int64_t)(g * r2 + hh) * n_ctx + j_lo;
        double sdb = 0.0;
        const int n8l = (int)max(min((int64_t)j_hi, n8) - j_lo, (int64_t)0); // the chunk's share of the row's full 8-groups (chunks start on multiples of 8)
        for (int gi = ht; gi < (my_len ? n8l / 8 : 0); gi += TPH) {
            float4 *p4 = reinterpret_cast<float4 *>(sc_row + gi * 8);
            const float4 xa = p4[0], xb = p4[1];
            float vv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int l = 0; l < 8; l++) vv[l] = ps_v_expf(__fadd_rn(vv[l], -mx));
            *reinterpret_cast<float4 *>(erow + gi * 8) = make_float4(vv[0], vv[1], vv[2], vv[3]);
            *reinterpret_cast<float4 *>(erow + gi * 8 + 4) = make_float4(vv[4], vv[5], vv[6], vv[7]);
            const float r0 = __fadd_rn(vv[4], vv[0]), r1 = __fadd_rn(vv[5], vv[1]), r2_ = __fadd_rn(vv[6], vv[2]), r3 = __fadd_rn(vv[7], vv[3]);
            sdb += (double)__fadd_rn(__fadd_rn(r0, r2_), __fadd_rn(r1, r3));
        }
        for (int j = n8l + ht; j < my_len; j += TPH) { // the row's scalar tail (only in the chunk that holds it): libm expf
            const float vv = ps_expf_glibc(__fadd_rn(sc_row[j], -mx));
            erow[j] = vv;
            sdb += (double)vv;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) sdb += __shfl_xor_sync(PS_FULL, sdb, o);
        if (lane == 0) shd[warp] = sdb;
        __syncthreads();
        if (ht == 0 && hh < r2) {
            double sum = shd[hh * WPH];
#pragma unroll
            for (int t = 1; t < WPH; t++) sum += shd[hh * WPH + t];
            part_sum[((size_t)g * NC + c) * R2 + hh] = sum;
        }
    }
    PS_AG_PROBE(6);
    ps_ag_meet(ctr, &s_target, 1, NC, err);
    PS_AG_PROBE(7);
    // ---- phase 3: gather the R2 rows of exponentials, scale, P.V for this CTA's eight output dims
    {
        const int64_t n4 = (n_kv + 3) >> 2;
        for (int h2 = 0; h2 < r2; h2++) {
            const float4 *src = reinterpret_cast<const float4 *>(ex + (int64_t)(g * r2 + h2) * n_ctx);
            float4 *dst = reinterpret_cast<float4 *>(s_p + (size_t)h2 * stride);
            for (int64_t t = tid; t < n4; t += PS_AG_THREADS) dst[t] = __ldcg(src + t);
        }
        double sum = 0.0;
        if (hh < r2)
            for (int cc = 0; cc < NC; cc++) sum += __ldcg(part_sum + ((size_t)g * NC + cc) * R2 + hh);
        const float inv = (float)(1.0 / sum);
        __syncthreads();
        if (hh < r2) {
            float *pp = s_p + (size_t)hh * stride;
            for (int64_t j = ht; j < n_kv; j += TPH) pp[j] = __fmul_rn(pp[j], inv);
        }
    }
    __syncthreads();
    ps_tl_max(tl, 3);
    const int d = c * 8 + warp;
    if (warp < 8 && d < hs) {
        const int64_t np = n_kv & ~(int64_t)31;
        float sum[R2];
#pragma unroll
        for (int h2 = 0; h2 < R2; h2++) sum[h2] = 0.f;
        const int ntail = (int)(n_kv - np);
        ps_mbar_wait(&bar_v, 0);
        const float *vs = s_v + warp * stride;
        const float vtail = (lane < ntail) ? vs[np + lane] : 0.f;
#pragma unroll 8
        for (int64_t s0 = 0; s0 < np; s0 += 32) { // the FMA chains stay in position order
            const float v = vs[s0 + lane];
#pragma unroll
            for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fmaf_rn(v, s_p[h2 * stride + s0 + lane], sum[h2]);
        }
        ps_f32x8_reduce_n<R2>(sum);
        for (int t = 0; t < ntail; t++) { // leftovers: mul, then add, in order
            const float v = __shfl_sync(PS_FULL, vtail, t);
#pragma unroll
            for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fadd_rn(sum[h2], __fmul_rn(v, s_p[h2 * stride + np + t]));
        }
        if (lane < r2) {
            float v = sum[0];
#pragma unroll
            for (int h2 = 1; h2 < R2; h2++)
                if (lane == h2) v = sum[h2];
            att[(int64_t)(g * r2 + lane) * hs + d] = v;
        }
    }
    ps_tl_max(tl, 1);
}

// GGMLBackend::get_embedding for the token held in device memory (decode feedback loop)
__global__ void __launch_bounds__(256) ps_k_embed_dev(float *__restrict__ dst, const uint8_t *__restrict__ w, int type, int64_t dim,
                                                      const int32_t *__restrict__ tokens, long long *tl) {
    ps_tl_min(tl, 0);
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    // identical arithmetic to ps_k_get_embedding (one token)
    const int64_t tok = tokens[0];
    const uint8_t *row = w + tok * ps_row_bytes(type, dim);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < dim; e += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t *blk = row + (e / 256) * PS_Q4_K_BYTES;
        const int r = (int)(e % 256), j = r / 32, el = r % 32;
        const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
        const float mn = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 2));
        const uint8_t *scp = blk + 4;
        int s, m;
        if (j < 4) { s = scp[j] & 63; m = scp[j + 4] & 63; }
        else { s = (scp[j + 4] & 0xF) | ((scp[j - 4] >> 6) << 4); m = (scp[j + 4] >> 4) | ((scp[j] >> 6) << 4); }
        const uint8_t qb = blk[16 + 32 * (j / 2) + el];
        const int qv = (j & 1) ? (qb >> 4) : (qb & 0xF);
        dst[e] = __fmaf_rn(__fmul_rn(d, (float)s), (float)qv, -__fmul_rn(mn, (float)m));
    }
    ps_tl_max(tl, 1);
}

// greedy pick (Model::decode with top_k = 1: ProbArray + greedy_sample, src/model/llama/llama_model.cpp:124-128), stage 2:
// reduce the lm_head kernel's per-CTA partial maxima (first maximum wins) and do the device-side step bookkeeping:
// ids[*ctr] = argmax, token feedback, position and counter advance.
__global__ void __launch_bounds__(256) ps_k_argmax_step(const float *__restrict__ part_val, const int *__restrict__ part_idx, int n_part,
                                                        int32_t *__restrict__ ids, int32_t *__restrict__ ctr, int32_t *__restrict__ next_token,
                                                        int32_t *__restrict__ pos, long long *tl, const PsTpIn *tpi) {
    __shared__ float sv[8];
    __shared__ int si[8];
    ps_tl_min(tl, 0);
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    if (tpi) {
        if (threadIdx.x == 0) ps_tp_wait(tpi);
        __syncthreads();
    }
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int t = threadIdx.x; t < n_part; t += blockDim.x) {
        const float v = part_val[t];
        const int i = part_idx[t];
        if (v > best || (v == best && i < bi)) { best = v; bi = i; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(PS_FULL, best, o);
        const int oi = __shfl_xor_sync(PS_FULL, bi, o);
        if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int t = 1; t < (int)(blockDim.x >> 5); t++)
            if (sv[t] > best || (sv[t] == best && si[t] < bi)) { best = sv[t]; bi = si[t]; }
        if (bi == 0x7fffffff) bi = 0;
        ids[*ctr] = bi;
        *ctr += 1;
        *next_token = bi;
        *pos += 1;
    }
    ps_tl_max(tl, 1);
}

// ====================================================================================================================
// Decode attention in ONE kernel (replaces ps_k_attn1 + ps_k_attn2 when the rows fit shared memory): thread-block CLUSTERS
// of eight CTAs (the portable size; clusters of sixteen ran one or two at a time on the B200 and serialised the layer).
// A cluster serves one kv head and 64 of its output dims (grid = (8, hs / 64, kv heads)): its CTAs split the cache positions
// for the scores, exchange the row maxima, the exponentials and the partial sums through distributed shared memory, and
// then each takes eight output dims of the P.V product with its V^T rows staged by TMA.  The scores never visit global
// memory, the soft-max row is built once per cluster (it was rebuilt by each of the 16 CTAs of a kv group before) and one
// kernel boundary per layer disappears.  Arithmetic and summation orders are those of ps_k_attn1 / ps_k_attn2 (ggml_vec_dot_f32 lane
// chains, GGML_F32x8_REDUCE, ggml_v_expf on full 8-groups of the row / libm expf on its tail, double sum, position-ordered
// FMA chains for P.V) - bit-identical results.  norm_attention.cpp:115-151, ggml.c:2092-2131, 14846-14940.
// ====================================================================================================================
#include <cooperative_groups.h>
#define PS_AF_THREADS 512
template <int R2, int STEPS>
__global__ void __launch_bounds__(PS_AF_THREADS) ps_k_attn_fused(float *__restrict__ att, const float *__restrict__ kc, const float *__restrict__ vct,
                                                                 const float *__restrict__ q, float *__restrict__ ex, const int32_t *__restrict__ pos_dev,
                                                                 int n_kv_heads, int n_ctx, float scale, int s_cap, long long *tl, const PsTpOut *tpo) {
    namespace cg = cooperative_groups;
    constexpr int hs = 32 * STEPS, NW = PS_AF_THREADS / 32;
    extern __shared__ __align__(128) float s_af[]; // [R2][stride] probabilities | [8][stride] V^T rows | [R2][s_cap] this CTA's scores / exponentials
    __shared__ float s_q[R2][hs];
    __shared__ float s_max[R2];
    __shared__ double s_sum[R2];
    __shared__ float shf[NW][R2];
    __shared__ double shd[NW];
    __shared__ __align__(8) uint64_t bar_v;
    cg::cluster_group cluster = cg::this_cluster();
    const int CL = (int)cluster.num_blocks(), c = (int)cluster.block_rank(), g = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, kvd = hs * n_kv_heads;
    constexpr int dpc = 8;                               // output dims of this CTA
    const int d_base = ((int)blockIdx.y * CL + c) * dpc; // first of them
    if (tid == 0) {
        ps_mbar_init(&bar_v, 1);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    ps_tl_min(tl, 0);
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    ps_tl_min(tl, 2);
    const uint32_t ll_epoch = (tpo && tpo->peer_ll[0]) ? ps_tp_ll_epoch(tpo) : 0;
    const long long t_dep = (tl && tid == 0) ? ps_globaltimer() : 0;
#define PS_AF_PROBE(k)                                                                                                   \
    do {                                                                                                                 \
        if (tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(tl + (k)), (unsigned long long)(ps_globaltimer() - t_dep)); \
    } while (0)
    const int n_kv = pos_dev[0] + 1;
    const int stride = (n_kv + 31) & ~31;
    const int S = (((n_kv + CL - 1) / CL) + 7) & ~7;              // positions per CTA, a multiple of 8: the SIMD-exp groups never straddle CTAs
    const int j_lo = min(c * S, n_kv), j_hi = min(n_kv, j_lo + S);
    float *s_p = s_af, *s_v = s_af + R2 * stride, *s_sc = s_v + dpc * stride;
    // the exponentials travel between the CTAs of the cluster through global memory (L2): rows of this cluster's own scratch
    float *ex_rows = ex + ((size_t)blockIdx.y * gridDim.z + g) * R2 * (size_t)n_ctx;
    __syncthreads(); // barrier word initialised
    if (tid == 0) {  // this CTA's V^T rows: one bulk copy each, in flight during the whole soft-max
        const uint32_t bytes = (uint32_t)(((n_kv + 3) & ~3) * 4);
        ps_mbar_expect_tx(&bar_v, bytes * dpc);
        for (int w = 0; w < dpc; w++) ps_bulk_g2s(s_v + w * stride, vct + ((int64_t)g * hs + d_base + w) * n_ctx, bytes, &bar_v);
    }
    // ---- scores of positions [j_lo, j_hi) against the R2 query heads of the group (ps_k_attn1)
    for (int idx = tid; idx < R2 * hs; idx += PS_AF_THREADS) s_q[idx / hs][idx % hs] = q[(int64_t)g * R2 * hs + idx];
    __syncthreads();
    float qv[R2][STEPS];
#pragma unroll
    for (int hh = 0; hh < R2; hh++)
#pragma unroll
        for (int s = 0; s < STEPS; s++) qv[hh][s] = s_q[hh][32 * s + lane];
    float lmax = -INFINITY; // lane hh < R2: running maximum of head hh over this warp's positions
    for (int j0 = j_lo + warp * 8; j0 < j_hi; j0 += NW * 8) {
        float kv[8][STEPS];
#pragma unroll
        for (int t = 0; t < 8; t++)
#pragma unroll
            for (int s = 0; s < STEPS; s++) kv[t][s] = (j0 + t < j_hi) ? kc[(int64_t)(j0 + t) * kvd + g * hs + 32 * s + lane] : 0.f;
#pragma unroll
        for (int t = 0; t < 8; t++) {
            float sum[R2];
#pragma unroll
            for (int hh = 0; hh < R2; hh++) {
                sum[hh] = 0.f;
#pragma unroll
                for (int s = 0; s < STEPS; s++) sum[hh] = __fmaf_rn(kv[t][s], qv[hh][s], sum[hh]); // ggml_vec_dot_f32 lane chain
            }
            ps_f32x8_reduce_n<R2>(sum);
            if (lane < R2 && j0 + t < j_hi) {
                float v = sum[0];
#pragma unroll
                for (int hh = 1; hh < R2; hh++)
                    if (lane == hh) v = sum[hh];
                v = __fadd_rn(__fmul_rn(v, scale), 0.0f);
                s_sc[lane * s_cap + (j0 + t - j_lo)] = v;
                lmax = fmaxf(lmax, v);
            }
        }
    }
    if (lane < R2) shf[warp][lane] = lmax;
    __syncthreads();
    if (tid < R2) {
        float m = shf[0][tid];
        for (int w = 1; w < NW; w++) m = fmaxf(m, shf[w][tid]);
        s_max[tid] = m;
    }
    PS_AF_PROBE(4);
    cluster.sync(); // every CTA's maxima are in its shared memory
    PS_AF_PROBE(5);
    // ---- soft-max: row maximum over the cluster, exponentials of the own slice, partial sums in double (ps_k_attn2)
    constexpr int TPH = PS_AF_THREADS / R2, WPH = TPH / 32; // threads / warps per head
    const int hh = tid / TPH, ht = tid % TPH;
    float mx = -INFINITY;
    for (int r = 0; r < CL; r++) mx = fmaxf(mx, *cluster.map_shared_rank(&s_max[hh], r));
    {
        float *pp = s_sc + hh * s_cap;
        float *gp = ex_rows + (size_t)hh * n_ctx + j_lo;
        const int n8 = n_kv & ~7, len = j_hi - j_lo;
        const int g_end = (min(j_hi, n8) - j_lo) >> 3; // full 8-groups of the ROW inside this slice (may be <= 0)
        double s = 0.0;
        for (int gi = ht; gi < g_end; gi += TPH) {
            const float4 *p4 = reinterpret_cast<const float4 *>(pp + gi * 8);
            const float4 xa = p4[0], xb = p4[1];
            float vv[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
#pragma unroll
            for (int l = 0; l < 8; l++) vv[l] = ps_v_expf(__fadd_rn(vv[l], -mx));
            float4 *g4 = reinterpret_cast<float4 *>(gp + gi * 8); // rows and slices are 32-byte aligned (n_ctx % 8 == 0, S % 8 == 0)
            g4[0] = make_float4(vv[0], vv[1], vv[2], vv[3]);
            g4[1] = make_float4(vv[4], vv[5], vv[6], vv[7]);
            const float r0 = __fadd_rn(vv[4], vv[0]), r1 = __fadd_rn(vv[5], vv[1]), r2_ = __fadd_rn(vv[6], vv[2]), r3 = __fadd_rn(vv[7], vv[3]);
            s += (double)__fadd_rn(__fadd_rn(r0, r2_), __fadd_rn(r1, r3));
        }
        for (int j = max(n8 - j_lo, 0) + ht; j < len; j += TPH) { // scalar tail of the row: libm expf
            const float vv = ps_expf_glibc(__fadd_rn(pp[j], -mx));
            gp[j] = vv;
            s += (double)vv;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(PS_FULL, s, o);
        if (lane == 0) shd[warp] = s;
        __syncthreads();
        if (ht == 0) {
            double t = 0.0;
            for (int w = 0; w < WPH; w++) t += shd[hh * WPH + w];
            s_sum[hh] = t;
        }
    }
    PS_AF_PROBE(6);
    cluster.sync(); // release / acquire at cluster scope: every CTA's exponentials (global) and partial sums (shared) are visible
    PS_AF_PROBE(7);
    {
        double sum = 0.0;
        for (int r = 0; r < CL; r++) sum += *cluster.map_shared_rank(&s_sum[hh], r); // rank order: the same value on every CTA
        const float inv = (float)(1.0 / sum);
        float *dstp = s_p + hh * stride;
        const float *srcp = ex_rows + (size_t)hh * n_ctx;
        const int n4 = n_kv >> 2;
#pragma unroll 4
        for (int j4 = ht; j4 < n4; j4 += TPH) { // the whole row from L2, normalised on the way
            const float4 e = __ldcg(reinterpret_cast<const float4 *>(srcp) + j4);
            reinterpret_cast<float4 *>(dstp)[j4] = make_float4(__fmul_rn(e.x, inv), __fmul_rn(e.y, inv), __fmul_rn(e.z, inv), __fmul_rn(e.w, inv));
        }
        const int j = (n4 << 2) + ht;
        if (j < n_kv) dstp[j] = __fmul_rn(__ldcg(srcp + j), inv);
    }
    __syncthreads();
    ps_tl_max(tl, 3);
    // ---- P.V: warp w < 8 walks V^T row w of this CTA once for all R2 heads (ggml_vec_dot_f32 order)
    const int np = n_kv & ~31, ntail = n_kv - np;
    if (warp < dpc) {
        ps_mbar_wait(&bar_v, 0);
        const float *vs = s_v + warp * stride;
        const int d = d_base + warp;
        float sum[R2];
#pragma unroll
        for (int h2 = 0; h2 < R2; h2++) sum[h2] = 0.f;
        const float vtail = (lane < ntail) ? vs[np + lane] : 0.f;
#pragma unroll 8
        for (int s0 = 0; s0 < np; s0 += 32) { // the FMA chains stay in position order
            const float v = vs[s0 + lane];
#pragma unroll
            for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fmaf_rn(v, s_p[h2 * stride + s0 + lane], sum[h2]);
        }
        ps_f32x8_reduce_n<R2>(sum);
        for (int t = 0; t < ntail; t++) { // leftovers: mul, then add, in order (every lane computes the same chain)
            const float v = __shfl_sync(PS_FULL, vtail, t);
#pragma unroll
            for (int h2 = 0; h2 < R2; h2++) sum[h2] = __fadd_rn(sum[h2], __fmul_rn(v, s_p[h2 * stride + np + t]));
        }
        if (lane < R2) {
            float v = sum[0];
#pragma unroll
            for (int h2 = 1; h2 < R2; h2++)
                if (lane == h2) v = sum[h2];
            att[(int64_t)(g * R2 + lane) * hs + d] = v;
            if (tpo) { // all-gather by peer stores
                if (tpo->peer_ll[0]) ps_tp_ll_store(tpo, (int64_t)(g * R2 + lane) * hs + d, v, ll_epoch);
                else
                    for (int p = 0; p < tpo->n; p++) tpo->peer_dst[p][(int64_t)(g * R2 + lane) * hs + d] = v;
            }
        }
    }
    if (tpo) {
        if (tpo->peer_ll[0]) {
            if (tid == 0) ps_tp_ll_done(tpo, (int)(gridDim.x * gridDim.y * gridDim.z), ll_epoch);
        } else {
            __syncthreads();
            if (tid == 0) ps_tp_signal(tpo, (int)(gridDim.x * gridDim.y * gridDim.z));
        }
    }
    cluster.sync(); // nobody leaves while a peer may still read its maxima / sums
    ps_tl_max(tl, 1);
}
