// ps_rw.cuh — "row-walker" Q4_K mat-vec for decode (bs = 1): the fused decode path's weight-streaming kernel.
//
// Replaces powerserve_compute_forward_mul_mat -> ggml_vec_dot_q4_K_q8_K (libs/ggml/src/ggml.c:13344-13432,
// ggml-quants.c:7809-7872) for one activation column, with the RMSNorm + Q8_K activation quantisation of the op before
// it (ggml.c:12667-12721, ggml-quants.c:3799-3837) as prologue and the bias / residual / SiLU*up ops after it
// (ggml.c:10042-10112, src/backend/ggml/ggml.cpp:115-129) as epilogue.  Bit-identical to the table-op kernels.
//
// Why this shape (B200): the op is an HBM stream of 144-byte blocks with ~85 warp instructions of work per 8 row-blocks -
// at 6.5 TB/s an SM must retire a block every ~6 cycles, so the kernel is as much issue-bound as bandwidth-bound.  The
// earlier design (one thread per block, shared-memory hand-off to separate fp32 chain threads, CTA-wide barriers per
// tile, 8 warps) measured 20-27 % issue utilisation and 3x instruction overhead (profiles/r01b_*).  Here:
//   * FOUR threads own one weight row: thread q of the quad owns AVX lanes 2q, 2q+1 of the reference's __m256 accumulator
//     (and lane q of the __m128 mins accumulator) and walks the row's super-blocks IN ORDER, so the fp32 FMA chains of
//     the reference live in three registers per thread — no hand-off, no chain phase, no barrier inside the stream.
//   * a warp owns an OCTET of rows.  Weights are re-laid at bind time (quant bytes permuted, the 6-bit scales / mins
//     expanded to bytes once) so that the eight rows' 16-byte headers and 32-byte quant groups of one super-block are
//     adjacent: a warp's LDS are conflict-free and a pipeline stage (kb super-blocks of an octet) is ONE contiguous bulk copy.
//   * every warp runs its own TMA ring (cp.async.bulk + mbarrier, lane 0 re-arms a slot right after the warp drained
//     it), so warps never wait for each other; the first slots are requested before griddepcontrol.wait, i.e. while the
//     previous kernel in the PDL chain is still draining.
#pragma once
#include "ps_decode.cuh"

#define PS_RW_WARPS 16
#define PS_RW_THREADS (PS_RW_WARPS * 32)
#define PS_RW_OCTET_BLOCK 1184                 // 8 rows x 148 bytes: the 144 bytes of a Q4_K block with its 6-bit scales / mins expanded to bytes
#define PS_RW_HDR2 128                         // offset of the second header array (mins 4..7, 4 bytes per row)
#define PS_RW_QS 160                           // offset of the four 256-byte quant groups
#define PS_RW_MAX_NS 8
#define PS_RW_KS_KB 4                            // K-split: most blocks per stage (their factors wait in registers for the token)

enum { PS_RW_OUT_PLAIN = 0, PS_RW_OUT_ROPE = 1, PS_RW_OUT_ROPE_KCACHE = 2, PS_RW_OUT_VCACHE_T = 3 };

struct PsRwSeg {
    float *dst;         // output rows of this segment (indexed by row - row_begin)
    const float *bias;  // optional
    int row_begin, row_end;
    int mode;           // PS_RW_OUT_*: the q / k / v epilogues fuse ROPE and the two KV-cache COPY ops (norm_attention.cpp:72-105)
};

struct PsRwArgs {
    const uint8_t *w;      // repacked weights: [n_oct][nb][rpt][1184]
    int n_oct;             // row octets (pairs of octets when rpt == 2)
    int K;                 // contraction length, multiple of 256
    int kb;                // super-blocks per pipeline stage
    int ns;                // stages per warp ring
    int n_act;             // warps of a CTA that own octets (ring slots exist only for these)
    PsRwSeg seg[3];
    int n_seg;
    const float *x;        // fp32 activation [K]
    const float *norm_w;   // non-null: quantise rmsnorm(x) * norm_w
    float eps;
    const float *residual; // PS_EPI_RESIDUAL
    // fused ROPE / KV store (decode): position from device memory, cos/sin table row = pos (ggml.c:15342-15356)
    const int32_t *pos_dev;
    const float *rope_table;
    int hs, kvd, n_ctx;
    // producer-side activation quantisation (gate/up -> down): the SiLU epilogue's last-arriving warp of every 256-row
    // block quantises it to Q8_K straight into the consumer's shared-memory image (xq_out, [K bytes][nb x 32]); the
    // consumer (xq_in) then needs one bulk copy instead of a quantisation prologue.
    uint8_t *xq_out;
    int *blk_cnt;          // one arrival counter per 256-row block, left at zero
    const uint8_t *xq_in;
    const float *next_norm_w; // norm weights of the NEXT kernel in the chain: pulled into L2 here so that its prologue
    int next_norm_n;          // does not wait on DRAM behind its own weight prefetch
    double inv_k;          // 1 / K when K is a power of two (the mean is then an exact scaling), else 0
    float *part_val;       // optional (lm_head): per-CTA partial arg-max of the produced rows, [gridDim.x]
    int *part_idx;
    const PsTpOut *tpo;    // tensor parallel (else null): the output rows (or the arg-max partials) go into every rank's buffer
    const PsTpIn *tpi;     // tensor parallel (else null): x is a gathered vector — wait for the peers' shards
    const unsigned long long *x_ll; // tensor parallel, in-band flags (else null): x as (value, epoch) words; replaces `x` and `tpi`
    const uint32_t *x_epoch;        //   the epoch those words must carry (local counter of the producing slot)
    int *tp_err;                    //   set to 1 if a poll gave up
    int unroll2;           // walk two blocks per loop trip (launches with few octets per CTA: latency-bound lone warps)
    int ksplit;            // K-split kernels: warps per octet (each walks nb / ksplit consecutive blocks of the row); else 0
    int defer;             // the helper warp requests the weight stream only after the activation vector has arrived
    int idx_offset;        // added to the row index stored in part_idx (tensor parallel: first vocabulary row of this rank)
    long long *tl;         // optional timeline slot (option "trace")
    long long *cta_tl;     // optional per-CTA stream trace: [grid][8] (options "trace" + "cta_trace" = launch kind)
};

// ---------------------------------------------------------------------------------------------------- repack
// GGUF rows [n_rows][nb][144] -> [octet][block][slot][1184] where 1184 = 8 headers (16 B) + 8 x 4 B (mins 4..7) + 4 groups x 8 rows x 32 B.
// `slot`/`n_slots` interleave several matrices per octet-block (gate | up).  Rows beyond n_rows are zero blocks.
__global__ void ps_k_rw_repack(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n_rows, int64_t nb, int64_t oct0, int slot,
                               int n_slots) {
    const int64_t n_oct = (n_rows + 7) / 8;
    const int64_t total = n_oct * nb * 72; // 16-byte chunks of the source blocks
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % 9);
        const int r = (int)((t / 9) % 8);
        const int64_t i = (t / 72) % nb;
        const int64_t o = t / (72 * nb);
        const int64_t row = o * 8 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < n_rows) v = *reinterpret_cast<const uint4 *>(src + (row * nb + i) * PS_Q4_K_BYTES + 16 * c);
        uint8_t *ob = dst + (((oct0 + o) * nb + i) * n_slots + slot) * PS_RW_OCTET_BLOCK;
        if (c == 0) {
            // d | dmin, then the twelve packed bytes as eight scale bytes and eight min bytes: the utmp shuffle of
            // ggml_vec_dot_q4_K_q8_K (ggml-quants.c:7816-7826) == get_scale_min_k4, done once here instead of per block per token
            const uint32_t k1 = 0x3f3f3f3fu, k2 = 0x0f0f0f0fu, k3 = 0x03030303u;
            const uint32_t scA = v.y & k1, scB = (v.w & k2) | (((v.y >> 6) & k3) << 4);
            const uint32_t mA = v.z & k1, mB = ((v.w >> 4) & k2) | (((v.z >> 6) & k3) << 4);
            *reinterpret_cast<uint4 *>(ob + 16 * r) = make_uint4(v.x, scA, scB, mA);
            *reinterpret_cast<uint32_t *>(ob + PS_RW_HDR2 + 4 * r) = mB;
        } else {
            *reinterpret_cast<uint4 *>(ob + PS_RW_QS + 256 * ((c - 1) >> 1) + 32 * r + 16 * ((c - 1) & 1)) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------- block math
// One super-block of one row for the quad thread q: the two int32 lanes (2q, 2q+1) of `sumi` and lane q of `prod`
// exactly as one iteration of the AVX2 loop leaves them (ggml-quants.c:7828-7860), then the three FMAs (:7858, :7834).
struct PsRwAcc {
    float a0, a1, am;
};

// byte k of w as an int (one PRMT)
PS_D int ps_rw_byte(uint32_t w, int k) { return (int)__byte_perm(w, 0, 0x4440 + k); }
// offset, inside an octet block, of the 16-bit pair of mins (2q, 2q+1) of row r
PS_D int ps_rw_mins_off(int r, int q) { return (q < 2) ? 16 * r + 12 + 2 * q : PS_RW_HDR2 + 4 * r + 2 * (q - 2); }
// prod lane q: mins(2q) * bsums(2q) + mins(2q+1) * bsums(2q+1)  (_mm_madd_epi16(mins, q8s), :7830-7833); exact integers
PS_D int ps_rw_mins_dot(uint32_t bsums_pair, uint32_t mins_pair) {
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(bsums_pair), "r"(mins_pair), "r"(0));
    return d;
}

// the exact integer sums of one super-block as floats (S of lanes 2q, 2q+1, P of mins lane q) and the two fp32 factors
// d = yd * d_x, dm = -yd * dmin_x - everything the FMA chains need from the block
PS_D void ps_rw_block_factors(const uint8_t *ob, int r, int q, int moff, const uint4 *qa, const uint2 meta, float &f0, float &f1, float &fp, float &d,
                              float &dm) {
    const uint4 h = *reinterpret_cast<const uint4 *>(ob + 16 * r); // d | dmin, scales 0..3, scales 4..7, mins 0..3
    const uint32_t mp = *reinterpret_cast<const uint16_t *>(ob + moff);
    int S0 = 0, S1 = 0, H0 = 0, H1 = 0;
#pragma unroll
    for (int j2 = 0; j2 < 4; j2++) {
        const uint2 w = *reinterpret_cast<const uint2 *>(ob + PS_RW_QS + 256 * j2 + 32 * r + 8 * q);
        const uint4 a = qa[j2 * 4 + q];
        const uint32_t scw = (j2 < 2) ? h.y : h.z;
        const int s_lo = ps_rw_byte(scw, 2 * (j2 & 1)), s_hi = ps_rw_byte(scw, 2 * (j2 & 1) + 1);
        S0 += s_lo * __dp4a((int)(w.x & 0x0f0f0f0fu), (int)a.x, 0);
        S1 += s_lo * __dp4a((int)(w.y & 0x0f0f0f0fu), (int)a.y, 0);
        H0 += s_hi * ps_dp4a_us(w.x & 0xf0f0f0f0u, (int)a.z, 0);   // 16 x the high-nibble dot
        H1 += s_hi * ps_dp4a_us(w.y & 0xf0f0f0f0u, (int)a.w, 0);
    }
    S0 += H0 >> 4;
    S1 += H1 >> 4;
    const int P = ps_rw_mins_dot(meta.y, mp);
    const float yd = __uint_as_float(meta.x);
    d = __fmul_rn(yd, ps_half_bits_to_float(h.x & 0xffffu));
    dm = __fmul_rn(-yd, ps_half_bits_to_float(h.x >> 16));
    f0 = __int2float_rn(S0);
    f1 = __int2float_rn(S1);
    fp = __int2float_rn(P);
}

PS_D void ps_rw_block(const uint8_t *ob, int r, int q, int moff, const uint4 *qa, const uint2 meta, PsRwAcc &acc) {
    float f0, f1, fp, d, dm;
    ps_rw_block_factors(ob, r, q, moff, qa, meta, f0, f1, fp, d, dm);
    acc.a0 = __fmaf_rn(d, f0, acc.a0);
    acc.a1 = __fmaf_rn(d, f1, acc.a1);
    acc.am = __fmaf_rn(dm, fp, acc.am);
}

// hsum_float_8(acc) + the movehl/movehdup sum of acc_m (ggml-quants.c:62-68, 7862-7871) across the quad; every lane
// of the quad ends with the row result.
PS_D float ps_rw_row_result(const PsRwAcc &acc) {
    // quad thread q holds lanes (2q, 2q+1): r_l = x[l+4] + x[l]
    const float p0 = __shfl_xor_sync(PS_FULL, acc.a0, 2), p1 = __shfl_xor_sync(PS_FULL, acc.a1, 2);
    const float ra = __fadd_rn(p0, acc.a0), rb = __fadd_rn(p1, acc.a1); // q=0: r0,r1 ; q=1: r2,r3 (q=2,3 mirror them)
    const float oa = __shfl_xor_sync(PS_FULL, ra, 1), ob = __shfl_xor_sync(PS_FULL, rb, 1);
    const float hs = __fadd_rn(__fadd_rn(ra, oa), __fadd_rn(rb, ob));   // (r0 + r2) + (r1 + r3)
    const float pm = __shfl_xor_sync(PS_FULL, acc.am, 2);
    const float ma = __fadd_rn(acc.am, pm);                             // q=0: m0+m2 ; q=1: m1+m3
    const float mb = __shfl_xor_sync(PS_FULL, ma, 1);
    return __fadd_rn(hs, __fadd_rn(ma, mb));
}

// ---------------------------------------------------------------------------------------------------- the kernel
// quantise one 256-block held as e[8] per lane (see ps_quant_block_q8k_regs) into the shared-memory image the block
// math reads: qa[i][j2][q] = {sub-block 2*j2 words 2q, 2q+1 ; sub-block 2*j2+1 words 2q, 2q+1}, meta[i][q] = {d, bsums pair q}
// (the eight elements travel BY VALUE: a pointer to the caller's register array would force it through local memory - two
// STL.128 + two LDL.128 per call, L2 round trips when the 24 KB of L1 left beside the shared-memory carve-out miss)
__device__ __noinline__ void ps_rw_quant_store_v(float4 ea, float4 eb, int lane, uint32_t *qw, uint2 *meta) {
    uint32_t words[2], bsp4;
    float yd;
    const float v[8] = {ea.x, ea.y, ea.z, ea.w, eb.x, eb.y, eb.z, eb.w};
    ps_quant_block_q8k_regs(v, lane, words, yd, bsp4);
    const int l = lane & 7, jA = lane >> 3, jB = 4 + (lane >> 3); // natural words: (sub-block L/8, word L%8), (4 + L/8, L%8)
    qw[((jA >> 1) * 4 + (l >> 1)) * 4 + (jA & 1) * 2 + (l & 1)] = words[0];
    qw[((jB >> 1) * 4 + (l >> 1)) * 4 + (jB & 1) * 2 + (l & 1)] = words[1];
    if (lane < 4) meta[lane] = make_uint2(__float_as_uint(yd), bsp4);
}

PS_D void ps_rw_quant_store(const float *e, int lane, uint32_t *qw, uint2 *meta) {
    ps_rw_quant_store_v(make_float4(e[0], e[1], e[2], e[3]), make_float4(e[4], e[5], e[6], e[7]), lane, qw, meta);
}

// the same eight elements from an in-band-flag vector: poll until all eight words carry epoch `ep`
PS_D void ps_rw_load8_ll(const unsigned long long *p, int lane, float e[8], uint32_t ep, int *err) {
    const unsigned long long *p0 = p + 4 * lane, *p1 = p + 128 + 4 * lane;
    int spins = 0;
    for (;;) {
        const bool ok = ps_tp_ll_load2(p0, ep, e[0], e[1]) & ps_tp_ll_load2(p0 + 2, ep, e[2], e[3]) & ps_tp_ll_load2(p1, ep, e[4], e[5]) &
                        ps_tp_ll_load2(p1 + 2, ep, e[6], e[7]);
        if (ok) break;
        if (++spins > PS_TP_LL_SPINS || ((spins & 1023) == 0 && *reinterpret_cast<volatile int *>(err))) { *err = 1; break; } // one give-up ends them all
    }
}
PS_D void ps_rw_load8(const float *p, int lane, float e[8]) {
    const float4 v0 = *reinterpret_cast<const float4 *>(p + 4 * lane);
    const float4 v1 = *reinterpret_cast<const float4 *>(p + 128 + 4 * lane);
    e[0] = v0.x; e[1] = v0.y; e[2] = v0.z; e[3] = v0.w; e[4] = v1.x; e[5] = v1.y; e[6] = v1.z; e[7] = v1.w;
}

// Threads: PS_RW_WARPS compute warps + one helper warp that initialises the mbarriers and fills every warp's ring
// (lane w serves warp w) and then exits, so no compute warp ever stalls on the TMA queue during the prologue.
// Dynamic shared memory: [qa: K bytes][meta: nb x 4 x 8][rings: n_act x ns x stage_bytes][bars: n_act x ns x 8]
//
// KSPLIT (q|k|v, o, down: matrices with only 3-6 row octets per SM, where a lone warp per octet walks 16-56 blocks
// serially at ~230 cycles each): `ksplit` warps share an octet.  The row's stages (kb blocks each) are dealt round-robin to
// the warps of the group; a warp does the integer work of its stage (exact sums + the fp32 factors, kept in registers),
// then waits for the octet's TOKEN - the three running FMA accumulators of every row plus a stage counter, one 16-byte
// shared-memory word per lane - applies its kb blocks and passes the token on.  The chains therefore still advance block by
// block in row order (the arithmetic is that of the one-warp walk); only the integer work of different stages overlaps.
template <int EPI, bool KSPLIT = false>
__global__ void __launch_bounds__(PS_RW_THREADS + 32, 1) ps_k_rw_matvec(const PsRwArgs a) {
    constexpr int RPT = (EPI == PS_EPI_SILU) ? 2 : 1;
    static_assert(!(KSPLIT && RPT != 1), "K-split walks one matrix");
    extern __shared__ __align__(128) uint8_t ps_rw_smem[];
    __shared__ double sh_red[PS_RW_WARPS];
    __shared__ __align__(8) uint64_t xbar;
    __shared__ uint4 ks_tok[KSPLIT ? PS_RW_WARPS * 32 : 1]; // K-split: the tokens, [group][lane] = {a0, a1, am, stages applied}
    const int K = a.K, nb = K / 256, kb = a.kb, ns = a.ns;
    const uint32_t stage_bytes = (uint32_t)kb * RPT * PS_RW_OCTET_BLOCK;
    uint4 *s_qa = reinterpret_cast<uint4 *>(ps_rw_smem);
    uint2 *s_meta = reinterpret_cast<uint2 *>(ps_rw_smem + K);
    uint8_t *s_ring = ps_rw_smem + K + (size_t)nb * 32;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + (size_t)a.n_act * ns * stage_bytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, q = lane & 3;
    // this CTA's octets [o0, o1); warp w walks o0 + w, o0 + w + n_act, ...
    const int o0 = (int)(((long long)blockIdx.x * a.n_oct) / gridDim.x), o1 = (int)(((long long)(blockIdx.x + 1) * a.n_oct) / gridDim.x);
    // K-split: a group of `ks` warps per octet, `seg` blocks each; otherwise one warp per octet walking all nb blocks
    const int ks = KSPLIT ? a.ksplit : 1, n_grp = a.n_act / ks;
    const int spo = nb / kb / ks;            // stages per (octet, warp); K-split: warp k of a group owns stages k, k + ks, ...
    const size_t oct_bytes = (size_t)nb * RPT * PS_RW_OCTET_BLOCK;
    auto stages_of = [&](int w) { return (w < a.n_act && o0 + w / ks < o1) ? ((o1 - o0 - w / ks - 1) / n_grp + 1) * spo : 0; };
    auto issue = [&](int w, int s) { // request stage #s of warp w's stream into slot s % ns of its ring
        const int oct = o0 + w / ks + (s / spo) * n_grp;
        const uint8_t *src = a.w + (size_t)oct * oct_bytes + (size_t)((s % spo) * ks + w % ks) * stage_bytes;
        uint64_t *bar = s_bar + w * ns + (s % ns);
        ps_mbar_expect_tx(bar, stage_bytes);
        ps_bulk_g2s(s_ring + ((size_t)w * ns + (s % ns)) * stage_bytes, src, stage_bytes, bar);
    };
    ps_tl_min(a.tl, 0);

    if (warp == PS_RW_WARPS) {
        // ===== helper warp: weights never depend on the previous kernel, so the rings fill while it drains
        const int n = stages_of(lane);
        if (n > 0) {
            for (int s = 0; s < ns; s++) ps_mbar_init(s_bar + lane * ns + s, 1);
        }
        if (lane == 0) ps_mbar_init(&xbar, 1);
        if (KSPLIT)
            for (int t = lane; t < PS_RW_WARPS * 32; t += 32) ks_tok[t] = make_uint4(0, 0, 0, 0);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        asm volatile("bar.arrive 1, %0;" ::"n"(PS_RW_THREADS + 32) : "memory"); // barriers are live
        // barrier 3: the compute warps' norm-weight loads are on their way (DRAM queues are FIFO); with `defer` they arrive
        // only once the activation vector is in their registers, so its few KB do not queue behind ~200 KB of weights per SM
        ps_bar_sync(3, PS_RW_THREADS + 32);
        for (int s = 0; s < ns && s < n; s++) issue(lane, s);
        return;
    }

    // ===== compute warps
    // the norm weights do not depend on the previous kernel either
    float wv0[8];
    const bool early_w = a.norm_w != nullptr && warp < nb;
    if (early_w) ps_rw_load8(a.norm_w + warp * 256, lane, wv0);
    if (!a.defer) asm volatile("bar.arrive 3, %0;" ::"n"(PS_RW_THREADS + 32) : "memory");
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    ps_tl_min(a.tl, 2);
    const long long t_dep = (a.tl && tid == 0) ? ps_globaltimer() : 0;
#define PS_RW_PROBE(k)                                                                                                   \
    do {                                                                                                                 \
        if (a.tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(a.tl + (k)), (unsigned long long)(ps_globaltimer() - t_dep)); \
    } while (0)
    if (a.tpi && tid == 0) ps_tp_wait(a.tpi); // the gathered activation vector is complete on this rank
    const uint32_t ll_epoch = (a.tpo && a.tpo->peer_ll[0]) ? ps_tp_ll_epoch(a.tpo) : 0; // in-band-flag exchange: this launch's epoch (never 0)
    ps_bar_sync(1, PS_RW_THREADS + 32);         // mbarriers initialised by the helper (and the wait above is over)
    if (a.next_norm_w && blockIdx.x == 0)
        for (int i = tid * 32; i < a.next_norm_n; i += PS_RW_THREADS * 32)
            asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(a.next_norm_w + i));

    // ---- prologue: the Q8_K image of the activation vector in shared memory
    if (a.xq_in) { // quantised by the producer kernel: one bulk copy
        if (tid == 0) {
            const uint32_t bytes = (uint32_t)K + (uint32_t)nb * 32;
            ps_mbar_expect_tx(&xbar, bytes);
            ps_bulk_g2s(ps_rw_smem, a.xq_in, bytes, &xbar);
        }
        ps_mbar_wait(&xbar, 0);
        if (a.defer) asm volatile("bar.arrive 3, %0;" ::"n"(PS_RW_THREADS + 32) : "memory");
        PS_RW_PROBE(4);
    } else { // (RMSNorm) + quantize_row_q8_K, one warp per 256-block
        float e0[8];
        const bool have0 = warp < nb;
        const uint32_t x_ep = a.x_ll ? *reinterpret_cast<const volatile uint32_t *>(a.x_epoch) : 0;
        auto load_x = [&](int i, float (&e)[8]) {
            if (a.x_ll) ps_rw_load8_ll(a.x_ll + (size_t)i * 256, lane, e, x_ep, a.tp_err);
            else ps_rw_load8(a.x + i * 256, lane, e);
        };
        if (have0) load_x(warp, e0);
        if (a.defer) { // the first block of x is here (the sum below consumes it) before the weight stream is requested
            float sink = 0.f;
            if (have0) {
#pragma unroll
                for (int t = 0; t < 8; t++) sink += e0[t];
            }
            asm volatile("" ::"f"(sink) : "memory");
            asm volatile("bar.arrive 3, %0;" ::"n"(PS_RW_THREADS + 32) : "memory");
        }
        float nscale = 1.f;
        if (a.norm_w) {
            double ss = 0.0;
            if (have0) {
#pragma unroll
                for (int t = 0; t < 8; t++) ss += (double)__fmul_rn(e0[t], e0[t]);
            }
#pragma unroll 1
            for (int i = warp + PS_RW_WARPS; i < nb; i += PS_RW_WARPS) {
                float e[8];
                load_x(i, e);
#pragma unroll
                for (int t = 0; t < 8; t++) ss += (double)__fmul_rn(e[t], e[t]);
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(PS_FULL, ss, o);
            if (lane == 0) sh_red[warp] = ss;
            PS_RW_PROBE(4);
            ps_bar_sync(2, PS_RW_THREADS);
            double t = (lane < PS_RW_WARPS) ? sh_red[lane] : 0.0;
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(PS_FULL, t, o);
            const float mean = (float)(a.inv_k != 0.0 ? t * a.inv_k : t / (double)K);
            nscale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, a.eps)));
        }
        if (have0) {
            if (a.norm_w) {
#pragma unroll
                for (int t = 0; t < 8; t++) e0[t] = __fmul_rn(e0[t], __fmul_rn(wv0[t], nscale)); // y = x * (w * scale)
            }
            PS_RW_PROBE(5);
            ps_rw_quant_store(e0, lane, reinterpret_cast<uint32_t *>(s_qa) + (size_t)warp * 64, s_meta + warp * 4);
        }
#pragma unroll 1
        for (int i = warp + PS_RW_WARPS; i < nb; i += PS_RW_WARPS) {
            float e[8];
            load_x(i, e);
            if (a.norm_w) {
                float wv[8];
                ps_rw_load8(a.norm_w + i * 256, lane, wv);
#pragma unroll
                for (int t = 0; t < 8; t++) e[t] = __fmul_rn(e[t], __fmul_rn(wv[t], nscale));
            }
            ps_rw_quant_store(e, lane, reinterpret_cast<uint32_t *>(s_qa) + (size_t)i * 64, s_meta + i * 4);
        }
        PS_RW_PROBE(6);
        ps_bar_sync(2, PS_RW_THREADS);
    }
    PS_RW_PROBE(3); // slowest CTA's prologue

    // ---- the stream
    const int n_stages = stages_of(warp);
    const int n_mine = n_stages / spo;
    uint8_t *my_ring = s_ring + (size_t)warp * ns * stage_bytes;
    uint64_t *my_bar = s_bar + warp * ns;
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    int s = 0, slot = 0;
    uint32_t phase = 0;
    const int moff = ps_rw_mins_off(r, q);
    long long wcyc = 0; // trace: cycles this warp spent waiting for weight stages
    const long long c_begin = a.tl ? clock64() : 0, t_begin = a.tl ? ps_globaltimer() : 0;
    const int grp = warp / ks, kk = warp - grp * ks; // K-split: octet group of the warp, its segment of the row
#pragma unroll 1
    for (int m = 0; m < n_mine; m++) {
        const int oct = o0 + grp + m * n_grp;
        PsRwAcc acc[RPT];
#pragma unroll
        for (int t = 0; t < RPT; t++) acc[t].a0 = acc[t].a1 = acc[t].am = 0.f;
        if constexpr (KSPLIT) {
            const int G = spo * ks; // stages of a row
            uint4 *tok = ks_tok + grp * 32 + lane; // this lane's word of the group's token: {a0, a1, am, stages applied so far}
#pragma unroll 1
            for (int j = 0; j < spo; j++, s++) {
                const int g = j * ks + kk; // this warp's stage of the row: blocks [g * kb, (g + 1) * kb)
                const long long c0 = a.tl ? clock64() : 0;
                ps_mbar_wait(&my_bar[slot], phase);
                if (a.tl) wcyc += clock64() - c0;
                const uint8_t *st = my_ring + (size_t)slot * stage_bytes;
                // the stage's integer work, kept in registers (kb <= PS_RW_KS_KB) until the token arrives
                float f0[PS_RW_KS_KB], f1[PS_RW_KS_KB], fp[PS_RW_KS_KB], fd[PS_RW_KS_KB], fm[PS_RW_KS_KB];
#pragma unroll
                for (int b = 0; b < PS_RW_KS_KB; b++) {
                    if (b < kb) {
                        const int i = g * kb + b;
                        ps_rw_block_factors(st + (size_t)b * PS_RW_OCTET_BLOCK, r, q, moff, s_qa + (size_t)i * 16, s_meta[i * 4 + q], f0[b], f1[b], fp[b], fd[b],
                                            fm[b]);
                    }
                }
                __syncwarp(); // the slot is drained: re-arm it
                if (lane == 0 && s + ns < n_stages) issue(warp, s + ns);
                if (++slot == ns) { slot = 0; phase ^= 1; }
                // the token: stages 0 .. g - 1 of this octet (and all stages of the group's earlier octets) are applied.  Every
                // lane polls its own 16-byte word, whose last field is the count - value and flag travel in ONE shared-memory
                // store, so the hop needs no fence.  (A new octet's first stage waits too: its token write must not overtake
                // the previous octet's last reader.)
                const uint32_t want = (uint32_t)(m * G + g);
                if (want > 0) {
                    uint4 t;
                    do {
                        asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(t.x), "=r"(t.y), "=r"(t.z), "=r"(t.w) : "r"(ps_smem_u32(tok)) : "memory");
                    } while (t.w != want);
                    if (g > 0) { acc[0].a0 = __uint_as_float(t.x); acc[0].a1 = __uint_as_float(t.y); acc[0].am = __uint_as_float(t.z); }
                }
#pragma unroll
                for (int b = 0; b < PS_RW_KS_KB; b++) {
                    if (b < kb) {
                        acc[0].a0 = __fmaf_rn(fd[b], f0[b], acc[0].a0);
                        acc[0].a1 = __fmaf_rn(fd[b], f1[b], acc[0].a1);
                        acc[0].am = __fmaf_rn(fm[b], fp[b], acc[0].am);
                    }
                }
                asm volatile("st.volatile.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(ps_smem_u32(tok)), "r"(__float_as_uint(acc[0].a0)), "r"(__float_as_uint(acc[0].a1)),
                             "r"(__float_as_uint(acc[0].am)), "r"(want + 1)
                             : "memory");
            }
            if (kk != ks - 1) continue; // the warp that applied the last stage owns the epilogue
        } else {
#pragma unroll 1
        for (int sb = 0; sb < spo; sb++, s++) {
            const long long c0 = a.tl ? clock64() : 0;
            ps_mbar_wait(&my_bar[slot], phase);
            if (a.tl) wcyc += clock64() - c0;
            const uint8_t *st = my_ring + (size_t)slot * stage_bytes;
            if (RPT == 1 && a.unroll2) {
                // few octets per CTA (q|k|v, o, down): a warp walks its row alone on its scheduler and every block is a chain of
                // dependent shared-memory loads - two blocks per trip let the loads of one overlap the integer math of the other
                // (the fp32 chains still advance block by block, in row order)
#pragma unroll 2
                for (int b = 0; b < kb; b++) {
                    const int i = sb * kb + b;
                    ps_rw_block(st + (size_t)b * PS_RW_OCTET_BLOCK, r, q, moff, s_qa + (size_t)i * 16, s_meta[i * 4 + q], acc[0]);
                }
            } else {
#pragma unroll 1
                for (int b = 0; b < kb; b++) {
                    const int i = sb * kb + b;
                    const uint4 *qa = s_qa + (size_t)i * 16;
                    const uint2 meta = s_meta[i * 4 + q];
#pragma unroll
                    for (int t = 0; t < RPT; t++) ps_rw_block(st + (size_t)(b * RPT + t) * PS_RW_OCTET_BLOCK, r, q, moff, qa, meta, acc[t]);
                }
            }
            __syncwarp();
            if (lane == 0 && s + ns < n_stages) issue(warp, s + ns); // the slot is drained: re-arm it
            if (++slot == ns) { slot = 0; phase ^= 1; }
        }
        }
        // ---- epilogue
        const int row = oct * 8 + r;
        if (EPI == PS_EPI_SILU) {
            const float g = ps_rw_row_result(acc[0]);
            const float u = ps_rw_row_result(acc[RPT - 1]);
            if (q == 0 && row < a.seg[0].row_end) {
                const float hv = ps_silu_mul(g, u);
                a.seg[0].dst[row] = hv;
                if (a.tpo) { // all-gather by peer stores
                    if (ll_epoch) ps_tp_ll_store(a.tpo, row, hv, ll_epoch);
                    else
                        for (int p = 0; p < a.tpo->n; p++) a.tpo->peer_dst[p][row] = hv;
                }
            }
            if (a.xq_out) {
                const int i = oct >> 5; // 32 octets per 256-row block
                __syncwarp();
                int old = 0;
                if (lane == 0) // release: the warp's h stores above; acquire: the other warps' before we read the block
                    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;" : "=r"(old) : "l"(a.blk_cnt + i) : "memory");
                old = __shfl_sync(PS_FULL, old, 0);
                const int expect = min(32, a.n_oct - i * 32);
                if (old == expect - 1) { // this warp completed the block: quantise it (quantize_row_q8_K_ref, one warp)
                    const float4 v0 = __ldcg(reinterpret_cast<const float4 *>(a.seg[0].dst + i * 256 + 4 * lane));
                    const float4 v1 = __ldcg(reinterpret_cast<const float4 *>(a.seg[0].dst + i * 256 + 128 + 4 * lane));
                    const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
                    uint32_t words[2], bsp4;
                    float yd;
                    ps_quant_block_q8k_regs(e, lane, words, yd, bsp4);
                    uint32_t *qw = reinterpret_cast<uint32_t *>(a.xq_out) + (size_t)i * 64;
                    const int l = lane & 7, jA = lane >> 3, jB = 4 + (lane >> 3);
                    qw[((jA >> 1) * 4 + (l >> 1)) * 4 + (jA & 1) * 2 + (l & 1)] = words[0];
                    qw[((jB >> 1) * 4 + (l >> 1)) * 4 + (jB & 1) * 2 + (l & 1)] = words[1];
                    if (lane < 4) reinterpret_cast<uint2 *>(a.xq_out + (size_t)a.n_oct * 8)[i * 4 + lane] = make_uint2(__float_as_uint(yd), bsp4);
                    if (lane == 0) a.blk_cnt[i] = 0;
                }
            }
        } else {
            float res = ps_rw_row_result(acc[0]);
            int sg = 0;
            if (a.n_seg > 1 && row >= a.seg[1].row_begin) sg = 1;
            if (a.n_seg > 2 && row >= a.seg[2].row_begin) sg = 2;
            const bool live = row < a.seg[sg].row_end;
            const int n = row - a.seg[sg].row_begin;
            const int mode = a.seg[sg].mode;      // warp-uniform: segments are octet-aligned
            if (live && a.seg[sg].bias) res = __fadd_rn(res, a.seg[sg].bias[n]);
            if (EPI == PS_EPI_STORE && mode != PS_RW_OUT_PLAIN) {
                const int pos = a.pos_dev[0];
                if (mode == PS_RW_OUT_VCACHE_T) {         // V cache is stored transposed: [kv_dim][n_ctx]
                    if (q == 0 && live) a.seg[sg].dst[(size_t)n * a.n_ctx + pos] = res;
                } else {
                    // ggml_compute_forward_rope_f32, adjacent pairs (ggml.c:15455-15486): rows (2p, 2p+1) sit in
                    // neighbouring quads of the octet; products rounded separately, as the reference does
                    const float other = __shfl_xor_sync(PS_FULL, res, 4);
                    const int i0 = (n % a.hs) & ~1;
                    const float c = live ? a.rope_table[(size_t)pos * a.hs + i0] : 0.f, sn = live ? a.rope_table[(size_t)pos * a.hs + i0 + 1] : 0.f;
                    const float x0 = (r & 1) ? other : res, x1 = (r & 1) ? res : other;
                    const float out = (r & 1) ? __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c)) : __fadd_rn(__fmul_rn(x0, c), -__fmul_rn(x1, sn));
                    if (q == 0 && live) a.seg[sg].dst[(mode == PS_RW_OUT_ROPE_KCACHE ? (size_t)pos * a.kvd : 0) + n] = out;
                }
            } else if (q == 0 && live) {
                if (EPI == PS_EPI_RESIDUAL) res = __fadd_rn(a.residual[n], res);
                a.seg[sg].dst[n] = res;
                if (a.tpo && (EPI == PS_EPI_RESIDUAL || !a.part_val)) { // all-gather by peer stores
                    if (ll_epoch) ps_tp_ll_store(a.tpo, n, res, ll_epoch);
                    else
                        for (int p = 0; p < a.tpo->n; p++) a.tpo->peer_dst[p][n] = res;
                }
                if (res > best_v || (res == best_v && n < best_i)) { best_v = res; best_i = n; } // first maximum wins
            }
        }
    }
    if (a.tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(a.tl + 7), (unsigned long long)wcyc);
    if (a.cta_tl && tid == 0) { // per-CTA stream trace of warp 0 (tools/timeline.py --cta=KIND)
        long long *c = a.cta_tl + (size_t)blockIdx.x * 8;
        c[0] = t_dep; c[1] = t_begin; c[2] = ps_globaltimer(); c[3] = clock64() - c_begin; c[4] = wcyc; c[5] = n_mine * nb; c[6] = o1 - o0;
    }
    if (EPI == PS_EPI_STORE && a.part_val) { // greedy pick, stage 1: the CTA's best (value, lowest index)
        __shared__ float sv[PS_RW_WARPS];
        __shared__ int si[PS_RW_WARPS];
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(PS_FULL, best_v, o);
            const int oi = __shfl_xor_sync(PS_FULL, best_i, o);
            if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        if (lane == 0) { sv[warp] = best_v; si[warp] = best_i; }
        ps_bar_sync(2, PS_RW_THREADS);
        if (tid == 0) {
            for (int t = 1; t < PS_RW_WARPS; t++)
                if (sv[t] > best_v || (sv[t] == best_v && si[t] < best_i)) { best_v = sv[t]; best_i = si[t]; }
            const int gi = (best_i == 0x7fffffff) ? best_i : best_i + a.idx_offset;
            a.part_val[blockIdx.x] = best_v;
            a.part_idx[blockIdx.x] = gi;
            if (a.tpo)
                for (int p = 0; p < a.tpo->n; p++) { a.tpo->peer_dst[p][blockIdx.x] = best_v; a.tpo->peer_idx[p][blockIdx.x] = gi; }
        }
    }
    ps_bar_sync(2, PS_RW_THREADS);
    if (a.tpo && tid == 0) {
        if (ll_epoch) ps_tp_ll_done(a.tpo, (int)gridDim.x, ll_epoch);
        else ps_tp_signal(a.tpo, (int)gridDim.x);
    }
    ps_tl_max(a.tl, 1);
}

// ====================================================================================================================
// Multi-column row-walker: dst{N, bs} = W{K, N} . x{K, bs} for prefill chunks and speculative-verify batches.
// Same exact arithmetic per column as the mat-vec above (each column owns its FMA chains), same octet-interleaved weights
// and per-warp TMA rings; a weight block is unpacked ONCE and then multiplied with the C columns of the current column
// group, whose Q8_K images sit in shared memory.  The CTA's weight slice is re-streamed (from L2) once per column group.
// ====================================================================================================================
struct PsRwmArgs {
    const uint8_t *w;      // repacked weights: [n_oct][nb][n_slots][1184]
    int n_oct, K, kb, ns, n_act;
    int tile;              // octets per work unit (<= PS_RW_WARPS): warps >= tile idle
    int slot, n_slots;     // which interleaved matrix of the buffer (gate | up)
    PsRwSeg seg[3];        // dst of a segment is [bs][rows of the segment]; mode unused
    int n_seg;
    const uint8_t *x_img;  // Q8_K images of the activation columns: [bs][K + nb * 32]
    int bs;
    const float *residual; // optional, same layout as seg[0].dst (single-segment calls only)
};

// (RMSNorm output or any fp32 activation) -> Q8_K shared-memory image, one warp per (256-block, column)
__global__ void __launch_bounds__(128) ps_k_rw_quant_img(const float *__restrict__ x, int64_t K, uint8_t *__restrict__ img) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nb = K / 256, i = (int64_t)blockIdx.x * 4 + warp, col = blockIdx.y;
    if (i >= nb) return;
    float e[8];
    ps_rw_load8(x + col * K + i * 256, lane, e);
    uint8_t *base = img + col * (K + nb * 32);
    ps_rw_quant_store(e, lane, reinterpret_cast<uint32_t *>(base) + i * 64, reinterpret_cast<uint2 *>(base + K) + i * 4);
}

// Work decomposition: a UNIT is (tile of `tile` <= PS_RW_WARPS octets, column group of C columns); units are dealt
// round-robin to the persistent CTAs, and inside a unit warp w < tile owns octet w of the tile.  The host picks the
// smallest tile that still fits the units into one wave of CTAs: a narrow batch over a 4096-row matrix then runs 4 walking
// warps on each of 128 SMs (one per scheduler) instead of 16 on 32 SMs - the walk is issue-bound per scheduler.
// A warp's weight stream is simply the concatenation of its octets over the CTA's units.
template <int C>
__global__ void __launch_bounds__(PS_RW_THREADS + 32, 1) ps_k_rw_matmul(const PsRwmArgs a) {
    extern __shared__ __align__(128) uint8_t ps_rw_smem[];
    __shared__ __align__(8) uint64_t xbar;
    const int K = a.K, nb = K / 256, kb = a.kb, ns = a.ns;
    const uint32_t img_bytes = (uint32_t)K + (uint32_t)nb * 32;
    const uint32_t stage_bytes = (uint32_t)kb * PS_RW_OCTET_BLOCK;
    uint8_t *s_act = ps_rw_smem;                                   // [C][img_bytes]
    uint8_t *s_ring = ps_rw_smem + (size_t)C * img_bytes;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + (size_t)PS_RW_WARPS * ns * stage_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, q = lane & 3;
    const int spo = nb / kb;
    const int tile = a.tile;
    const int n_cg = (a.bs + C - 1) / C, n_tiles = (a.n_oct + tile - 1) / tile;
    const int n_units = n_cg * n_tiles;                            // unit u = (tile u / n_cg, column group u % n_cg)
    const int my_units = (n_units > (int)blockIdx.x) ? (n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    // stage #s of warp w: unit s / spo of this CTA (octets beyond n_oct are skipped by both sides: they own no stages)
    auto issue = [&](int w, int s) {
        const int u = (int)blockIdx.x + (s / spo) * (int)gridDim.x;
        const int oct = (u / n_cg) * tile + w;
        uint64_t *bar = s_bar + w * ns + (s % ns);
        uint8_t *dst = s_ring + ((size_t)w * ns + (s % ns)) * stage_bytes;
        ps_mbar_expect_tx(bar, stage_bytes);
        for (int b = 0; b < kb; b++) {
            const int oo = min(oct, a.n_oct - 1); // a ragged last tile re-reads a valid octet; its results are discarded
            const uint8_t *src = a.w + (((size_t)oo * nb + (size_t)(s % spo) * kb + b) * a.n_slots + a.slot) * PS_RW_OCTET_BLOCK;
            ps_bulk_g2s(dst + (size_t)b * PS_RW_OCTET_BLOCK, src, PS_RW_OCTET_BLOCK, bar);
        }
    };
    const int n_stages = my_units * spo;
    if (warp == PS_RW_WARPS) { // helper warp: barrier init + first ring fill (lane w serves warp w)
        if (lane < tile)
            for (int s = 0; s < ns; s++) ps_mbar_init(s_bar + lane * ns + s, 1);
        if (lane == 0) ps_mbar_init(&xbar, 1);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        asm volatile("bar.arrive 1, %0;" ::"n"(PS_RW_THREADS + 32) : "memory");
        if (lane < tile)
            for (int s = 0; s < ns && s < n_stages; s++) issue(lane, s);
        return;
    }
    ps_bar_sync(1, PS_RW_THREADS + 32);
    uint8_t *my_ring = s_ring + (size_t)warp * ns * stage_bytes;
    uint64_t *my_bar = s_bar + warp * ns;
    int s = 0;
#pragma unroll 1
    for (int mu = 0; mu < my_units; mu++) {
        const int u = (int)blockIdx.x + mu * (int)gridDim.x;
        const int col0 = (u % n_cg) * C, ncol = min(C, a.bs - col0);
        const int oct = (u / n_cg) * tile + warp;
        ps_bar_sync(2, PS_RW_THREADS); // everyone is done with the previous unit's images
        if (tid == 0) {
            ps_mbar_expect_tx(&xbar, (uint32_t)ncol * img_bytes);
            for (int c = 0; c < ncol; c++) ps_bulk_g2s(s_act + (size_t)c * img_bytes, a.x_img + (size_t)(col0 + c) * img_bytes, img_bytes, &xbar);
        }
        ps_mbar_wait(&xbar, mu & 1);
        if (warp < tile) {
            PsRwAcc acc[C];
#pragma unroll
            for (int c = 0; c < C; c++) acc[c].a0 = acc[c].a1 = acc[c].am = 0.f;
#pragma unroll 1
            for (int sb = 0; sb < spo; sb++, s++) {
                const int slot = s % ns;
                const int moff = ps_rw_mins_off(r, q);
                ps_mbar_wait(&my_bar[slot], (s / ns) & 1);
                const uint8_t *st = my_ring + (size_t)slot * stage_bytes;
#pragma unroll 1
                for (int b = 0; b < kb; b++) {
                    const int i = sb * kb + b;
                    const uint8_t *ob = st + (size_t)b * PS_RW_OCTET_BLOCK;
                    // ---- unpack the weight block once (see ps_rw_block)
                    const uint4 h = *reinterpret_cast<const uint4 *>(ob + 16 * r);
                    const uint32_t mp = *reinterpret_cast<const uint16_t *>(ob + moff);
                    uint32_t lo[8], hi[8];
                    int s_lo[4], s_hi[4];
#pragma unroll
                    for (int j2 = 0; j2 < 4; j2++) {
                        const uint2 w = *reinterpret_cast<const uint2 *>(ob + PS_RW_QS + 256 * j2 + 32 * r + 8 * q);
                        lo[2 * j2] = w.x & 0x0f0f0f0fu; lo[2 * j2 + 1] = w.y & 0x0f0f0f0fu;
                        hi[2 * j2] = w.x & 0xf0f0f0f0u; hi[2 * j2 + 1] = w.y & 0xf0f0f0f0u;
                        const uint32_t scw = (j2 < 2) ? h.y : h.z;
                        s_lo[j2] = ps_rw_byte(scw, 2 * (j2 & 1));
                        s_hi[j2] = ps_rw_byte(scw, 2 * (j2 & 1) + 1);
                    }
                    const float xd = ps_half_bits_to_float(h.x & 0xffffu), xmin = ps_half_bits_to_float(h.x >> 16);
                    // ---- every column of the group
#pragma unroll
                    for (int c = 0; c < C; c++) {
                        if (c < ncol) {
                            const uint4 *qa = reinterpret_cast<const uint4 *>(s_act + (size_t)c * img_bytes) + (size_t)i * 16;
                            const uint2 meta = reinterpret_cast<const uint2 *>(s_act + (size_t)c * img_bytes + K)[i * 4 + q];
                            int S0 = 0, S1 = 0, H0 = 0, H1 = 0;
#pragma unroll
                            for (int j2 = 0; j2 < 4; j2++) {
                                const uint4 av = qa[j2 * 4 + q];
                                S0 += s_lo[j2] * __dp4a((int)lo[2 * j2], (int)av.x, 0);
                                S1 += s_lo[j2] * __dp4a((int)lo[2 * j2 + 1], (int)av.y, 0);
                                H0 += s_hi[j2] * ps_dp4a_us(hi[2 * j2], (int)av.z, 0);
                                H1 += s_hi[j2] * ps_dp4a_us(hi[2 * j2 + 1], (int)av.w, 0);
                            }
                            S0 += H0 >> 4;
                            S1 += H1 >> 4;
                            const int P = ps_rw_mins_dot(meta.y, mp);
                            const float yd = __uint_as_float(meta.x);
                            const float d = __fmul_rn(yd, xd), dm = __fmul_rn(-yd, xmin);
                            acc[c].a0 = __fmaf_rn(d, __int2float_rn(S0), acc[c].a0);
                            acc[c].a1 = __fmaf_rn(d, __int2float_rn(S1), acc[c].a1);
                            acc[c].am = __fmaf_rn(dm, __int2float_rn(P), acc[c].am);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0 && s + ns < n_stages) issue(warp, s + ns);
            }
            // ---- epilogue: bias / residual, dst[col][row]
            const int row = oct * 8 + r;
            int sg = 0;
            if (a.n_seg > 1 && row >= a.seg[1].row_begin) sg = 1;
            if (a.n_seg > 2 && row >= a.seg[2].row_begin) sg = 2;
            const bool live = row < a.seg[sg].row_end;
            const int n = row - a.seg[sg].row_begin, ld = a.seg[sg].row_end - a.seg[sg].row_begin;
            const float bias = (live && a.seg[sg].bias) ? a.seg[sg].bias[n] : 0.f;
#pragma unroll
            for (int c = 0; c < C; c++) {
                if (c < ncol) {
                    float res = ps_rw_row_result(acc[c]);
                    if (q == 0 && live) {
                        const size_t o = (size_t)(col0 + c) * ld + n;
                        if (a.seg[sg].bias) res = __fadd_rn(res, bias);
                        if (a.residual) res = __fadd_rn(a.residual[o], res);
                        a.seg[sg].dst[o] = res;
                    }
                }
            }
        }
    }
}
