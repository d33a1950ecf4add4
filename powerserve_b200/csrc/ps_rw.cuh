// ps_rw.cuh — "row-walker" Q4_K mat-vec for decode (bs = 1): the fused decode path's weight-streaming kernel.
//
// Replaces powerserve_compute_forward_mul_mat -> ggml_vec_dot_q4_K_q8_K (libs/ggml/src/ggml.c:13344-13432,
// ggml-quants.c:7809-7872) for one activation column, with the RMSNorm + Q8_K activation quantisation of the op before
// it (ggml.c:12667-12721, ggml-quants.c:3799-3837) as prologue and the bias / residual / SiLU*up ops after it
// (ggml.c:10042-10112, src/backend/ggml/ggml.cpp:115-129) as epilogue.  Bit-identical to the table-op kernels.
//
// Why this shape (B200): the op is an HBM stream of 144-byte blocks with ~250 integer instructions of work per block —
// at 6.5 TB/s an SM must retire a block every ~6 cycles, so the kernel is as much issue-bound as bandwidth-bound.  The
// earlier design (one thread per block, shared-memory hand-off to separate fp32 chain threads, CTA-wide barriers per
// tile, 8 warps) measured 20-27 % issue utilisation and 3x instruction overhead (profiles/r01b_*).  Here:
//   * FOUR threads own one weight row: thread q of the quad owns AVX lanes 2q, 2q+1 of the reference's __m256 accumulator
//     (and lane q of the __m128 mins accumulator) and walks the row's super-blocks IN ORDER, so the fp32 FMA chains of
//     the reference live in three registers per thread — no hand-off, no chain phase, no barrier inside the stream.
//   * a warp owns an OCTET of rows.  Weights are re-laid at bind time (same bytes, permuted) so that the eight rows'
//     16-byte headers and 32-byte quant groups of one super-block are adjacent: a warp's LDS are conflict-free and a
//     pipeline stage (kb super-blocks of an octet) is ONE contiguous bulk copy.
//   * every warp runs its own TMA ring (cp.async.bulk + mbarrier, lane 0 re-arms a slot right after the warp drained
//     it), so warps never wait for each other; the first slots are requested before griddepcontrol.wait, i.e. while the
//     previous kernel in the PDL chain is still draining.
#pragma once
#include "ps_decode.cuh"

#define PS_RW_WARPS 16
#define PS_RW_THREADS (PS_RW_WARPS * 32)
#define PS_RW_OCTET_BLOCK 1152                 // 8 rows x 144 bytes
#define PS_RW_MAX_NS 8

enum { PS_RW_OUT_PLAIN = 0, PS_RW_OUT_ROPE = 1, PS_RW_OUT_ROPE_KCACHE = 2, PS_RW_OUT_VCACHE_T = 3 };

struct PsRwSeg {
    float *dst;         // output rows of this segment (indexed by row - row_begin)
    const float *bias;  // optional
    int row_begin, row_end;
    int mode;           // PS_RW_OUT_*: the q / k / v epilogues fuse ROPE and the two KV-cache COPY ops (norm_attention.cpp:72-105)
};

struct PsRwArgs {
    const uint8_t *w;      // repacked weights: [n_oct][nb][rpt][1152]
    int n_oct;             // row octets (pairs of octets when rpt == 2)
    int K;                 // contraction length, multiple of 256
    int kb;                // super-blocks per pipeline stage
    int ns;                // stages per warp ring
    int n_act;             // warps of a CTA that own octets (ring slots exist only for these)
    int pre;               // ring slots requested before griddepcontrol.wait; the rest follow once x has been read, so the
                           // activation loads do not queue behind ~200 KB per SM of weight prefetch
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
    double inv_k;          // 1 / K when K is a power of two (the mean is then an exact scaling), else 0
    float *part_val;       // optional (lm_head): per-CTA partial arg-max of the produced rows, [gridDim.x]
    int *part_idx;
    long long *tl;         // optional timeline slot (option "trace")
};

// ---------------------------------------------------------------------------------------------------- repack
// GGUF rows [n_rows][nb][144] -> [octet][block][slot][1152] where 1152 = 8 headers (16 B) + 4 groups x 8 rows x 32 B.
// `slot`/`n_slots` interleave several matrices per octet-block (gate | up).  Rows beyond n_rows are zero blocks.
__global__ void ps_k_rw_repack(uint8_t *__restrict__ dst, const uint8_t *__restrict__ src, int64_t n_rows, int64_t nb, int64_t oct0, int slot,
                               int n_slots) {
    const int64_t n_oct = (n_rows + 7) / 8;
    const int64_t total = n_oct * nb * 72; // 16-byte chunks
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int c = (int)(t % 9);
        const int r = (int)((t / 9) % 8);
        const int64_t i = (t / 72) % nb;
        const int64_t o = t / (72 * nb);
        const int64_t row = o * 8 + r;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (row < n_rows) v = *reinterpret_cast<const uint4 *>(src + (row * nb + i) * PS_Q4_K_BYTES + 16 * c);
        uint8_t *ob = dst + (((oct0 + o) * nb + i) * n_slots + slot) * PS_RW_OCTET_BLOCK;
        const int off = (c == 0) ? 16 * r : 128 + 256 * ((c - 1) >> 1) + 32 * r + 16 * ((c - 1) & 1);
        *reinterpret_cast<uint4 *>(ob + off) = v;
    }
}

// ---------------------------------------------------------------------------------------------------- block math
// One super-block of one row for the quad thread q: the two int32 lanes (2q, 2q+1) of `sumi` and lane q of `prod`
// exactly as one iteration of the AVX2 loop leaves them (ggml-quants.c:7828-7860), then the three FMAs (:7858, :7834).
struct PsRwAcc {
    float a0, a1, am;
};

PS_D void ps_rw_block(const uint8_t *ob, int r, int q, const uint4 *qa, const uint2 meta, PsRwAcc &acc) {
    const uint4 h = *reinterpret_cast<const uint4 *>(ob + 16 * r);
    const uint32_t k1 = 0x3f3f3f3fu, k2 = 0x0f0f0f0fu, k3 = 0x03030303u;
    const uint32_t scA = h.y & k1, scB = (h.w & k2) | (((h.y >> 6) & k3) << 4);   // utmp shuffle, :7816-7826
    const uint32_t mA = h.z & k1, mB = ((h.w >> 4) & k2) | (((h.z >> 6) & k3) << 4);
    int S0 = 0, S1 = 0, H0 = 0, H1 = 0;
#pragma unroll
    for (int j2 = 0; j2 < 4; j2++) {
        const uint2 w = *reinterpret_cast<const uint2 *>(ob + 128 + 256 * j2 + 32 * r + 8 * q);
        const uint4 a = qa[j2 * 4 + q];
        const uint32_t scw = (j2 < 2) ? scA : scB;
        const int s_lo = (scw >> (16 * (j2 & 1))) & 0xff, s_hi = (scw >> (16 * (j2 & 1) + 8)) & 0xff;
        S0 += s_lo * __dp4a((int)(w.x & 0x0f0f0f0fu), (int)a.x, 0);
        S1 += s_lo * __dp4a((int)(w.y & 0x0f0f0f0fu), (int)a.y, 0);
        H0 += s_hi * ps_dp4a_us(w.x & 0xf0f0f0f0u, (int)a.z, 0);   // 16 x the high-nibble dot
        H1 += s_hi * ps_dp4a_us(w.y & 0xf0f0f0f0u, (int)a.w, 0);
    }
    S0 += H0 >> 4;
    S1 += H1 >> 4;
    const uint32_t mw = (q < 2) ? mA : mB;
    const int m0 = (mw >> (16 * (q & 1))) & 0xff, m1 = (mw >> (16 * (q & 1) + 8)) & 0xff;
    const int P = m0 * (int)(short)(meta.y & 0xffffu) + m1 * (int)(short)(meta.y >> 16);
    const float yd = __uint_as_float(meta.x);
    const float d = __fmul_rn(yd, ps_half_bits_to_float(h.x & 0xffffu));
    const float dm = __fmul_rn(-yd, ps_half_bits_to_float(h.x >> 16));
    acc.a0 = __fmaf_rn(d, __int2float_rn(S0), acc.a0);
    acc.a1 = __fmaf_rn(d, __int2float_rn(S1), acc.a1);
    acc.am = __fmaf_rn(dm, __int2float_rn(P), acc.am);
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
// Dynamic shared memory: [qa: K bytes][meta: nb x 4 x 8][rings: n_act x ns x stage_bytes][bars: n_act x ns x 8]
template <int EPI>
__global__ void __launch_bounds__(PS_RW_THREADS, 1) ps_k_rw_matvec(const PsRwArgs a) {
    constexpr int RPT = (EPI == PS_EPI_SILU) ? 2 : 1;
    extern __shared__ __align__(128) uint8_t ps_rw_smem[];
    __shared__ double sh_red[PS_RW_WARPS];
    const int K = a.K, nb = K / 256, kb = a.kb, ns = a.ns;
    const uint32_t stage_bytes = (uint32_t)kb * RPT * PS_RW_OCTET_BLOCK;
    uint4 *s_qa = reinterpret_cast<uint4 *>(ps_rw_smem);
    uint2 *s_meta = reinterpret_cast<uint2 *>(ps_rw_smem + K);
    uint8_t *s_ring = ps_rw_smem + K + (size_t)nb * 32;
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + (size_t)a.n_act * ns * stage_bytes);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, q = lane & 3;
    // this CTA's octets [o0, o1); warp w walks o0 + w, o0 + w + n_act, ...
    const int o0 = (int)(((long long)blockIdx.x * a.n_oct) / gridDim.x), o1 = (int)(((long long)(blockIdx.x + 1) * a.n_oct) / gridDim.x);
    const int n_mine = (warp < a.n_act && o0 + warp < o1) ? (o1 - o0 - warp - 1) / a.n_act + 1 : 0;
    const int spo = nb / kb;                 // stages per octet
    const int n_stages = n_mine * spo;
    uint8_t *my_ring = s_ring + (size_t)warp * ns * stage_bytes;
    uint64_t *my_bar = s_bar + warp * ns;
    const size_t oct_bytes = (size_t)nb * RPT * PS_RW_OCTET_BLOCK;

    auto issue = [&](int s) { // lane 0: request stage #s of this warp's stream into slot s % ns
        const int oct = o0 + warp + (s / spo) * a.n_act;
        const uint8_t *src = a.w + (size_t)oct * oct_bytes + (size_t)(s % spo) * stage_bytes;
        const int slot = s % ns;
        ps_mbar_expect_tx(&my_bar[slot], stage_bytes);
        ps_bulk_g2s(my_ring + (size_t)slot * stage_bytes, src, stage_bytes, &my_bar[slot]);
    };
    ps_tl_min(a.tl, 0);
    // the norm weights do not depend on the previous kernel either: request them before the weight prefetch floods HBM
    float wv0[8];
    const bool early_w = a.norm_w != nullptr && warp < nb;
    if (early_w) {
        const float4 w0 = *reinterpret_cast<const float4 *>(a.norm_w + warp * 256 + 4 * lane);
        const float4 w1 = *reinterpret_cast<const float4 *>(a.norm_w + warp * 256 + 128 + 4 * lane);
        wv0[0] = w0.x; wv0[1] = w0.y; wv0[2] = w0.z; wv0[3] = w0.w; wv0[4] = w1.x; wv0[5] = w1.y; wv0[6] = w1.z; wv0[7] = w1.w;
    }
    if (n_stages > 0 && lane == 0) {
        for (int s = 0; s < ns; s++) ps_mbar_init(&my_bar[s], 1);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < a.pre && s < n_stages; s++) issue(s); // weights never depend on the previous kernel
    }
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    ps_tl_min(a.tl, 2);
    const long long t_dep = (a.tl && tid == 0) ? ps_globaltimer() : 0;

    // ---- prologue: (RMSNorm) + Q8_K quantisation of x into shared memory, one warp per 256-block
    {
        float e[4][8];
        double ss = 0.0;
        const int per_warp = (nb + PS_RW_WARPS - 1) / PS_RW_WARPS; // <= 4 (K <= 16384)
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = warp + u * PS_RW_WARPS;
            if (u < per_warp && i < nb) {
                const float4 v0 = *reinterpret_cast<const float4 *>(a.x + i * 256 + 4 * lane);
                const float4 v1 = *reinterpret_cast<const float4 *>(a.x + i * 256 + 128 + 4 * lane);
                e[u][0] = v0.x; e[u][1] = v0.y; e[u][2] = v0.z; e[u][3] = v0.w;
                e[u][4] = v1.x; e[u][5] = v1.y; e[u][6] = v1.z; e[u][7] = v1.w;
                if (a.norm_w) {
#pragma unroll
                    for (int t = 0; t < 8; t++) ss += (double)__fmul_rn(e[u][t], e[u][t]);
                }
            }
        }
        // the rest of the ring, once this warp's x values have arrived (the compare consumes a loaded register)
        if (n_stages > 0 && lane == 0 && (warp >= nb || __float_as_uint(e[0][0]) != 0xffc0dead))
            for (int s = a.pre; s < ns && s < n_stages; s++) issue(s);
        float nscale = 1.f;
        if (a.norm_w) {
#pragma unroll
            for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(PS_FULL, ss, o);
            if (lane == 0) sh_red[warp] = ss;
            __syncthreads();
            double t = (lane < PS_RW_WARPS) ? sh_red[lane] : 0.0;
#pragma unroll
            for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(PS_FULL, t, o);
            const float mean = (float)(a.inv_k != 0.0 ? t * a.inv_k : t / (double)K);
            nscale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, a.eps)));
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
            const int i = warp + u * PS_RW_WARPS;
            if (u < per_warp && i < nb) {
                if (a.norm_w) {
                    float wv[8];
                    if (u == 0) {
#pragma unroll
                        for (int t = 0; t < 8; t++) wv[t] = wv0[t];
                    } else {
                        const float4 w0 = *reinterpret_cast<const float4 *>(a.norm_w + i * 256 + 4 * lane);
                        const float4 w1 = *reinterpret_cast<const float4 *>(a.norm_w + i * 256 + 128 + 4 * lane);
                        wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
                    }
#pragma unroll
                    for (int t = 0; t < 8; t++) e[u][t] = __fmul_rn(e[u][t], __fmul_rn(wv[t], nscale)); // y = x * (w * scale)
                }
                // natural-order words: lane L -> (sub-block L/8, word L%8) and (sub-block 4 + L/8, word L%8)
                uint32_t words[2];
                float yd;
                uint32_t bsp4;
                ps_quant_block_q8k_regs(e[u], lane, words, yd, bsp4);
                uint32_t *qw = reinterpret_cast<uint32_t *>(s_qa) + (size_t)i * 64;
                const int l = lane & 7, jA = lane >> 3, jB = 4 + (lane >> 3);
                qw[((jA >> 1) * 4 + (l >> 1)) * 4 + (jA & 1) * 2 + (l & 1)] = words[0];
                qw[((jB >> 1) * 4 + (l >> 1)) * 4 + (jB & 1) * 2 + (l & 1)] = words[1];
                if (lane < 4) s_meta[i * 4 + lane] = make_uint2(__float_as_uint(yd), bsp4);
            }
        }
    }
    __syncthreads();
    if (a.tl && tid == 0) atomicMax(reinterpret_cast<unsigned long long *>(a.tl + 3), (unsigned long long)(ps_globaltimer() - t_dep)); // slowest CTA's prologue

    // ---- the stream
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    int s = 0;
    for (int m = 0; m < n_mine; m++) {
        const int oct = o0 + warp + m * a.n_act;
        PsRwAcc acc[RPT];
#pragma unroll
        for (int t = 0; t < RPT; t++) acc[t].a0 = acc[t].a1 = acc[t].am = 0.f;
        for (int sb = 0; sb < spo; sb++, s++) {
            const int slot = s % ns;
            ps_mbar_wait(&my_bar[slot], (s / ns) & 1);
            const uint8_t *st = my_ring + (size_t)slot * stage_bytes;
            for (int b = 0; b < kb; b++) {
                const int i = sb * kb + b;
                const uint4 *qa = s_qa + (size_t)i * 16;
                const uint2 meta = s_meta[i * 4 + q];
#pragma unroll
                for (int t = 0; t < RPT; t++) ps_rw_block(st + (size_t)(b * RPT + t) * PS_RW_OCTET_BLOCK, r, q, qa, meta, acc[t]);
            }
            __syncwarp();
            if (lane == 0 && s + ns < n_stages) issue(s + ns); // the slot is drained: re-arm it
        }
        // ---- epilogue
        const int row = oct * 8 + r;
        if (EPI == PS_EPI_SILU) {
            const float g = ps_rw_row_result(acc[0]);
            const float u = ps_rw_row_result(acc[RPT - 1]);
            if (q == 0 && row < a.seg[0].row_end) a.seg[0].dst[row] = ps_silu_mul(g, u);
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
                if (res > best_v || (res == best_v && n < best_i)) { best_v = res; best_i = n; } // first maximum wins
            }
        }
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
        __syncthreads();
        if (tid == 0) {
            for (int t = 1; t < PS_RW_WARPS; t++)
                if (sv[t] > best_v || (sv[t] == best_v && si[t] < best_i)) { best_v = sv[t]; best_i = si[t]; }
            a.part_val[blockIdx.x] = best_v;
            a.part_idx[blockIdx.x] = best_i;
        }
    }
    __syncthreads();
    ps_tl_max(a.tl, 1);
}
