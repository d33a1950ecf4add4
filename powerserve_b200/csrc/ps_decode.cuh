// ps_decode.cuh — the fused decode path (bs = 1): a persistent, TMA-fed Q4_K mat-vec with fused prologue
// (RMSNorm + Q8_K activation quantisation) and epilogue (bias / residual / SiLU·up), bit-identical to the table-op
// kernels in ps_kernels.cuh (tests/test_gpu_decode.py proves it op by op and end to end).
//
// Design (B200): the kernel is HBM-bound integer work, so the SM's job is to keep ~64 KB of weight bytes in flight and
// to spend as few issue slots per byte as possible.
//   * weights: one elected thread streams contiguous row tiles (R rows x K/256 x 144 B, 16-B aligned) from HBM into a
//     4-stage shared-memory ring with cp.async.bulk (TMA, 1-D) completing on mbarriers; the first stages are issued
//     BEFORE griddepcontrol.wait, so with programmatic dependent launch the weight stream of kernel N+1 starts while
//     kernel N drains (weights never depend on activations).
//   * integer work: ONE thread owns ONE 144-byte Q4_K block of a row (9 x LDS.128, conflict-free at the 144-B stride)
//     and keeps the Q8_K activation block it pairs with in 64 registers for the whole kernel (thread t always meets
//     block index t % nb).  It produces exactly what one AVX2 iteration of ggml_vec_dot_q4_K_q8_K leaves in the eight
//     int32 lanes of `sumi` and the four lanes of `prod` (ggml-quants.c:7828-7860): 64 dp4a + 64 IMAD per block.
//   * fp32 chains: 12 threads per row (8 acc lanes + 4 acc_m lanes) replay the reference's per-block FMAs in row order
//     from a shared-memory hand-off buffer, then reduce in hsum_float_8 order (ggml-quants.c:7862-7871).
#pragma once
#include "ps_kernels.cuh"

#define PS_MV_THREADS 288          // 8 compute warps + 1 producer warp
#define PS_MV_COMPUTE 256
#define PS_MV_STAGES 4

enum { PS_EPI_STORE = 0, PS_EPI_RESIDUAL = 1, PS_EPI_SILU = 2 };

struct PsMvSeg {
    const uint8_t *w;  // Q4_K rows, row-major
    float *dst;        // output vector of this segment
    const float *bias; // optional (Qwen2 q/k/v)
    int n_rows;
    int tile0;         // first tile index of the segment
};

struct PsMvArgs {
    PsMvSeg seg[3];
    int n_seg;
    int n_tiles;
    int K;                 // contraction length (multiple of 256)
    int R;                 // rows per tile / pass
    const float *x;        // fp32 activation [K]
    const float *norm_w;   // non-null: xn = rmsnorm(x) * norm_w is what gets quantised
    float eps;
    const float *residual; // PS_EPI_RESIDUAL: dst[n] = residual[n] + r
    int epi;
    long long *trace;      // optional per-CTA timestamps (globaltimer ns), 16 slots per CTA; nullptr in production
};

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
#define PS_TRACE(slot)                                                                   \
    do {                                                                                 \
        if (a.trace && (threadIdx.x & 31) == 0 && (threadIdx.x >> 5) == 0) a.trace[blockIdx.x * 16 + (slot)] = ps_globaltimer(); \
    } while (0)
PS_D int ps_dp4a_us(uint32_t a, int b, int c) { // unsigned bytes of a  x  signed bytes of b
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// ---------------------------------------------------------------------------------------------------- prologue
// quantize_row_q8_K_ref (ggml-quants.c:3799-3837) of one 256-block by one warp; e[0..3] = elements 4*lane..+3,
// e[4..7] = elements 128+4*lane..+3.  Writes the natural-order words (word w = elements 4w..4w+3), d and the four
// int16 pairs of sub-block sums.
PS_D void ps_quant_block_q8k_warp(const float e[8], int lane, uint32_t *qs_words, float *d_out, uint32_t *bsp_out) {
    const int idx0 = 4 * lane, idx1 = 128 + 4 * lane;
    float amax = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++) amax = fmaxf(amax, fabsf(e[t]));
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(PS_FULL, amax, o));
    int first = 1 << 20;
#pragma unroll
    for (int t = 7; t >= 0; t--)
        if (fabsf(e[t]) == amax) first = (t < 4 ? idx0 + t : idx1 + t - 4);
#pragma unroll
    for (int o = 16; o; o >>= 1) first = min(first, __shfl_xor_sync(PS_FULL, first, o));
    float mx = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++)
        if ((t < 4 ? idx0 + t : idx1 + t - 4) == first) mx = e[t];
    const unsigned owner = __ballot_sync(PS_FULL, (first >= idx0 && first < idx0 + 4) || (first >= idx1 && first < idx1 + 4));
    mx = __shfl_sync(PS_FULL, mx, __ffs(owner) - 1);
    if (amax == 0.f) {
        qs_words[lane] = 0;
        qs_words[32 + lane] = 0;
        if (lane == 0) *d_out = 0.f;
        if (lane < 4) bsp_out[lane] = 0;
        return;
    }
    const float iscale = __fdiv_rn(-127.f, mx);
    int q[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const float val = __fadd_rn(__fmul_rn(iscale, e[t]), 12582912.f);
        q[t] = min(127, (int)(ps_f2u(val) & 0x007fffffu) - 0x00400000);
    }
    qs_words[lane] = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    qs_words[32 + lane] = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    int s0 = q[0] + q[1] + q[2] + q[3], s1 = q[4] + q[5] + q[6] + q[7]; // sub-blocks lane/8 and 4 + lane/8
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        s0 += __shfl_xor_sync(PS_FULL, s0, o);
        s1 += __shfl_xor_sync(PS_FULL, s1, o);
    }
    int sj[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        sj[j] = __shfl_sync(PS_FULL, s0, j * 8);
        sj[j + 4] = __shfl_sync(PS_FULL, s1, j * 8);
    }
    if (lane < 4) bsp_out[lane] = ((uint32_t)sj[2 * lane] & 0xffffu) | ((uint32_t)sj[2 * lane + 1] << 16);
    if (lane == 0) *d_out = __fdiv_rn(1.f, iscale);
}

// ---------------------------------------------------------------------------------------------------- the kernel
// Dynamic shared memory carve-up (bytes):
//   [stages]   PS_MV_STAGES x stage_bytes, stage_bytes = R * nb * 144 (PS_EPI_SILU: R/2 gate rows then R/2 up rows)
//   [q8]       K                                   quantised activation, natural word order
//   [yd]       nb x 4, [bsp] nb x 16
//   [chain]    2 x R x (nb*16 + 16) x 4            per block: S[8], P[4], d, dmin, pad x2 (double-buffered hand-off)
//   [res]      R x 4
//   [bars]     2 x PS_MV_STAGES x 8
__global__ void __launch_bounds__(PS_MV_THREADS, 1) ps_k_matvec_q4k_tma(const PsMvArgs a) {
    extern __shared__ __align__(128) uint8_t ps_mv_smem[];
    uint8_t *smem = ps_mv_smem;
    const int K = a.K, nb = K / 256, R = a.R;
    const uint32_t row_bytes = (uint32_t)nb * PS_Q4_K_BYTES;
    const uint32_t stage_bytes = (uint32_t)R * row_bytes;
    uint8_t *s_stage = smem;
    uint32_t *s_q8 = reinterpret_cast<uint32_t *>(smem + (size_t)PS_MV_STAGES * stage_bytes);
    float *s_yd = reinterpret_cast<float *>(s_q8 + K / 4);
    uint32_t *s_bsp = reinterpret_cast<uint32_t *>(s_yd + ((nb + 3) & ~3));
    const int chain_row = nb * 16 + 16;
    uint32_t *s_chain = s_bsp + nb * 4;                                       // 16-byte aligned: every term above is
    float *s_res = reinterpret_cast<float *>(s_chain + (size_t)2 * R * chain_row);             // two hand-off buffers
    uint64_t *s_full = reinterpret_cast<uint64_t *>(s_res + ((R + 3) & ~3));
    uint64_t *s_empty = s_full + PS_MV_STAGES;
    __shared__ double sh_red[32];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_producer = (warp == PS_MV_COMPUTE / 32);
    const int my_tiles = (a.n_tiles > (int)blockIdx.x) ? (a.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int half = R / 2;

    // tile -> (segment, first row, rows) ; SILU tiles cover `half` gate rows + the same `half` up rows
    auto tile_info = [&](int tile, int &seg, int &row0, int &nrows) {
        if (a.epi == PS_EPI_SILU) {
            seg = 0;
            row0 = tile * half;
            nrows = min(half, a.seg[0].n_rows - row0);
        } else {
            seg = 0;
            if (a.n_seg > 1 && tile >= a.seg[1].tile0) seg = 1;
            if (a.n_seg > 2 && tile >= a.seg[2].tile0) seg = 2;
            row0 = (tile - a.seg[seg].tile0) * R;
            nrows = min(R, a.seg[seg].n_rows - row0);
        }
    };
    auto issue_tile = [&](int k) { // producer lane 0: stream tile #k of this CTA into stage k % STAGES
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int st = k % PS_MV_STAGES;
        int seg, row0, nrows;
        tile_info(tile, seg, row0, nrows);
        uint8_t *dst = s_stage + (size_t)st * stage_bytes;
        if (a.epi == PS_EPI_SILU) {
            const uint32_t bytes = (uint32_t)nrows * row_bytes;
            ps_mbar_expect_tx(&s_full[st], 2 * bytes);
            ps_bulk_g2s(dst, a.seg[0].w + (size_t)row0 * row_bytes, bytes, &s_full[st]);
            ps_bulk_g2s(dst + (size_t)half * row_bytes, a.seg[1].w + (size_t)row0 * row_bytes, bytes, &s_full[st]);
        } else {
            const uint32_t bytes = (uint32_t)nrows * row_bytes;
            ps_mbar_expect_tx(&s_full[st], bytes);
            ps_bulk_g2s(dst, a.seg[seg].w + (size_t)row0 * row_bytes, bytes, &s_full[st]);
        }
    };

    PS_TRACE(0);
    if (tid == 0) {
        for (int s = 0; s < PS_MV_STAGES; s++) {
            ps_mbar_init(&s_full[s], 1);
            ps_mbar_init(&s_empty[s], PS_MV_COMPUTE / 32);
        }
        ps_fence_barrier_init();
    }
    __syncthreads();

    if (is_producer) {
        // ===== producer warp: weights do not depend on the previous kernel -> start streaming immediately
        if (lane == 0) {
            int k = 0;
            for (; k < my_tiles && k < PS_MV_STAGES; k++) issue_tile(k);
            for (; k < my_tiles; k++) {
                const int st = k % PS_MV_STAGES;
                ps_mbar_wait(&s_empty[st], ((k / PS_MV_STAGES) - 1) & 1);
                issue_tile(k);
            }
        }
        return;
    }

    // ===== compute warps
    PS_TRACE(1);
    ps_grid_dep_wait();
    PS_TRACE(2);        // activations come from the previous kernel in the stream / graph
    ps_grid_dep_launch();      // let the next kernel's CTAs start their weight prefetch as SMs free up
    // ---- prologue: (RMSNorm) + Q8_K quantisation of the activation vector, redundantly per CTA (K <= 14336 floats)
    float nscale = 1.f;
    if (a.norm_w) {
        double s = 0.0;
        for (int e = tid; e < K; e += PS_MV_COMPUTE) s += (double)__fmul_rn(a.x[e], a.x[e]);
#pragma unroll
        for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(PS_FULL, s, o);
        if (lane == 0) sh_red[warp] = s;
        ps_bar_sync(1, PS_MV_COMPUTE);
        double t = (lane < PS_MV_COMPUTE / 32) ? sh_red[lane] : 0.0;
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(PS_FULL, t, o);
        const float mean = (float)(t / (double)K);
        nscale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, a.eps)));
    }
    for (int i = warp; i < nb; i += PS_MV_COMPUTE / 32) {
        const float *xb = a.x + i * 256;
        float4 v0 = *reinterpret_cast<const float4 *>(xb + 4 * lane);
        float4 v1 = *reinterpret_cast<const float4 *>(xb + 128 + 4 * lane);
        float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        if (a.norm_w) {
            const float4 w0 = *reinterpret_cast<const float4 *>(a.norm_w + i * 256 + 4 * lane);
            const float4 w1 = *reinterpret_cast<const float4 *>(a.norm_w + i * 256 + 128 + 4 * lane);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int t = 0; t < 8; t++) e[t] = __fmul_rn(e[t], __fmul_rn(wv[t], nscale)); // y = x * (w * scale), ggml.c:2466
        }
        ps_quant_block_q8k_warp(e, lane, s_q8 + i * 64, s_yd + i, s_bsp + i * 4);
    }
    ps_bar_sync(1, PS_MV_COMPUTE);

    PS_TRACE(3);
    // ---- this thread's fixed (row-in-tile, block) slot and its activation block in registers
    const int r_slot = tid / nb, i_blk = tid % nb;
    const bool active = tid < R * nb;
    uint32_t q8[64];
    float yd = 0.f;
    uint32_t bsp[4] = {0, 0, 0, 0};
    if (active) {
        const uint4 *src = reinterpret_cast<const uint4 *>(s_q8 + i_blk * 64);
#pragma unroll
        for (int u = 0; u < 16; u++) {
            const uint4 v = src[u];
            q8[4 * u + 0] = v.x; q8[4 * u + 1] = v.y; q8[4 * u + 2] = v.z; q8[4 * u + 3] = v.w;
        }
        yd = s_yd[i_blk];
#pragma unroll
        for (int k = 0; k < 4; k++) bsp[k] = s_bsp[i_blk * 4 + k];
    }

    // chain + epilogue of tile #kk of this CTA (reads hand-off buffer kk & 1).  16 lanes per chain group: roles 0-7 are
    // the acc lanes, 8-11 the acc_m lanes.  STORE / RESIDUAL: group c owns row slot c.  SILU: group c owns the gate slot c
    // and the up slot half + c, so the group leader ends up holding both values and no second barrier is needed.
    auto chain_tile = [&](int kk) {
        const int tile = (int)blockIdx.x + kk * (int)gridDim.x;
        int seg, row0, nrows;
        tile_info(tile, seg, row0, nrows);
        const int grp = tid >> 4, role = tid & 15;
        const uint32_t *cbuf = s_chain + (size_t)(kk & 1) * R * chain_row;
        const int base = lane & 16;
        const bool g_ok = grp < nrows;
        float res[2] = {0.f, 0.f};
        const int nchains = (a.epi == PS_EPI_SILU) ? 2 : 1;
        for (int c = 0; c < nchains; c++) {
            float acc = 0.f;
            if (g_ok && role < 12) {
                const uint32_t *cb = cbuf + (size_t)(grp + c * half) * chain_row;
                const int dsel = (role < 8) ? 12 : 13;
                int i = 0;
                for (; i + 8 <= nb; i += 8) {
                    int v[8];
                    float dd[8];
#pragma unroll
                    for (int u = 0; u < 8; u++) { v[u] = (int)cb[(i + u) * 16 + role]; dd[u] = __uint_as_float(cb[(i + u) * 16 + dsel]); }
#pragma unroll
                    for (int u = 0; u < 8; u++) acc = __fmaf_rn(dd[u], __int2float_rn(v[u]), acc);
                }
                for (; i < nb; i++) acc = __fmaf_rn(__uint_as_float(cb[i * 16 + dsel]), __int2float_rn((int)cb[i * 16 + role]), acc);
            }
            float x[12];
#pragma unroll
            for (int t = 0; t < 12; t++) x[t] = __shfl_sync(PS_FULL, acc, base + t);
            const float r0 = __fadd_rn(x[4], x[0]), r1 = __fadd_rn(x[5], x[1]), r2 = __fadd_rn(x[6], x[2]), r3 = __fadd_rn(x[7], x[3]);
            const float hsum = __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
            res[c] = __fadd_rn(hsum, __fadd_rn(__fadd_rn(x[8], x[10]), __fadd_rn(x[9], x[11])));
        }
        if (g_ok && role == 0) {
            const int n = row0 + grp;
            if (a.epi == PS_EPI_SILU) {
                a.seg[0].dst[n] = ps_silu_mul(res[0], res[1]);
            } else {
                float r = res[0];
                if (a.seg[seg].bias) r = __fadd_rn(r, a.seg[seg].bias[n]);
                if (a.epi == PS_EPI_RESIDUAL) r = __fadd_rn(a.residual[n], r);
                a.seg[seg].dst[n] = r;
            }
        }
    };

    for (int k = 0; k < my_tiles; k++) {
        const int tile = (int)blockIdx.x + k * (int)gridDim.x;
        const int st = k % PS_MV_STAGES;
        int seg, row0, nrows;
        tile_info(tile, seg, row0, nrows);
        if (k < 4) PS_TRACE(4 + 2 * k);
        ps_mbar_wait(&s_full[st], (k / PS_MV_STAGES) & 1);
        if (k < 4) PS_TRACE(5 + 2 * k);
        // ---------------- integer phase: one thread, one block
        const bool row_ok = active && ((a.epi == PS_EPI_SILU) ? (r_slot < nrows || (r_slot >= half && r_slot < half + nrows)) : (r_slot < nrows));
        if (row_ok) {
            const uint4 *blk = reinterpret_cast<const uint4 *>(s_stage + (size_t)st * stage_bytes + ((size_t)r_slot * nb + i_blk) * PS_Q4_K_BYTES);
            const uint4 h = blk[0];
            const uint32_t k1 = 0x3f3f3f3fu, k2 = 0x0f0f0f0fu, k3 = 0x03030303u;
            const uint32_t mB = ((h.w >> 4) & k2) | (((h.z >> 6) & k3) << 4), mA = h.z & k1;
            const uint32_t scB = (h.w & k2) | (((h.y >> 6) & k3) << 4), scA = h.y & k1;
            int S[8], Sh[8];
#pragma unroll
            for (int l = 0; l < 8; l++) { S[l] = 0; Sh[l] = 0; }
#pragma unroll
            for (int j2 = 0; j2 < 4; j2++) {
                const uint32_t scw = (j2 < 2) ? scA : scB;
                const int s_lo = (scw >> (16 * (j2 & 1))) & 0xff, s_hi = (scw >> (16 * (j2 & 1) + 8)) & 0xff;
#pragma unroll
                for (int hh = 0; hh < 2; hh++) {
                    const uint4 qv = blk[1 + 2 * j2 + hh];
                    const uint32_t w4[4] = {qv.x, qv.y, qv.z, qv.w};
#pragma unroll
                    for (int t = 0; t < 4; t++) {
                        const int l = 4 * hh + t;
                        const int p0 = __dp4a((int)(w4[t] & 0x0f0f0f0fu), (int)q8[(2 * j2) * 8 + l], 0);
                        const int p1 = ps_dp4a_us(w4[t] & 0xf0f0f0f0u, (int)q8[(2 * j2 + 1) * 8 + l], 0); // 16 x the high-nibble dot
                        S[l] += s_lo * p0;
                        Sh[l] += s_hi * p1;
                    }
                }
            }
            uint4 o0, o1, o2, o3;
            o0.x = (uint32_t)(S[0] + (Sh[0] >> 4)); o0.y = (uint32_t)(S[1] + (Sh[1] >> 4));
            o0.z = (uint32_t)(S[2] + (Sh[2] >> 4)); o0.w = (uint32_t)(S[3] + (Sh[3] >> 4));
            o1.x = (uint32_t)(S[4] + (Sh[4] >> 4)); o1.y = (uint32_t)(S[5] + (Sh[5] >> 4));
            o1.z = (uint32_t)(S[6] + (Sh[6] >> 4)); o1.w = (uint32_t)(S[7] + (Sh[7] >> 4));
            int P[4]; // prod lanes: m_{2k} s_{2k} + m_{2k+1} s_{2k+1}
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const uint32_t mw = (kk < 2) ? mA : mB;
                const int m0 = (mw >> (16 * (kk & 1))) & 0xff, m1 = (mw >> (16 * (kk & 1) + 8)) & 0xff;
                P[kk] = m0 * (int)(short)(bsp[kk] & 0xffffu) + m1 * (int)(short)(bsp[kk] >> 16);
            }
            o2.x = (uint32_t)P[0]; o2.y = (uint32_t)P[1]; o2.z = (uint32_t)P[2]; o2.w = (uint32_t)P[3];
            const float xd = ps_half_bits_to_float(h.x & 0xffffu), xmin = ps_half_bits_to_float(h.x >> 16);
            o3.x = __float_as_uint(__fmul_rn(yd, xd));
            o3.y = __float_as_uint(__fmul_rn(-yd, xmin));
            o3.z = 0; o3.w = 0;
            uint4 *cbv = reinterpret_cast<uint4 *>(s_chain + (size_t)(k & 1) * R * chain_row + (size_t)r_slot * chain_row + i_blk * 16);
            cbv[0] = o0; cbv[1] = o1; cbv[2] = o2; cbv[3] = o3;
        }
        __syncwarp();
        if (lane == 0) ps_mbar_arrive(&s_empty[st]); // the weight bytes of this stage are consumed
        // the fp32 chains of the PREVIOUS tile run here, interleaved (across warps) with this tile's integer work
        if (k > 0) chain_tile(k - 1);
        ps_bar_sync(1, PS_MV_COMPUTE);
    }
    PS_TRACE(12);
    if (my_tiles > 0) chain_tile(my_tiles - 1);
    PS_TRACE(13);
}

// ====================================================================================================================
// Decode attention (bs = 1), two kernels, positions read from device memory so the step can be replayed as a graph
// ====================================================================================================================
// ATTN1 = ROPE(q), ROPE(k) + KV store + mat_mul(k_view, q) + scale/mask      (norm_attention.cpp:72-134, first half)
// One CTA (4 warps) per (32-position chunk, kv head).  Each CTA re-derives rope(q) for the r2 heads of its group (128
// floats each) instead of running a separate kernel; the CTA whose chunk holds the current position also ropes k,
// appends it to the K cache, stores v into the transposed V cache, and scores it from registers.
//   wp[h][j] = fl(fl(dot_f32(K[j], q_h) * scale) + 0.0f)   written to `sc` ({n_ctx} per head), j < n_kv = pos + 1.
__global__ void __launch_bounds__(128) ps_k_attn1(float *__restrict__ sc, float *__restrict__ kc, float *__restrict__ vct,
                                                  const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                                                  const int32_t *__restrict__ pos_dev, const float *__restrict__ table, int hs, int n_heads,
                                                  int n_kv_heads, int n_ctx, int neox, float scale) {
    __shared__ float s_q[8][256];  // roped q of the (<= 8) heads of this group
    __shared__ float s_k[256];     // roped k of the current position
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    const int pos = pos_dev[0];
    const int64_t n_kv = (int64_t)pos + 1;
    const int chunk = blockIdx.x, g = blockIdx.y;
    if ((int64_t)chunk * 32 >= n_kv) return;
    const int r2 = n_heads / n_kv_heads, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const float *cache = table + (int64_t)pos * hs;
    const bool has_cur = (pos / 32) == chunk;
    // rope, pair by pair (ggml.c:15455-15486): d[a] = x0*c - x1*s ; d[b] = x0*s + x1*c, products rounded separately
    for (int idx = tid; idx < (r2 + (has_cur ? 1 : 0)) * (hs / 2); idx += blockDim.x) {
        const int hh = idx / (hs / 2), p = idx % (hs / 2);
        const float *src = (hh < r2) ? q + (int64_t)(g * r2 + hh) * hs : k + (int64_t)g * hs;
        float *dst = (hh < r2) ? s_q[hh] : s_k;
        const int i0 = 2 * p;
        const float c = cache[i0], sn = cache[i0 + 1];
        const int ia = neox ? p : i0, ib = neox ? p + hs / 2 : i0 + 1;
        const float x0 = src[ia], x1 = src[ib];
        dst[ia] = __fadd_rn(__fmul_rn(x0, c), -__fmul_rn(x1, sn));
        dst[ib] = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
    }
    __syncthreads();
    if (has_cur) { // KV store (norm_attention.cpp:79-105): K row `pos`, V column `pos` of the transposed cache
        for (int e = tid; e < hs; e += blockDim.x) {
            kc[(int64_t)pos * (hs * n_kv_heads) + g * hs + e] = s_k[e];
            vct[((int64_t)g * hs + e) * n_ctx + pos] = v[g * hs + e];
        }
    }
    const int steps = hs / 32;
    // all (<= 8) K rows of this warp are requested before any is used: the loop is latency- not bandwidth-bound
    float kv[8][8];
    const int64_t j0 = (int64_t)chunk * 32 + warp * 8;
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int64_t j = j0 + t;
#pragma unroll
        for (int s = 0; s < 8; s++) {
            kv[t][s] = 0.f;
            if (s < steps && j < n_kv) kv[t][s] = (j == pos) ? s_k[32 * s + lane] : kc[j * (int64_t)(hs * n_kv_heads) + g * hs + 32 * s + lane];
        }
    }
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const int64_t j = j0 + t;
        if (j < n_kv) {
            for (int hh = 0; hh < r2; hh++) {
                float sum = 0.f;
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (s < steps) sum = __fmaf_rn(kv[t][s], s_q[hh][32 * s + lane], sum);
                sum = ps_f32x8_reduce(sum);
                if (lane == 0) sc[(int64_t)(g * r2 + hh) * n_ctx + j] = __fadd_rn(__fmul_rn(sum, scale), 0.0f);
            }
        }
    }
}

// ATTN2 = softmax_ext + mat_mul(v_view, kq) + permute/cont                    (norm_attention.cpp:133-151)
// One CTA (8 warps) per (group of 8 output dims d, kv head): it rebuilds the soft-max row of each of the r2 heads of the
// group in shared memory (max, ggml_v_expf / expf tail, double sum, scale: ggml.c:14846-14940, 2814-2868) and then each
// warp streams one V^T row once for all r2 heads (ggml_vec_dot_f32 lane order, leftovers in order).
__global__ void __launch_bounds__(256) ps_k_attn2(float *__restrict__ att, const float *__restrict__ sc, const float *__restrict__ vct,
                                                  const int32_t *__restrict__ pos_dev, int hs, int n_heads, int n_kv_heads, int n_ctx) {
    extern __shared__ float s_p[]; // [r2][n_kv_pad]
    __shared__ double sh[32];
    __shared__ float shf[32];
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    const int64_t n_kv = (int64_t)pos_dev[0] + 1;
    const int g = blockIdx.y, r2 = n_heads / n_kv_heads, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t stride = (n_kv + 31) & ~(int64_t)31;
    const int64_t n8 = n_kv & ~(int64_t)7;
    for (int hh = 0; hh < r2; hh++) {
        const float *wp = sc + (int64_t)(g * r2 + hh) * n_ctx;
        float *pp = s_p + hh * stride;
        float mx = -INFINITY;
        for (int64_t j = tid; j < n_kv; j += blockDim.x) {
            const float vv = wp[j];
            pp[j] = vv;
            mx = fmaxf(mx, vv);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PS_FULL, mx, o));
        __syncthreads();
        if (lane == 0) shf[warp] = mx;
        __syncthreads();
        mx = shf[0];
        for (int t = 1; t < 8; t++) mx = fmaxf(mx, shf[t]);
        double s = 0.0;
        for (int64_t gi = tid; gi < n8 / 8; gi += blockDim.x) {
            float vv[8];
#pragma unroll
            for (int l = 0; l < 8; l++) {
                vv[l] = ps_v_expf(__fadd_rn(pp[gi * 8 + l], -mx));
                pp[gi * 8 + l] = vv[l];
            }
            const float r0 = __fadd_rn(vv[4], vv[0]), r1 = __fadd_rn(vv[5], vv[1]), r2_ = __fadd_rn(vv[6], vv[2]), r3 = __fadd_rn(vv[7], vv[3]);
            s += (double)__fadd_rn(__fadd_rn(r0, r2_), __fadd_rn(r1, r3));
        }
        for (int64_t j = n8 + tid; j < n_kv; j += blockDim.x) {
            const float vv = ps_expf_glibc(__fadd_rn(pp[j], -mx));
            pp[j] = vv;
            s += (double)vv;
        }
        const double sum = ps_block_sum_double(s, sh);
        const float inv = (float)(1.0 / sum);
        for (int64_t j = tid; j < n_kv; j += blockDim.x) pp[j] = __fmul_rn(pp[j], inv);
    }
    __syncthreads();
    const int d = blockIdx.x * 8 + warp;
    if (d >= hs) return;
    const float *vrow = vct + ((int64_t)g * hs + d) * n_ctx;
    const int64_t np = n_kv & ~(int64_t)31;
    float sum[8];
#pragma unroll
    for (int hh = 0; hh < 8; hh++) sum[hh] = 0.f;
    int64_t s0 = 0;
    for (; s0 + 256 <= np; s0 += 256) { // 8 independent V loads in flight per lane; the FMA chains stay in position order
        float vv[8];
#pragma unroll
        for (int u = 0; u < 8; u++) vv[u] = vrow[s0 + 32 * u + lane];
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int hh = 0; hh < 8; hh++)
                if (hh < r2) sum[hh] = __fmaf_rn(vv[u], s_p[hh * stride + s0 + 32 * u + lane], sum[hh]);
    }
    for (; s0 < np; s0 += 32) {
        const float vv = vrow[s0 + lane];
#pragma unroll
        for (int hh = 0; hh < 8; hh++)
            if (hh < r2) sum[hh] = __fmaf_rn(vv, s_p[hh * stride + s0 + lane], sum[hh]);
    }
#pragma unroll
    for (int hh = 0; hh < 8; hh++) {
        if (hh < r2) {
            float r = ps_f32x8_reduce(sum[hh]);
            if (lane == 0) {
                for (int64_t j = np; j < n_kv; j++) r = __fadd_rn(r, __fmul_rn(vrow[j], s_p[hh * stride + j]));
                att[(int64_t)(g * r2 + hh) * hs + d] = r;
            }
        }
    }
}

// GGMLBackend::get_embedding for the token held in device memory (decode feedback loop)
__global__ void __launch_bounds__(256) ps_k_embed_dev(float *__restrict__ dst, const uint8_t *__restrict__ w, int type, int64_t dim,
                                                      const int32_t *__restrict__ tokens) {
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
}

// greedy pick + device-side step bookkeeping: ids[*ctr] = argmax, token feedback, position and counter advance
__global__ void __launch_bounds__(1024) ps_k_argmax_step(const float *__restrict__ logits, int64_t n, int32_t *__restrict__ ids,
                                                         int32_t *__restrict__ ctr, int32_t *__restrict__ next_token, int32_t *__restrict__ pos) {
    __shared__ float sv[32];
    __shared__ int si[32];
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int64_t t = threadIdx.x; t < n; t += blockDim.x) {
        const float v = logits[t];
        if (v > best || (v == best && (int)t < bi)) { best = v; bi = (int)t; }
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
}
