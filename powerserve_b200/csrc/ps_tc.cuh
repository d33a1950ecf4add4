// ps_tc.cuh — tcgen05 (5th-gen tensor core) GEMM for prefill chunks: dst{N, bs} = W{K, N} . x{K, bs}, Q4_K weights x Q8_K
// activations, BIT-IDENTICAL to ggml_vec_dot_q4_K_q8_K per column (ggml-quants.c:7809-7872).
//
// The reference's result is not a plain dot product: per super-block i and AVX lane l it forms the INTEGER
//   S_l = sum_{j<8} sc_j * sum_{t<4} q4[j][4l+t] * q8[j][4l+t]            (and P_k = m_2k*bs_2k + m_2k+1*bs_2k+1 for the mins)
// and then advances 8 + 4 fp32 FMA chains, acc_l = fma(d_x*d_y, (float)S_l, acc_l).  The integer part is a dense
// contraction, and it is EXACT on the tensor cores in fp16/fp32: the operands sc_j*q4 (<= 945) and q8 (|.| <= 127) are
// integers that fp16 represents exactly, every product is exact in fp32 and a 32-term sum stays below 2^24.  So:
//   * A operand (weights): per (128-row tile, super-block, lane l) a K = 32 fp16 tile of sc_j*q4, expanded ONCE at bind time
//     into HBM in the UMMA canonical K-major layout (2 B / weight: 15 GB for Llama-3.1-8B — the 180 GB part pays for zero
//     dequantisation work on the prefill path); the mins use one shared K = 16 tile of the m_j.
//   * B operand (activations): per (16-column group, super-block, lane) the q8 bytes as fp16, plus half-block sums for the
//     mins, written per forward call by ps_k_tc_prep_b (same quantiser as everywhere else).
//   * per super-block: 16 + 4 tcgen05.mma (M = 128, N = 16, K = 16) into 12 x 16 TMEM columns (double-buffered); 16
//     epilogue warps read the exact integers back (12 tcgen05.ld in flight, one wait) and advance the FMA chains in
//     registers, one (row, column) pair's 12 chains in ONE thread — the final hsum_float_8 order needs no shuffles.
//     The tile is 128 x 16 because the chains, not the MMAs, set the budget: 12 fp32 accumulators per output element
//     (24.5 K registers per tile) plus the 24.5 K registers the TMEM read-back lands in fill the register file.
// Pipeline: persistent CTAs over (row tile, column group) units; TMA producer warp (2 stages, one bulk copy per operand
// block) -> single-thread MMA issuer -> epilogue warps, linked by mbarriers (smem full/empty, tmem full/empty x 2).
#pragma once
#include "ps_rw.cuh"

#define PS_TC_M 128
#define PS_TC_N 16
#define PS_TC_EPI_WARPS 16
#define PS_TC_THREADS ((PS_TC_EPI_WARPS + 2) * 32)
// operand blocks in global / shared memory (bytes)
#define PS_TC_A_TILE (PS_TC_M * 16 * 2)                    // one MMA's A operand: 128 rows x K16 fp16 = 4096
#define PS_TC_B_TILE (PS_TC_N * 16 * 2)                    // one MMA's B operand: 16 cols x K16 fp16 = 512
#define PS_TC_A_BLOCK (16 * PS_TC_A_TILE + PS_TC_A_TILE + PS_TC_M * 8)  // 8 lanes x 2 + mins tile + (xd, xmin) per row = 70656
#define PS_TC_B_BLOCK (16 * PS_TC_B_TILE + 4 * PS_TC_B_TILE + PS_TC_N * 4) // 8 lanes x 2 + 4 mins tiles + yd per column = 10304
#define PS_TC_STAGE (PS_TC_A_BLOCK + PS_TC_B_BLOCK)        // 80960
#define PS_TC_STAGES 2

// canonical K-major, no-swizzle UMMA tile: [k-chunk (2)][8-row group][8 rows][8 fp16]  (cute: ((8,n),2):((1,SBO),LBO))
PS_HD int ps_tc_tile_off(int row, int k16, int n_rows) { return ((k16 >> 3) * (n_rows >> 3) + (row >> 3)) * 128 + (row & 7) * 16 + (k16 & 7) * 2; }

// ---------------------------------------------------------------------------------------------------- operand builders
// A blocks of a matrix: [row tile][super-block][PS_TC_A_BLOCK].  One CTA per (row tile, super-block); rows >= n_rows are zero.
__global__ void __launch_bounds__(256) ps_k_tc_expand_a(uint8_t *__restrict__ dst, const uint8_t *__restrict__ w, int64_t n_rows, int64_t nb,
                                                        int64_t tile0) {
    const int64_t i = blockIdx.x, rt = blockIdx.y;
    uint8_t *out = dst + ((tile0 + rt) * nb + i) * PS_TC_A_BLOCK;
    for (int idx = threadIdx.x; idx < PS_TC_M * 8; idx += blockDim.x) { // (row, sub-block j)
        const int row = idx >> 3, j = idx & 7;
        const int64_t n = rt * PS_TC_M + row;
        const bool live = n < n_rows;
        const uint8_t *blk = w + (n * nb + i) * PS_Q4_K_BYTES;
        int sc = 0, mn = 0;
        if (live) { // get_scale_min_k4 (ggml-quants.c:1912-1919)
            const uint8_t *s = blk + 4;
            if (j < 4) { sc = s[j] & 63; mn = s[j + 4] & 63; }
            else { sc = (s[j + 4] & 0xF) | ((s[j - 4] >> 6) << 4); mn = (s[j + 4] >> 4) | ((s[j] >> 6) << 4); }
        }
        for (int e = 0; e < 32; e++) { // element e of sub-block j -> lane l = e / 4, t = e % 4
            int q = 0;
            if (live) {
                const uint8_t b = blk[16 + 32 * (j >> 1) + e];
                q = (j & 1) ? (b >> 4) : (b & 0xF);
            }
            const int l = e >> 2, t = e & 3, k32 = 4 * j + t;
            uint8_t *tile = out + (size_t)(2 * l + (k32 >> 4)) * PS_TC_A_TILE;
            *reinterpret_cast<__half *>(tile + ps_tc_tile_off(row, k32 & 15, PS_TC_M)) = __int2half_rn(sc * q); // <= 945: exact
        }
        // mins tile (shared by the four mins MMAs): K position 4k + u holds m_{2k + u/2}
        uint8_t *mt = out + (size_t)16 * PS_TC_A_TILE;
        const int k = j >> 1, u0 = (j & 1) * 2;
        *reinterpret_cast<__half *>(mt + ps_tc_tile_off(row, 4 * k + u0, PS_TC_M)) = __int2half_rn(mn);
        *reinterpret_cast<__half *>(mt + ps_tc_tile_off(row, 4 * k + u0 + 1, PS_TC_M)) = __int2half_rn(mn);
        if (j == 0) {
            float *xs = reinterpret_cast<float *>(out + (size_t)17 * PS_TC_A_TILE) + 2 * row;
            xs[0] = live ? ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk)) : 0.f;
            xs[1] = live ? ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 2)) : 0.f;
        }
    }
}

// B blocks of an activation batch: [column group][super-block][PS_TC_B_BLOCK]; one warp per (column, super-block):
// quantize_row_q8_K (the shared warp quantiser), then the q8 bytes as fp16 in the lane tiles, the half-block sums in the
// mins tiles (zero outside lane k's four K positions) and d in the tail.  Columns >= bs are zero.
__global__ void __launch_bounds__(128) ps_k_tc_prep_b(uint8_t *__restrict__ dst, const float *__restrict__ x, int64_t K, int bs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nb = K / 256, i = (int64_t)blockIdx.x * 4 + warp;
    const int col = blockIdx.y, cg = col / PS_TC_N, c = col % PS_TC_N;
    if (i >= nb) return;
    uint8_t *out = dst + ((int64_t)cg * nb + i) * PS_TC_B_BLOCK;
    uint32_t words[2] = {0, 0}, bsp = 0;
    float yd = 0.f;
    if (col < bs) {
        float e[8];
        ps_rw_load8(x + (int64_t)col * K + i * 256, lane, e);
        ps_quant_block_q8k_regs(e, lane, words, yd, bsp);
    }
    // natural word `lane` = (sub-block j = lane / 8, AVX lane l = lane % 8); second word: sub-block 4 + lane / 8
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int j = 4 * h + (lane >> 3), l = lane & 7;
        int hsum = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const int q = (int)(signed char)((words[h] >> (8 * t)) & 0xff);
            const int k32 = 4 * j + t;
            uint8_t *tile = out + (size_t)(2 * l + (k32 >> 4)) * PS_TC_B_TILE;
            *reinterpret_cast<__half *>(tile + ps_tc_tile_off(c, k32 & 15, PS_TC_N)) = __int2half_rn(q);
            hsum += q;
        }
        // half-block sums of sub-block j: lanes l = 0..3 -> elements 0..15, l = 4..7 -> elements 16..31
        hsum += __shfl_xor_sync(PS_FULL, hsum, 1);
        hsum += __shfl_xor_sync(PS_FULL, hsum, 2);
        const int k = j >> 1, u = (j & 1) * 2 + (l >> 2);
        if ((l & 3) == 0) { // lane k's tile gets the value at K position 4k + u; the other K positions of that tile stay zero
            uint8_t *tile = out + (size_t)(16 + k) * PS_TC_B_TILE;
            *reinterpret_cast<__half *>(tile + ps_tc_tile_off(c, 4 * k + u, PS_TC_N)) = __int2half_rn(hsum); // |.| <= 2032: exact
        }
    }
    if (lane == 0) reinterpret_cast<float *>(out + (size_t)20 * PS_TC_B_TILE)[c] = yd;
}

// ---------------------------------------------------------------------------------------------------- tcgen05 helpers
PS_D uint64_t ps_tc_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    // cute::UMMA::SmemDescriptor: start address [0,14) >> 4, leading byte offset [16,30) >> 4, stride byte offset [32,46) >> 4,
    // version [46,48) = 1 (Blackwell), layout type [61,64) = 0 (no swizzle)
    return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) | (1ull << 46);
}
PS_D void ps_tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
PS_D void ps_tc_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(ps_smem_u32(bar)) : "memory");
}
PS_D void ps_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
PS_D void ps_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
PS_D void ps_tc_ld8(uint32_t taddr, float v[8]) { // 32 lanes x 8 consecutive columns, one row (lane) per thread
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
#pragma unroll
    for (int t = 0; t < 8; t++) v[t] = __uint_as_float(r[t]);
}
PS_D void ps_tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// bounded mbarrier wait: a mis-sequenced pipeline must never hang the GPU (returns false after ~2^22 polls)
PS_D bool ps_tc_wait(uint64_t *bar, uint32_t parity) {
    for (int spin = 0; spin < (1 << 22); spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(ps_smem_u32(bar)), "r"(parity)
            : "memory");
        if (ok) return true;
    }
    return false;
}

struct PsTcArgs {
    const uint8_t *a;   // [row tiles][nb][PS_TC_A_BLOCK]
    const uint8_t *b;   // [column groups][nb][PS_TC_B_BLOCK]
    int nb, n_cg, n_units; // units = row tiles x column groups
    PsRwSeg seg[3];     // dst of a segment is [bs][rows of the segment]
    int n_seg, bs;
    const float *residual;
    int *err;           // set to 1 if a pipeline wait timed out (results invalid)
};

// Persistent CTAs; unit u = (row tile u / n_cg, column group u % n_cg), so that the column groups sharing an A tile run at
// the same time on neighbouring CTAs (A comes out of L2 for all but the first).
__global__ void __launch_bounds__(PS_TC_THREADS, 1) ps_k_tc_gemm(const PsTcArgs a) {
    extern __shared__ __align__(1024) uint8_t ps_tc_smem[];
    __shared__ __align__(8) uint64_t bar_full[PS_TC_STAGES], bar_empty[PS_TC_STAGES], bar_tmem_full[2], bar_tmem_empty[2];
    __shared__ uint32_t tmem_base_smem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nb = a.nb;
    const int my_units = (a.n_units > (int)blockIdx.x) ? (a.n_units - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int n_steps = my_units * nb; // (unit, super-block) steps of this CTA, in order

    if (tid == 0) {
        for (int s = 0; s < PS_TC_STAGES; s++) {
            ps_mbar_init(&bar_full[s], 1);
            ps_mbar_init(&bar_empty[s], 1 + PS_TC_EPI_WARPS); // the MMA commit + every epilogue warp (they read xd / yd from the stage)
        }
        for (int b = 0; b < 2; b++) {
            ps_mbar_init(&bar_tmem_full[b], 1);
            ps_mbar_init(&bar_tmem_empty[b], PS_TC_EPI_WARPS);
        }
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == PS_TC_EPI_WARPS + 1) { // the MMA warp owns the tensor memory: 512 columns (2 x 12 x 16 used)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(ps_smem_u32(&tmem_base_smem)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    ps_tc_fence_before();
    __syncthreads();
    ps_tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;

    if (warp == PS_TC_EPI_WARPS) {
        // ===== TMA producer
        if (lane == 0) {
            for (int g = 0; g < n_steps; g++) {
                const int u = (int)blockIdx.x + (g / nb) * (int)gridDim.x, i = g % nb, s = g % PS_TC_STAGES;
                if (g >= PS_TC_STAGES && !ps_tc_wait(&bar_empty[s], ((g / PS_TC_STAGES) - 1) & 1)) { *a.err = 1; break; }
                uint8_t *st = ps_tc_smem + (size_t)s * PS_TC_STAGE;
                ps_mbar_expect_tx(&bar_full[s], PS_TC_STAGE);
                ps_bulk_g2s(st, a.a + ((size_t)(u / a.n_cg) * nb + i) * PS_TC_A_BLOCK, PS_TC_A_BLOCK, &bar_full[s]);
                ps_bulk_g2s(st + PS_TC_A_BLOCK, a.b + ((size_t)(u % a.n_cg) * nb + i) * PS_TC_B_BLOCK, PS_TC_B_BLOCK, &bar_full[s]);
            }
        }
    } else if (warp == PS_TC_EPI_WARPS + 1) {
        // ===== MMA issuer (one thread)
        if (lane == 0) {
            // kind::f16, D = F32 (c_format 1 @ bit 4), A = B = F16 (0), K-major both, N >> 3 @ bit 17, M >> 4 @ bit 24
            const uint32_t idesc = (1u << 4) | ((uint32_t)(PS_TC_N >> 3) << 17) | ((uint32_t)(PS_TC_M >> 4) << 24);
            for (int g = 0; g < n_steps; g++) {
                const int s = g % PS_TC_STAGES, tb = g & 1;
                if (!ps_tc_wait(&bar_full[s], (g / PS_TC_STAGES) & 1)) { *a.err = 2; break; }
                if (g >= 2 && !ps_tc_wait(&bar_tmem_empty[tb], ((g >> 1) - 1) & 1)) { *a.err = 3; break; } // that accumulator buffer has been drained
                ps_tc_fence_after();
                const uint32_t sa = ps_smem_u32(ps_tc_smem + (size_t)s * PS_TC_STAGE), sb = sa + PS_TC_A_BLOCK;
                const uint32_t td = tmem_base + tb * (12 * PS_TC_N);
#pragma unroll 1
                for (int l = 0; l < 8; l++)
#pragma unroll
                    for (int h = 0; h < 2; h++)
                        ps_tc_mma_f16(td + l * PS_TC_N, ps_tc_smem_desc(sa + (2 * l + h) * PS_TC_A_TILE, (PS_TC_M / 8) * 128, 128),
                                      ps_tc_smem_desc(sb + (2 * l + h) * PS_TC_B_TILE, (PS_TC_N / 8) * 128, 128), idesc, h);
#pragma unroll 1
                for (int k = 0; k < 4; k++)
                    ps_tc_mma_f16(td + (8 + k) * PS_TC_N, ps_tc_smem_desc(sa + 16 * PS_TC_A_TILE, (PS_TC_M / 8) * 128, 128),
                                  ps_tc_smem_desc(sb + (16 + k) * PS_TC_B_TILE, (PS_TC_N / 8) * 128, 128), idesc, 0);
                ps_tc_commit(&bar_empty[s]);        // operands consumed
                ps_tc_commit(&bar_tmem_full[tb]);   // accumulators of this super-block complete
            }
        }
    } else {
        // ===== epilogue: warp w owns TMEM lanes 32 * (w % 4) .. +31 (rows) and columns 4 * (w / 4) .. +3 of every accumulator
        const int row = 32 * (warp & 3) + lane, c0 = 4 * (warp >> 2);
        const uint32_t t_lane = tmem_base + ((uint32_t)(32 * (warp & 3)) << 16) + c0;
        int g = 0;
        bool ok = true;
        for (int mu = 0; mu < my_units && ok; mu++) {
            const int u = (int)blockIdx.x + mu * (int)gridDim.x, rt = u / a.n_cg, cg = u % a.n_cg;
            float acc[4][12];
#pragma unroll
            for (int c = 0; c < 4; c++)
#pragma unroll
                for (int l = 0; l < 12; l++) acc[c][l] = 0.f;
            for (int i = 0; i < nb; i++, g++) {
                const int s = g % PS_TC_STAGES, tb = g & 1;
                if (!ps_tc_wait(&bar_tmem_full[tb], (g >> 1) & 1)) { *a.err = 4; ok = false; break; }
                ps_tc_fence_after();
                const uint8_t *st = ps_tc_smem + (size_t)s * PS_TC_STAGE;
                const float2 xs = reinterpret_cast<const float2 *>(st + (size_t)17 * PS_TC_A_TILE)[row];
                const float4 yd = *reinterpret_cast<const float4 *>(st + PS_TC_A_BLOCK + (size_t)20 * PS_TC_B_TILE + 4 * c0);
                __syncwarp();
                if (lane == 0) ps_mbar_arrive(&bar_empty[s]); // this warp no longer needs the stage
                uint32_t v[12][4];
#pragma unroll
                for (int l = 0; l < 12; l++)
                    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(v[l][0]), "=r"(v[l][1]), "=r"(v[l][2]), "=r"(v[l][3])
                                 : "r"(t_lane + tb * (12 * PS_TC_N) + l * PS_TC_N)
                                 : "memory");
                ps_tc_ld_wait();
                ps_tc_fence_before();
                __syncwarp();
                if (lane == 0) ps_mbar_arrive(&bar_tmem_empty[tb]); // the integers are in registers: the buffer can be refilled
                const float ydc[4] = {yd.x, yd.y, yd.z, yd.w};
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const float d = __fmul_rn(ydc[c], xs.x), dm = __fmul_rn(-ydc[c], xs.y);
#pragma unroll
                    for (int l = 0; l < 12; l++) // v is the exact integer S_l (l < 8) / P_k (l >= 8)
                        acc[c][l] = __fmaf_rn(l < 8 ? d : dm, __uint_as_float(v[l][c]), acc[c][l]);
                }
            }
            if (!ok) break;
            // ---- hsum_float_8 + mins sum (ggml-quants.c:62-68, 7862-7871), bias / residual, dst[col][row]
            const int grow = rt * PS_TC_M + row;
            int sg = 0;
            if (a.n_seg > 1 && grow >= a.seg[1].row_begin) sg = 1;
            if (a.n_seg > 2 && grow >= a.seg[2].row_begin) sg = 2;
            if (grow < a.seg[sg].row_end) {
                const int n = grow - a.seg[sg].row_begin, ld = a.seg[sg].row_end - a.seg[sg].row_begin;
                const float bias = a.seg[sg].bias ? a.seg[sg].bias[n] : 0.f;
#pragma unroll
                for (int c = 0; c < 4; c++) {
                    const int col = cg * PS_TC_N + c0 + c;
                    if (col < a.bs) {
                        const float r0 = __fadd_rn(acc[c][4], acc[c][0]), r1 = __fadd_rn(acc[c][5], acc[c][1]);
                        const float r2 = __fadd_rn(acc[c][6], acc[c][2]), r3 = __fadd_rn(acc[c][7], acc[c][3]);
                        float res = __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
                        res = __fadd_rn(res, __fadd_rn(__fadd_rn(acc[c][8], acc[c][10]), __fadd_rn(acc[c][9], acc[c][11])));
                        const size_t o = (size_t)col * ld + n;
                        if (a.seg[sg].bias) res = __fadd_rn(res, bias);
                        if (a.residual) res = __fadd_rn(a.residual[o], res);
                        a.seg[sg].dst[o] = res;
                    }
                }
            }
        }
    }
    ps_tc_fence_before();
    __syncthreads();
    if (warp == PS_TC_EPI_WARPS + 1) {
        ps_tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    }
}
