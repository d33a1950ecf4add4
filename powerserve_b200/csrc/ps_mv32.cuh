// ps_mv32.cuh — fused decode mat-vec for the 32-element block formats (Q4_0, Q8_0 weights x Q8_0 activations):
// the counterpart of the Q4_K row-walker (ps_rw.cuh) for the models BASELINE configs[0] names (Qwen2-0.5B Q4_0) and for
// Q8_0 files.  Replaces powerserve_compute_forward_mul_mat -> ggml_vec_dot_q4_0_q8_0 / ggml_vec_dot_q8_0_q8_0
// (libs/ggml/src/ggml.c:13344-13432, ggml-quants.c:4205-4228, 5761-5782) with quantize_row_q8_0 (ggml-quants.c:957-1017),
// RMSNorm (ggml.c:12667-12721) or SiLU.up (src/backend/ggml/ggml.cpp:115-129) fused into the prologue and bias /
// residual / greedy-pick partials into the epilogue.
//
// Arithmetic (bit-exact, the same chains as PsBlk<2> / PsBlk<8> in ps_kernels.cuh): per row the reference keeps one
// __m256 accumulator; lane l of block b adds fl(d_x * d_y) * float(sum of the four int8 products 4l..4l+3) with ONE FMA,
// blocks in row order; hsum_float_8 at the end.  Four threads own a row (thread t = AVX lanes t and t + 4: for Q4_0 both
// come out of the same 32-bit word of nibbles, low nibbles = lane t, high nibbles = lane t + 4), a warp owns an octet of
// rows and walks the row's blocks in order.
//
// Weights are repacked once at bind time into an OCTET layout: [octet][block][8 rows x QB quant bytes | 8 x fp16 d]
// (144 B for Q4_0, 272 B for Q8_0 - a permutation of the GGUF bytes), so that a warp's stream is contiguous, every
// pipeline stage is one TMA bulk copy into the warp's private shared-memory ring and every warp-wide shared-memory load
// is conflict free.  Rings are filled BEFORE the dependency wait (weights never depend on the previous kernel).
#pragma once
#include "ps_decode.cuh"

#define PS_MV_WARPS 8
#define PS_MV_THREADS 256
#define PS_MV_MAX_NS 4
#define PS_MV_STAGE_CAP 6144

template <int TYPE> struct PsMv32;
template <> struct PsMv32<2> { static constexpr int QB = 16, BLK = 144, SRC = 18; };
template <> struct PsMv32<8> { static constexpr int QB = 32, BLK = 272, SRC = 34; };

// GGUF rows of 18- / 34-byte blocks -> octet layout (rows of a partial last octet stay zero: the buffer is cleared first)
template <int TYPE>
__global__ void ps_k_mv32_repack(uint8_t *__restrict__ dst, const uint8_t *__restrict__ w, int64_t n_rows, int64_t nb, int64_t oct0) {
    using G = PsMv32<TYPE>;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n_rows * nb; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = t / nb, b = t % nb;
        const uint8_t *src = w + (row * nb + b) * G::SRC;
        uint8_t *o = dst + ((oct0 + row / 8) * nb + b) * G::BLK;
        const int r = (int)(row % 8);
        for (int k = 0; k < G::QB; k++) o[r * G::QB + k] = src[2 + k];
        o[8 * G::QB + 2 * r] = src[0];
        o[8 * G::QB + 2 * r + 1] = src[1];
    }
}

struct PsMvSeg {
    float *dst;            // [rows of the segment]
    const float *bias;     // optional
    int row_begin, row_end;
};
enum { PS_MV_PRO_PLAIN = 0, PS_MV_PRO_RMSNORM = 1, PS_MV_PRO_SILU = 2 };
struct PsMvArgs {
    const uint8_t *w;      // [n_oct][nb][BLK]
    int n_oct, K, sb, ns;  // sb = blocks per ring stage (divides nb), ns = ring stages per warp
    int kpar;              // block-parallel mode (one octet per CTA, long rows): slice capacity in bytes per warp, 0 = off
    const float *x;        // activation vector (PRO_SILU: the gate vector)
    const float *x2;       // PRO_SILU: the up vector
    const float *norm_w;   // PRO_RMSNORM
    float eps;
    double inv_k;
    int pro;
    PsMvSeg seg[3];        // q | k | v, gate | up, or a single matrix
    int n_seg;
    const float *residual; // optional, indexed like seg[0].dst (single-segment launches)
    float *part_val;       // optional greedy-pick partials, one per CTA
    int *part_idx;
};

// quantize_row_q8_0, AVX2 branch, for the four 32-blocks a warp holds (lane = 8 b + l owns elements 4l..4l+3 of block b):
// returns the packed word and d = fp16(max|x| / 127) widened back
PS_D void ps_mv_quant4(const float4 v, uint32_t &word, float &d_out) {
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(PS_FULL, amax, o));
    const float d = __fdiv_rn(amax, 127.f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
    const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
    word = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
    d_out = __half2float(__float2half_rn(d));
}

// prologue of every launch: (RMSNorm | SiLU.up) + quantize_row_q8_0 of the activation vector into shared memory (all eight warps;
// ends with a CTA barrier)
PS_D void ps_mv_prologue(const PsMvArgs &a, uint32_t *s_qs, float *s_d, double *sh_red, int K, int nb, int tid, int warp, int lane) {
    float nscale = 1.f;
    if (a.pro == PS_MV_PRO_RMSNORM) {
        double ss = 0.0;
        for (int e = tid * 4; e < K; e += PS_MV_THREADS * 4) {
            const float4 v = *reinterpret_cast<const float4 *>(a.x + e);
            ss += (double)__fmul_rn(v.x, v.x);
            ss += (double)__fmul_rn(v.y, v.y);
            ss += (double)__fmul_rn(v.z, v.z);
            ss += (double)__fmul_rn(v.w, v.w);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(PS_FULL, ss, o);
        if (lane == 0) sh_red[warp] = ss;
        __syncthreads();
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < PS_MV_WARPS; w++) tot += sh_red[w];
        const float mean = (float)(a.inv_k != 0.0 ? tot * a.inv_k : tot / (double)K);
        nscale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, a.eps)));
    }
    for (int i0 = warp * 4; i0 < nb; i0 += PS_MV_WARPS * 4) { // four blocks per warp pass
        const int i = i0 + (lane >> 3), l = lane & 7;
        const bool valid = i < nb;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (valid) {
            v = *reinterpret_cast<const float4 *>(a.x + i * 32 + 4 * l);
            if (a.pro == PS_MV_PRO_RMSNORM) {
                const float4 wv = *reinterpret_cast<const float4 *>(a.norm_w + i * 32 + 4 * l);
                v.x = __fmul_rn(v.x, __fmul_rn(wv.x, nscale)); // y = x * (w * scale)
                v.y = __fmul_rn(v.y, __fmul_rn(wv.y, nscale));
                v.z = __fmul_rn(v.z, __fmul_rn(wv.z, nscale));
                v.w = __fmul_rn(v.w, __fmul_rn(wv.w, nscale));
            } else if (a.pro == PS_MV_PRO_SILU) {
                const float4 u = *reinterpret_cast<const float4 *>(a.x2 + i * 32 + 4 * l);
                v.x = ps_silu_mul(v.x, u.x);
                v.y = ps_silu_mul(v.y, u.y);
                v.z = ps_silu_mul(v.z, u.z);
                v.w = ps_silu_mul(v.w, u.w);
            }
        }
        uint32_t word;
        float d;
        ps_mv_quant4(v, word, d);
        if (valid) {
            s_qs[i * 8 + l] = word;
            if (l == 0) s_d[i] = d;
        }
    }
    __syncthreads();

}

// epilogue of one octet: hsum_float_8 across the row's four threads, bias / residual, store, running arg-max
PS_D void ps_mv_epilogue(const PsMvArgs &a, int oct, int r, int t, float a_lo, float a_hi, float &best_v, int &best_i) {
    // hsum_float_8 (ggml-quants.c:62-68): (x4 + x0, x5 + x1, x6 + x2, x7 + x3) -> (r0 + r2) + (r1 + r3)
    float res = __fadd_rn(a_hi, a_lo);
    res = __fadd_rn(res, __shfl_xor_sync(PS_FULL, res, 2));
    res = __fadd_rn(res, __shfl_xor_sync(PS_FULL, res, 1));
    const int row = oct * 8 + r;
    int sg = 0;
    if (a.n_seg > 1 && row >= a.seg[1].row_begin) sg = 1;
    if (a.n_seg > 2 && row >= a.seg[2].row_begin) sg = 2;
    if (t == 0 && row < a.seg[sg].row_end) {
        const int n = row - a.seg[sg].row_begin;
        if (a.seg[sg].bias) res = __fadd_rn(res, a.seg[sg].bias[n]);
        if (a.residual) res = __fadd_rn(a.residual[n], res);
        a.seg[sg].dst[n] = res;
        if (res > best_v || (res == best_v && n < best_i)) { best_v = res; best_i = n; } // first maximum wins
    }
}

// Dynamic shared memory: [s_qs: K bytes][s_d: nb floats, padded to 128][rings: 8 warps x ns x stage_bytes][bars: 8 x ns x 8]
//
// Block-parallel mode (a.kpar > 0; matrices with at most one octet per CTA and long rows, e.g. Qwen2's down projection:
// 112 octets of 152 blocks): a lone warp would walk the whole row serially, so the row's blocks are dealt to the CTA's eight
// warps instead.  Each warp streams its slice with one bulk copy, does the INTEGER work of its blocks and leaves, per block and
// lane, the three numbers the chains need - fl(d_x d_y), float(S_lo), float(S_hi), all exact - in shared memory; after one CTA
// barrier warp 0 advances the two FMA chains over all blocks IN ROW ORDER (the arithmetic of the one-warp walk) and runs
// the epilogue.  Shared memory: [s_qs][s_d][slices: 8 x kpar bytes][factors: nb x 3 x 32 floats][bars: 8 x 8].
template <int TYPE>
__global__ void __launch_bounds__(PS_MV_THREADS) ps_k_mv32(const PsMvArgs a) {
    using G = PsMv32<TYPE>;
    extern __shared__ __align__(128) uint8_t ps_mv_smem[];
    __shared__ double sh_red[PS_MV_WARPS];
    __shared__ float sv[PS_MV_WARPS];
    __shared__ int si[PS_MV_WARPS];
    const int K = a.K, nb = K / 32, sb = a.sb, ns = a.ns;
    const uint32_t stage_bytes = (uint32_t)sb * G::BLK;
    uint32_t *s_qs = reinterpret_cast<uint32_t *>(ps_mv_smem);
    float *s_d = reinterpret_cast<float *>(ps_mv_smem + K);
    uint8_t *s_ring = ps_mv_smem + (((size_t)K + (size_t)nb * 4 + 127) & ~(size_t)127);
    uint64_t *s_bar = reinterpret_cast<uint64_t *>(s_ring + (size_t)PS_MV_WARPS * ns * stage_bytes);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, r = lane >> 2, t = lane & 3;
    const int o0 = (int)(((long long)blockIdx.x * a.n_oct) / gridDim.x), o1 = (int)(((long long)(blockIdx.x + 1) * a.n_oct) / gridDim.x);
    const int spo = nb / sb; // stages per octet
    const int n_mine = (o0 + warp < o1) ? (o1 - o0 - warp - 1) / PS_MV_WARPS + 1 : 0;
    const int n_stages = n_mine * spo;
    if (a.kpar) { // ---- block-parallel mode (see above)
        uint8_t *slice = s_ring + (size_t)warp * a.kpar;
        float *s_fac = reinterpret_cast<float *>(s_ring + (size_t)PS_MV_WARPS * a.kpar);
        uint64_t *bar = reinterpret_cast<uint64_t *>(s_fac + (size_t)nb * 96) + warp;
        const bool own = o0 < o1;                                       // CTAs beyond the octet count only take part in the dependency chain
        const int b_lo = nb * warp / PS_MV_WARPS, b_hi = nb * (warp + 1) / PS_MV_WARPS;
        if (lane == 0) {
            ps_mbar_init(bar, 1);
            ps_fence_barrier_init();
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (own && b_hi > b_lo) {
                const uint32_t bytes = (uint32_t)(b_hi - b_lo) * G::BLK;
                ps_mbar_expect_tx(bar, bytes);
                ps_bulk_g2s(slice, a.w + ((size_t)o0 * nb + b_lo) * G::BLK, bytes, bar);
            }
        }
        __syncwarp();
        ps_grid_dep_wait();
        ps_grid_dep_launch();
        ps_mv_prologue(a, s_qs, s_d, sh_red, K, nb, tid, warp, lane);
        if (!own) return;
        if (b_hi > b_lo) ps_mbar_wait(bar, 0);
        for (int i = b_lo; i < b_hi; i++) {
            const uint8_t *blk = slice + (size_t)(i - b_lo) * G::BLK;
            int S_lo, S_hi;
            if (TYPE == 2) {
                const uint32_t w = *reinterpret_cast<const uint32_t *>(blk + r * 16 + 4 * t);
                const uint32_t lo = __vsub4(w & 0x0f0f0f0fu, 0x08080808u), hi = __vsub4((w >> 4) & 0x0f0f0f0fu, 0x08080808u);
                S_lo = __dp4a((int)lo, (int)s_qs[i * 8 + t], 0);
                S_hi = __dp4a((int)hi, (int)s_qs[i * 8 + 4 + t], 0);
            } else {
                const uint32_t w0 = *reinterpret_cast<const uint32_t *>(blk + r * 32 + 4 * t);
                const uint32_t w1 = *reinterpret_cast<const uint32_t *>(blk + r * 32 + 16 + 4 * t);
                S_lo = __dp4a((int)w0, (int)s_qs[i * 8 + t], 0);
                S_hi = __dp4a((int)w1, (int)s_qs[i * 8 + 4 + t], 0);
            }
            const float xd = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 8 * G::QB + 2 * r));
            float *f = s_fac + (size_t)i * 96 + lane;
            f[0] = __fmul_rn(xd, s_d[i]);
            f[32] = __int2float_rn(S_lo);
            f[64] = __int2float_rn(S_hi);
        }
        __syncthreads();
        if (warp != 0) return;
        float a_lo = 0.f, a_hi = 0.f;
#pragma unroll 4
        for (int i = 0; i < nb; i++) { // the two chains of this thread's AVX lanes, block by block in row order
            const float *f = s_fac + (size_t)i * 96 + lane;
            const float d = f[0];
            a_lo = __fmaf_rn(d, f[32], a_lo);
            a_hi = __fmaf_rn(d, f[64], a_hi);
        }
        float best_v = -INFINITY;
        int best_i = 0x7fffffff;
        ps_mv_epilogue(a, o0, r, t, a_lo, a_hi, best_v, best_i);
        if (a.part_val) {
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                const float ov = __shfl_xor_sync(PS_FULL, best_v, o);
                const int oi = __shfl_xor_sync(PS_FULL, best_i, o);
                if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
            }
            if (lane == 0) { a.part_val[blockIdx.x] = best_v; a.part_idx[blockIdx.x] = best_i; }
        }
        return;
    }
    uint8_t *my_ring = s_ring + (size_t)warp * ns * stage_bytes;
    uint64_t *my_bar = s_bar + warp * ns;
    auto issue = [&](int s) { // lane 0: request stage #s of this warp's stream into slot s % ns
        const int oct = o0 + warp + (s / spo) * PS_MV_WARPS;
        const uint8_t *src = a.w + ((size_t)oct * nb + (size_t)(s % spo) * sb) * G::BLK;
        uint64_t *bar = my_bar + (s % ns);
        ps_mbar_expect_tx(bar, stage_bytes);
        ps_bulk_g2s(my_ring + (size_t)(s % ns) * stage_bytes, src, stage_bytes, bar);
    };
    // every warp runs its own ring: barriers are warp-private, so no CTA-wide synchronisation guards them
    if (lane == 0) {
        for (int s = 0; s < ns; s++) ps_mbar_init(my_bar + s, 1);
        ps_fence_barrier_init();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int s = 0; s < ns && s < n_stages; s++) issue(s);
    }
    __syncwarp();
    ps_grid_dep_wait();
    ps_grid_dep_launch();

    ps_mv_prologue(a, s_qs, s_d, sh_red, K, nb, tid, warp, lane);

    // ---- the stream
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;
    int s = 0, slot = 0;
    uint32_t phase = 0;
#pragma unroll 1
    for (int m = 0; m < n_mine; m++) {
        const int oct = o0 + warp + m * PS_MV_WARPS;
        float a_lo = 0.f, a_hi = 0.f; // AVX lanes t and t + 4 of this row's accumulator
#pragma unroll 1
        for (int ss = 0; ss < spo; ss++, s++) {
            ps_mbar_wait(&my_bar[slot], phase);
            const uint8_t *st = my_ring + (size_t)slot * stage_bytes;
            const int ib = ss * sb;
#pragma unroll 4
            for (int b = 0; b < sb; b++) {
                const uint8_t *blk = st + (size_t)b * G::BLK;
                const int i = ib + b;
                int S_lo, S_hi;
                if (TYPE == 2) {
                    const uint32_t w = *reinterpret_cast<const uint32_t *>(blk + r * 16 + 4 * t);
                    const uint32_t lo = __vsub4(w & 0x0f0f0f0fu, 0x08080808u), hi = __vsub4((w >> 4) & 0x0f0f0f0fu, 0x08080808u);
                    S_lo = __dp4a((int)lo, (int)s_qs[i * 8 + t], 0);
                    S_hi = __dp4a((int)hi, (int)s_qs[i * 8 + 4 + t], 0);
                } else {
                    const uint32_t w0 = *reinterpret_cast<const uint32_t *>(blk + r * 32 + 4 * t);
                    const uint32_t w1 = *reinterpret_cast<const uint32_t *>(blk + r * 32 + 16 + 4 * t);
                    S_lo = __dp4a((int)w0, (int)s_qs[i * 8 + t], 0);
                    S_hi = __dp4a((int)w1, (int)s_qs[i * 8 + 4 + t], 0);
                }
                const float xd = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 8 * G::QB + 2 * r));
                const float d = __fmul_rn(xd, s_d[i]);
                a_lo = __fmaf_rn(d, __int2float_rn(S_lo), a_lo);
                a_hi = __fmaf_rn(d, __int2float_rn(S_hi), a_hi);
            }
            __syncwarp(); // the slot is drained by every lane: re-arm it
            if (lane == 0 && s + ns < n_stages) issue(s + ns);
            if (++slot == ns) { slot = 0; phase ^= 1; }
        }
        ps_mv_epilogue(a, oct, r, t, a_lo, a_hi, best_v, best_i);
    }
    if (a.part_val) { // greedy pick, stage 1: the CTA's best (value, lowest index)
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ov = __shfl_xor_sync(PS_FULL, best_v, o);
            const int oi = __shfl_xor_sync(PS_FULL, best_i, o);
            if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
        }
        if (lane == 0) { sv[warp] = best_v; si[warp] = best_i; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < PS_MV_WARPS; w++)
                if (sv[w] > best_v || (sv[w] == best_v && si[w] < best_i)) { best_v = sv[w]; best_i = si[w]; }
            a.part_val[blockIdx.x] = best_v;
            a.part_idx[blockIdx.x] = best_i;
        }
    }
}

// ROPE(q), ROPE(k) and the two KV-cache COPY ops of NormAttention::build (norm_attention.cpp:76-105) in one launch for the
// token at pos_dev[0]: blocks [0, n_heads) rotate a query head into `qr`; the next n_kv_heads rotate a key head straight
// into row `pos` of the K cache; the last n_kv_heads write a value head into column `pos` of the transposed V cache.
// Rotation arithmetic = ps_k_rope (ggml_compute_forward_rope_f32, ggml.c:15368-15497; NORM pairs (2p, 2p+1), NEOX pairs
// (p, p + n_dims/2); products rounded separately).
__global__ void __launch_bounds__(64) ps_k_rope_kv(float *__restrict__ qr, const float *__restrict__ q, const float *__restrict__ k, const float *__restrict__ v,
                                                   float *__restrict__ kc, float *__restrict__ vct, int head_size, int n_heads, int n_kv_heads, int n_dims,
                                                   int neox, const int32_t *__restrict__ pos_dev, const float *__restrict__ table, int64_t n_ctx) {
    ps_grid_dep_wait();
    ps_grid_dep_launch();
    const int pos = pos_dev[0];
    const int b = blockIdx.x;
    if (b >= n_heads + n_kv_heads) { // value head
        const int h = b - n_heads - n_kv_heads;
        for (int e = threadIdx.x; e < head_size; e += blockDim.x) vct[((int64_t)h * head_size + e) * n_ctx + pos] = v[h * head_size + e];
        return;
    }
    const bool is_q = b < n_heads;
    const int h = is_q ? b : b - n_heads;
    const float *s = (is_q ? q : k) + (int64_t)h * head_size;
    float *d = is_q ? qr + (int64_t)h * head_size : kc + (int64_t)pos * head_size * n_kv_heads + (int64_t)h * head_size;
    const float *cache = table + (int64_t)pos * head_size;
    for (int p = threadIdx.x; p < head_size / 2; p += blockDim.x) {
        const int i0 = 2 * p;
        if (i0 < n_dims) {
            const float c = cache[i0], sn = cache[i0 + 1];
            const int ia = neox ? p : i0, ib = neox ? p + n_dims / 2 : i0 + 1;
            const float x0 = s[ia], x1 = s[ib];
            d[ia] = __fadd_rn(__fmul_rn(x0, c), -__fmul_rn(x1, sn));
            d[ib] = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));
        } else {
            d[i0] = s[i0];
            d[i0 + 1] = s[i0 + 1];
        }
    }
}
