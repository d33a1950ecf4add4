// ps_kernels.cuh — hand-written sm_100a kernels for PowerServe's decode / prefill hot path (table-op granularity).
//
// Every kernel cites the reference function it replaces (paths relative to /root/reference).  All of them are
// HBM/L2-bound integer / fp32 work; the data-path rule is the same everywhere: a warp lane owns one "AVX lane"
// (four consecutive bytes of a 32-byte group == one dp4a), partial sums stay integer until the reference converts
// them, and the fp32 accumulation chains are replayed in the reference's order (see ps_math.cuh).
#pragma once
#include "ps_math.cuh"

#define PS_FULL 0xffffffffu

// ggml block sizes in bytes (libs/ggml/src/ggml-common.h:158-162, 200-204, 299-310, 335-340)
#define PS_Q4_0_BYTES 18
#define PS_Q8_0_BYTES 34
#define PS_Q4_K_BYTES 144
#define PS_Q5_K_BYTES 176
#define PS_Q6_K_BYTES 210

__host__ __device__ inline int64_t ps_row_bytes(int type, int64_t k) {
    switch (type) {
    case 0: return k * 4;
    case 1: return k * 2;
    case 2: return k / 32 * PS_Q4_0_BYTES;
    case 8: return k / 32 * PS_Q8_0_BYTES;
    case 12: return k / 256 * PS_Q4_K_BYTES;
    case 13: return k / 256 * PS_Q5_K_BYTES;
    case 14: return k / 256 * PS_Q6_K_BYTES;
    default: return 0;
    }
}

PS_D float ps_half_bits_to_float(uint32_t h16) { return __half2float(__ushort_as_half((unsigned short)h16)); }
PS_D uint32_t ps_ld_u32_a2(const uint8_t *p) { // 4 bytes from a 2-byte-aligned address
    const unsigned short *q = reinterpret_cast<const unsigned short *>(p);
    return (uint32_t)q[0] | ((uint32_t)q[1] << 16);
}

// hsum_float_8 (libs/ggml/src/ggml-quants.c:62-68) over the values held by lanes 0..7 of the warp; all lanes call.
PS_D float ps_hsum8_lanes(float v) {
    float x[8];
#pragma unroll
    for (int l = 0; l < 8; l++) x[l] = __shfl_sync(PS_FULL, v, l);
    const float r0 = __fadd_rn(x[4], x[0]), r1 = __fadd_rn(x[5], x[1]), r2 = __fadd_rn(x[6], x[2]), r3 = __fadd_rn(x[7], x[3]);
    return __fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
}

// GGML_F32x8_REDUCE (libs/ggml/src/ggml.c:1354-1372) when lane t = 8*j + l holds sum[j][l]: the five butterfly
// steps below are exactly x0+=x2, x1+=x3; x0+=x1; lo128+hi128; hadd; hadd (fp add is commutative, so every lane ends
// with the same value the reference leaves in element 0).
PS_D float ps_f32x8_reduce(float v) {
    v = __fadd_rn(v, __shfl_xor_sync(PS_FULL, v, 16));
    v = __fadd_rn(v, __shfl_xor_sync(PS_FULL, v, 8));
    v = __fadd_rn(v, __shfl_xor_sync(PS_FULL, v, 4));
    v = __fadd_rn(v, __shfl_xor_sync(PS_FULL, v, 1));
    v = __fadd_rn(v, __shfl_xor_sync(PS_FULL, v, 2));
    return v;
}

// ====================================================================================================================
// Activation quantisers  (replaces the from_float step of powerserve_compute_forward_mul_mat, ggml.c:13502-13530)
// ====================================================================================================================
// Device layout of a Q8_K-quantised activation column (K elements, nb = K/256 blocks) — "lane-major" so that the
// matvec's lane l reads its eight dp4a operands of block i with two 16-byte shared-memory loads:
//   qs  [nb][2][8] uint4 : word (h, l, w) = bytes of elements 32*(4h+w) + 4l .. +3      (K bytes)
//   d   [nb] float       : block scale (block_q8_K::d)
//   bsp [nb][4] uint32   : int16 pair (s_{2k}, s_{2k+1}), s_j = sum of the 32 quants of sub-block j
//                          (== hadd_epi16 of block_q8_K::bsums, ggml-quants.c:7829-7830)

// quantize_row_q8_K_ref (libs/ggml/src/ggml-quants.c:3799-3837).  One warp per 256-block.
__global__ void __launch_bounds__(128) ps_k_quantize_q8k(const float *__restrict__ x, int64_t K, uint32_t *__restrict__ qs,
                                                         float *__restrict__ dq, uint32_t *__restrict__ bsp) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nb = K / 256;
    const int64_t i = (int64_t)blockIdx.x * 4 + warp;
    const int64_t col = blockIdx.y;
    if (i >= nb) return;
    const int l = lane & 7, jj = lane >> 3;
    const float *xb = x + col * K + i * 256;
    const float4 v0 = *reinterpret_cast<const float4 *>(xb + 32 * jj + 4 * l);
    const float4 v1 = *reinterpret_cast<const float4 *>(xb + 32 * (jj + 4) + 4 * l);
    const float e[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
    const int idx0 = 32 * jj + 4 * l, idx1 = 32 * (jj + 4) + 4 * l;
    float amax = 0.f;
#pragma unroll
    for (int t = 0; t < 8; t++) amax = fmaxf(amax, fabsf(e[t]));
#pragma unroll
    for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(PS_FULL, amax, o));
    // `if (ax > amax) { amax = ax; max = x[j]; }` keeps the FIRST element that attains the maximum magnitude
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
    {   // broadcast the signed maximum from its owner
        const unsigned owner = __ballot_sync(PS_FULL, (first >= idx0 && first < idx0 + 4) || (first >= idx1 && first < idx1 + 4));
        mx = __shfl_sync(PS_FULL, mx, __ffs(owner) - 1);
    }
    uint32_t *qcol = qs + (col * nb + i) * 64; // 64 words per block
    if (amax == 0.f) {
        qcol[(0 * 8 + l) * 4 + jj] = 0;
        qcol[(1 * 8 + l) * 4 + jj] = 0;
        if (lane == 0) dq[col * nb + i] = 0.f;
        if (lane < 4) bsp[(col * nb + i) * 4 + lane] = 0;
        return;
    }
    const float iscale = __fdiv_rn(-127.f, mx);
    int q[8];
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const float val = __fadd_rn(__fmul_rn(iscale, e[t]), 12582912.f); // nearest_int, ggml-quants.c:1653-1658
        const int v = (int)(ps_f2u(val) & 0x007fffffu) - 0x00400000;
        q[t] = min(127, v);
    }
    const uint32_t w0 = (uint32_t)(q[0] & 0xff) | ((uint32_t)(q[1] & 0xff) << 8) | ((uint32_t)(q[2] & 0xff) << 16) | ((uint32_t)(q[3] & 0xff) << 24);
    const uint32_t w1 = (uint32_t)(q[4] & 0xff) | ((uint32_t)(q[5] & 0xff) << 8) | ((uint32_t)(q[6] & 0xff) << 16) | ((uint32_t)(q[7] & 0xff) << 24);
    qcol[(0 * 8 + l) * 4 + jj] = w0; // sub-block j = jj      -> half 0, word jj
    qcol[(1 * 8 + l) * 4 + jj] = w1; // sub-block j = jj + 4  -> half 1, word jj
    int s0 = q[0] + q[1] + q[2] + q[3], s1 = q[4] + q[5] + q[6] + q[7];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        s0 += __shfl_xor_sync(PS_FULL, s0, o);
        s1 += __shfl_xor_sync(PS_FULL, s1, o);
    }
    // lane (jj, l=0) holds s_jj and s_{jj+4}; pair k packs (s_{2k}, s_{2k+1})
    int sj[8];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        sj[j] = __shfl_sync(PS_FULL, s0, j * 8);
        sj[j + 4] = __shfl_sync(PS_FULL, s1, j * 8);
    }
    if (lane < 4) bsp[(col * nb + i) * 4 + lane] = ((uint32_t)sj[2 * lane] & 0xffffu) | ((uint32_t)sj[2 * lane + 1] << 16);
    if (lane == 0) dq[col * nb + i] = __fdiv_rn(1.f, iscale);
}

// quantize_row_q8_0, AVX2 branch (libs/ggml/src/ggml-quants.c:957-1017): d = max|x| / 127 stored as fp16,
// id = 127 / max|x|, round-half-to-even.  One warp handles four 32-blocks (lane = 8*b + l owns word l of block b).
// Layout: qs [nb32][8] uint32 natural order, d [nb32] float holding fp16(d) widened back (what the dot product reads).
__global__ void __launch_bounds__(128) ps_k_quantize_q80(const float *__restrict__ x, int64_t K, uint32_t *__restrict__ qs,
                                                         float *__restrict__ dq) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t nb = K / 32;
    const int64_t i = ((int64_t)blockIdx.x * 4 + warp) * 4 + (lane >> 3);
    const int64_t col = blockIdx.y;
    const int l = lane & 7;
    const bool valid = i < nb;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (valid) v = *reinterpret_cast<const float4 *>(x + col * K + i * 32 + 4 * l);
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) amax = fmaxf(amax, __shfl_xor_sync(PS_FULL, amax, o));
    if (!valid) return;
    const float d = __fdiv_rn(amax, 127.f);
    const float id = (amax != 0.0f) ? __fdiv_rn(127.f, amax) : 0.0f;
    const int q0 = __float2int_rn(__fmul_rn(v.x, id)), q1 = __float2int_rn(__fmul_rn(v.y, id));
    const int q2 = __float2int_rn(__fmul_rn(v.z, id)), q3 = __float2int_rn(__fmul_rn(v.w, id));
    qs[(col * nb + i) * 8 + l] = (uint32_t)(q0 & 0xff) | ((uint32_t)(q1 & 0xff) << 8) | ((uint32_t)(q2 & 0xff) << 16) | ((uint32_t)(q3 & 0xff) << 24);
    if (l == 0) dq[col * nb + i] = __half2float(__float2half_rn(d));
}

// ====================================================================================================================
// Quantised weight x quantised activation  (replaces powerserve_compute_forward_mul_mat -> vec_dot, ggml.c:13344-13432)
// ====================================================================================================================
// Lane t = 8*b + l of a warp owns AVX lane l of block (4*it + b) of the row: its integer partial S (and, for Q4_K, the
// mins product P) are what one lane of the reference's __m256i sumi / __m128i prod holds after the block.  The fp32
// accumulators acc[l] (and acc_m[k]) are then advanced block by block IN ROW ORDER with one FMA each, the values
// travelling between lanes by shuffle, and reduced once at the end in hsum_float_8 order.

struct PsActQ8K { const uint4 *qs; const float *d; const uint32_t *bsp; };  // shared-memory views of one column
struct PsActQ80 { const uint32_t *qs; const float *d; };

template <int TYPE> struct PsBlk;

// ggml_vec_dot_q4_K_q8_K, AVX2 branch (libs/ggml/src/ggml-quants.c:7809-7872)
template <> struct PsBlk<12> {
    static constexpr int BYTES = PS_Q4_K_BYTES, ELEMS = 256;
    static constexpr bool HAS_MIN = true, MIN_SCALAR = false;
    uint32_t q4[4], scA, scB, mA, mB;
    float xd, xmin;
    PS_D void load(const uint8_t *blk, int l) {
        const uint4 h = *reinterpret_cast<const uint4 *>(blk);
        xd = ps_half_bits_to_float(h.x & 0xffffu);
        xmin = ps_half_bits_to_float(h.x >> 16);
        // the utmp / kmask shuffle of :7816-7826 (== get_scale_min_k4 for all eight j)
        const uint32_t k1 = 0x3f3f3f3fu, k2 = 0x0f0f0f0fu, k3 = 0x03030303u;
        mB = ((h.w >> 4) & k2) | (((h.z >> 6) & k3) << 4);
        mA = h.z & k1;
        scB = (h.w & k2) | (((h.y >> 6) & k3) << 4);
        scA = h.y & k1;
        const uint32_t *q = reinterpret_cast<const uint32_t *>(blk + 16) + l;
#pragma unroll
        for (int j = 0; j < 4; j++) q4[j] = q[8 * j];
    }
    PS_D void partial(const PsActQ8K &a, int64_t i, int l, int &S, float &d, int &P, float &dm) const {
        const uint4 lo = a.qs[(i * 2 + 0) * 8 + l], hi = a.qs[(i * 2 + 1) * 8 + l];
        const int q8[8] = {(int)lo.x, (int)lo.y, (int)lo.z, (int)lo.w, (int)hi.x, (int)hi.y, (int)hi.z, (int)hi.w};
        S = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int p0 = __dp4a((int)(q4[j] & 0x0f0f0f0fu), q8[2 * j], 0);
            const int p1 = __dp4a((int)((q4[j] >> 4) & 0x0f0f0f0fu), q8[2 * j + 1], 0);
            const uint32_t scw = (j < 2) ? scA : scB;
            const int s0 = (scw >> (16 * (j & 1))) & 0xff, s1 = (scw >> (16 * (j & 1) + 8)) & 0xff;
            S += s0 * p0 + s1 * p1;
        }
        const float yd = a.d[i];
        d = __fmul_rn(yd, xd);
        dm = __fmul_rn(-yd, xmin);
        // prod lane k = l & 3: m_{2k} * s_{2k} + m_{2k+1} * s_{2k+1}  (madd_epi16 of mins with the hadd'ed bsums)
        const int k = l & 3;
        const uint32_t mw = (k < 2) ? mA : mB;
        const int m0 = (mw >> (16 * (k & 1))) & 0xff, m1 = (mw >> (16 * (k & 1) + 8)) & 0xff;
        const uint32_t bs = a.bsp[i * 4 + k];
        P = m0 * (int)(short)(bs & 0xffffu) + m1 * (int)(short)(bs >> 16);
    }
};

// ggml_vec_dot_q5_K_q8_K, AVX2 branch (libs/ggml/src/ggml-quants.c:8382-8460): block_q5_K = d, dmin, scales[12], qh[32], qs[128]
// (ggml-common.h); the lane sums of Q4_K with the fifth bit of sub-block s taken from bit s of qh[e]; the mins go through a
// SCALAR float, summs += dmin * sum_k prod_k - a separate multiply and add in the reference build (vmulss + vaddss; pinned against the compiled reference on the CPU).
template <> struct PsBlk<13> {
    static constexpr int BYTES = PS_Q5_K_BYTES, ELEMS = 256;
    static constexpr bool HAS_MIN = true, MIN_SCALAR = true;
    uint32_t q4[4], qh, scA, scB, mA, mB;
    float xd, xmin;
    PS_D void load(const uint8_t *blk, int l) {
        const uint4 h = *reinterpret_cast<const uint4 *>(blk);
        xd = ps_half_bits_to_float(h.x & 0xffffu);
        xmin = ps_half_bits_to_float(h.x >> 16);
        const uint32_t k1 = 0x3f3f3f3fu, k2 = 0x0f0f0f0fu, k3 = 0x03030303u;
        mB = ((h.w >> 4) & k2) | (((h.z >> 6) & k3) << 4);
        mA = h.z & k1;
        scB = (h.w & k2) | (((h.y >> 6) & k3) << 4);
        scA = h.y & k1;
        qh = reinterpret_cast<const uint32_t *>(blk + 16)[l];
        const uint32_t *q = reinterpret_cast<const uint32_t *>(blk + 48) + l;
#pragma unroll
        for (int j = 0; j < 4; j++) q4[j] = q[8 * j];
    }
    PS_D void partial(const PsActQ8K &a, int64_t i, int l, int &S, float &d, int &P, float &dm) const {
        const uint4 lo = a.qs[(i * 2 + 0) * 8 + l], hi = a.qs[(i * 2 + 1) * 8 + l];
        const int q8[8] = {(int)lo.x, (int)lo.y, (int)lo.z, (int)lo.w, (int)hi.x, (int)hi.y, (int)hi.z, (int)hi.w};
        S = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t v0 = (q4[j] & 0x0f0f0f0fu) | (((qh >> (2 * j)) & 0x01010101u) << 4);
            const uint32_t v1 = ((q4[j] >> 4) & 0x0f0f0f0fu) | (((qh >> (2 * j + 1)) & 0x01010101u) << 4);
            const int p0 = __dp4a((int)v0, q8[2 * j], 0), p1 = __dp4a((int)v1, q8[2 * j + 1], 0); // 5-bit values: fine as signed bytes
            const uint32_t scw = (j < 2) ? scA : scB;
            const int s0 = (scw >> (16 * (j & 1))) & 0xff, s1 = (scw >> (16 * (j & 1) + 8)) & 0xff;
            S += s0 * p0 + s1 * p1;
        }
        const float yd = a.d[i];
        d = __fmul_rn(yd, xd);
        dm = __fmul_rn(-yd, xmin);
        const int k = l & 3;
        const uint32_t mw = (k < 2) ? mA : mB;
        const int m0 = (mw >> (16 * (k & 1))) & 0xff, m1 = (mw >> (16 * (k & 1) + 8)) & 0xff;
        const uint32_t bs = a.bsp[i * 4 + k];
        P = m0 * (int)(short)(bs & 0xffffu) + m1 * (int)(short)(bs >> 16);
    }
};

// ggml_vec_dot_q6_K_q8_K, AVX2 branch (libs/ggml/src/ggml-quants.c:9039-9116)
template <> struct PsBlk<14> {
    static constexpr int BYTES = PS_Q6_K_BYTES, ELEMS = 256;
    static constexpr bool HAS_MIN = false, MIN_SCALAR = false;
    uint32_t ql[4], qh[2];
    int sc[8];
    float xd;
    PS_D void load(const uint8_t *blk, int l) {
#pragma unroll
        for (int jh = 0; jh < 2; jh++) {
            ql[2 * jh + 0] = ps_ld_u32_a2(blk + 64 * jh + 4 * l);
            ql[2 * jh + 1] = ps_ld_u32_a2(blk + 64 * jh + 32 + 4 * l);
            qh[jh] = ps_ld_u32_a2(blk + 128 + 32 * jh + 4 * l);
        }
#pragma unroll
        for (int j = 0; j < 8; j++) sc[j] = (int)(signed char)blk[192 + 2 * j + (l >= 4)];
        xd = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 208));
    }
    PS_D void partial(const PsActQ8K &a, int64_t i, int l, int &S, float &d, int &P, float &dm) const {
        const uint4 lo = a.qs[(i * 2 + 0) * 8 + l], hi = a.qs[(i * 2 + 1) * 8 + l];
        const int q8[8] = {(int)lo.x, (int)lo.y, (int)lo.z, (int)lo.w, (int)hi.x, (int)hi.y, (int)hi.z, (int)hi.w};
        S = 0;
#pragma unroll
        for (int jh = 0; jh < 2; jh++)
#pragma unroll
            for (int g = 0; g < 4; g++) {
                const uint32_t lw = ql[2 * jh + (g & 1)];
                const uint32_t nib = (g < 2) ? (lw & 0x0f0f0f0fu) : ((lw >> 4) & 0x0f0f0f0fu);
                const uint32_t q = __vsub4(nib | (((qh[jh] >> (2 * g)) & 0x03030303u) << 4), 0x20202020u);
                S += sc[4 * jh + g] * __dp4a((int)q, q8[4 * jh + g], 0);
            }
        d = __fmul_rn(a.d[i], xd);
        P = 0;
        dm = 0.f;
    }
};

// ggml_vec_dot_q4_0_q8_0, AVX2 branch (libs/ggml/src/ggml-quants.c:4205-4228)
template <> struct PsBlk<2> {
    static constexpr int BYTES = PS_Q4_0_BYTES, ELEMS = 32;
    static constexpr bool HAS_MIN = false, MIN_SCALAR = false;
    uint32_t q;
    float xd;
    PS_D void load(const uint8_t *blk, int l) {
        const uint32_t w = ps_ld_u32_a2(blk + 2 + 4 * (l & 3));
        q = __vsub4((l < 4) ? (w & 0x0f0f0f0fu) : ((w >> 4) & 0x0f0f0f0fu), 0x08080808u);
        xd = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
    }
    PS_D void partial(const PsActQ80 &a, int64_t i, int l, int &S, float &d, int &P, float &dm) const {
        S = __dp4a((int)q, (int)a.qs[i * 8 + l], 0);
        d = __fmul_rn(xd, a.d[i]);
        P = 0;
        dm = 0.f;
    }
};

// ggml_vec_dot_q8_0_q8_0, AVX2 branch (libs/ggml/src/ggml-quants.c:5761-5782)
template <> struct PsBlk<8> {
    static constexpr int BYTES = PS_Q8_0_BYTES, ELEMS = 32;
    static constexpr bool HAS_MIN = false, MIN_SCALAR = false;
    uint32_t q;
    float xd;
    PS_D void load(const uint8_t *blk, int l) {
        q = ps_ld_u32_a2(blk + 2 + 4 * l);
        xd = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
    }
    PS_D void partial(const PsActQ80 &a, int64_t i, int l, int &S, float &d, int &P, float &dm) const {
        S = __dp4a((int)q, (int)a.qs[i * 8 + l], 0);
        d = __fmul_rn(xd, a.d[i]);
        P = 0;
        dm = 0.f;
    }
};

// dst{N, bs} = W{K, N} . x{K, bs}; one warp per weight row, C activation columns per pass (gridDim.y passes).
// Shared memory holds the C quantised columns.  Row results are bit-identical for every C (each column has its own
// chain), which tests/test_gpu_ops.py checks.
template <int TYPE, int C>
__global__ void __launch_bounds__(256) ps_k_matmul_q(const uint8_t *__restrict__ w, int64_t K, int64_t N, int64_t bs,
                                                     const uint32_t *__restrict__ aqs, const float *__restrict__ ad,
                                                     const uint32_t *__restrict__ absp, float *__restrict__ dst,
                                                     const float *__restrict__ bias, const float *__restrict__ residual) {
    using B = PsBlk<TYPE>;
    constexpr bool KQ = (B::ELEMS == 256);
    extern __shared__ __align__(16) uint8_t smem[];
    const int64_t nb = K / B::ELEMS;
    const int64_t col0 = (int64_t)blockIdx.y * C;
    const int ncol = (int)min((int64_t)C, bs - col0);
    // ---- stage the quantised activation columns
    const int64_t qwords = K / 4;                       // words of quants per column
    uint32_t *s_qs = reinterpret_cast<uint32_t *>(smem);                      // [C][K/4]
    float *s_d = reinterpret_cast<float *>(s_qs + (int64_t)C * qwords);       // [C][nb]
    uint32_t *s_bsp = reinterpret_cast<uint32_t *>(s_d + (int64_t)C * nb);    // [C][nb*4] (K-quants only)
    for (int c = 0; c < ncol; c++) {
        const uint4 *src = reinterpret_cast<const uint4 *>(aqs + (col0 + c) * qwords);
        uint4 *dstq = reinterpret_cast<uint4 *>(s_qs + c * qwords);
        for (int64_t t = threadIdx.x; t < qwords / 4; t += blockDim.x) dstq[t] = src[t];
        for (int64_t t = threadIdx.x; t < nb; t += blockDim.x) s_d[c * nb + t] = ad[(col0 + c) * nb + t];
        if (KQ)
            for (int64_t t = threadIdx.x; t < nb * 4; t += blockDim.x) s_bsp[c * nb * 4 + t] = absp[(col0 + c) * nb * 4 + t];
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, l = lane & 7, b = lane >> 3;
    const int64_t n = (int64_t)blockIdx.x * 8 + warp;
    if (n >= N) return;
    const uint8_t *wrow = w + n * nb * B::BYTES;
    float acc[C], accm[C];
#pragma unroll
    for (int c = 0; c < C; c++) { acc[c] = 0.f; accm[c] = 0.f; }
    for (int64_t i0 = 0; i0 < nb; i0 += 4) {
        const int64_t i = i0 + b;
        const bool valid = i < nb;
        B blk;
        if (valid) blk.load(wrow + i * B::BYTES, l);
        const int nvalid = (int)min((int64_t)4, nb - i0);
#pragma unroll
        for (int c = 0; c < C; c++) {
            if (c < ncol) {
                int S = 0, P = 0;
                float d = 0.f, dm = 0.f;
                if (valid) {
                    if constexpr (KQ) {
                        PsActQ8K a{reinterpret_cast<const uint4 *>(s_qs + c * qwords), s_d + c * nb, s_bsp + c * nb * 4};
                        blk.partial(a, i, l, S, d, P, dm);
                    } else {
                        PsActQ80 a{s_qs + c * qwords, s_d + c * nb};
                        blk.partial(a, i, l, S, d, P, dm);
                    }
                }
                // advance the 8 (+4) accumulator chains over the (up to) four blocks of this step, in row order
                for (int bb = 0; bb < nvalid; bb++) {
                    const int Sb = __shfl_sync(PS_FULL, S, bb * 8 + l);
                    const float db = __shfl_sync(PS_FULL, d, bb * 8 + l);
                    acc[c] = __fmaf_rn(db, __int2float_rn(Sb), acc[c]);
                    if constexpr (B::HAS_MIN && B::MIN_SCALAR) { // summs += dmin * hsum(prod): multiply, then add (Q5_K)
                        int Pt = 0;
#pragma unroll
                        for (int k = 0; k < 4; k++) Pt += __shfl_sync(PS_FULL, P, bb * 8 + k);
                        const float dmb = __shfl_sync(PS_FULL, dm, bb * 8 + l);
                        accm[c] = __fadd_rn(accm[c], __fmul_rn(dmb, __int2float_rn(Pt)));
                    } else if constexpr (B::HAS_MIN) {
                        const int Pb = __shfl_sync(PS_FULL, P, bb * 8 + (l & 3));
                        const float dmb = __shfl_sync(PS_FULL, dm, bb * 8 + l);
                        accm[c] = __fmaf_rn(dmb, __int2float_rn(Pb), accm[c]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; c++) {
        if (c < ncol) {
            float r = ps_hsum8_lanes(acc[c]);
            if constexpr (B::HAS_MIN && B::MIN_SCALAR) {
                r = __fadd_rn(r, accm[c]); // hsum_float_8(acc) + summs
            } else if constexpr (B::HAS_MIN) {
                // acc_m = add(acc_m, movehl(acc_m)); add_ss(acc_m, movehdup(acc_m))   (ggml-quants.c:7868-7871)
                const float m0 = __shfl_sync(PS_FULL, accm[c], 0), m1 = __shfl_sync(PS_FULL, accm[c], 1);
                const float m2 = __shfl_sync(PS_FULL, accm[c], 2), m3 = __shfl_sync(PS_FULL, accm[c], 3);
                r = __fadd_rn(r, __fadd_rn(__fadd_rn(m0, m2), __fadd_rn(m1, m3)));
            }
            if (lane == 0) {
                const int64_t o = (col0 + c) * N + n;
                if (bias) r = __fadd_rn(r, bias[n]);          // GGMLBackend::add with row broadcast (Qwen2 q/k/v bias)
                if (residual) r = __fadd_rn(residual[o], r);  // residual add: x + W.h  (powerserve_compute_forward_add)
                dst[o] = r;
            }
        }
    }
}

// ====================================================================================================================
// Small fp32 operators
// ====================================================================================================================
// GGMLBackend::get_embedding (src/backend/ggml/ggml_wrapper.cpp:181-211) + dequantize_row_* (ggml-quants.c:1536,
// 1630, 2569, 2991).  One CTA per token row.
__global__ void __launch_bounds__(256) ps_k_get_embedding(float *__restrict__ dst, const uint8_t *__restrict__ w, int type,
                                                          int64_t dim, const int32_t *__restrict__ tokens) {
    const int64_t tok = tokens[blockIdx.x];
    const uint8_t *row = w + tok * ps_row_bytes(type, dim);
    float *y = dst + (int64_t)blockIdx.x * dim;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) {
        float v;
        if (type == 0) {
            v = reinterpret_cast<const float *>(row)[e];
        } else if (type == 2) {
            const uint8_t *blk = row + (e / 32) * PS_Q4_0_BYTES;
            const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
            const int j = (int)(e % 32);
            const int q = (j < 16) ? (blk[2 + j] & 0x0F) - 8 : (blk[2 + j - 16] >> 4) - 8;
            v = __fmul_rn((float)q, d);
        } else if (type == 8) {
            const uint8_t *blk = row + (e / 32) * PS_Q8_0_BYTES;
            const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
            v = __fmul_rn((float)(int)(signed char)blk[2 + e % 32], d);
        } else if (type == 12) {
            const uint8_t *blk = row + (e / 256) * PS_Q4_K_BYTES;
            const int r = (int)(e % 256), j = r / 32, el = r % 32;
            const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
            const float mn = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 2));
            const uint8_t *sc = blk + 4;
            int s, m; // get_scale_min_k4, ggml-quants.c:1912-1919
            if (j < 4) { s = sc[j] & 63; m = sc[j + 4] & 63; }
            else { s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
            const uint8_t qb = blk[16 + 32 * (j / 2) + el];
            const int q = (j & 1) ? (qb >> 4) : (qb & 0xF);
            // `d1 * q - m1` is one FMA in the reference build (gcc -O3 -mfma contracts it)
            v = __fmaf_rn(__fmul_rn(d, (float)s), (float)q, -__fmul_rn(mn, (float)m));
        } else if (type == 1) { // F16 table (GGML_FP16_TO_FP32)
            v = ps_half_bits_to_float(reinterpret_cast<const unsigned short *>(row)[e]);
        } else if (type == 13) { // dequantize_row_q5_K (ggml-quants.c:2776-2803)
            const uint8_t *blk = row + (e / 256) * PS_Q5_K_BYTES;
            const int r = (int)(e % 256), j = r / 32, el = r % 32;
            const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk));
            const float mn = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 2));
            const uint8_t *sc = blk + 4;
            int s, m;
            if (j < 4) { s = sc[j] & 63; m = sc[j + 4] & 63; }
            else { s = (sc[j + 4] & 0xF) | ((sc[j - 4] >> 6) << 4); m = (sc[j + 4] >> 4) | ((sc[j] >> 6) << 4); }
            const uint8_t qb = blk[48 + 32 * (j / 2) + el];
            const int q = ((j & 1) ? (qb >> 4) : (qb & 0xF)) + (((blk[16 + el] >> j) & 1) << 4);
            v = __fmaf_rn(__fmul_rn(d, (float)s), (float)q, -__fmul_rn(mn, (float)m));
        } else { // 14: Q6_K
            const uint8_t *blk = row + (e / 256) * PS_Q6_K_BYTES;
            const int r = (int)(e % 256), half = r / 128, rr = r % 128, g = rr / 32, lq = rr % 32;
            const float d = ps_half_bits_to_float(*reinterpret_cast<const unsigned short *>(blk + 208));
            const uint8_t lb = blk[64 * half + (g & 1) * 32 + lq], hb = blk[128 + 32 * half + lq];
            const int q = (int)(signed char)(((g < 2) ? (lb & 0xF) : (lb >> 4)) | (((hb >> (2 * g)) & 3) << 4)) - 32;
            const int s = (int)(signed char)blk[192 + 8 * half + 2 * g + lq / 16];
            v = __fmul_rn(__fmul_rn(d, (float)s), (float)q);
        }
        y[e] = v;
    }
}

// block-wide sum of doubles (tree order; see DESIGN.md "double-precision sums")
PS_D double ps_block_sum_double(double v, double *sh) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(PS_FULL, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane == 0) sh[warp] = v;
    __syncthreads();
    double t = (threadIdx.x < nw) ? sh[threadIdx.x] : 0.0;
    if (warp == 0) {
#pragma unroll
        for (int o = 16; o; o >>= 1) t += __shfl_xor_sync(PS_FULL, t, o);
        if (lane == 0) sh[0] = t;
    }
    __syncthreads();
    return sh[0];
}

// powerserve_compute_forward_rms_norm_f32 (libs/ggml/src/ggml.c:12667-12721): y = x * (w * 1/sqrtf(mean(x^2)+eps)),
// the sum of fp32 squares accumulated in double.  One CTA per row.
__global__ void __launch_bounds__(256) ps_k_rmsnorm(float *__restrict__ dst, const float *__restrict__ x, const float *__restrict__ w,
                                                    int64_t dim, float eps) {
    __shared__ double sh[32];
    const float *xr = x + (int64_t)blockIdx.x * dim;
    float *yr = dst + (int64_t)blockIdx.x * dim;
    double s = 0.0;
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) s += (double)__fmul_rn(xr[e], xr[e]);
    const double sum = ps_block_sum_double(s, sh);
    const float mean = (float)(sum / (double)dim);
    const float scale = __fdiv_rn(1.0f, __fsqrt_rn(__fadd_rn(mean, eps)));
    for (int64_t e = threadIdx.x; e < dim; e += blockDim.x) yr[e] = __fmul_rn(xr[e], __fmul_rn(w[e], scale));
}

// ggml_compute_forward_rope_f32 (libs/ggml/src/ggml.c:15368-15497) with the cos/sin cache of ggml_rope_cache_init
// (:15342-15356) precomputed for every position by the host (ps_cuda.cu: build_rope_table) — table[pos][i0] = cos,
// table[pos][i0+1] = sin.  src/dst {head_size, n_heads, bs}; grid (n_heads, bs).
__global__ void ps_k_rope(float *__restrict__ dst, const float *__restrict__ src, int head_size, int n_dims, int neox,
                          const int32_t *__restrict__ pos, const float *__restrict__ table) {
    const int64_t row = (int64_t)blockIdx.y * gridDim.x + blockIdx.x;
    const float *s = src + row * head_size;
    float *d = dst + row * head_size;
    const float *cache = table + (int64_t)pos[blockIdx.y] * head_size;
    for (int p = threadIdx.x; p < head_size / 2; p += blockDim.x) {
        const int i0 = 2 * p;
        if (i0 < n_dims) {
            const float c = cache[i0], sn = cache[i0 + 1];
            const int a = neox ? p : i0, bidx = neox ? p + n_dims / 2 : i0 + 1;
            const float x0 = s[a], x1 = s[bidx];
            d[a] = __fadd_rn(__fmul_rn(x0, c), -__fmul_rn(x1, sn));    // x0*c - x1*s  (not fused in the reference)
            d[bidx] = __fadd_rn(__fmul_rn(x0, sn), __fmul_rn(x1, c));  // x0*s + x1*c
        } else {
            d[i0] = s[i0];
            d[i0 + 1] = s[i0 + 1];
        }
    }
}

// KV store: the two COPY ops of NormAttention::build (src/model/module/norm_attention.cpp:79-105).
// K cache [n_ctx][kv_dim] <- rope(k) rows at pos0..; V cache TRANSPOSED [kv_dim][n_ctx] <- v columns at pos0..
__global__ void ps_k_kv_store(float *__restrict__ kc, float *__restrict__ vct, const float *__restrict__ k, const float *__restrict__ v,
                              int64_t kv_dim, int64_t n_ctx, const int32_t *__restrict__ pos, int64_t bs) {
    const int64_t pos0 = pos[0];
    const int64_t total = kv_dim * bs;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / kv_dim, e = t % kv_dim;
        kc[(pos0 + i) * kv_dim + e] = k[t];
        vct[e * n_ctx + pos0 + i] = v[t];
    }
}

// ---- speculative decode: the KV cache as KVCacheInterface sees it (src/core/kv_cache.hpp:97-276) - cache
// SLOTS decoupled from token positions, a per-slot mask, and the last batch's K / V kept aside for `copy`.
// KV store of a batch at cache slots base .. base + bs - 1 (save_tokens) + a copy into the per-layer staging rows
__global__ void ps_k_kv_store_at(float *__restrict__ kc, float *__restrict__ vct, float *__restrict__ k_stage, float *__restrict__ v_stage,
                                 const float *__restrict__ k, const float *__restrict__ v, int64_t kv_dim, int64_t n_ctx, int64_t base, int64_t bs) {
    const int64_t total = kv_dim * bs;
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / kv_dim, e = t % kv_dim;
        kc[(base + i) * kv_dim + e] = k[t];
        vct[e * n_ctx + base + i] = v[t];
        k_stage[t] = k[t];
        v_stage[t] = v[t];
    }
}
// KVCacheInterface::copy (kv_cache.hpp:120-127, 188-203): token `src` of the last batch -> cache slot `dst`, every layer.
// kv = [2 * n_layers] device pointers (K caches, then transposed V caches); stage = [n_layers][max_batch][kv_dim] x 2.
__global__ void ps_k_kv_copy_slot(float *const *__restrict__ kv, int n_layers, const float *__restrict__ k_stage, const float *__restrict__ v_stage,
                                  int64_t layer_stride, int64_t kv_dim, int64_t n_ctx, int64_t dst, int64_t src) {
    const int L = blockIdx.y;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < kv_dim; e += (int64_t)gridDim.x * blockDim.x) {
        kv[L][dst * kv_dim + e] = k_stage[L * layer_stride + src * kv_dim + e];
        kv[n_layers + L][e * n_ctx + dst] = v_stage[L * layer_stride + src * kv_dim + e];
    }
}
// KVCacheInterface::move (kv_cache.hpp:205-221): cache slot `src` -> cache slot `dst`, every layer
__global__ void ps_k_kv_move_slot(float *const *__restrict__ kv, int n_layers, int64_t kv_dim, int64_t n_ctx, int64_t dst, int64_t src) {
    const int L = blockIdx.y;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < kv_dim; e += (int64_t)gridDim.x * blockDim.x) {
        kv[L][dst * kv_dim + e] = kv[L][src * kv_dim + e];
        kv[n_layers + L][e * n_ctx + dst] = kv[n_layers + L][e * n_ctx + src];
    }
}
// attention bias of a tree batch (CausalLM::fill_attention_mask, src/backend/qnn/causal_models.cpp:215-230, with the per-slot
// mask of the cache in front): row i sees cache slot j < base unless the slot is masked, and batch token j - base iff
// tree[i][j - base].  mask {n_kv, bs}, 0 or -inf, the layout softmax_ext expects.
__global__ void ps_k_tree_mask(float *__restrict__ mask, const uint8_t *__restrict__ slot_mask, const uint8_t *__restrict__ tree, int64_t base,
                               int64_t bs, int64_t n_kv) {
    const int64_t i = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_kv; j += (int64_t)gridDim.x * blockDim.x) {
        const bool vis = j < base ? !slot_mask[j] : tree[i * bs + (j - base)] != 0;
        mask[i * n_kv + j] = vis ? 0.f : -INFINITY;
    }
}

// GET_MASK (src/executor/executor.cpp:210-224)
__global__ void ps_k_get_mask(float *__restrict__ mask, int64_t n_kv, const int32_t *__restrict__ pos) {
    const int64_t i = blockIdx.y;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n_kv; j += (int64_t)gridDim.x * blockDim.x)
        mask[j + i * n_kv] = (j <= (int64_t)pos[i]) ? 0.f : -INFINITY;
}

// mat_mul(k_view, q) (norm_attention.cpp:115-129) -> ggml_vec_dot_f32 (ggml.c:2092-2131), head_size a multiple of 32.
// One warp per (cache position j, kv head): the K row is read once and dotted with every q head of the group and
// every batch column.  kq {n_kv, bs, n_heads}.
__global__ void __launch_bounds__(128) ps_k_attn_scores(float *__restrict__ kq, const float *__restrict__ kc, const float *__restrict__ q,
                                                        int hs, int n_heads, int n_kv_heads, int64_t n_kv, int bs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t j = (int64_t)blockIdx.x * 4 + warp;
    const int g = blockIdx.y;
    if (j >= n_kv) return;
    const int r2 = n_heads / n_kv_heads;
    const float *krow = kc + j * (int64_t)(hs * n_kv_heads) + g * hs;
    float kv[8];
    const int steps = hs / 32;
#pragma unroll
    for (int s = 0; s < 8; s++) kv[s] = (s < steps) ? krow[32 * s + lane] : 0.f;
    // (query, head) pairs in batches of 8: eight independent FMA chains and eight interleaved butterfly reductions keep
    // the shuffle latency covered (the arithmetic of every dot product is unchanged)
    const int n_pairs = bs * r2;
    for (int p0 = 0; p0 < n_pairs; p0 += 8) {
        float sum[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int p = min(p0 + u, n_pairs - 1), i = p / r2, h = g * r2 + p % r2;
            const float *qv = q + ((int64_t)i * n_heads + h) * hs;
            float qe[8];
#pragma unroll
            for (int s = 0; s < 8; s++) qe[s] = (s < steps) ? qv[32 * s + lane] : 0.f; // all loads of the batch in flight
            sum[u] = 0.f;
#pragma unroll
            for (int s = 0; s < 8; s++)
                if (s < steps) sum[u] = __fmaf_rn(kv[s], qe[s], sum[u]);
        }
#pragma unroll
        for (int k = 0; k < 5; k++) { // GGML_F32x8_REDUCE order (see ps_f32x8_reduce)
            const int step = (k == 0) ? 16 : (k == 1) ? 8 : (k == 2) ? 4 : (k == 3) ? 1 : 2;
            float o[8];
#pragma unroll
            for (int u = 0; u < 8; u++) o[u] = __shfl_xor_sync(PS_FULL, sum[u], step);
#pragma unroll
            for (int u = 0; u < 8; u++) sum[u] = __fadd_rn(sum[u], o[u]);
        }
        if (lane < 8 && p0 + lane < n_pairs) {
            float v = sum[0];
#pragma unroll
            for (int u = 1; u < 8; u++)
                if (lane == u) v = sum[u];
            const int p = p0 + lane, i = p / r2, h = g * r2 + p % r2;
            kq[((int64_t)h * bs + i) * n_kv + j] = v;
        }
    }
}

// ggml_compute_forward_soft_max_f32 (ggml.c:14846-14940) + ggml_vec_soft_max_f32 AVX2 branch (:2814-2868).
// One CTA per row of {ne0, ne1, ne2}.  mask == nullptr -> the position mask of GET_MASK is applied from `pos`; both
// nullptr -> no mask at all (GGMLBackend::softmax: ggml_compute_forward_soft_max_f32 with src1 == NULL).
__global__ void __launch_bounds__(256) ps_k_softmax_ext(float *__restrict__ dst, const float *__restrict__ x, const float *__restrict__ mask,
                                                        const int32_t *__restrict__ pos, int64_t ne0, int64_t ne1, float scale) {
    extern __shared__ float wp[];
    __shared__ double sh[32];
    __shared__ float shf[32];
    const int64_t row = blockIdx.x, i1 = row % ne1;
    const float *sp = x + row * ne0;
    float *dp = dst + row * ne0;
    float mx = -INFINITY;
    for (int64_t j = threadIdx.x; j < ne0; j += blockDim.x) {
        float v = __fmul_rn(sp[j], scale);
        if (mask) v = __fadd_rn(v, mask[i1 * ne0 + j]);
        else if (pos) v = __fadd_rn(v, (j <= (int64_t)pos[i1]) ? 0.f : -INFINITY); // ggml_vec_scale_f32 then `wp[i] += slope*mp[i]`, slope == 1
        wp[j] = v;
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PS_FULL, mx, o));
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = shf[0];
    for (int t = 1; t < (int)((blockDim.x + 31) >> 5); t++) mx = fmaxf(mx, shf[t]);
    const int64_t n8 = ne0 & ~(int64_t)7;
    double s = 0.0;
    for (int64_t gi = threadIdx.x; gi < n8 / 8; gi += blockDim.x) {
        float v[8];
#pragma unroll
        for (int l = 0; l < 8; l++) {
            v[l] = ps_v_expf(__fadd_rn(wp[gi * 8 + l], -mx));
            wp[gi * 8 + l] = v[l]; // the exponentials stay in shared memory: the row goes to global memory once, scaled
        }
        const float r0 = __fadd_rn(v[4], v[0]), r1 = __fadd_rn(v[5], v[1]), r2 = __fadd_rn(v[6], v[2]), r3 = __fadd_rn(v[7], v[3]);
        s += (double)__fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
    }
    for (int64_t j = n8 + threadIdx.x; j < ne0; j += blockDim.x) { // scalar tail: libm expf
        const float v = ps_expf_glibc(__fadd_rn(wp[j], -mx));
        wp[j] = v;
        s += (double)v;
    }
    const double sum = ps_block_sum_double(s, sh);
    const float inv = (float)(1.0 / sum);
    for (int64_t j = threadIdx.x; j < ne0; j += blockDim.x) dp[j] = __fmul_rn(wp[j], inv);
}

// mat_mul(v_view, kq) + permute + cont (norm_attention.cpp:138-151): out[i][h*hs + d] = vec_dot_f32(n_kv, Vt row, P row).
// One warp per (kv head g, d): the V^T row is streamed once for all heads of the group and all columns.
__global__ void __launch_bounds__(128) ps_k_attn_pv(float *__restrict__ out, const float *__restrict__ vct, const float *__restrict__ p,
                                                    int hs, int n_heads, int n_kv_heads, int64_t n_kv, int64_t n_ctx, int bs) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int d = blockIdx.x * 4 + warp, g = blockIdx.y;
    if (d >= hs) return;
    const int r2 = n_heads / n_kv_heads;
    const float *vrow = vct + ((int64_t)g * hs + d) * n_ctx;
    const int64_t np = n_kv & ~(int64_t)31;
    for (int i = 0; i < bs; i++)
        for (int hh = 0; hh < r2; hh++) {
            const int h = g * r2 + hh;
            const float *pr = p + ((int64_t)h * bs + i) * n_kv;
            float sum = 0.f;
            for (int64_t s = 0; s < np; s += 32) sum = __fmaf_rn(vrow[s + lane], pr[s + lane], sum);
            sum = ps_f32x8_reduce(sum);
            if (lane == 0) {
                for (int64_t j = np; j < n_kv; j++) sum = __fadd_rn(sum, __fmul_rn(vrow[j], pr[j])); // leftovers: mul, then add
                out[((int64_t)i * n_heads + h) * hs + d] = sum;
            }
        }
}

// Scores for batches (prefill chunks): one CTA per (32 cache positions, kv head); each warp keeps 8 K rows in registers,
// the queries of the group come through shared memory in batches, and every query row read from shared memory is
// dotted with the warp's 8 K rows at once — eight independent FMA chains and eight interleaved butterfly reductions
// (ggml_vec_dot_f32 lane order and GGML_F32x8_REDUCE order, exactly as in ps_k_attn_scores).
__global__ void __launch_bounds__(128) ps_k_attn_scores_batch(float *__restrict__ kq, const float *__restrict__ kc, const float *__restrict__ q,
                                                              int hs, int n_heads, int n_kv_heads, int64_t n_kv, int bs, int qb) {
    extern __shared__ float s_qb[]; // [qb * r2][hs]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = blockIdx.y, r2 = n_heads / n_kv_heads, steps = hs / 32;
    const int64_t j0 = (int64_t)blockIdx.x * 32 + warp * 8;
    float kv[8][8];
#pragma unroll
    for (int t = 0; t < 8; t++)
#pragma unroll
        for (int s = 0; s < 8; s++) kv[t][s] = (s < steps && j0 + t < n_kv) ? kc[(j0 + t) * (int64_t)(hs * n_kv_heads) + g * hs + 32 * s + lane] : 0.f;
    {   // one query batch per CTA (blockIdx.z): the r2 heads of a group are adjacent in q, so a query is one contiguous chunk
        const int i0 = blockIdx.z * qb;
        const int nq = min(qb, bs - i0);
        const int chunk4 = r2 * hs / 4; // float4s per query
        for (int idx = tid; idx < nq * chunk4; idx += 128) {
            const int qi = idx / chunk4, e = idx % chunk4;
            reinterpret_cast<float4 *>(s_qb)[idx] = reinterpret_cast<const float4 *>(q + ((int64_t)(i0 + qi) * n_heads + g * r2) * hs)[e];
        }
        __syncthreads();
#pragma unroll 2
        for (int p = 0; p < nq * r2; p++) {
            float qe[8];
#pragma unroll
            for (int s = 0; s < 8; s++) qe[s] = (s < steps) ? s_qb[p * hs + 32 * s + lane] : 0.f;
            float sum[8];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                sum[t] = 0.f;
#pragma unroll
                for (int s = 0; s < 8; s++)
                    if (s < steps) sum[t] = __fmaf_rn(kv[t][s], qe[s], sum[t]);
            }
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const int step = (k == 0) ? 16 : (k == 1) ? 8 : (k == 2) ? 4 : (k == 3) ? 1 : 2;
                float o[8];
#pragma unroll
                for (int t = 0; t < 8; t++) o[t] = __shfl_xor_sync(PS_FULL, sum[t], step);
#pragma unroll
                for (int t = 0; t < 8; t++) sum[t] = __fadd_rn(sum[t], o[t]);
            }
            if (lane < 8 && j0 + lane < n_kv) { // lane t stores position j0 + t: 8 consecutive floats of the score row
                float v = sum[0];
#pragma unroll
                for (int t = 1; t < 8; t++)
                    if (lane == t) v = sum[t];
                const int i = i0 + p / r2, h = g * r2 + p % r2;
                kq[((int64_t)h * bs + i) * n_kv + j0 + lane] = v;
            }
        }
    }
}

// Same operator for batches (prefill chunks): one CTA per (block of 8 queries, head).  The 8 probability rows sit in
// shared memory for the whole kernel; each warp then streams V^T rows (one output dim d at a time, 16 loads in flight)
// against all 8 queries, so P is read once and V^T is re-read from L2 only once per 8 queries.  Per-lane FMA chains,
// reduction and leftovers are those of ggml_vec_dot_f32 (ggml.c:2092-2131), exactly as in ps_k_attn_pv.
#define PS_PV_QB 8
__global__ void __launch_bounds__(256) ps_k_attn_pv_batch(float *__restrict__ out, const float *__restrict__ vct, const float *__restrict__ p,
                                                          int hs, int n_heads, int n_kv_heads, int64_t n_kv, int64_t n_ctx, int bs) {
    extern __shared__ float s_pq[]; // [PS_PV_QB][n_kv]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int i0 = blockIdx.x * PS_PV_QB, h = blockIdx.y, g = h / (n_heads / n_kv_heads);
    const int nq = min(PS_PV_QB, bs - i0);
    for (int qi = 0; qi < nq; qi++) {
        const float *pr = p + ((int64_t)h * bs + i0 + qi) * n_kv;
        for (int64_t j = tid; j < n_kv; j += 256) s_pq[qi * n_kv + j] = pr[j];
    }
    __syncthreads();
    const int64_t np = n_kv & ~(int64_t)31;
    const int ntail = (int)(n_kv - np);
    for (int d = warp; d < hs; d += 8) {
        const float *vrow = vct + ((int64_t)g * hs + d) * n_ctx;
        float sum[PS_PV_QB];
#pragma unroll
        for (int qi = 0; qi < PS_PV_QB; qi++) sum[qi] = 0.f;
        const float vtail = (lane < ntail) ? vrow[np + lane] : 0.f;
        for (int64_t s0 = 0; s0 < np; s0 += 512) {
            float vv[16];
#pragma unroll
            for (int u = 0; u < 16; u++) vv[u] = (s0 + 32 * u < np) ? vrow[s0 + 32 * u + lane] : 0.f;
#pragma unroll
            for (int u = 0; u < 16; u++)
                if (s0 + 32 * u < np) {
#pragma unroll
                    for (int qi = 0; qi < PS_PV_QB; qi++)
                        if (qi < nq) sum[qi] = __fmaf_rn(vv[u], s_pq[qi * n_kv + s0 + 32 * u + lane], sum[qi]);
                }
        }
#pragma unroll
        for (int k = 0; k < 5; k++) { // GGML_F32x8_REDUCE butterfly (see ps_f32x8_reduce), interleaved over the queries
            const int step = (k == 0) ? 16 : (k == 1) ? 8 : (k == 2) ? 4 : (k == 3) ? 1 : 2;
            float o[PS_PV_QB];
#pragma unroll
            for (int qi = 0; qi < PS_PV_QB; qi++) o[qi] = __shfl_xor_sync(PS_FULL, sum[qi], step);
#pragma unroll
            for (int qi = 0; qi < PS_PV_QB; qi++) sum[qi] = __fadd_rn(sum[qi], o[qi]);
        }
        for (int t = 0; t < ntail; t++) { // leftovers: mul, then add, in order
            const float v = __shfl_sync(PS_FULL, vtail, t);
#pragma unroll
            for (int qi = 0; qi < PS_PV_QB; qi++)
                if (qi < nq) sum[qi] = __fadd_rn(sum[qi], __fmul_rn(v, s_pq[qi * n_kv + np + t]));
        }
        if (lane < nq) {
            float v = sum[0];
#pragma unroll
            for (int qi = 1; qi < PS_PV_QB; qi++)
                if (lane == qi) v = sum[qi];
            out[((int64_t)(i0 + lane) * n_heads + h) * hs + d] = v;
        }
    }
}

// Register-tiled variants of the two batched attention kernels (prefill chunks, verify batches).  The arithmetic of
// every dot product is that of ps_k_attn_scores_batch / ps_k_attn_pv_batch - ggml_vec_dot_f32 lane chains
// (ggml.c:2092-2131), GGML_F32x8_REDUCE order 16, 8, 4, 1, 2 - only the data movement differs:
//  * the 32-lane reductions run as a reduce-scatter: at every butterfly stage a lane keeps half of its sums and hands
//    the other half to its partner, so 8 sums cost 4 + 2 + 1 + 1 + 1 shuffles instead of 40 (a + b == b + a, so the value
//    a lane ends up with is the one the full butterfly leaves in it);
//  * the scores kernel lets lane code c = lane >> 2 hold cache row (t ^ c) in register slot t, which makes the kept
//    half "slots 0..n/2" on every lane - no selects;
//  * the P.V kernel gives a warp an 8 (output dims) x 8 (queries) accumulator tile, so every probability read from
//    shared memory and every V^T element read from L2 feeds 8 FMAs instead of 1 (the old kernel issued one
//    shared-memory load per FMA and was bound by that pipe at ~16 % of the FMA rate).
template <int STEPS> // head size / 32: the loops below carry no run-time predicates
__global__ void __launch_bounds__(128) ps_k_attn_scores_tile(float *__restrict__ kq, const float *__restrict__ kc, const float *__restrict__ q,
                                                             int n_heads, int n_kv_heads, int64_t n_kv, int bs, int qb) {
    extern __shared__ float s_qb[]; // [qb * r2][hs]
    constexpr int hs = 32 * STEPS;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, c = lane >> 2;
    const int g = blockIdx.y, r2 = n_heads / n_kv_heads;
    const int64_t j0 = (int64_t)blockIdx.x * 32 + warp * 8;
    float kv[8][STEPS];
#pragma unroll
    for (int t = 0; t < 8; t++) {
        const bool in = j0 + (t ^ c) < n_kv;
        const float *kr = kc + (j0 + (t ^ c)) * (int64_t)(hs * n_kv_heads) + g * hs + lane;
#pragma unroll
        for (int s = 0; s < STEPS; s++) kv[t][s] = in ? kr[32 * s] : 0.f;
    }
    const int i0 = blockIdx.z * qb;
    const int nq = min(qb, bs - i0);
    const int chunk4 = r2 * hs / 4; // float4s per query: the r2 heads of a group are adjacent in q
    for (int idx = tid; idx < nq * chunk4; idx += 128) {
        const int qi = idx / chunk4, e = idx % chunk4;
        reinterpret_cast<float4 *>(s_qb)[idx] = reinterpret_cast<const float4 *>(q + ((int64_t)(i0 + qi) * n_heads + g * r2) * hs)[e];
    }
    __syncthreads();
    const bool writer = (lane & 3) == 0 && j0 + c < n_kv; // slot 0 of lane code c ends up as cache row j0 + c
    const float *sq = s_qb + lane;
    for (int hh = 0; hh < r2; hh++) {
        float *dst = kq + ((int64_t)(g * r2 + hh) * bs + i0) * n_kv + j0 + c;
#pragma unroll 2
        for (int qi = 0; qi < nq; qi++) {
            const float *qp = sq + (qi * r2 + hh) * hs;
            float qe[STEPS];
#pragma unroll
            for (int s = 0; s < STEPS; s++) qe[s] = qp[32 * s];
            float sum[8];
#pragma unroll
            for (int t = 0; t < 8; t++) {
                sum[t] = 0.f;
#pragma unroll
                for (int s = 0; s < STEPS; s++) sum[t] = __fmaf_rn(kv[t][s], qe[s], sum[t]);
            }
#pragma unroll
            for (int t = 0; t < 4; t++) sum[t] = __fadd_rn(sum[t], __shfl_xor_sync(PS_FULL, sum[t + 4], 16));
#pragma unroll
            for (int t = 0; t < 2; t++) sum[t] = __fadd_rn(sum[t], __shfl_xor_sync(PS_FULL, sum[t + 2], 8));
            sum[0] = __fadd_rn(sum[0], __shfl_xor_sync(PS_FULL, sum[1], 4));
            sum[0] = __fadd_rn(sum[0], __shfl_xor_sync(PS_FULL, sum[0], 1));
            sum[0] = __fadd_rn(sum[0], __shfl_xor_sync(PS_FULL, sum[0], 2));
            if (writer) dst[(int64_t)qi * n_kv] = sum[0];
        }
    }
}

// one reduce-scatter stage over N sums: the lane whose `up` bit is set keeps the upper half, the other the lower half
template <int N>
PS_D void ps_rs_stage(float (&a)[N], bool up, int mask) {
#pragma unroll
    for (int i = 0; i < N / 2; i++) {
        const float keep = up ? a[i + N / 2] : a[i], send = up ? a[i] : a[i + N / 2];
        a[i] = __fadd_rn(keep, __shfl_xor_sync(PS_FULL, send, mask));
    }
}

#define PS_PVT_Q 8 // queries per CTA
#define PS_PVT_D 8 // output dims per warp tile
#define PS_PVT_DEPTH 4 // V^T chunks in flight per warp (cp.async ring in shared memory: no registers held by loads in flight)
#define PS_PVT_RING (8 * PS_PVT_DEPTH * PS_PVT_D * 32 * 4) // bytes: 8 warps x depth x 8 rows x 32 positions
PS_D void ps_cp_async4(float *dst_smem, const float *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(dst_smem)), "l"(src) : "memory");
}
PS_D void ps_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PS_D void ps_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__global__ void __launch_bounds__(256, 2) ps_k_attn_pv_tile(float *__restrict__ out, const float *__restrict__ vct, const float *__restrict__ p,
                                                            int hs, int n_heads, int n_kv_heads, int64_t n_kv, int64_t n_ctx, int bs) {
    // [ring: 8 warps][PS_PVT_DEPTH][PS_PVT_D][32] V^T chunks, then the probabilities, position-major: s_p4[j] = queries 0..3
    // at position j, s_p4[n_kv + j] = queries 4..7, so a lane fetches its eight probabilities of a chunk with two
    // conflict-free 16-byte loads off one address register
    extern __shared__ __align__(16) float4 s_pvt[];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    float *ring = reinterpret_cast<float *>(s_pvt) + warp * (PS_PVT_DEPTH * PS_PVT_D * 32) + lane; // every lane reads back only what it copied itself
    float4 *s_p4 = s_pvt + PS_PVT_RING / 16;
    const int i0 = blockIdx.x * PS_PVT_Q, h = blockIdx.y, g = h / (n_heads / n_kv_heads);
    const int nq = min(PS_PVT_Q, bs - i0);
    const int nkv = (int)n_kv;
    const int nc = nkv >> 5; // full 32-position chunks; the rest are the leftovers of ggml_vec_dot_f32
    const int np = nc << 5;
    // hs is a multiple of PS_PVT_D (launch condition); gridDim.z CTAs share the dim groups of a (query block, head) when the batch is narrow
    const int d_first = ((int)blockIdx.z * 8 + warp) * PS_PVT_D, d_step = (int)gridDim.z * 8 * PS_PVT_D;
    const float *vp[PS_PVT_D]; // one running pointer per V^T row: the copies below only add immediates
    auto issue = [&](bool in, int slot, int off) { // a chunk of the current dim group -> ring slot (always one group per call, possibly empty)
        if (in) {
#pragma unroll
            for (int di = 0; di < PS_PVT_D; di++) ps_cp_async4(ring + (slot * PS_PVT_D + di) * 32, vp[di] + off);
        }
        ps_cp_async_commit();
    };
    if (d_first < hs) { // the first dim group's V^T chunks are on their way while the probabilities are staged
#pragma unroll
        for (int di = 0; di < PS_PVT_D; di++) vp[di] = vct + ((int64_t)g * hs + d_first + di) * n_ctx + lane;
#pragma unroll
        for (int k = 0; k < PS_PVT_DEPTH - 1; k++) issue(k < nc, k, 32 * k);
    }
    {
        const float *pr = p + ((int64_t)h * bs + i0) * n_kv;
        for (int j = tid; j < nkv; j += 256) {
            float e[PS_PVT_Q];
#pragma unroll
            for (int qi = 0; qi < PS_PVT_Q; qi++) e[qi] = (qi < nq) ? pr[(int64_t)qi * n_kv + j] : 0.f;
            s_p4[j] = make_float4(e[0], e[1], e[2], e[3]);
            s_p4[nkv + j] = make_float4(e[4], e[5], e[6], e[7]);
        }
    }
    __syncthreads();
    for (int d0 = d_first; d0 < hs; d0 += d_step) {
        const float4 *sp = s_p4 + lane;
        float acc[PS_PVT_D * PS_PVT_Q]; // [di][qi]
#pragma unroll
        for (int k = 0; k < PS_PVT_D * PS_PVT_Q; k++) acc[k] = 0.f;
        for (int c0 = 0; c0 < nc; c0 += PS_PVT_DEPTH) {
#pragma unroll
            for (int k = 0; k < PS_PVT_DEPTH; k++) {
                const int c = c0 + k;
                if (c < nc) { // warp-uniform
                    issue(c + PS_PVT_DEPTH - 1 < nc, (k + PS_PVT_DEPTH - 1) % PS_PVT_DEPTH, 32 * (k + PS_PVT_DEPTH - 1)); // refills the slot consumed one chunk ago
                    ps_cp_async_wait<PS_PVT_DEPTH - 1>();                                // chunk c has landed
                    float v[PS_PVT_D];
#pragma unroll
                    for (int di = 0; di < PS_PVT_D; di++) v[di] = ring[(k * PS_PVT_D + di) * 32];
                    const float4 lo = sp[32 * k], hi = sp[nkv + 32 * k];
                    const float pp[PS_PVT_Q] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
                    for (int di = 0; di < PS_PVT_D; di++)
#pragma unroll
                        for (int qi = 0; qi < PS_PVT_Q; qi++) acc[di * PS_PVT_Q + qi] = __fmaf_rn(v[di], pp[qi], acc[di * PS_PVT_Q + qi]);
                }
            }
#pragma unroll
            for (int di = 0; di < PS_PVT_D; di++) vp[di] += 32 * PS_PVT_DEPTH;
            sp += 32 * PS_PVT_DEPTH;
        }
        ps_cp_async_wait<0>();
        if (d0 + d_step < hs) { // next dim group of this warp: its first chunks fly during the reduction below
            const int64_t adv = (int64_t)d_step * n_ctx - (int64_t)((nc + PS_PVT_DEPTH - 1) / PS_PVT_DEPTH) * (32 * PS_PVT_DEPTH);
#pragma unroll
            for (int di = 0; di < PS_PVT_D; di++) vp[di] += adv;
#pragma unroll
            for (int k = 0; k < PS_PVT_DEPTH - 1; k++) issue(k < nc, k, 32 * k);
        }
        // GGML_F32x8_REDUCE as a reduce-scatter: 64 sums -> 2 per lane.  Flat index = qi + 8 * di; the stages peel index
        // bits 5, 4, 3 (lane bits 4, 3, 2), then bit 2 (lane bit 0, stage xor 1) and bit 1 (lane bit 1, stage xor 2).
        ps_rs_stage<64>(acc, lane & 16, 16);
        float a32[32];
#pragma unroll
        for (int k = 0; k < 32; k++) a32[k] = acc[k];
        ps_rs_stage<32>(a32, lane & 8, 8);
        float a16[16];
#pragma unroll
        for (int k = 0; k < 16; k++) a16[k] = a32[k];
        ps_rs_stage<16>(a16, lane & 4, 4);
        float a8[8];
#pragma unroll
        for (int k = 0; k < 8; k++) a8[k] = a16[k];
        ps_rs_stage<8>(a8, lane & 1, 1);
        float a4[4];
#pragma unroll
        for (int k = 0; k < 4; k++) a4[k] = a8[k];
        ps_rs_stage<4>(a4, lane & 2, 2);
        const int di = lane >> 2, qb = 4 * (lane & 1) + (lane & 2);
        const float *vrow = vct + ((int64_t)g * hs + d0 + di) * n_ctx;
        const float *sf = reinterpret_cast<const float *>(s_p4);
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int qi = qb + e;
            const float *pq = sf + (size_t)(qi >> 2) * nkv * 4 + (qi & 3);
            float sum = a4[e];
            for (int t = np; t < nkv; t++) sum = __fadd_rn(sum, __fmul_rn(vrow[t], pq[4 * t])); // leftovers: mul, then add, in order
            if (qi < nq) out[((int64_t)(i0 + qi) * n_heads + h) * hs + d0 + di] = sum;
        }
    }
}

// GGMLBackend::silu_hadamard (src/backend/ggml/ggml.cpp:115-129)
__global__ void ps_k_silu_hadamard(float *__restrict__ dst, const float *__restrict__ g, const float *__restrict__ u, int64_t n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = ps_silu_mul(g[t], u[t]);
}

// powerserve_compute_forward_add_f32 (libs/ggml/src/ggml.c:10042-10112), src1 row-broadcast
__global__ void ps_k_add(float *__restrict__ dst, const float *__restrict__ a, const float *__restrict__ b, int64_t n, int64_t nb) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x)
        dst[t] = __fadd_rn(a[t], b[t % nb]);
}

// powerserve_compute_forward_dup (ggml.c:9519-9558) for 2-D fp32 views with byte strides
__global__ void ps_k_copy_2d(uint8_t *__restrict__ dst, int64_t ds0, int64_t ds1, const uint8_t *__restrict__ src, int64_t ss0, int64_t ss1,
                             int64_t ne0, int64_t ne1) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < ne0 * ne1; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i0 = t % ne0, i1 = t / ne0;
        *reinterpret_cast<float *>(dst + i0 * ds0 + i1 * ds1) = *reinterpret_cast<const float *>(src + i0 * ss0 + i1 * ss1);
    }
}

// powerserve_compute_forward_dup (ggml.c:9519-9558) for fp32 views of up to four dims with byte strides and possibly
// different shapes (GGMLBackend::copy / cont): element t of the source in ITS row-major order goes to element t of the
// destination in the destination's order - what dup does when the shapes differ but the element counts agree.
struct PsNd { int64_t ne[4], nb[4]; };
__global__ void ps_k_copy_4d(uint8_t *__restrict__ dst, const PsNd d, const uint8_t *__restrict__ src, const PsNd s, int64_t n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = t, so = 0, doff = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) { so += (r % s.ne[k]) * s.nb[k]; r /= s.ne[k]; }
        r = t;
#pragma unroll
        for (int k = 0; k < 4; k++) { doff += (r % d.ne[k]) * d.nb[k]; r /= d.ne[k]; }
        *reinterpret_cast<float *>(dst + doff) = *reinterpret_cast<const float *>(src + so);
    }
}

// GGMLBackend::matmul with an FP32 src0 (the two attention products over strided cache views, norm_attention.cpp:117-147 ->
// powerserve_compute_forward_mul_mat with vec_dot_f32, ggml.c:13344-13432, 2092-2131): dst{ne01, ne11, ne12} (contiguous) =
// dot over ne00 of src0 row (i01, i12 / r2) and src1 column (i11, i12); innermost dims contiguous, byte strides otherwise.
// One warp per output element: lane t = 8 j + l is lane l of accumulator j (GGML_F32_STEP 32, GGML_F32_EPR 8), then
// GGML_F32x8_REDUCE and the leftovers in order (mul, then add).  The op-by-op executor path: a bring-up / debug path.
__global__ void __launch_bounds__(128) ps_k_matmul_f32(float *__restrict__ dst, const uint8_t *__restrict__ a, int64_t ne00, int64_t ne01, int64_t ne02,
                                                       int64_t nb01, int64_t nb02, const uint8_t *__restrict__ b, int64_t ne11, int64_t ne12,
                                                       int64_t nb11, int64_t nb12) {
    const int lane = threadIdx.x & 31;
    const int64_t n_out = ne01 * ne11 * ne12, r2 = ne12 / ne02, np = ne00 & ~(int64_t)31;
    for (int64_t o = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5); o < n_out; o += (int64_t)gridDim.x * 4) {
        const int64_t i01 = o % ne01, i11 = (o / ne01) % ne11, i12 = o / (ne01 * ne11);
        const float *x = reinterpret_cast<const float *>(a + i01 * nb01 + (i12 / r2) * nb02);
        const float *y = reinterpret_cast<const float *>(b + i11 * nb11 + i12 * nb12);
        float sum = 0.f;
        for (int64_t s = 0; s < np; s += 32) sum = __fmaf_rn(x[s + lane], y[s + lane], sum);
        sum = ps_f32x8_reduce(sum);
        if (lane == 0) {
            for (int64_t j = np; j < ne00; j++) sum = __fadd_rn(sum, __fmul_rn(x[j], y[j]));
            dst[o] = sum;
        }
    }
}

// ====================================================================================================================
// Server-side batching: one token of each of n independent SESSIONS per forward pass.  The weight
// products run once for the n columns (one weight stream); attention is per column over that session's OWN cache at its OWN
// position.  Column `c` = blockIdx.z; kc_ptrs / vct_ptrs hold the session's cache of the current layer, pos[c] its
// position.  The arithmetic of every column is that of the bs = 1 operator-table kernels above (ps_k_kv_store,
// ps_k_attn_scores, ps_k_softmax_ext with the position mask, ps_k_attn_pv), so a session's result does not depend on which
// other sessions share the batch.
// ====================================================================================================================
__global__ void ps_k_sess_kv_store(float *const *__restrict__ kc_ptrs, float *const *__restrict__ vct_ptrs, const float *__restrict__ k,
                                   const float *__restrict__ v, int64_t kv_dim, int64_t n_ctx, const int32_t *__restrict__ pos, int n) {
    for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < kv_dim * n; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t c = t / kv_dim, e = t % kv_dim, p = pos[c];
        kc_ptrs[c][p * kv_dim + e] = k[t];
        vct_ptrs[c][e * n_ctx + p] = v[t];
    }
}

// scores of column c: kq_c {n_kv_c, 1, n_heads} at kq + c * col_stride; one warp per (cache position, kv head)
__global__ void __launch_bounds__(128) ps_k_sess_scores(float *__restrict__ kq, const float *const *__restrict__ kc_ptrs, const float *__restrict__ q, int hs,
                                                        int n_heads, int n_kv_heads, const int32_t *__restrict__ pos, int64_t col_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.z, g = blockIdx.y;
    const int64_t n_kv = (int64_t)pos[c] + 1, j = (int64_t)blockIdx.x * 4 + warp;
    if (j >= n_kv) return;
    const int r2 = n_heads / n_kv_heads, steps = hs / 32;
    const float *krow = kc_ptrs[c] + j * (int64_t)(hs * n_kv_heads) + g * hs;
    float kv[8];
#pragma unroll
    for (int s = 0; s < 8; s++) kv[s] = (s < steps) ? krow[32 * s + lane] : 0.f;
    for (int hh = 0; hh < r2; hh++) {
        const int h = g * r2 + hh;
        const float *qv = q + ((int64_t)c * n_heads + h) * hs;
        float sum = 0.f;
#pragma unroll
        for (int s = 0; s < 8; s++)
            if (s < steps) sum = __fmaf_rn(kv[s], qv[32 * s + lane], sum); // ggml_vec_dot_f32 lane chain
        sum = ps_f32x8_reduce(sum);
        if (lane == 0) kq[c * col_stride + (int64_t)h * n_kv + j] = sum;
    }
}

// softmax_ext of column c with the position mask of GET_MASK (every j <= pos: + 0.0f), rows of n_kv_c; grid (n_heads, n)
__global__ void __launch_bounds__(256) ps_k_sess_softmax(float *__restrict__ kq, const int32_t *__restrict__ pos, int64_t col_stride, float scale) {
    extern __shared__ float wp[];
    __shared__ double sh[32];
    __shared__ float shf[32];
    const int c = blockIdx.y;
    const int64_t ne0 = (int64_t)pos[c] + 1;
    float *dp = kq + c * col_stride + (int64_t)blockIdx.x * ne0;
    float mx = -INFINITY;
    for (int64_t j = threadIdx.x; j < ne0; j += blockDim.x) {
        const float v = __fadd_rn(__fmul_rn(dp[j], scale), 0.f);
        wp[j] = v;
        mx = fmaxf(mx, v);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(PS_FULL, mx, o));
    if ((threadIdx.x & 31) == 0) shf[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = shf[0];
    for (int t = 1; t < (int)((blockDim.x + 31) >> 5); t++) mx = fmaxf(mx, shf[t]);
    const int64_t n8 = ne0 & ~(int64_t)7;
    double s = 0.0;
    for (int64_t gi = threadIdx.x; gi < n8 / 8; gi += blockDim.x) {
        float v[8];
#pragma unroll
        for (int l = 0; l < 8; l++) {
            v[l] = ps_v_expf(__fadd_rn(wp[gi * 8 + l], -mx));
            dp[gi * 8 + l] = v[l];
        }
        const float r0 = __fadd_rn(v[4], v[0]), r1 = __fadd_rn(v[5], v[1]), r2 = __fadd_rn(v[6], v[2]), r3 = __fadd_rn(v[7], v[3]);
        s += (double)__fadd_rn(__fadd_rn(r0, r2), __fadd_rn(r1, r3));
    }
    for (int64_t j = n8 + threadIdx.x; j < ne0; j += blockDim.x) {
        const float v = ps_expf_glibc(__fadd_rn(wp[j], -mx));
        dp[j] = v;
        s += (double)v;
    }
    const double sum = ps_block_sum_double(s, sh);
    const float inv = (float)(1.0 / sum);
    for (int64_t j = threadIdx.x; j < ne0; j += blockDim.x) dp[j] = __fmul_rn(dp[j], inv);
}

// P.V of column c: out[c][h * hs + d]; one warp per (kv head g, d)
__global__ void __launch_bounds__(128) ps_k_sess_pv(float *__restrict__ out, const float *const *__restrict__ vct_ptrs, const float *__restrict__ kq, int hs,
                                                    int n_heads, int n_kv_heads, const int32_t *__restrict__ pos, int64_t n_ctx, int64_t col_stride) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, c = blockIdx.z;
    const int d = blockIdx.x * 4 + warp, g = blockIdx.y;
    if (d >= hs) return;
    const int r2 = n_heads / n_kv_heads;
    const int64_t n_kv = (int64_t)pos[c] + 1, np = n_kv & ~(int64_t)31;
    const float *vrow = vct_ptrs[c] + ((int64_t)g * hs + d) * n_ctx;
    for (int hh = 0; hh < r2; hh++) {
        const int h = g * r2 + hh;
        const float *pr = kq + c * col_stride + (int64_t)h * n_kv;
        float sum = 0.f;
        for (int64_t s = 0; s < np; s += 32) sum = __fmaf_rn(vrow[s + lane], pr[s + lane], sum);
        sum = ps_f32x8_reduce(sum);
        if (lane == 0) {
            for (int64_t j = np; j < n_kv; j++) sum = __fadd_rn(sum, __fmul_rn(vrow[j], pr[j])); // leftovers: mul, then add
            out[((int64_t)c * n_heads + h) * hs + d] = sum;
        }
    }
}

// ====================================================================================================================
// Device-side top-k: TopKSampler (src/sampler/sampler.cpp:39-56: partial_sort by logit, descending) on the
// device, so that k (logit, token) pairs cross PCIe instead of the vocabulary's logits and the host never builds a
// vocabulary-sized ProbArray (prob_array.hpp:43-49).  Two stages of repeated arg-max: every CTA extracts the k largest of
// its slice (kept in shared memory), then one CTA merges the candidates.  Order: logit descending, equal logits by
// ascending token id (std::partial_sort leaves that order unspecified).
// ====================================================================================================================
#define PS_TOPK_MAX 64
#define PS_TOPK_SLICE 2048
// the block's (largest value, lowest index) among the entries of v[0..n) that are not NaN (NaN marks "already taken");
// returns the position in every thread, -1 if nothing is left
PS_D bool ps_topk_better(float x, int id, float best, int bi) { return x == x && (!(best == best) || x > best || (x == best && id < bi)); }
PS_D int ps_block_argmax_smem(const float *v, const int *idx, int n, float *sv, int *si, int *sp) {
    float best = __int_as_float(0x7fc00000);
    int bi = 0x7fffffff, bp = -1;
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
        const float x = v[t];
        const int id = idx ? idx[t] : t;
        if (ps_topk_better(x, id, best, bi)) { best = x; bi = id; bp = t; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const float ov = __shfl_xor_sync(PS_FULL, best, o);
        const int oi = __shfl_xor_sync(PS_FULL, bi, o), op = __shfl_xor_sync(PS_FULL, bp, o);
        if (ps_topk_better(ov, oi, best, bi)) { best = ov; bi = oi; bp = op; }
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { sv[warp] = best; si[warp] = bi; sp[warp] = bp; }
    __syncthreads();
    best = sv[0]; bi = si[0]; bp = sp[0];
    for (int w = 1; w < (int)(blockDim.x >> 5); w++)
        if (ps_topk_better(sv[w], si[w], best, bi)) { best = sv[w]; bi = si[w]; bp = sp[w]; }
    __syncthreads();
    return bp;
}
// stage 1: CTA b -> cand_val / cand_idx [b][k] = the k largest of logits[b * SLICE, (b + 1) * SLICE) (padded with -inf, id 0x7fffffff)
__global__ void __launch_bounds__(256) ps_k_topk_stage1(const float *__restrict__ logits, int n, int k, float *__restrict__ cand_val, int *__restrict__ cand_idx) {
    __shared__ float s_v[PS_TOPK_SLICE];
    __shared__ float sv[8];
    __shared__ int si[8], sp[8];
    const int lo = blockIdx.x * PS_TOPK_SLICE, m = min(PS_TOPK_SLICE, n - lo);
    for (int t = threadIdx.x; t < m; t += blockDim.x) s_v[t] = logits[lo + t];
    __syncthreads();
    for (int r = 0; r < k; r++) {
        const int p = ps_block_argmax_smem(s_v, nullptr, m, sv, si, sp); // ids inside a slice ascend with the position
        if (threadIdx.x == 0) {
            const bool live = p >= 0;
            cand_val[blockIdx.x * k + r] = live ? s_v[p] : __int_as_float(0x7fc00000); // NaN = no candidate (slice shorter than k)
            cand_idx[blockIdx.x * k + r] = live ? lo + p : 0x7fffffff;
            if (live) s_v[p] = __int_as_float(0x7fc00000); // a quiet NaN marks "taken"
        }
        __syncthreads();
    }
}
// stage 2: one CTA merges n_cand candidates into out_val / out_idx [k]
__global__ void __launch_bounds__(256) ps_k_topk_stage2(float *__restrict__ cand_val, const int *__restrict__ cand_idx, int n_cand, int k, float *__restrict__ out_val,
                                                        int *__restrict__ out_idx) {
    __shared__ float sv[8];
    __shared__ int si[8], sp[8];
    for (int r = 0; r < k; r++) {
        const int p = ps_block_argmax_smem(cand_val, cand_idx, n_cand, sv, si, sp);
        if (threadIdx.x == 0) {
            out_val[r] = p >= 0 ? cand_val[p] : -INFINITY;
            out_idx[r] = p >= 0 ? cand_idx[p] : -1;
            if (p >= 0) cand_val[p] = __int_as_float(0x7fc00000);
        }
        __syncthreads();
    }
}

// greedy pick (Model::decode with top_k = 1: ProbArray + greedy_sample, src/model/llama/llama_model.cpp:124-128):
// first maximum wins.  One CTA; writes the id to `out[step]` and to `next_token` (device feedback for the next step).
__global__ void __launch_bounds__(1024) ps_k_argmax(const float *__restrict__ logits, int64_t n, int32_t *__restrict__ out, int32_t *__restrict__ next_token) {
    __shared__ float sv[32];
    __shared__ int si[32];
    // one CTA per logits row (gridDim.x rows: the columns of a session batch); a single row for the decode loop
    logits += (int64_t)blockIdx.x * n;
    out += blockIdx.x;
    next_token += blockIdx.x;
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
        *out = bi;
        *next_token = bi;
    }
}
