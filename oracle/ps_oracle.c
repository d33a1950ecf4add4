/* oracle/ps_oracle.c — TEST INFRASTRUCTURE ONLY (see ps_oracle.h).
 *
 * Scalar C restatement of the reference's x86 AVX2+FMA code paths.  Where the reference uses 8-lane AVX vectors the
 * lanes are spelled out as arrays of 8 floats/ints and every fused multiply-add of the reference is an explicit
 * fmaf() here; the file is compiled with -ffp-contract=off so nothing else is fused.  All `file:line` citations are
 * relative to /root/reference/libs/ggml/src unless they start with src/.
 */
#include "ps_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

/* tiny parallel-for over independent output rows (order-safe: every output element is produced by one call) */
typedef void (*par_fn)(int64_t begin, int64_t end, void *ctx);
typedef struct { par_fn fn; void *ctx; int64_t begin, end; } par_job;
static void *par_tramp(void *p) { par_job *j = (par_job *)p; j->fn(j->begin, j->end, j->ctx); return NULL; }
static int par_threads(void) {
    const char *e = getenv("PS_ORACLE_THREADS");
    long n = e ? atol(e) : sysconf(_SC_NPROCESSORS_ONLN);
    return n < 1 ? 1 : (n > 64 ? 64 : (int)n);
}
static void par_for(int64_t n, int64_t min_chunk, par_fn fn, void *ctx) {
    int nt = par_threads();
    if (n / (min_chunk > 0 ? min_chunk : 1) < nt) nt = (int)(n / (min_chunk > 0 ? min_chunk : 1));
    if (nt <= 1) { fn(0, n, ctx); return; }
    pthread_t th[64];
    par_job jobs[64];
    for (int t = 0; t < nt; t++) {
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].begin = n * t / nt; jobs[t].end = n * (t + 1) / nt;
        pthread_create(&th[t], NULL, par_tramp, &jobs[t]);
    }
    for (int t = 0; t < nt; t++) pthread_join(th[t], NULL);
}

#define QK_K 256
#define QK8_0 32

/* ------------------------------------------------------------------------------------------------ block layouts
 * ggml-common.h:158-162 (q4_0), :200-204 (q8_0), :299-310 (q4_K), :335-340 (q6_K), :344-348 (q8_K) */
#pragma pack(push, 1)
typedef struct { uint16_t d; uint8_t qs[16]; } blk_q4_0;                                  /* 18 B / 32 */
typedef struct { uint16_t d; int8_t qs[32]; } blk_q8_0;                                   /* 34 B / 32 */
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qs[128]; } blk_q4_K; /* 144 B / 256 */
typedef struct { uint16_t d; uint16_t dmin; uint8_t scales[12]; uint8_t qh[32]; uint8_t qs[128]; } blk_q5_K; /* 176 B / 256 (ggml-common.h: block_q5_K) */
typedef struct { uint8_t ql[128]; uint8_t qh[64]; int8_t scales[16]; uint16_t d; } blk_q6_K; /* 210 B / 256 */
typedef struct { float d; int8_t qs[256]; int16_t bsums[16]; } blk_q8_K;                  /* 292 B / 256 */
#pragma pack(pop)

/* ------------------------------------------------------------------------------------------------ fp16
 * GGML_FP16_TO_FP32 / GGML_FP32_TO_FP16 compile to F16C vcvtph2ps / vcvtps2ph (round-to-nearest-even) on this
 * build (ggml-impl.h, __F16C__ branch). */
float ps_or_fp16_to_fp32(uint16_t h) {
    uint32_t s = (uint32_t)(h & 0x8000u) << 16, e = (h >> 10) & 0x1fu, m = h & 0x3ffu, b;
    if (e == 0) {
        if (m == 0) {
            b = s;
        } else {
            int sh = 0;
            while (!(m & 0x400u)) { m <<= 1; sh++; }
            m &= 0x3ffu;
            b = s | ((uint32_t)(113 - sh) << 23) | (m << 13);
        }
    } else if (e == 31) {
        b = s | 0x7f800000u | (m << 13);
    } else {
        b = s | ((e + 112u) << 23) | (m << 13);
    }
    float f;
    memcpy(&f, &b, 4);
    return f;
}

uint16_t ps_or_fp32_to_fp16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint16_t sign = (uint16_t)((x >> 16) & 0x8000u);
    const uint32_t a = x & 0x7fffffffu;
    if (a > 0x7f800000u) return (uint16_t)(sign | 0x7e00u | ((a >> 13) & 0x3ffu));
    const uint32_t e = a >> 23;
    if (e >= 143) return (uint16_t)(sign | 0x7c00u);
    if (e >= 113) {
        const uint32_t m = a & 0x7fffffu;
        uint32_t h = ((e - 112u) << 10) | (m >> 13);
        const uint32_t rem = m & 0x1fffu;
        if (rem > 0x1000u || (rem == 0x1000u && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    if (e >= 102) {
        const uint32_t m = (a & 0x7fffffu) | 0x800000u;
        const int shift = 126 - (int)e;
        uint32_t h = m >> shift;
        const uint32_t rem = m & ((1u << shift) - 1u), half = 1u << (shift - 1);
        if (rem > half || (rem == half && (h & 1u))) h++;
        return (uint16_t)(sign | h);
    }
    return sign;
}

size_t ps_or_row_size(int type, int64_t k) {
    switch (type) {
    case PS_OR_F32: return (size_t)k * 4;
    case PS_OR_F16: return (size_t)k * 2;
    case PS_OR_Q4_0: return (size_t)(k / 32) * sizeof(blk_q4_0);
    case PS_OR_Q8_0: return (size_t)(k / 32) * sizeof(blk_q8_0);
    case PS_OR_Q4_K: return (size_t)(k / 256) * sizeof(blk_q4_K);
    case PS_OR_Q5_K: return (size_t)(k / 256) * sizeof(blk_q5_K);
    case PS_OR_Q6_K: return (size_t)(k / 256) * sizeof(blk_q6_K);
    case PS_OR_Q8_K: return (size_t)(k / 256) * sizeof(blk_q8_K);
    default: return 0;
    }
}

/* type_traits[].vec_dot_type, ggml.c:734-900 */
int ps_or_vec_dot_type(int wtype) {
    switch (wtype) {
    case PS_OR_Q4_0: case PS_OR_Q8_0: return PS_OR_Q8_0;
    case PS_OR_Q4_K: case PS_OR_Q5_K: case PS_OR_Q6_K: return PS_OR_Q8_K;
    default: return PS_OR_F32;
    }
}

/* hsum_float_8, ggml-quants.c:62-68: ((x0+x4)+(x2+x6)) + ((x1+x5)+(x3+x7)) */
static float hsum8(const float x[8]) {
    const float r0 = x[4] + x[0], r1 = x[5] + x[1], r2 = x[6] + x[2], r3 = x[7] + x[3];
    const float s0 = r0 + r2, s1 = r1 + r3;
    return s0 + s1;
}

/* ------------------------------------------------------------------------------------------------ quantisers */
/* nearest_int, ggml-quants.c:1653-1658 */
static int nearest_int(float fval) {
    float val = fval + 12582912.f;
    int i;
    memcpy(&i, &val, sizeof(int));
    return (i & 0x007fffff) - 0x00400000;
}

/* quantize_row_q8_K_ref, ggml-quants.c:3799-3837 (quantize_row_q8_K :3849 forwards to it on every ISA) */
void ps_or_quantize_row_q8_K(const float *x, void *vy, int64_t k) {
    blk_q8_K *y = (blk_q8_K *)vy;
    const int64_t nb = k / QK_K;
    for (int64_t i = 0; i < nb; i++) {
        float max = 0, amax = 0;
        for (int j = 0; j < QK_K; ++j) {
            float ax = fabsf(x[j]);
            if (ax > amax) { amax = ax; max = x[j]; }
        }
        if (!amax) {
            /* NOTE: the reference leaves bsums untouched (stale scratch) here; its dot product multiplies them by
             * d == 0 so they never matter.  We zero them to keep the oracle's bytes deterministic. */
            y[i].d = 0;
            memset(y[i].qs, 0, QK_K);
            memset(y[i].bsums, 0, sizeof(y[i].bsums));
            x += QK_K;
            continue;
        }
        const float iscale = -127.f / max;
        for (int j = 0; j < QK_K; ++j) {
            int v = nearest_int(iscale * x[j]);
            y[i].qs[j] = (int8_t)(v < 127 ? v : 127);
        }
        for (int j = 0; j < QK_K / 16; ++j) {
            int sum = 0;
            for (int ii = 0; ii < 16; ++ii) sum += y[i].qs[j * 16 + ii];
            y[i].bsums[j] = (int16_t)sum;
        }
        y[i].d = 1 / iscale;
        x += QK_K;
    }
}

/* quantize_row_q8_0, AVX2 branch, ggml-quants.c:957-1017: d = max|x|/127 (stored fp16), id = 127/max|x|,
 * _mm256_round_ps(NEAREST) = round-half-to-even. */
void ps_or_quantize_row_q8_0(const float *x, void *vy, int64_t k) {
    blk_q8_0 *y = (blk_q8_0 *)vy;
    const int64_t nb = k / QK8_0;
    for (int64_t i = 0; i < nb; i++) {
        float maxabs = 0.0f;
        for (int j = 0; j < 32; j++) {
            const float a = fabsf(x[j]);
            if (a > maxabs) maxabs = a;
        }
        const float d = maxabs / 127.f;
        y[i].d = ps_or_fp32_to_fp16(d);
        const float id = (maxabs != 0.0f) ? 127.f / maxabs : 0.0f;
        for (int j = 0; j < 32; j++) {
            const float v = x[j] * id;
            y[i].qs[j] = (int8_t)(int)nearbyintf(v); /* default rounding mode = nearest-even */
        }
        x += 32;
    }
}

void ps_or_quantize_row(int qtype, const float *x, void *y, int64_t k) {
    if (qtype == PS_OR_Q8_K) ps_or_quantize_row_q8_K(x, y, k);
    else if (qtype == PS_OR_Q8_0) ps_or_quantize_row_q8_0(x, y, k);
    else memcpy(y, x, (size_t)k * 4);
}

/* ------------------------------------------------------------------------------------------------ de-quantisers */
/* get_scale_min_k4, ggml-quants.c:1912-1919 */
static void get_scale_min_k4(int j, const uint8_t *q, uint8_t *d, uint8_t *m) {
    if (j < 4) {
        *d = q[j] & 63; *m = q[j + 4] & 63;
    } else {
        *d = (uint8_t)((q[j + 4] & 0xF) | ((q[j - 4] >> 6) << 4));
        *m = (uint8_t)((q[j + 4] >> 4) | ((q[j - 0] >> 6) << 4));
    }
}

void ps_or_dequantize_row(int type, const void *vx, float *y, int64_t k) {
    switch (type) {
    case PS_OR_F32: memcpy(y, vx, (size_t)k * 4); break;
    case PS_OR_Q4_0: { /* dequantize_row_q4_0, ggml-quants.c:1536-1554 */
        const blk_q4_0 *x = (const blk_q4_0 *)vx;
        for (int64_t i = 0; i < k / 32; i++) {
            const float d = ps_or_fp16_to_fp32(x[i].d);
            for (int j = 0; j < 16; ++j) {
                const int x0 = (x[i].qs[j] & 0x0F) - 8, x1 = (x[i].qs[j] >> 4) - 8;
                y[i * 32 + j] = x0 * d;
                y[i * 32 + j + 16] = x1 * d;
            }
        }
    } break;
    case PS_OR_Q8_0: { /* dequantize_row_q8_0, ggml-quants.c:1630-1643 */
        const blk_q8_0 *x = (const blk_q8_0 *)vx;
        for (int64_t i = 0; i < k / 32; i++) {
            const float d = ps_or_fp16_to_fp32(x[i].d);
            for (int j = 0; j < 32; ++j) y[i * 32 + j] = x[i].qs[j] * d;
        }
    } break;
    case PS_OR_Q4_K: { /* dequantize_row_q4_K, ggml-quants.c:2569-2591.  `d1*q - m1` is contracted to one FMA by
                          gcc -O3 -mfma in the reference build (checked against oracle/_ref). */
        const blk_q4_K *x = (const blk_q4_K *)vx;
        for (int64_t i = 0; i < k / QK_K; i++) {
            const uint8_t *q = x[i].qs;
            const float d = ps_or_fp16_to_fp32(x[i].d), min = ps_or_fp16_to_fp32(x[i].dmin);
            int is = 0;
            uint8_t sc, m;
            for (int j = 0; j < QK_K; j += 64) {
                get_scale_min_k4(is + 0, x[i].scales, &sc, &m);
                const float d1 = d * sc, m1 = min * m;
                get_scale_min_k4(is + 1, x[i].scales, &sc, &m);
                const float d2 = d * sc, m2 = min * m;
                for (int l = 0; l < 32; ++l) *y++ = fmaf(d1, (float)(q[l] & 0xF), -m1);
                for (int l = 0; l < 32; ++l) *y++ = fmaf(d2, (float)(q[l] >> 4), -m2);
                q += 32; is += 2;
            }
        }
    } break;
    case PS_OR_Q5_K: { /* dequantize_row_q5_K, ggml-quants.c:2771-2798: y = d1 * ((q & 0xF) + (high bit ? 16 : 0)) - m1, the
                          same contracted FMA as Q4_K (checked against oracle/_ref) */
        const blk_q5_K *x = (const blk_q5_K *)vx;
        for (int64_t i = 0; i < k / QK_K; i++) {
            const uint8_t *ql = x[i].qs, *qh = x[i].qh;
            const float d = ps_or_fp16_to_fp32(x[i].d), min = ps_or_fp16_to_fp32(x[i].dmin);
            int is = 0;
            uint8_t sc, m, u1 = 1, u2 = 2;
            for (int j = 0; j < QK_K; j += 64) {
                get_scale_min_k4(is + 0, x[i].scales, &sc, &m);
                const float d1 = d * sc, m1 = min * m;
                get_scale_min_k4(is + 1, x[i].scales, &sc, &m);
                const float d2 = d * sc, m2 = min * m;
                for (int l = 0; l < 32; ++l) *y++ = fmaf(d1, (float)((ql[l] & 0xF) + (qh[l] & u1 ? 16 : 0)), -m1);
                for (int l = 0; l < 32; ++l) *y++ = fmaf(d2, (float)((ql[l] >> 4) + (qh[l] & u2 ? 16 : 0)), -m2);
                ql += 32; is += 2;
                u1 <<= 2; u2 <<= 2;
            }
        }
    } break;
    case PS_OR_Q6_K: { /* dequantize_row_q6_K, ggml-quants.c:2991-3019 */
        const blk_q6_K *x = (const blk_q6_K *)vx;
        for (int64_t i = 0; i < k / QK_K; i++) {
            const float d = ps_or_fp16_to_fp32(x[i].d);
            const uint8_t *ql = x[i].ql, *qh = x[i].qh;
            const int8_t *sc = x[i].scales;
            for (int n = 0; n < QK_K; n += 128) {
                for (int l = 0; l < 32; ++l) {
                    int is = l / 16;
                    const int8_t q1 = (int8_t)((ql[l + 0] & 0xF) | (((qh[l] >> 0) & 3) << 4)) - 32;
                    const int8_t q2 = (int8_t)((ql[l + 32] & 0xF) | (((qh[l] >> 2) & 3) << 4)) - 32;
                    const int8_t q3 = (int8_t)((ql[l + 0] >> 4) | (((qh[l] >> 4) & 3) << 4)) - 32;
                    const int8_t q4 = (int8_t)((ql[l + 32] >> 4) | (((qh[l] >> 6) & 3) << 4)) - 32;
                    y[l + 0] = d * sc[is + 0] * q1;
                    y[l + 32] = d * sc[is + 2] * q2;
                    y[l + 64] = d * sc[is + 4] * q3;
                    y[l + 96] = d * sc[is + 6] * q4;
                }
                y += 128; ql += 64; qh += 32; sc += 8;
            }
        }
    } break;
    default: break;
    }
}

/* ------------------------------------------------------------------------------------------------ dot products
 * Common shape of every AVX2 kernel below: per block, eight int32 lane sums (lane l = the four consecutive
 * bytes 4l..4l+3 of each 32-byte group, i.e. exactly one dp4a), converted to float, and ONE fused multiply-add per
 * lane and block into an 8-lane fp32 accumulator that is carried sequentially over the blocks of the row; the lanes
 * are summed by hsum_float_8 only at the end. */

static int dot4_u8_s8(const uint8_t *a, const int8_t *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]; }

/* ggml_vec_dot_q4_K_q8_K, AVX2 branch, ggml-quants.c:7809-7872 */
static float vec_dot_q4_K_q8_K(int64_t n, const blk_q4_K *x, const blk_q8_K *y) {
    const int64_t nb = n / QK_K;
    float acc[8] = {0}, acc_m[4] = {0};
    for (int64_t i = 0; i < nb; ++i) {
        const float d = y[i].d * ps_or_fp16_to_fp32(x[i].d);
        const float dmin = -y[i].d * ps_or_fp16_to_fp32(x[i].dmin);
        uint8_t sc[8], mn[8];
        for (int j = 0; j < 8; j++) get_scale_min_k4(j, x[i].scales, &sc[j], &mn[j]); /* == the utmp/kmask shuffle */
        /* mins: q8s = hadd_epi16(bsums lo, bsums hi) -> sub-block sums; prod = madd_epi16(mins, q8s) -> 4 lanes */
        for (int kk = 0; kk < 4; kk++) {
            const int16_t s0 = (int16_t)(y[i].bsums[4 * kk + 0] + y[i].bsums[4 * kk + 1]);
            const int16_t s1 = (int16_t)(y[i].bsums[4 * kk + 2] + y[i].bsums[4 * kk + 3]);
            const int32_t prod = (int32_t)mn[2 * kk] * s0 + (int32_t)mn[2 * kk + 1] * s1;
            acc_m[kk] = fmaf(dmin, (float)prod, acc_m[kk]);
        }
        int32_t sumi[8] = {0};
        for (int j = 0; j < QK_K / 64; ++j) {
            const uint8_t *q4 = x[i].qs + 32 * j;
            const int8_t *q8l = y[i].qs + 64 * j, *q8h = q8l + 32;
            for (int l = 0; l < 8; l++) {
                uint8_t lo[4], hi[4];
                for (int b = 0; b < 4; b++) { lo[b] = q4[4 * l + b] & 0xF; hi[b] = q4[4 * l + b] >> 4; }
                sumi[l] += (int32_t)sc[2 * j] * dot4_u8_s8(lo, q8l + 4 * l) + (int32_t)sc[2 * j + 1] * dot4_u8_s8(hi, q8h + 4 * l);
            }
        }
        for (int l = 0; l < 8; l++) acc[l] = fmaf(d, (float)sumi[l], acc[l]);
    }
    /* acc_m = add(acc_m, movehl) ; add_ss(acc_m, movehdup) */
    const float m02 = acc_m[0] + acc_m[2], m13 = acc_m[1] + acc_m[3];
    return hsum8(acc) + (m02 + m13);
}

/* ggml_vec_dot_q5_K_q8_K, AVX2 branch, ggml-quants.c:8382-8460: the lane sums of Q4_K with the fifth bit from qh (sub-block s
 * takes bit s of qh[e]), ONE fused multiply-add per block into the 8-lane accumulator - and the mins through a SCALAR float:
 * summs += dmin * (sum_j m_j * (bsums_2j + bsums_2j+1)), a separate multiply and add in the reference build. */
static float vec_dot_q5_K_q8_K(int64_t n, const blk_q5_K *x, const blk_q8_K *y) {
    const int64_t nb = n / QK_K;
    float acc[8] = {0}, summs = 0.f;
    for (int64_t i = 0; i < nb; ++i) {
        const float d = y[i].d * ps_or_fp16_to_fp32(x[i].d);
        const float dmin = -y[i].d * ps_or_fp16_to_fp32(x[i].dmin);
        uint8_t sc[8], mn[8];
        for (int j = 0; j < 8; j++) get_scale_min_k4(j, x[i].scales, &sc[j], &mn[j]);
        int32_t hsum = 0; /* hadd_epi32(hadd_epi32(madd_epi16(mins, hadd_epi16(bsums)))) element 0 = the sum of all four products */
        for (int kk = 0; kk < 4; kk++) {
            const int16_t s0 = (int16_t)(y[i].bsums[4 * kk + 0] + y[i].bsums[4 * kk + 1]);
            const int16_t s1 = (int16_t)(y[i].bsums[4 * kk + 2] + y[i].bsums[4 * kk + 3]);
            hsum += (int32_t)mn[2 * kk] * s0 + (int32_t)mn[2 * kk + 1] * s1;
        }
        summs += dmin * (float)hsum; /* NOT contracted in the reference build: vmulss + vaddss (checked against oracle/_ref) */
        int32_t sumi[8] = {0};
        for (int j = 0; j < QK_K / 64; ++j) {
            const uint8_t *q5 = x[i].qs + 32 * j;
            const int8_t *q8l = y[i].qs + 64 * j, *q8h = q8l + 32;
            for (int l = 0; l < 8; l++) {
                uint8_t lo[4], hi[4];
                for (int b = 0; b < 4; b++) {
                    const uint8_t hb = x[i].qh[4 * l + b];
                    lo[b] = (uint8_t)((q5[4 * l + b] & 0xF) + (((hb >> (2 * j)) & 1) << 4));
                    hi[b] = (uint8_t)((q5[4 * l + b] >> 4) + (((hb >> (2 * j + 1)) & 1) << 4));
                }
                sumi[l] += (int32_t)sc[2 * j] * dot4_u8_s8(lo, q8l + 4 * l) + (int32_t)sc[2 * j + 1] * dot4_u8_s8(hi, q8h + 4 * l);
            }
        }
        for (int l = 0; l < 8; l++) acc[l] = fmaf(d, (float)sumi[l], acc[l]);
    }
    return hsum8(acc) + summs;
}

/* ggml_vec_dot_q6_K_q8_K, AVX2 branch, ggml-quants.c:9039-9116 */
static float vec_dot_q6_K_q8_K(int64_t n, const blk_q6_K *x, const blk_q8_K *y) {
    const int64_t nb = n / QK_K;
    float acc[8] = {0};
    for (int64_t i = 0; i < nb; ++i) {
        const float d = y[i].d * ps_or_fp16_to_fp32(x[i].d);
        int32_t sumi[8] = {0};
        for (int j = 0; j < 2; j++) {               /* 128-element halves */
            const uint8_t *ql = x[i].ql + 64 * j, *qh = x[i].qh + 32 * j;
            const int8_t *sc = x[i].scales + 8 * j, *q8 = y[i].qs + 128 * j;
            for (int g = 0; g < 4; g++) {           /* 32-element groups, dequantize_row_q6_K order */
                for (int l = 0; l < 8; l++) {
                    int s = 0;
                    for (int b = 0; b < 4; b++) {
                        const int e = 4 * l + b;
                        const int lo = (g & 1) ? ql[e + 32] : ql[e];
                        const int q = (((g < 2) ? (lo & 0xF) : (lo >> 4)) | (((qh[e] >> (2 * g)) & 3) << 4)) - 32;
                        s += q * q8[32 * g + e];
                    }
                    sumi[l] += (int32_t)sc[2 * g + (l >= 4)] * s;
                }
            }
        }
        for (int l = 0; l < 8; l++) acc[l] = fmaf(d, (float)sumi[l], acc[l]);
    }
    return hsum8(acc);
}

/* ggml_vec_dot_q4_0_q8_0, AVX2 branch, ggml-quants.c:4205-4228 (+ bytes_from_nibbles_32 :103-110,
 * mul_sum_i8_pairs_float :131-142): lanes 0-3 = low nibbles of qs[0..15] with y.qs[0..15], lanes 4-7 = high nibbles
 * with y.qs[16..31]. */
static float vec_dot_q4_0_q8_0(int64_t n, const blk_q4_0 *x, const blk_q8_0 *y) {
    const int64_t nb = n / 32;
    float acc[8] = {0};
    for (int64_t ib = 0; ib < nb; ++ib) {
        const float d = ps_or_fp16_to_fp32(x[ib].d) * ps_or_fp16_to_fp32(y[ib].d);
        for (int l = 0; l < 8; l++) {
            int s = 0;
            for (int b = 0; b < 4; b++) {
                const int e = 4 * (l & 3) + b;
                const int q = (l < 4) ? (x[ib].qs[e] & 0xF) - 8 : (x[ib].qs[e] >> 4) - 8;
                s += q * y[ib].qs[(l < 4 ? 0 : 16) + e];
            }
            acc[l] = fmaf(d, (float)s, acc[l]);
        }
    }
    return hsum8(acc);
}

/* ggml_vec_dot_q8_0_q8_0, AVX2 branch, ggml-quants.c:5761-5782 */
static float vec_dot_q8_0_q8_0(int64_t n, const blk_q8_0 *x, const blk_q8_0 *y) {
    const int64_t nb = n / 32;
    float acc[8] = {0};
    for (int64_t ib = 0; ib < nb; ++ib) {
        const float d = ps_or_fp16_to_fp32(x[ib].d) * ps_or_fp16_to_fp32(y[ib].d);
        for (int l = 0; l < 8; l++) {
            int s = 0;
            for (int b = 0; b < 4; b++) s += x[ib].qs[4 * l + b] * y[ib].qs[4 * l + b];
            acc[l] = fmaf(d, (float)s, acc[l]);
        }
    }
    return hsum8(acc);
}

float ps_or_vec_dot(int wtype, int64_t k, const void *w, const void *xq) {
    switch (wtype) {
    case PS_OR_Q4_K: return vec_dot_q4_K_q8_K(k, (const blk_q4_K *)w, (const blk_q8_K *)xq);
    case PS_OR_Q5_K: return vec_dot_q5_K_q8_K(k, (const blk_q5_K *)w, (const blk_q8_K *)xq);
    case PS_OR_Q6_K: return vec_dot_q6_K_q8_K(k, (const blk_q6_K *)w, (const blk_q8_K *)xq);
    case PS_OR_Q4_0: return vec_dot_q4_0_q8_0(k, (const blk_q4_0 *)w, (const blk_q8_0 *)xq);
    case PS_OR_Q8_0: return vec_dot_q8_0_q8_0(k, (const blk_q8_0 *)w, (const blk_q8_0 *)xq);
    case PS_OR_F32: return ps_or_vec_dot_f32(k, (const float *)w, (const float *)xq);
    default: return NAN;
    }
}

/* ggml_vec_dot_f32, AVX branch: GGML_F32_STEP 32, 4 accumulators x 8 lanes, GGML_F32x8_REDUCE (ggml.c:1354-1372),
 * scalar leftovers (ggml.c:2092-2131).  The leftover `sumf += x[i]*y[i]` is NOT fused in the reference build: gcc
 * vectorises the products (vmulps) and adds them to sumf one by one, in order (vaddss) — objdump of oracle/_ref. */
float ps_or_vec_dot_f32(int64_t n, const float *x, const float *y) {
    const int64_t np = n & ~(int64_t)31;
    float sum[4][8];
    memset(sum, 0, sizeof(sum));
    for (int64_t i = 0; i < np; i += 32)
        for (int j = 0; j < 4; j++)
            for (int l = 0; l < 8; l++) sum[j][l] = fmaf(x[i + j * 8 + l], y[i + j * 8 + l], sum[j][l]);
    float t0[4];
    for (int l = 0; l < 8; l++) {
        sum[0][l] = sum[0][l] + sum[2][l];
        sum[1][l] = sum[1][l] + sum[3][l];
    }
    for (int l = 0; l < 8; l++) sum[0][l] = sum[0][l] + sum[1][l];
    for (int l = 0; l < 4; l++) t0[l] = sum[0][l] + sum[0][l + 4];
    const float t10 = t0[0] + t0[1], t11 = t0[2] + t0[3];
    float sumf = t10 + t11;
    for (int64_t i = np; i < n; ++i) sumf += x[i] * y[i];
    return sumf;
}

/* ------------------------------------------------------------------------------------------------ operators */
/* powerserve_compute_forward_mul_mat, ggml.c:13434-13648: quantise every src1 row to vec_dot_type (:13502-13530),
 * then dst[n, b] = vec_dot(W row n, xq row b) (:13344-13432).  Each output is produced by one vec_dot call, so the
 * result is independent of the thread count. */
typedef struct { int wtype; const void *w; int64_t K, N, bs; float *dst; const char *xq; size_t xrow, wrow; } mm_ctx;
static void mm_rows(int64_t n0, int64_t n1, void *p) {
    const mm_ctx *c = (const mm_ctx *)p;
    for (int64_t n = n0; n < n1; n++)
        for (int64_t b = 0; b < c->bs; b++)
            c->dst[b * c->N + n] = ps_or_vec_dot(c->wtype, c->K, (const char *)c->w + n * c->wrow, c->xq + b * c->xrow);
}
void ps_or_matmul(int wtype, const void *w, int64_t K, int64_t N, const float *x, int64_t bs, float *dst) {
    const int qt = ps_or_vec_dot_type(wtype);
    const size_t xrow = ps_or_row_size(qt, K), wrow = ps_or_row_size(wtype, K);
    char *xq = (char *)malloc(xrow * (size_t)bs);
    for (int64_t b = 0; b < bs; b++) ps_or_quantize_row(qt, x + b * K, xq + b * xrow, K);
    mm_ctx c = {wtype, w, K, N, bs, dst, xq, xrow, wrow};
    par_for(N, 16, mm_rows, &c);
    free(xq);
}

/* powerserve_compute_forward_rms_norm_f32, ggml.c:12667-12721 (+ ggml_vec_scale_f32_weight :2442-2470):
 * double-precision sequential sum of fp32 squares; y = x * (w * scale). */
void ps_or_rmsnorm(float *dst, const float *x, const float *w, int64_t dim, int64_t bs, float eps) {
    for (int64_t b = 0; b < bs; b++) {
        const float *xr = x + b * dim;
        float *yr = dst + b * dim;
        double sum = 0.0;
        for (int64_t i = 0; i < dim; i++) sum += (double)(xr[i] * xr[i]);
        const float mean = (float)(sum / (double)dim);
        const float scale = 1.0f / sqrtf(mean + eps);
        for (int64_t i = 0; i < dim; i++) yr[i] = xr[i] * (w[i] * scale);
    }
}

/* ggml_compute_forward_rope_f32, ggml.c:15368-15497 with ggml_rope_cache_init :15342-15356 and rope_yarn
 * :15319-15336 at ext_factor = 0, freq_factors = NULL.  src/dst: {head_size, n_heads, bs}.
 * The rotation `x0*c - x1*s`, `x0*s + x1*c` is NOT fused in the reference build: gcc emits four vmulss/vmulps and
 * a vsubss/vaddss (vaddsubps when vectorised) — objdump of oracle/_ref; so every product is rounded to fp32. */
void ps_or_rope(float *dst, const float *src, int64_t head_size, int64_t n_heads, int64_t bs, const int32_t *pos,
                int n_dims, int mode, float freq_base, float freq_scale, float attn_factor) {
    const float theta_scale = powf(freq_base, -2.0f / n_dims);
    const int is_neox = mode & 2;
    float *cache = (float *)malloc(sizeof(float) * (size_t)head_size);
    for (int64_t i2 = 0; i2 < bs; i2++) {
        float theta = (float)pos[i2];
        for (int64_t i0 = 0; i0 < head_size; i0 += 2) {
            const float th = freq_scale * theta;
            cache[i0 + 0] = cosf(th) * attn_factor;
            cache[i0 + 1] = sinf(th) * attn_factor;
            cache[i0 + 1] *= 1.0f;
            theta *= theta_scale;
        }
        for (int64_t i1 = 0; i1 < n_heads; i1++) {
            const float *s = src + (i2 * n_heads + i1) * head_size;
            float *d = dst + (i2 * n_heads + i1) * head_size;
            if (!is_neox) {
                for (int64_t i0 = 0; i0 < n_dims; i0 += 2) {
                    const float c = cache[i0], sn = cache[i0 + 1], x0 = s[i0], x1 = s[i0 + 1];
                    d[i0] = x0 * c - x1 * sn;
                    d[i0 + 1] = x0 * sn + x1 * c;
                }
            } else {
                for (int64_t i0 = 0; i0 < n_dims; i0 += 2) {
                    const int64_t ic = i0 / 2;
                    const float c = cache[i0], sn = cache[i0 + 1], x0 = s[ic], x1 = s[ic + n_dims / 2];
                    d[ic] = x0 * c - x1 * sn;
                    d[ic + n_dims / 2] = x0 * sn + x1 * c;
                }
            }
            for (int64_t i0 = n_dims; i0 < head_size; i0++) d[i0] = s[i0];
        }
    }
    free(cache);
}

/* GET_MASK, src/executor/executor.cpp:210-224: positions only, the tree mask is ignored */
void ps_or_get_mask(float *mask, int64_t n_kv, int64_t bs, const int32_t *pos) {
    for (int64_t i = 0; i < bs; i++)
        for (int64_t j = 0; j < n_kv; j++) mask[j + i * n_kv] = (j <= (int64_t)pos[i]) ? 0.f : -INFINITY;
}

/* one lane of ggml_v_expf, AVX2+FMA branch, ggml.c:2685-2722 */
float ps_or_v_expf(float x) {
    const float r = 0x1.8p23f;
    const float z = fmaf(x, 0x1.715476p+0f, r);
    const float n = z - r;
    const float b = fmaf(-n, 0x1.7f7d1cp-20f, fmaf(-n, 0x1.62e4p-1f, x));
    uint32_t zb, one = 0x3f800000u;
    memcpy(&zb, &z, 4);
    const uint32_t e = zb << 23;
    uint32_t kb = e + one;
    float k;
    memcpy(&k, &kb, 4);
    const int c = fabsf(n) > 126.0f;
    const float u = b * b;
    const float j = fmaf(fmaf(fmaf(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u, fmaf(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)), u,
                         0x1.ffffecp-1f * b);
    if (!c) return fmaf(j, k, k);
    const uint32_t g = (n <= 0.0f) ? 0x82000000u : 0u;
    uint32_t s1b = g + 0x7f000000u, s2b = e - g;
    float s1, s2;
    memcpy(&s1, &s1b, 4);
    memcpy(&s2, &s2b, 4);
    const int d = fabsf(n) > 192.0f;
    if (d) return s1 * s1;
    return fmaf(s2, j, s2) * s1;
}

/* ggml_compute_forward_soft_max_f32, ggml.c:14846-14940 with ggml_vec_soft_max_f32 :2814-2868 (AVX2 branch: 8-wide
 * ggml_v_expf + in-register hsum added to a double; tail with libm expf) — x: {ne0, ne1, ne2}, mask: {ne0, ne1}. */
void ps_or_softmax_ext(float *dst, const float *x, const float *mask, int64_t ne0, int64_t ne1, int64_t ne2, float scale) {
    float *wp = (float *)malloc(sizeof(float) * (size_t)ne0);
    for (int64_t i1 = 0; i1 < ne1 * ne2; i1++) {
        const float *sp = x + i1 * ne0;
        float *dp = dst + i1 * ne0;
        const float *mp = mask ? mask + (i1 % ne1) * ne0 : NULL;
        for (int64_t i = 0; i < ne0; i++) wp[i] = sp[i] * scale;
        if (mp) for (int64_t i = 0; i < ne0; i++) wp[i] += 1.0f * mp[i];
        float max = -INFINITY;
        for (int64_t i = 0; i < ne0; i++) max = (max > wp[i]) ? max : wp[i];
        double sum = 0;
        int64_t i = 0;
        for (; i + 7 < ne0; i += 8) {
            float v[8];
            for (int l = 0; l < 8; l++) { v[l] = ps_or_v_expf(wp[i + l] - max); dp[i + l] = v[l]; }
            sum += (double)hsum8(v);
        }
        for (; i < ne0; ++i) {
            float val = expf(wp[i] - max);
            sum += (double)val;
            dp[i] = val;
        }
        const float inv = (float)(1.0 / sum);
        for (int64_t k = 0; k < ne0; k++) dp[k] *= inv;
    }
    free(wp);
}

/* powerserve_compute_forward_add_f32, ggml.c:10042-10112: dst = a + b, b broadcast row-wise */
void ps_or_add(float *dst, const float *a, const float *b, int64_t n, int64_t nb) {
    for (int64_t i = 0; i < n; i++) dst[i] = a[i] + b[i % nb];
}

/* GGMLBackend::silu_hadamard, src/backend/ggml/ggml.cpp:115-129 */
void ps_or_silu_hadamard(float *dst, const float *gate, const float *up, int64_t n) {
    for (int64_t j = 0; j < n; j++) {
        float val = gate[j];
        val *= (1.0f / (1.0f + expf(-val)));
        val *= up[j];
        dst[j] = val;
    }
}

/* GGMLBackend::get_embedding, src/backend/ggml/ggml_wrapper.cpp:181-211 (+ Q4_K/Q6_K enablement) */
void ps_or_get_embedding(float *dst, const void *w, int wtype, int64_t dim, const int32_t *tokens, int64_t bs) {
    const size_t rb = ps_or_row_size(wtype, dim);
    for (int64_t i = 0; i < bs; i++) ps_or_dequantize_row(wtype, (const char *)w + rb * (size_t)tokens[i], dst + i * dim, dim);
}

/* mat_mul(k_view, q): src/model/module/norm_attention.cpp:115-129; kq is {n_kv, bs, n_heads}; GQA broadcast
 * i02 = i12 / r2 (ggml.c:13365,13399-13400).  q is {head_size, n_heads, bs} (rope output, read through the permute). */
void ps_or_attn_scores(float *kq, const float *k_cache, const float *q, int64_t hs, int64_t n_heads, int64_t n_kv_heads,
                       int64_t n_kv, int64_t bs) {
    const int64_t r2 = n_heads / n_kv_heads, kv_dim = hs * n_kv_heads;
    for (int64_t h = 0; h < n_heads; h++)
        for (int64_t i = 0; i < bs; i++)
            for (int64_t j = 0; j < n_kv; j++)
                kq[(h * bs + i) * n_kv + j] =
                    ps_or_vec_dot_f32(hs, k_cache + j * kv_dim + (h / r2) * hs, q + (i * n_heads + h) * hs);
}

/* mat_mul(v_view, kq) + permute + cont: norm_attention.cpp:138-151; out is {dim, bs} = [i][h*hs + d] */
void ps_or_attn_pv(float *out, const float *v_cache_t, const float *p, int64_t hs, int64_t n_heads, int64_t n_kv_heads,
                   int64_t n_kv, int64_t n_ctx, int64_t bs) {
    const int64_t r2 = n_heads / n_kv_heads;
    for (int64_t h = 0; h < n_heads; h++)
        for (int64_t i = 0; i < bs; i++)
            for (int64_t d = 0; d < hs; d++)
                out[(i * n_heads + h) * hs + d] =
                    ps_or_vec_dot_f32(n_kv, v_cache_t + ((h / r2) * hs + d) * n_ctx, p + (h * bs + i) * n_kv);
}

/* ------------------------------------------------------------------------------------------------ model */
struct ps_or_model {
    ps_or_config c;
    ps_or_weights w;
    ps_or_layer *layers;
    float **k_cache; /* [L][n_ctx][kv_dim]  (ggml_kv_cache.cpp:35-58, norm_attention.cpp:82-91) */
    float **v_cache; /* [L][kv_dim][n_ctx]  transposed (norm_attention.cpp:93-104) */
    int position;
    /* taps of the last forward */
    int tap_bs;
    float **tap[4];
};

ps_or_model *ps_or_model_create(const ps_or_config *cfg, const ps_or_weights *w) {
    ps_or_model *m = (ps_or_model *)calloc(1, sizeof(*m));
    m->c = *cfg;
    m->w = *w;
    m->layers = (ps_or_layer *)malloc(sizeof(ps_or_layer) * (size_t)cfg->n_layers);
    memcpy(m->layers, w->layers, sizeof(ps_or_layer) * (size_t)cfg->n_layers);
    m->w.layers = m->layers;
    const size_t kvn = (size_t)cfg->n_ctx * (size_t)(cfg->n_kv_heads * cfg->head_size);
    m->k_cache = (float **)calloc((size_t)cfg->n_layers, sizeof(float *));
    m->v_cache = (float **)calloc((size_t)cfg->n_layers, sizeof(float *));
    for (int t = 0; t < 4; t++) m->tap[t] = (float **)calloc((size_t)cfg->n_layers, sizeof(float *));
    for (int L = 0; L < cfg->n_layers; L++) {
        m->k_cache[L] = (float *)calloc(kvn, sizeof(float));
        m->v_cache[L] = (float *)calloc(kvn, sizeof(float));
    }
    return m;
}

void ps_or_model_free(ps_or_model *m) {
    if (!m) return;
    for (int L = 0; L < m->c.n_layers; L++) {
        free(m->k_cache[L]); free(m->v_cache[L]);
        for (int t = 0; t < 4; t++) free(m->tap[t][L]);
    }
    for (int t = 0; t < 4; t++) free(m->tap[t]);
    free(m->k_cache); free(m->v_cache); free(m->layers); free(m);
}

void ps_or_model_reset(ps_or_model *m) { m->position = 0; }
int ps_or_model_position(const ps_or_model *m) { return m->position; }
void ps_or_model_set_position(ps_or_model *m, int pos) { m->position = pos; }
const float *ps_or_model_k_cache(ps_or_model *m, int layer) { return m->k_cache[layer]; }
const float *ps_or_model_v_cache(ps_or_model *m, int layer) { return m->v_cache[layer]; }

static void tap_store(ps_or_model *m, int which, int L, const float *src, size_t n) {
    m->tap[which][L] = (float *)realloc(m->tap[which][L], n * sizeof(float));
    memcpy(m->tap[which][L], src, n * sizeof(float));
}

int64_t ps_or_model_tap(ps_or_model *m, int layer, int which, float *out) {
    if (layer < 0 || layer >= m->c.n_layers || which < 0 || which > 3 || !m->tap[which][layer]) return 0;
    const int64_t n = (int64_t)m->tap_bs * m->c.dim;
    if (out) memcpy(out, m->tap[which][layer], (size_t)n * sizeof(float));
    return n;
}

/* LlamaModel::forward / Qwen2Model::forward: src/model/llama/llama_model.cpp:52-117, qwen2_model.cpp (bias adds),
 * NormAttention::build src/model/module/norm_attention.cpp:26-160, FFN::build src/model/module/ffn.cpp:22-42. */
int ps_or_model_forward(ps_or_model *m, const int32_t *tokens, const int32_t *pos, int bs, int lm_head, float *logits) {
    const ps_or_config *c = &m->c;
    const int64_t dim = c->dim, hs = c->head_size, nh = c->n_heads, nkv = c->n_kv_heads, kv_dim = hs * nkv;
    const int64_t qdim = nh * hs, ffn = c->ffn_dim, n_ctx = c->n_ctx;
    if (bs <= 0 || pos[0] + bs > n_ctx) return -1;
    m->tap_bs = bs;
    float *x = (float *)malloc(sizeof(float) * (size_t)(dim * bs));
    float *xn = (float *)malloc(sizeof(float) * (size_t)(dim * bs));
    float *q = (float *)malloc(sizeof(float) * (size_t)(qdim * bs));
    float *k = (float *)malloc(sizeof(float) * (size_t)(kv_dim * bs));
    float *v = (float *)malloc(sizeof(float) * (size_t)(kv_dim * bs));
    float *qr = (float *)malloc(sizeof(float) * (size_t)(qdim * bs));
    float *kr = (float *)malloc(sizeof(float) * (size_t)(kv_dim * bs));
    float *att = (float *)malloc(sizeof(float) * (size_t)(qdim * bs));
    float *tmp = (float *)malloc(sizeof(float) * (size_t)(dim * bs));
    float *g = (float *)malloc(sizeof(float) * (size_t)(ffn * bs));
    float *u = (float *)malloc(sizeof(float) * (size_t)(ffn * bs));
    const int64_t n_kv = (int64_t)pos[bs - 1] + 1;
    const int64_t cur_pos = pos[0];
    float *kq = (float *)malloc(sizeof(float) * (size_t)(n_kv * bs * nh));
    float *mask = (float *)malloc(sizeof(float) * (size_t)(n_kv * bs));
    const float kq_scale = 1.0f / sqrtf((float)hs);

    ps_or_get_embedding(x, m->w.token_embd.data, m->w.token_embd.type, dim, tokens, bs);
    for (int L = 0; L < c->n_layers; L++) {
        const ps_or_layer *lw = &m->layers[L];
        tap_store(m, 0, L, x, (size_t)(dim * bs));
        ps_or_rmsnorm(xn, x, (const float *)lw->attn_norm.data, dim, bs, c->norm_eps);
        ps_or_matmul(lw->attn_q.type, lw->attn_q.data, dim, qdim, xn, bs, q);
        ps_or_matmul(lw->attn_k.type, lw->attn_k.data, dim, kv_dim, xn, bs, k);
        ps_or_matmul(lw->attn_v.type, lw->attn_v.data, dim, kv_dim, xn, bs, v);
        if (c->qkv_bias) {
            ps_or_add(q, q, (const float *)lw->q_bias.data, qdim * bs, qdim);
            ps_or_add(k, k, (const float *)lw->k_bias.data, kv_dim * bs, kv_dim);
            ps_or_add(v, v, (const float *)lw->v_bias.data, kv_dim * bs, kv_dim);
        }
        ps_or_rope(qr, q, hs, nh, bs, pos, c->rope_n_dims, c->rope_type, c->rope_freq_base, c->rope_freq_scale, c->rope_attn_factor);
        ps_or_rope(kr, k, hs, nkv, bs, pos, c->rope_n_dims, c->rope_type, c->rope_freq_base, c->rope_freq_scale, c->rope_attn_factor);
        /* KV store: K rows at cur_pos.., V transposed columns at cur_pos.. */
        memcpy(m->k_cache[L] + cur_pos * kv_dim, kr, sizeof(float) * (size_t)(kv_dim * bs));
        for (int64_t i = 0; i < bs; i++)
            for (int64_t e = 0; e < kv_dim; e++) m->v_cache[L][e * n_ctx + cur_pos + i] = v[i * kv_dim + e];
        ps_or_attn_scores(kq, m->k_cache[L], qr, hs, nh, nkv, n_kv, bs);
        ps_or_get_mask(mask, n_kv, bs, pos);
        ps_or_softmax_ext(kq, kq, mask, n_kv, bs, nh, kq_scale);
        ps_or_attn_pv(att, m->v_cache[L], kq, hs, nh, nkv, n_kv, n_ctx, bs);
        tap_store(m, 1, L, att, (size_t)(qdim * bs));
        ps_or_matmul(lw->attn_output.type, lw->attn_output.data, qdim, dim, att, bs, tmp);
        ps_or_add(x, x, tmp, dim * bs, dim * bs);
        tap_store(m, 2, L, x, (size_t)(dim * bs));
        ps_or_rmsnorm(xn, x, (const float *)lw->ffn_norm.data, dim, bs, c->norm_eps);
        ps_or_matmul(lw->ffn_gate.type, lw->ffn_gate.data, dim, ffn, xn, bs, g);
        ps_or_matmul(lw->ffn_up.type, lw->ffn_up.data, dim, ffn, xn, bs, u);
        ps_or_silu_hadamard(g, g, u, ffn * bs);
        ps_or_matmul(lw->ffn_down.type, lw->ffn_down.data, ffn, dim, g, bs, tmp);
        ps_or_add(x, x, tmp, dim * bs, dim * bs);
        tap_store(m, 3, L, x, (size_t)(dim * bs));
    }
    if (lm_head) {
        ps_or_rmsnorm(xn, x, (const float *)m->w.output_norm.data, dim, bs, c->norm_eps);
        ps_or_matmul(m->w.output.type, m->w.output.data, dim, c->vocab_size, xn, bs, logits);
    }
    m->position += bs; /* m_kv->advance(batch_size), llama_model.cpp:109 */
    free(x); free(xn); free(q); free(k); free(v); free(qr); free(kr); free(att); free(tmp); free(g); free(u); free(kq); free(mask);
    return 0;
}
