// ps_math.cuh — exact-arithmetic building blocks shared by every kernel.
//
// Numerics contract: the CUDA backend reproduces the reference's x86 AVX2+FMA results BIT FOR BIT.  The reference
// quantises activations before every weight matmul, so a 1-ulp upstream difference is amplified to
// percent-level logit noise in one forward pass (F13); bit-identity is the only stable parity regime.  Rules used
// throughout:
//   * every fp32 operation is spelled with an explicit round-to-nearest intrinsic (__fmul_rn / __fadd_rn / __fmaf_rn /
//     __fdiv_rn / __fsqrt_rn) so nvcc can neither fuse nor reorder it; FMAs appear exactly where the reference's
//     AVX2 code (or gcc's contraction of its scalar code, checked in the compiled reference) has one;
//   * libm calls of the reference (expf) are restated from the published glibc algorithm in double precision;
//     cosf/sinf/powf only feed the RoPE table, which the host builds with the platform libm at context creation;
//   * 8-lane AVX accumulators are kept as 8 separate chains and reduced in hsum_float_8 order.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define PS_HD __host__ __device__ __forceinline__
#define PS_D __device__ __forceinline__
#else
#define PS_HD inline
#define PS_D inline
#endif

// ---------------------------------------------------------------------------------------------------- host/device fp
PS_HD float ps_mul(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fmul_rn(a, b);
#else
    volatile float r = a * b;
    return r;
#endif
}
PS_HD float ps_add(float a, float b) {
#ifdef __CUDA_ARCH__
    return __fadd_rn(a, b);
#else
    volatile float r = a + b;
    return r;
#endif
}
PS_HD float ps_fma(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#else
    return fmaf(a, b, c);
#endif
}
PS_HD double ps_dfma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return fma(a, b, c);
#endif
}
PS_HD uint32_t ps_f2u(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}
PS_HD float ps_u2f(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}

// glibc expf (sysdeps/ieee754/flt-32/e_expf.c, the ARM "optimized routines" algorithm, EXP2F_TABLE_BITS = 5; the
// x86-64 ifunc picks the FMA build): z = x*N/ln2, k = round(z), r = z-k, exp = 2^(k/N) * P(r) evaluated in double
// and rounded once to float.  The reference calls it in GGMLBackend::silu_hadamard (src/backend/ggml/ggml.cpp:124)
// and in the scalar tail of ggml_vec_soft_max_f32 (ggml.c:2862).  Pinned against the platform libm by
// tests/test_host_math.py (not a GPU test).
#if defined(__CUDACC__)
__device__ __constant__ uint64_t ps_exp2f_tab_dev[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, 0x3fef72b83c7d517bULL,
    0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, 0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL,
    0x3feedea64c123422ULL, 0x3feece086061892dULL, 0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL,
    0x3feea47eb03a5585ULL, 0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, 0x3feee89f995ad3adULL,
    0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, 0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL,
    0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};
#endif
static const uint64_t ps_exp2f_tab_host[32] = {
    0x3ff0000000000000ULL, 0x3fefd9b0d3158574ULL, 0x3fefb5586cf9890fULL, 0x3fef9301d0125b51ULL, 0x3fef72b83c7d517bULL,
    0x3fef54873168b9aaULL, 0x3fef387a6e756238ULL, 0x3fef1e9df51fdee1ULL, 0x3fef06fe0a31b715ULL, 0x3feef1a7373aa9cbULL,
    0x3feedea64c123422ULL, 0x3feece086061892dULL, 0x3feebfdad5362a27ULL, 0x3feeb42b569d4f82ULL, 0x3feeab07dd485429ULL,
    0x3feea47eb03a5585ULL, 0x3feea09e667f3bcdULL, 0x3fee9f75e8ec5f74ULL, 0x3feea11473eb0187ULL, 0x3feea589994cce13ULL,
    0x3feeace5422aa0dbULL, 0x3feeb737b0cdc5e5ULL, 0x3feec49182a3f090ULL, 0x3feed503b23e255dULL, 0x3feee89f995ad3adULL,
    0x3feeff76f2fb5e47ULL, 0x3fef199bdd85529cULL, 0x3fef3720dcef9069ULL, 0x3fef5818dcfba487ULL, 0x3fef7c97337b9b5fULL,
    0x3fefa4afa2a490daULL, 0x3fefd0765b6e4540ULL};

PS_HD float ps_expf_glibc(float x) {
    const double InvLn2N = 0x1.71547652b82fep+0 * 32, SHIFT = 0x1.8p+52;
    const double C0 = 0x1.c6af84b912394p-5 / 32 / 32 / 32, C1 = 0x1.ebfce50fac4f3p-3 / 32 / 32, C2 = 0x1.62e42ff0c52d6p-1 / 32;
    const uint32_t ux = ps_f2u(x);
    const uint32_t abstop = (ux >> 20) & 0x7ff;
    if (abstop >= 0x42b) { // top12(88.0f)
        if (ux == 0xff800000u) return 0.0f;
        if (abstop >= 0x7f8) return ps_add(x, x);
        if (x > 0x1.62e42ep6f) return ps_u2f(0x7f800000u);
        if (x < -0x1.9fe368p6f) return 0.0f;
    }
    const double xd = (double)x;
#ifdef __CUDA_ARCH__
    const double z = __dmul_rn(InvLn2N, xd);
    double kd = __dadd_rn(z, SHIFT);
    const uint64_t ki = (uint64_t)__double_as_longlong(kd);
    kd = __dsub_rn(kd, SHIFT);
    const double r = __dsub_rn(z, kd);
    uint64_t t = ps_exp2f_tab_dev[ki & 31];
    t += ki << 47;
    const double s = __longlong_as_double((long long)t);
    const double zz = __fma_rn(C0, r, C1);
    const double r2 = __dmul_rn(r, r);
    double y = __fma_rn(C2, r, 1.0);
    y = __fma_rn(zz, r2, y);
    y = __dmul_rn(y, s);
    return __double2float_rn(y);
#else
    volatile double z = InvLn2N * xd;
    volatile double kd = z + SHIFT;
    uint64_t ki;
    {
        double k0 = kd;
        memcpy(&ki, &k0, 8);
    }
    kd = kd - SHIFT;
    volatile double r = z - kd;
    uint64_t t = ps_exp2f_tab_host[ki & 31];
    t += ki << 47;
    double s;
    memcpy(&s, &t, 8);
    const double zz = fma(C0, r, C1);
    volatile double r2 = r * r;
    double y = fma(C2, r, 1.0);
    y = fma(zz, r2, y);
    volatile double ys = y * s;
    return (float)ys;
#endif
}

// One lane of ggml_v_expf, AVX2+FMA branch (libs/ggml/src/ggml.c:2685-2722).
PS_HD float ps_v_expf(float x) {
    const float r = 0x1.8p23f;
    const float z = ps_fma(x, 0x1.715476p+0f, r);
    const float n = ps_add(z, -r);
    const float b = ps_fma(-n, 0x1.7f7d1cp-20f, ps_fma(-n, 0x1.62e4p-1f, x));
    const uint32_t e = ps_f2u(z) << 23;
    const float k = ps_u2f(e + 0x3f800000u);
    const bool c = fabsf(n) > 126.0f;
    const float u = ps_mul(b, b);
    const float j = ps_fma(ps_fma(ps_fma(0x1.0e4020p-7f, b, 0x1.573e2ep-5f), u, ps_fma(0x1.555e66p-3f, b, 0x1.fffdb6p-2f)), u,
                           ps_mul(0x1.ffffecp-1f, b));
    if (!c) return ps_fma(j, k, k);
    const uint32_t g = (n <= 0.0f) ? 0x82000000u : 0u;
    const float s1 = ps_u2f(g + 0x7f000000u);
    const float s2 = ps_u2f(e - g);
    if (fabsf(n) > 192.0f) return ps_mul(s1, s1);
    return ps_mul(ps_fma(s2, j, s2), s1);
}

// GGMLBackend::silu_hadamard (src/backend/ggml/ggml.cpp:115-129): val *= 1/(1+expf(-val)); val *= up
PS_HD float ps_silu_mul(float g, float u) {
#ifdef __CUDA_ARCH__
    const float e = ps_expf_glibc(-g);
    const float s = __fdiv_rn(1.0f, __fadd_rn(1.0f, e));
    return __fmul_rn(__fmul_rn(g, s), u);
#else
    const float e = ps_expf_glibc(-g);
    volatile float den = 1.0f + e;
    volatile float s = 1.0f / den;
    volatile float v = g * s;
    volatile float o = v * u;
    return o;
#endif
}
