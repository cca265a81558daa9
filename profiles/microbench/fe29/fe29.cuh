// 9 x 29-bit signed limbs, Montgomery domain R' = 2^261 -- prototype
#pragma once
#include <stdint.h>
#if defined(__CUDACC__)
#define F29_HD __host__ __device__ __forceinline__
#else
#define F29_HD inline
#endif
struct fe29 { int32_t v[9]; };
#define F29_MASK 0x1fffffff

F29_HD void f29_madw(int64_t& c, int32_t a, int32_t b)
{
#if defined(__CUDA_ARCH__)
    asm("mad.wide.s32 %0, %1, %2, %0;" : "+l"(c) : "r"(a), "r"(b));
#else
    c += (int64_t)a * b;
#endif
}

// columns c[0..16] (value a*b) -> Montgomery reduce by 2^261, carry, fold the bits above 2^256
#if defined(__CUDACC__)
__device__ __constant__ int32_t kF29K[4] = {1 << 9, 1 << 18, -(1 << 21), 1 << 24};
#endif
F29_HD int32_t f29_opaque(int32_t k)
{
#if defined(__CUDA_ARCH__)
    int32_t r; asm volatile("mov.s32 %0, %1;" : "=r"(r) : "r"(k)); return r;   // keeps ptxas from splitting mad.wide by a power of two into shifts
#else
    return k;
#endif
}
F29_HD fe29 f29_reduce(int64_t* c)
{
    #if defined(__CUDA_ARCH__)
    const int32_t k9 = kF29K[0], k18 = kF29K[1], k21 = kF29K[2], k24 = kF29K[3];   // from constant memory: ptxas cannot split the multiplications into shifts
#else
    const int32_t k9 = 1 << 9, k18 = 1 << 18, k21 = -(1 << 21), k24 = 1 << 24;
#endif
#pragma unroll
    for (int i = 0; i < 9; i++) {
        const int32_t m = (int32_t)((uint32_t)c[i] & F29_MASK);
        // + m * p * 2^(29 i),  p = 2^256 - 2^224 + 2^192 + 2^96 - 1
        f29_madw(c[i + 3], m, k9);
        f29_madw(c[i + 6], m, k18);
        f29_madw(c[i + 7], m, k21);
        f29_madw(c[i + 8], m, k24);
        c[i + 1] += c[i] >> 29;                 // (c[i] - m) / 2^29, exact
    }
    fe29 r;
#pragma unroll
    for (int k = 9; k < 17; k++) {
        r.v[k - 9] = (int32_t)((uint32_t)c[k] & F29_MASK);
        if (k < 16) c[k + 1] += c[k] >> 29;
    }
    int32_t top = (int32_t)(c[16] >> 29);      // limb 8 (bit 232 up), small
    const int32_t hi = top >> 24;              // bits >= 2^256
    top &= 0xffffff;
    r.v[8] = top;
    r.v[7] += hi << 21; r.v[6] -= hi << 18; r.v[3] -= hi << 9; r.v[0] += hi;
    return r;
}

F29_HD fe29 f29_mul(const fe29& a, const fe29& b)
{
    int64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) c[k] = 0;
#pragma unroll
    for (int i = 0; i < 9; i++)
#pragma unroll
        for (int j = 0; j < 9; j++) f29_madw(c[i + j], a.v[i], b.v[j]);
    return f29_reduce(c);
}

F29_HD fe29 f29_sqr(const fe29& a)
{
    int64_t c[17];
#pragma unroll
    for (int k = 0; k < 17; k++) c[k] = 0;
    int32_t d[9];
#pragma unroll
    for (int i = 0; i < 9; i++) d[i] = a.v[i] * 2;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        f29_madw(c[2 * i], a.v[i], a.v[i]);
#pragma unroll
        for (int j = i + 1; j < 9; j++) f29_madw(c[i + j], a.v[i], d[j]);
    }
    return f29_reduce(c);
}
F29_HD fe29 f29_sub(const fe29& a, const fe29& b) { fe29 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] - b.v[i]; return r; }
F29_HD fe29 f29_add(const fe29& a, const fe29& b) { fe29 r; for (int i = 0; i < 9; i++) r.v[i] = a.v[i] + b.v[i]; return r; }
