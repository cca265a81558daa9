// P-256 field (mod p) and scalar (mod n) arithmetic, 8 x 32-bit limbs in registers, Montgomery form.
//
// This is the arithmetic layer of the B200 replacement for the leaf the reference reaches at
// bccsp/sw/ecdsa.go:56 (Go crypto/ecdsa.Verify -> crypto/elliptic P-256).  Nothing here is a port of
// Go's amd64 assembly: limbs are 32-bit so that one product is one IMAD.WIDE.U32 on the sm_100a fma
// pipe and the carry chains run on the alu pipe beside it.
//
//   * device path: hand-written PTX (mad.lo.cc/madc.hi.cc pairs that ptxas fuses into IMAD.WIDE.U32
//     with predicate carries; add.cc/addc.cc chains);
//   * host path (#ifndef __CUDA_ARCH__): a plain uint64_t restatement of the same functions.  It exists
//     only so tests/host_sim can run the *same* point/verify code on the CPU build box; it is never
//     linked into libfabgpu_ecdsa.so's product entry points.
//
// Montgomery reduction mod p uses p = 2^256 - 2^224 + 2^192 + 2^96 - 1 == -1 (mod 2^64): the
// per-round multiplier is the low 64 bits of the accumulator itself and m*p is shifts/adds only.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FAB_HD __host__ __device__ __forceinline__
#define FAB_D __device__ __forceinline__
#else
#define FAB_HD inline
#endif

namespace fabgpu {

struct alignas(16) u256 { uint32_t v[8]; };   // little-endian limbs; a field/scalar element or a plain integer

// ---- constants (little-endian limbs) ----
#define FAB_P_LIMBS  {0xffffffffu, 0xffffffffu, 0xffffffffu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0xffffffffu}
#define FAB_N_LIMBS  {0xfc632551u, 0xf3b9cac2u, 0xa7179e84u, 0xbce6faadu, 0xffffffffu, 0xffffffffu, 0x00000000u, 0xffffffffu}
#define FAB_NPRIME_LIMBS {0xee00bc4fu, 0xccd1c8aau, 0x7d74d2e4u, 0x48c94408u, 0xc588c6f6u, 0x50fe77ecu, 0xa9d6281cu, 0x60d06633u} /* -n^-1 mod 2^256 */
#define FAB_ONE_MONT_P {0x00000001u, 0x00000000u, 0x00000000u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffffeu, 0x00000000u} /* 2^256 mod p */
#define FAB_R2_MOD_P {0x00000003u, 0x00000000u, 0xffffffffu, 0xfffffffbu, 0xfffffffeu, 0xffffffffu, 0xfffffffdu, 0x00000004u}
#define FAB_B_MONT   {0x29c4bddfu, 0xd89cdf62u, 0x78843090u, 0xacf005cdu, 0xf7212ed6u, 0xe5a220abu, 0x04874834u, 0xdc30061du}
#define FAB_GX_MONT  {0x18a9143cu, 0x79e730d4u, 0x5fedb601u, 0x75ba95fcu, 0x77622510u, 0x79fb732bu, 0xa53755c6u, 0x18905f76u}
#define FAB_GY_MONT  {0xce95560au, 0xddf25357u, 0xba19e45cu, 0x8b4ab8e4u, 0xdd21f325u, 0xd2e88688u, 0x25885d85u, 0x8571ff18u}
#define FAB_ONE_MONT_N {0x039cdaafu, 0x0c46353du, 0x58e8617bu, 0x43190552u, 0x00000000u, 0x00000000u, 0xffffffffu, 0x00000000u} /* 2^256 mod n */
#define FAB_R2_MOD_N {0xbe79eea2u, 0x83244c95u, 0x49bd6fa6u, 0x4699799cu, 0x2b6bec59u, 0x2845b239u, 0xf3d95620u, 0x66e12d94u}
#define FAB_P_MINUS_N {0x039cdaaeu, 0x0c46353du, 0x58e8617bu, 0x43190553u, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000000u}

FAB_HD u256 u256_const(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7)
{
    u256 r; r.v[0] = a0; r.v[1] = a1; r.v[2] = a2; r.v[3] = a3; r.v[4] = a4; r.v[5] = a5; r.v[6] = a6; r.v[7] = a7; return r;
}

FAB_HD u256 fe_p()        { return u256_const(0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u, 1u, 0xffffffffu); }
FAB_HD u256 sc_n()        { return u256_const(0xfc632551u, 0xf3b9cac2u, 0xa7179e84u, 0xbce6faadu, 0xffffffffu, 0xffffffffu, 0x00000000u, 0xffffffffu); }
FAB_HD u256 sc_nprime()   { return u256_const(0xee00bc4fu, 0xccd1c8aau, 0x7d74d2e4u, 0x48c94408u, 0xc588c6f6u, 0x50fe77ecu, 0xa9d6281cu, 0x60d06633u); }
FAB_HD u256 fe_one()      { return u256_const(0x00000001u, 0x00000000u, 0x00000000u, 0xffffffffu, 0xffffffffu, 0xffffffffu, 0xfffffffeu, 0x00000000u); }
FAB_HD u256 fe_r2()       { return u256_const(0x00000003u, 0x00000000u, 0xffffffffu, 0xfffffffbu, 0xfffffffeu, 0xffffffffu, 0xfffffffdu, 0x00000004u); }
FAB_HD u256 fe_b_mont()   { return u256_const(0x29c4bddfu, 0xd89cdf62u, 0x78843090u, 0xacf005cdu, 0xf7212ed6u, 0xe5a220abu, 0x04874834u, 0xdc30061du); }
FAB_HD u256 fe_gx_mont()  { return u256_const(0x18a9143cu, 0x79e730d4u, 0x5fedb601u, 0x75ba95fcu, 0x77622510u, 0x79fb732bu, 0xa53755c6u, 0x18905f76u); }
FAB_HD u256 fe_gy_mont()  { return u256_const(0xce95560au, 0xddf25357u, 0xba19e45cu, 0x8b4ab8e4u, 0xdd21f325u, 0xd2e88688u, 0x25885d85u, 0x8571ff18u); }
FAB_HD u256 sc_one()      { return u256_const(0x039cdaafu, 0x0c46353du, 0x58e8617bu, 0x43190552u, 0x00000000u, 0x00000000u, 0xffffffffu, 0x00000000u); }
FAB_HD u256 sc_r2()       { return u256_const(0xbe79eea2u, 0x83244c95u, 0x49bd6fa6u, 0x4699799cu, 0x2b6bec59u, 0x2845b239u, 0xf3d95620u, 0x66e12d94u); }
FAB_HD u256 p_minus_n()   { return u256_const(0x039cdaaeu, 0x0c46353du, 0x58e8617bu, 0x43190553u, 0u, 0u, 0u, 0u); }

// ---- plain 256-bit helpers (same code on host and device) ----
FAB_HD bool u256_is_zero(const u256& a)
{
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
FAB_HD bool u256_eq(const u256& a, const u256& b)
{
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i];
    return o == 0;
}
// a < b
FAB_HD bool u256_lt(const u256& a, const u256& b)
{
    bool lt = false;
#pragma unroll
    for (int i = 0; i < 8; i++) lt = (a.v[i] < b.v[i]) || (a.v[i] == b.v[i] && lt);
    return lt;
}
FAB_HD u256 u256_zero() { return u256_const(0, 0, 0, 0, 0, 0, 0, 0); }
FAB_HD u256 u256_select(bool c, const u256& a, const u256& b)   // c ? a : b
{
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = c ? a.v[i] : b.v[i];
    return r;
}
// big-endian 32 bytes -> limbs (device: two 128-bit loads when 16-byte aligned input is guaranteed by the caller)
FAB_HD u256 u256_from_be(const uint8_t* p)
{
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint8_t* q = p + 4 * (7 - i);
        r.v[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
    }
    return r;
}
FAB_HD void u256_to_be(const u256& a, uint8_t* p)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint8_t* q = p + 4 * (7 - i);
        q[0] = (uint8_t)(a.v[i] >> 24); q[1] = (uint8_t)(a.v[i] >> 16); q[2] = (uint8_t)(a.v[i] >> 8); q[3] = (uint8_t)a.v[i];
    }
}

// ======================================================================================================
//  Device primitives (PTX).  The CC flag is threaded through consecutive asm volatile statements; this is
//  the established idiom for multi-precision arithmetic on NVIDIA GPUs and ptxas keeps the chain intact.
// ======================================================================================================
#if defined(__CUDA_ARCH__)

FAB_D uint32_t ptx_add_cc(uint32_t a, uint32_t b)  { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;"  : "=r"(r) : "r"(a), "r"(b)); return r; }
FAB_D uint32_t ptx_addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FAB_D uint32_t ptx_addc(uint32_t a, uint32_t b)    { uint32_t r; asm volatile("addc.u32 %0, %1, %2;"    : "=r"(r) : "r"(a), "r"(b)); return r; }
FAB_D uint32_t ptx_sub_cc(uint32_t a, uint32_t b)  { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;"  : "=r"(r) : "r"(a), "r"(b)); return r; }
FAB_D uint32_t ptx_subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
FAB_D uint32_t ptx_subc(uint32_t a, uint32_t b)    { uint32_t r; asm volatile("subc.u32 %0, %1, %2;"    : "=r"(r) : "r"(a), "r"(b)); return r; }

// acc[0..7] += (a[0], a[2], a[4], a[6]) * b as four 64-bit products laid end to end; carry -> acc[8]
FAB_D void ptx_mad_row(uint32_t* acc, const uint32_t* a, uint32_t b)
{
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[2]), "+r"(acc[3]) : "r"(a[2]), "r"(b));
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[4]), "+r"(acc[5]) : "r"(a[4]), "r"(b));
    asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[6]), "+r"(acc[7]) : "r"(a[6]), "r"(b));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[8]));
}

// t[0..15] = a * b.  Even-aligned (E) and odd-aligned (O) 64-bit column accumulators keep every
// mad.lo/mad.hi pair on one IMAD.WIDE.U32 and every carry on a predicate; they are merged once at the end.
FAB_D void mul_8x8(uint32_t* t, const uint32_t* a, const uint32_t* b)
{
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if ((i & 1) == 0) { ptx_mad_row(E + i, a, b[i]);     ptx_mad_row(O + i, a + 1, b[i]); }
        else              { ptx_mad_row(O + i - 1, a, b[i]); ptx_mad_row(E + i + 1, a + 1, b[i]); }
    }
    // E holds words 0..16, O holds words 1..17 (O[k] is word k+1)
    t[0] = E[0];
    t[1] = ptx_add_cc(E[1], O[0]);
#pragma unroll
    for (int i = 2; i < 16; i++) t[i] = ptx_addc_cc(E[i], O[i - 1]);
}

// Montgomery reduction mod p of t[0..15] -> r in [0,p):  r = t * 2^-256 mod p.
// Four 64-bit rounds; round k cancels words 2k,2k+1 with m = (t[2k], t[2k+1]) and adds
//   m * (p+1)/2^64 = m * (2^32 + 2^128 * (1 - 2^32 + 2^64))   one 64-bit word higher.
FAB_D u256 fe_reduce(uint32_t* t)
{
    uint32_t c[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int b = 2 * k;
        const uint32_t m0 = t[b], m1 = t[b + 1];
        // W = m * (1 - 2^32 + 2^64)  (4 words, never negative)
        const uint32_t w1 = ptx_sub_cc(m1, m0);
        const uint32_t w2 = ptx_subc_cc(m0, m1);
        const uint32_t w3 = ptx_subc(m1, 0);
        t[b + 3] = ptx_add_cc(t[b + 3], m0);
        t[b + 4] = ptx_addc_cc(t[b + 4], m1);
        t[b + 5] = ptx_addc_cc(t[b + 5], 0);
        t[b + 6] = ptx_addc_cc(t[b + 6], m0);
        t[b + 7] = ptx_addc_cc(t[b + 7], w1);
        t[b + 8] = ptx_addc_cc(t[b + 8], w2);
        t[b + 9] = ptx_addc_cc(t[b + 9], w3);
        c[k] = ptx_addc(0, 0);               // belongs at word b+10; none of the later m's read that high
    }
    // fold the four deferred carries (words 10, 12, 14, 16)
    t[10] = ptx_add_cc(t[10], c[0]);
    t[11] = ptx_addc_cc(t[11], 0);
    t[12] = ptx_addc_cc(t[12], c[1]);
    t[13] = ptx_addc_cc(t[13], 0);
    t[14] = ptx_addc_cc(t[14], c[2]);
    t[15] = ptx_addc_cc(t[15], 0);
    const uint32_t top = ptx_addc(c[3], 0);
    // value = top:t[8..15] < 2p ; subtract p once if >= p
    u256 d;
    d.v[0] = ptx_sub_cc(t[8], 0xffffffffu);
    d.v[1] = ptx_subc_cc(t[9], 0xffffffffu);
    d.v[2] = ptx_subc_cc(t[10], 0xffffffffu);
    d.v[3] = ptx_subc_cc(t[11], 0u);
    d.v[4] = ptx_subc_cc(t[12], 0u);
    d.v[5] = ptx_subc_cc(t[13], 0u);
    d.v[6] = ptx_subc_cc(t[14], 1u);
    d.v[7] = ptx_subc_cc(t[15], 0xffffffffu);
    const uint32_t brw = ptx_subc(top, 0);   // 0 if value >= p, 0xffffffff if value < p
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = brw ? t[8 + i] : d.v[i];
    return r;
}

// FAB_MUL_CALL=1 (default): one out-of-line copy of the multiplier, reached by a register-convention call, so that
// the hot loop of the verify kernel (5 doublings + 1 addition = 56 field multiplications) fits the instruction cache
// instead of inlining 56 x 2.7 KB.  ncu on the fully inlined build showed "no_instruction" as the top stall.
#ifndef FAB_MUL_CALL
#define FAB_MUL_CALL 1
#endif
#if FAB_MUL_CALL
#define FAB_MUL_ATTR __device__ __noinline__
#else
#define FAB_MUL_ATTR __device__ __forceinline__
#endif
FAB_MUL_ATTR u256 fe_mul_dev(u256 a, u256 b)
{
    uint32_t t[16];
    mul_8x8(t, a.v, b.v);
    return fe_reduce(t);
}

// acc[0 .. 2*CNT) += (a[0], a[2], ..., a[2*(CNT-1)]) * b laid end to end; carry -> acc[2*CNT]
template <int CNT>
FAB_D void ptx_mad_chain(uint32_t* acc, const uint32_t* a, uint32_t b)
{
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(b));
#pragma unroll
    for (int k = 1; k < CNT; k++)
        asm volatile("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[2 * k]), "+r"(acc[2 * k + 1]) : "r"(a[2 * k]), "r"(b));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(acc[2 * CNT]));
}

// t[0..15] = a * a with 36 products instead of 64: the 28 off-diagonal products a_i a_j (i < j) are accumulated once
// (same even/odd 64-bit column scheme as mul_8x8), doubled, and the 8 squares a_i^2 are added in one carry chain.
FAB_D void sqr_8(uint32_t* t, const uint32_t* a)
{
    uint32_t E[18], O[18];
#pragma unroll
    for (int i = 0; i < 18; i++) { E[i] = 0; O[i] = 0; }
    // row i multiplies a_i by a_j, j > i.  Product a_i a_j sits at word i+j: even -> E[i+j], odd -> O[i+j-1].
    // j = i+1, i+3, ... (i+j odd) and j = i+2, i+4, ... (i+j even) are each one chain of consecutive 64-bit slots.
    ptx_mad_chain<4>(O + 0, a + 1, a[0]);   // j = 1,3,5,7  words 1,3,5,7
    ptx_mad_chain<3>(E + 2, a + 2, a[0]);   // j = 2,4,6    words 2,4,6
    ptx_mad_chain<3>(O + 2, a + 2, a[1]);   // j = 2,4,6    words 3,5,7
    ptx_mad_chain<3>(E + 4, a + 3, a[1]);   // j = 3,5,7    words 4,6,8
    ptx_mad_chain<3>(O + 4, a + 3, a[2]);   // j = 3,5,7    words 5,7,9
    ptx_mad_chain<2>(E + 6, a + 4, a[2]);   // j = 4,6      words 6,8
    ptx_mad_chain<2>(O + 6, a + 4, a[3]);   // j = 4,6      words 7,9
    ptx_mad_chain<2>(E + 8, a + 5, a[3]);   // j = 5,7      words 8,10
    ptx_mad_chain<2>(O + 8, a + 5, a[4]);   // j = 5,7      words 9,11
    ptx_mad_chain<1>(E + 10, a + 6, a[4]);  // j = 6        word 10
    ptx_mad_chain<1>(O + 10, a + 6, a[5]);  // j = 6        word 11
    ptx_mad_chain<1>(E + 12, a + 7, a[5]);  // j = 7        word 12
    ptx_mad_chain<1>(O + 12, a + 7, a[6]);  // j = 7        word 13
    // merge: off = E + (O << 32)   (word 0 is empty, word 15 at most a carry)
    uint32_t off[16];
    off[0] = 0;
    off[1] = O[0];
    off[2] = ptx_add_cc(E[2], O[1]);
#pragma unroll
    for (int i = 3; i < 16; i++) off[i] = ptx_addc_cc(E[i], O[i - 1]);
    // double
#pragma unroll
    for (int i = 15; i > 0; i--) off[i] = (off[i] << 1) | (off[i - 1] >> 31);
    // add the squares: a_i^2 at words 2i, 2i+1, one chain across all 16 words
    t[0] = a[0] * a[0];
    asm volatile("mad.hi.cc.u32 %0, %1, %1, %2;" : "=r"(t[1]) : "r"(a[0]), "r"(off[1]));
#pragma unroll
    for (int i = 1; i < 8; i++) {
        asm volatile("madc.lo.cc.u32 %0, %1, %1, %2;" : "=r"(t[2 * i]) : "r"(a[i]), "r"(off[2 * i]));
        asm volatile("madc.hi.cc.u32 %0, %1, %1, %2;" : "=r"(t[2 * i + 1]) : "r"(a[i]), "r"(off[2 * i + 1]));
    }
}

#ifndef FAB_SQR
#define FAB_SQR 1
#endif
FAB_MUL_ATTR u256 fe_sqr_dev(u256 a)
{
    uint32_t t[16];
    sqr_8(t, a.v);
    return fe_reduce(t);
}

// Two independent products in one routine: the scheduler interleaves the two carry-chain streams, which raises the
// issue rate of a single warp (at 64k-signature batches there are only ~3.5 warps per scheduler to hide latencies with).
struct u256x2 { u256 a, b; };
#ifndef FAB_MUL2
#define FAB_MUL2 1
#endif
FAB_MUL_ATTR u256x2 fe_mul2_dev(u256 a0, u256 b0, u256 a1, u256 b1)
{
    uint32_t t0[16], t1[16];
    mul_8x8(t0, a0.v, b0.v);
    mul_8x8(t1, a1.v, b1.v);
    u256x2 r;
    r.a = fe_reduce(t0);
    r.b = fe_reduce(t1);
    return r;
}

FAB_D u256 fe_add_dev(const u256& a, const u256& b)
{
    uint32_t s[8];
    s[0] = ptx_add_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) s[i] = ptx_addc_cc(a.v[i], b.v[i]);
    const uint32_t top = ptx_addc(0, 0);
    u256 d;
    d.v[0] = ptx_sub_cc(s[0], 0xffffffffu);
    d.v[1] = ptx_subc_cc(s[1], 0xffffffffu);
    d.v[2] = ptx_subc_cc(s[2], 0xffffffffu);
    d.v[3] = ptx_subc_cc(s[3], 0u);
    d.v[4] = ptx_subc_cc(s[4], 0u);
    d.v[5] = ptx_subc_cc(s[5], 0u);
    d.v[6] = ptx_subc_cc(s[6], 1u);
    d.v[7] = ptx_subc_cc(s[7], 0xffffffffu);
    const uint32_t brw = ptx_subc(top, 0);
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = brw ? s[i] : d.v[i];
    return r;
}

FAB_D u256 fe_sub_dev(const u256& a, const u256& b)
{
    u256 d;
    d.v[0] = ptx_sub_cc(a.v[0], b.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) d.v[i] = ptx_subc_cc(a.v[i], b.v[i]);
    const uint32_t m = ptx_subc(0, 0);        // 0xffffffff on borrow
    u256 r;
    r.v[0] = ptx_add_cc(d.v[0], m);
    r.v[1] = ptx_addc_cc(d.v[1], m);
    r.v[2] = ptx_addc_cc(d.v[2], m);
    r.v[3] = ptx_addc_cc(d.v[3], 0);
    r.v[4] = ptx_addc_cc(d.v[4], 0);
    r.v[5] = ptx_addc_cc(d.v[5], 0);
    r.v[6] = ptx_addc_cc(d.v[6], m & 1u);
    r.v[7] = ptx_addc(d.v[7], m);
    return r;
}

// Montgomery multiplication mod n (generic modulus): T = a*b; M = (T mod 2^256) * n' mod 2^256; r = (T + M*n) / 2^256
FAB_D u256 sc_mul_dev(const u256& a, const u256& b)
{
    uint32_t t[16], mm[16], u[16];
    mul_8x8(t, a.v, b.v);
    const u256 np = sc_nprime();
    const u256 n = sc_n();
    mul_8x8(mm, t, np.v);                    // only mm[0..7] is used
    mul_8x8(u, mm, n.v);
    // low halves cancel to zero by construction; the carry out of them is 1 unless t_lo == 0
    (void)ptx_add_cc(t[0], u[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) (void)ptx_addc_cc(t[i], u[i]);
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) s[i] = ptx_addc_cc(t[8 + i], u[8 + i]);
    const uint32_t top = ptx_addc(0, 0);
    u256 d;
    d.v[0] = ptx_sub_cc(s[0], n.v[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) d.v[i] = ptx_subc_cc(s[i], n.v[i]);
    const uint32_t brw = ptx_subc(top, 0);
    u256 r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = brw ? s[i] : d.v[i];
    return r;
}

#endif  // __CUDA_ARCH__

// ======================================================================================================
//  Host restatement of the same primitives (tests/host_sim only).
// ======================================================================================================
#if !defined(__CUDA_ARCH__)
namespace hostimpl {
inline void mul_full(uint32_t* t, const uint32_t* a, const uint32_t* b)
{
    for (int i = 0; i < 16; i++) t[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) {
            uint64_t cur = (uint64_t)a[j] * b[i] + t[i + j] + carry;
            t[i + j] = (uint32_t)cur; carry = cur >> 32;
        }
        t[i + 8] = (uint32_t)carry;
    }
}
// generic word-serial Montgomery reduction: t[0..15] (+ implicit top) -> [0, mod)
inline u256 mont_reduce(uint32_t* t, const uint32_t* mod, uint32_t m0inv)
{
    uint32_t top = 0;
    for (int i = 0; i < 8; i++) {
        uint32_t m = t[i] * m0inv;
        uint64_t carry = 0;
        for (int j = 0; j < 8; j++) {
            uint64_t cur = (uint64_t)m * mod[j] + t[i + j] + carry;
            t[i + j] = (uint32_t)cur; carry = cur >> 32;
        }
        for (int k = i + 8; k < 16 && carry; k++) { uint64_t cur = (uint64_t)t[k] + carry; t[k] = (uint32_t)cur; carry = cur >> 32; }
        top += (uint32_t)carry;
    }
    u256 r, d; uint64_t brw = 0;
    for (int i = 0; i < 8; i++) { r.v[i] = t[8 + i]; uint64_t cur = (uint64_t)r.v[i] - mod[i] - brw; d.v[i] = (uint32_t)cur; brw = (cur >> 32) & 1; }
    bool ge = (top != 0) || (brw == 0);
    return ge ? d : r;
}
}  // namespace hostimpl
#endif

// ---- public field API (dispatches to the device PTX or the host restatement) ----
FAB_HD u256 fe_mul(const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    return fe_mul_dev(a, b);
#else
    uint32_t t[16]; const uint32_t p[8] = FAB_P_LIMBS;
    hostimpl::mul_full(t, a.v, b.v);
    return hostimpl::mont_reduce(t, p, 1u);
#endif
}
// r0 = a0*b0, r1 = a1*b1 (independent)
FAB_HD void fe_mul2(const u256& a0, const u256& b0, const u256& a1, const u256& b1, u256& r0, u256& r1)
{
#if defined(__CUDA_ARCH__) && FAB_MUL2
    const u256x2 r = fe_mul2_dev(a0, b0, a1, b1);
    r0 = r.a; r1 = r.b;
#else
    r0 = fe_mul(a0, b0); r1 = fe_mul(a1, b1);
#endif
}
FAB_HD u256 fe_sqr(const u256& a)
{
#if defined(__CUDA_ARCH__) && FAB_SQR
    return fe_sqr_dev(a);
#else
    return fe_mul(a, a);
#endif
}
// The same three operations expanded in place (INL = true, device only): for a loop whose body is ONE point addition the
// expansion fits the instruction cache, saves the argument moves of 8 calls and lets the scheduler overlap independent
// products.  INL = false (and every host build) goes through the shared out-of-line copies above.
template <bool INL> FAB_HD u256 fe_mul_t(const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    if (INL) { uint32_t t[16]; mul_8x8(t, a.v, b.v); return fe_reduce(t); }
#endif
    return fe_mul(a, b);
}
template <bool INL> FAB_HD u256 fe_sqr_t(const u256& a)
{
#if defined(__CUDA_ARCH__)
    if (INL) { uint32_t t[16]; sqr_8(t, a.v); return fe_reduce(t); }
#endif
    return fe_sqr(a);
}
template <bool INL> FAB_HD void fe_mul2_t(const u256& a0, const u256& b0, const u256& a1, const u256& b1, u256& r0, u256& r1)
{
#if defined(__CUDA_ARCH__)
    if (INL) {
        uint32_t t0[16], t1[16];
        mul_8x8(t0, a0.v, b0.v);
        mul_8x8(t1, a1.v, b1.v);
        r0 = fe_reduce(t0); r1 = fe_reduce(t1);
        return;
    }
#endif
    fe_mul2(a0, b0, a1, b1, r0, r1);
}
FAB_HD u256 fe_add(const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    return fe_add_dev(a, b);
#else
    const uint32_t p[8] = FAB_P_LIMBS; u256 s, d; uint64_t c = 0, brw = 0;
    for (int i = 0; i < 8; i++) { uint64_t cur = (uint64_t)a.v[i] + b.v[i] + c; s.v[i] = (uint32_t)cur; c = cur >> 32; }
    for (int i = 0; i < 8; i++) { uint64_t cur = (uint64_t)s.v[i] - p[i] - brw; d.v[i] = (uint32_t)cur; brw = (cur >> 32) & 1; }
    return (c || !brw) ? d : s;
#endif
}
FAB_HD u256 fe_sub(const u256& a, const u256& b)
{
#if defined(__CUDA_ARCH__)
    return fe_sub_dev(a, b);
#else
    const uint32_t p[8] = FAB_P_LIMBS; u256 d, r; uint64_t brw = 0, c = 0;
    for (int i = 0; i < 8; i++) { uint64_t cur = (uint64_t)a.v[i] - b.v[i] - brw; d.v[i] = (uint32_t)cur; brw = (cur >> 32) & 1; }
    if (!brw) return d;
    for (int i = 0; i < 8; i++) { uint64_t cur = (uint64_t)d.v[i] + p[i] + c; r.v[i] = (uint32_t)cur; c = cur >> 32; }
    return r;
#endif
}
FAB_HD u256 fe_neg(const u256& a) { return fe_sub(u256_zero(), a); }
FAB_HD u256 fe_dbl(const u256& a) { return fe_add(a, a); }
FAB_HD u256 fe_to_mont(const u256& a) { return fe_mul(a, fe_r2()); }                 // a < p
FAB_HD u256 fe_from_mont(const u256& a) { u256 one = u256_const(1, 0, 0, 0, 0, 0, 0, 0); return fe_mul(a, one); }

// a^(p-2) mod p (Montgomery in/out).  p-2 is a constant, so the branch below is uniform across a warp.
FAB_HD u256 fe_inv(const u256& a)
{
    const uint32_t e[8] = {0xfffffffdu, 0xffffffffu, 0xffffffffu, 0x00000000u, 0x00000000u, 0x00000000u, 0x00000001u, 0xffffffffu};
    u256 r = fe_one();
    for (int i = 255; i >= 0; i--) {
        r = fe_sqr(r);
        if ((e[i >> 5] >> (i & 31)) & 1u) r = fe_mul(r, a);
    }
    return r;
}

// ---- scalar field mod n ----
FAB_HD u256 sc_mul(const u256& a, const u256& b)      // Montgomery product a*b*2^-256 mod n, inputs < n
{
#if defined(__CUDA_ARCH__)
    return sc_mul_dev(a, b);
#else
    uint32_t t[16]; const uint32_t n[8] = FAB_N_LIMBS;
    hostimpl::mul_full(t, a.v, b.v);
    return hostimpl::mont_reduce(t, n, 0xee00bc4fu);
#endif
}
// x - n if x >= n (x < 2^256 < 2n, so one conditional subtraction reduces fully)
FAB_HD u256 sc_reduce_once(const u256& x)
{
    const u256 n = sc_n();
    u256 d; uint32_t brw = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t xi = x.v[i], ni = n.v[i];
        uint32_t t1 = xi - ni; uint32_t b1 = xi < ni;
        uint32_t t2 = t1 - brw; uint32_t b2 = t1 < brw;
        d.v[i] = t2; brw = b1 | b2;
    }
    return u256_select(brw != 0, x, d);
}
// s^-1 mod n, plain in -> Montgomery out (s in [1, n-1]).  Fermat with the constant exponent n-2.
FAB_HD u256 sc_inv_to_mont(const u256& s)
{
    const uint32_t e[8] = {0xfc63254fu, 0xf3b9cac2u, 0xa7179e84u, 0xbce6faadu, 0xffffffffu, 0xffffffffu, 0x00000000u, 0xffffffffu};
    const u256 sm = sc_mul(s, sc_r2());
    u256 r = sc_one();
    for (int i = 255; i >= 0; i--) {
        r = sc_mul(r, r);
        if ((e[i >> 5] >> (i & 31)) & 1u) r = sc_mul(r, sm);
    }
    return r;
}

}  // namespace fabgpu
