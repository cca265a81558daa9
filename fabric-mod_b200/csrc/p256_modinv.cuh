// s^-1 mod n by Bernstein-Yang division steps ("safegcd"), variable time, batches of 30 steps.
//
// Replaces the Fermat ladder (s^(n-2): ~335 Montgomery products mod n, ~134 k instructions per signature -- 14 % of the
// first kernel) with ~18 batches x ~550 instructions.  Variable time is fine: everything in a verification is public.
//
// Numbers are 9 limbs of 30 bits (value = sum v[i] << 30 i); v[8] carries the sign.  The 2x2 transition matrices have
// |entries| <= 2^30, so every limb product fits an int64 accumulator without carries: on sm_100a these are plain
// IMAD.WIDE (full rate), unlike the carry-chained products of the 32-bit-limb multiplier.
//
// Algorithm (eprint 2019/266, "delta" form):   divstep(delta, f, g) =
//     (1 - delta, g, (g - f) / 2)            if delta > 0 and g odd
//     (1 + delta, f, (g + (g mod 2) f) / 2)  otherwise
// starting from (1, n, s); after enough steps g = 0 and f = +-gcd = +-1.  Alongside, (d, e) with f = d*s, g = e*s (mod n)
// are transformed by the same matrices, so at the end s^-1 = f * d (mod n).
#pragma once
#include "p256_fe.cuh"

namespace fabgpu {

struct s30x9 { int32_t v[9]; };

#define FAB_M30 0x3fffffff

FAB_HD s30x9 s30_from_u256(const u256& a)
{
    s30x9 r;
    r.v[0] = (int32_t)(a.v[0] & FAB_M30);
    r.v[1] = (int32_t)(((a.v[0] >> 30) | (a.v[1] << 2)) & FAB_M30);
    r.v[2] = (int32_t)(((a.v[1] >> 28) | (a.v[2] << 4)) & FAB_M30);
    r.v[3] = (int32_t)(((a.v[2] >> 26) | (a.v[3] << 6)) & FAB_M30);
    r.v[4] = (int32_t)(((a.v[3] >> 24) | (a.v[4] << 8)) & FAB_M30);
    r.v[5] = (int32_t)(((a.v[4] >> 22) | (a.v[5] << 10)) & FAB_M30);
    r.v[6] = (int32_t)(((a.v[5] >> 20) | (a.v[6] << 12)) & FAB_M30);
    r.v[7] = (int32_t)(((a.v[6] >> 18) | (a.v[7] << 14)) & FAB_M30);
    r.v[8] = (int32_t)(a.v[7] >> 16);
    return r;
}

// limbs must be normalised: v[0..7] in [0, 2^30), 0 <= value < 2^256
FAB_HD u256 s30_to_u256(const s30x9& a)
{
    u256 r;
    const uint32_t* v = reinterpret_cast<const uint32_t*>(a.v);
    r.v[0] = v[0] | (v[1] << 30);
    r.v[1] = (v[1] >> 2) | (v[2] << 28);
    r.v[2] = (v[2] >> 4) | (v[3] << 26);
    r.v[3] = (v[3] >> 6) | (v[4] << 24);
    r.v[4] = (v[4] >> 8) | (v[5] << 22);
    r.v[5] = (v[5] >> 10) | (v[6] << 20);
    r.v[6] = (v[6] >> 12) | (v[7] << 18);
    r.v[7] = (v[7] >> 14) | (v[8] << 16);
    return r;
}

// The modulus in 30-bit limbs and its inverse mod 2^30: the group order n (s^-1 of a signature) and the field prime p
// (the shared inversion of the batch-affine point additions, ecdsa_batchaffine.cuh).
struct ModN {
    static FAB_HD s30x9 m()
    {
        s30x9 r;
        r.v[0] = 0x3c632551; r.v[1] = 0x0ee72b0b; r.v[2] = 0x3179e84f; r.v[3] = 0x39beab69; r.v[4] = 0x3fffffbc;
        r.v[5] = 0x3fffffff; r.v[6] = 0x00000fff; r.v[7] = 0x3fffc000; r.v[8] = 0x0000ffff;
        return r;
    }
    static constexpr uint32_t inv30 = 0x11ff43b1u;
};
struct ModP {
    static FAB_HD s30x9 m()
    {
        s30x9 r;
        r.v[0] = 0x3fffffff; r.v[1] = 0x3fffffff; r.v[2] = 0x3fffffff; r.v[3] = 0x0000003f; r.v[4] = 0x00000000;
        r.v[5] = 0x00000000; r.v[6] = 0x00001000; r.v[7] = 0x3fffc000; r.v[8] = 0x0000ffff;
        return r;
    }
    static constexpr uint32_t inv30 = 0x3fffffffu;      // p == -1 (mod 2^96)
};
FAB_HD s30x9 s30_n() { return ModN::m(); }

// 30 division steps on the low words; returns the new delta and the matrix t = (u, v, q, r) with
//   2^30 * (f', g') = (u f + v g, q f + r g).
FAB_HD int32_t divsteps30(int32_t delta, uint32_t f, uint32_t g, int32_t* t)
{
    int32_t u = 1, v = 0, q = 0, r = 1;
#pragma unroll 6
    for (int i = 0; i < 30; i++) {
        const bool odd = (g & 1u) != 0;
        if (delta > 0 && odd) {
            const uint32_t tf = f; f = g; g = 0u - tf;
            const int32_t tu = u, tv = v; u = q; v = r; q = -tu; r = -tv;
            delta = -delta;
        }
        if (odd) { g += f; q += u; r += v; }
        delta += 1;
        g >>= 1;
        u += u; v += v;
    }
    t[0] = u; t[1] = v; t[2] = q; t[3] = r;
    return delta;
}

// (f, g) <- (u f + v g, q f + r g) / 2^30, exact
FAB_HD void update_fg30(s30x9& f, s30x9& g, const int32_t* t)
{
    const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
    int64_t cf = u * f.v[0] + v * g.v[0];
    int64_t cg = q * f.v[0] + r * g.v[0];
    cf >>= 30; cg >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        const int64_t fi = f.v[i], gi = g.v[i];
        cf += u * fi + v * gi;
        cg += q * fi + r * gi;
        f.v[i - 1] = (int32_t)(cf & FAB_M30); cf >>= 30;
        g.v[i - 1] = (int32_t)(cg & FAB_M30); cg >>= 30;
    }
    f.v[8] = (int32_t)cf;
    g.v[8] = (int32_t)cg;
}

// (d, e) <- (u d + v e, q d + r e) / 2^30 mod m, keeping d, e in (-2m, m)
template <class MOD> FAB_HD void update_de30(s30x9& d, s30x9& e, const int32_t* t)
{
    const s30x9 m = MOD::m();
    const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;          // all-ones when negative
    int32_t md = (u & sd) + (v & se);                             // adds n to a negative d / e before the linear map
    int32_t me = (q & sd) + (r & se);
    int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0];
    int64_t ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
    // choose the multiples of n that clear the low 30 bits
    md -= (int32_t)((MOD::inv30 * (uint32_t)cd + (uint32_t)md) & FAB_M30);
    me -= (int32_t)((MOD::inv30 * (uint32_t)ce + (uint32_t)me) & FAB_M30);
    cd += (int64_t)m.v[0] * md;
    ce += (int64_t)m.v[0] * me;
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cd += (int64_t)u * d.v[i] + (int64_t)v * e.v[i] + (int64_t)m.v[i] * md;
        ce += (int64_t)q * d.v[i] + (int64_t)r * e.v[i] + (int64_t)m.v[i] * me;
        d.v[i - 1] = (int32_t)(cd & FAB_M30); cd >>= 30;
        e.v[i - 1] = (int32_t)(ce & FAB_M30); ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}

// a in (-2m, m), optionally negated, brought to [0, m)
template <class MOD> FAB_HD s30x9 normalize30(const s30x9& a, bool negate)
{
    const s30x9 m = MOD::m();
    s30x9 r = a;
    if (negate) {
#pragma unroll
        for (int i = 0; i < 9; i++) r.v[i] = -r.v[i];
    }
    // carry-propagate into canonical limbs (v[8] keeps the sign)
    for (int pass = 0; pass < 3; pass++) {
        int32_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { r.v[i] += c; c = r.v[i] >> 30; r.v[i] &= FAB_M30; }
        r.v[8] += c;
        if (r.v[8] < 0) {                     // negative: add n
#pragma unroll
            for (int i = 0; i < 9; i++) r.v[i] += m.v[i];
        } else {
            // >= n ?  compare from the top
            bool ge = true, decided = false;
#pragma unroll
            for (int i = 8; i >= 0; i--) {
                if (!decided && r.v[i] != m.v[i]) { ge = r.v[i] > m.v[i]; decided = true; }
            }
            if (ge) {
#pragma unroll
                for (int i = 0; i < 9; i++) r.v[i] -= m.v[i];
            }
        }
    }
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i] += c; c = r.v[i] >> 30; r.v[i] &= FAB_M30; }
    r.v[8] += c;
    return r;
}

// s in [1, m-1]  ->  s^-1 mod m (plain integer, not Montgomery)
template <class MOD> FAB_HD u256 inv_safegcd(const u256& s)
{
    s30x9 f = MOD::m(), g = s30_from_u256(s);
    s30x9 d, e;
#pragma unroll
    for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    int32_t delta = 1;
    for (int it = 0; it < 25; it++) {          // 25 * 30 = 750 >= 741, the proven bound for 256-bit inputs
        int32_t t[4];
        delta = divsteps30(delta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), t);
        update_de30<MOD>(d, e, t);
        update_fg30(f, g, t);
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) nz |= g.v[i];
        if (nz == 0) break;
    }
    // f = +-1 ; s^-1 = f * d
    const s30x9 r = normalize30<MOD>(d, f.v[8] < 0);
    return s30_to_u256(r);
}

FAB_HD u256 sc_inv_safegcd(const u256& s) { return inv_safegcd<ModN>(s); }

// Field inversion by the same division steps: a R in -> a^-1 R out (Montgomery form on both sides).  The plain inverse of the
// residue a R is a^-1 R^-1; one Montgomery product with R^3 restores the domain.  ~5x cheaper than the Fermat ladder fe_inv.
FAB_HD u256 fe_inv_safegcd(const u256& a)
{
    const u256 r3 = u256_const(0x0000000au, 0xfffffffdu, 0xfffffff7u, 0xffffffedu, 0xfffffffcu, 0x00000005u, 0x00000001u, 0x00000018u);
    return fe_mul(inv_safegcd<ModP>(a), r3);
}

// plain in -> Montgomery out, the contract of sc_inv_to_mont
FAB_HD u256 sc_inv_to_mont_safegcd(const u256& s)
{
    return sc_mul(sc_inv_safegcd(s), sc_r2());
}

// ---- one compiled copy for both moduli (device code size): the modulus comes from a small table instead of a template ----
// row 0: n, row 1: p; [0..8] = 30-bit limbs, [9] = modulus^-1 mod 2^30
#define FAB_MODTAB_INIT {{0x3c632551, 0x0ee72b0b, 0x3179e84f, 0x39beab69, 0x3fffffbc, 0x3fffffff, 0x00000fff, 0x3fffc000, 0x0000ffff, 0x11ff43b1}, \
                         {0x3fffffff, 0x3fffffff, 0x3fffffff, 0x0000003f, 0x00000000, 0x00000000, 0x00001000, 0x3fffc000, 0x0000ffff, 0x3fffffff}}
#if defined(__CUDACC__)
__constant__ int32_t c_modtab[2][10] = FAB_MODTAB_INIT;
#endif
static const int32_t h_modtab[2][10] = FAB_MODTAB_INIT;
#if defined(__CUDA_ARCH__)
#define FAB_MODTAB c_modtab
#else
#define FAB_MODTAB h_modtab
#endif
struct ModRT {
    int which;
    FAB_HD s30x9 m() const { s30x9 r; for (int i = 0; i < 9; i++) r.v[i] = FAB_MODTAB[which][i]; return r; }
    FAB_HD uint32_t inv30() const { return (uint32_t)FAB_MODTAB[which][9]; }
};

FAB_HD void update_de30_rt(s30x9& d, s30x9& e, const int32_t* t, const s30x9& m, uint32_t minv)
{
    const int32_t u = t[0], v = t[1], q = t[2], r = t[3];
    const int32_t sd = d.v[8] >> 31, se = e.v[8] >> 31;
    int32_t md = (u & sd) + (v & se);
    int32_t me = (q & sd) + (r & se);
    int64_t cd = (int64_t)u * d.v[0] + (int64_t)v * e.v[0];
    int64_t ce = (int64_t)q * d.v[0] + (int64_t)r * e.v[0];
    md -= (int32_t)((minv * (uint32_t)cd + (uint32_t)md) & FAB_M30);
    me -= (int32_t)((minv * (uint32_t)ce + (uint32_t)me) & FAB_M30);
    cd += (int64_t)m.v[0] * md;
    ce += (int64_t)m.v[0] * me;
    cd >>= 30; ce >>= 30;
#pragma unroll
    for (int i = 1; i < 9; i++) {
        cd += (int64_t)u * d.v[i] + (int64_t)v * e.v[i] + (int64_t)m.v[i] * md;
        ce += (int64_t)q * d.v[i] + (int64_t)r * e.v[i] + (int64_t)m.v[i] * me;
        d.v[i - 1] = (int32_t)(cd & FAB_M30); cd >>= 30;
        e.v[i - 1] = (int32_t)(ce & FAB_M30); ce >>= 30;
    }
    d.v[8] = (int32_t)cd;
    e.v[8] = (int32_t)ce;
}

FAB_HD s30x9 normalize30_rt(const s30x9& a, bool negate, const s30x9& m)
{
    s30x9 r = a;
    if (negate) {
#pragma unroll
        for (int i = 0; i < 9; i++) r.v[i] = -r.v[i];
    }
    for (int pass = 0; pass < 3; pass++) {
        int32_t c = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { r.v[i] += c; c = r.v[i] >> 30; r.v[i] &= FAB_M30; }
        r.v[8] += c;
        if (r.v[8] < 0) {
#pragma unroll
            for (int i = 0; i < 9; i++) r.v[i] += m.v[i];
        } else {
            bool ge = true, decided = false;
#pragma unroll
            for (int i = 8; i >= 0; i--) {
                if (!decided && r.v[i] != m.v[i]) { ge = r.v[i] > m.v[i]; decided = true; }
            }
            if (ge) {
#pragma unroll
                for (int i = 0; i < 9; i++) r.v[i] -= m.v[i];
            }
        }
    }
    int32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) { r.v[i] += c; c = r.v[i] >> 30; r.v[i] &= FAB_M30; }
    r.v[8] += c;
    return r;
}

// a in [1, m-1] -> a^-1 mod m, m = n (which = 0) or p (which = 1)
FAB_HD u256 inv_safegcd_rt(const u256& a, int which)
{
    ModRT mod; mod.which = which;
    const s30x9 m = mod.m();
    const uint32_t minv = mod.inv30();
    s30x9 f = m, g = s30_from_u256(a);
    s30x9 d, e;
#pragma unroll
    for (int i = 0; i < 9; i++) { d.v[i] = 0; e.v[i] = 0; }
    e.v[0] = 1;
    int32_t delta = 1;
    for (int it = 0; it < 25; it++) {
        int32_t t[4];
        delta = divsteps30(delta, (uint32_t)f.v[0] | ((uint32_t)f.v[1] << 30), (uint32_t)g.v[0] | ((uint32_t)g.v[1] << 30), t);
        update_de30_rt(d, e, t, m, minv);
        update_fg30(f, g, t);
        int32_t nz = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) nz |= g.v[i];
        if (nz == 0) break;
    }
    return s30_to_u256(normalize30_rt(d, f.v[8] < 0, m));
}

}  // namespace fabgpu
