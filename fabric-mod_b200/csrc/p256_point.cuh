// P-256 group law in Jacobian coordinates over the Montgomery field of p256_fe.cuh.
// Infinity is Z == 0.  All routines are complete: P = Q, P = -Q and infinity operands are handled, because
// the reference's verifier (Go crypto/elliptic CombinedMult behind bccsp/sw/ecdsa.go:56) is complete and the
// bitmask must match it on adversarial inputs too.
#pragma once
#include "p256_fe.cuh"

namespace fabgpu {

struct jac { u256 X, Y, Z; };     // Jacobian, Montgomery form
struct aff { u256 x, y; };        // affine, Montgomery form (never infinity)

FAB_HD jac jac_infinity() { jac r; r.X = fe_one(); r.Y = fe_one(); r.Z = u256_zero(); return r; }
FAB_HD bool jac_is_infinity(const jac& p) { return u256_is_zero(p.Z); }
FAB_HD jac jac_from_aff(const aff& a) { jac r; r.X = a.x; r.Y = a.y; r.Z = fe_one(); return r; }
FAB_HD jac jac_neg(const jac& p) { jac r = p; r.Y = fe_neg(p.Y); return r; }

// dbl-2001-b (a = -3): 3M + 5S.  Infinity in -> infinity out (Z3 = 2*Y1*Z1 = 0).
FAB_HD jac jac_double(const jac& p)
{
    const u256 delta = fe_sqr(p.Z);
    const u256 gamma = fe_sqr(p.Y);
    const u256 beta = fe_mul(p.X, gamma);
    const u256 t0 = fe_sub(p.X, delta);
    const u256 t1 = fe_add(p.X, delta);
    const u256 t2 = fe_mul(t0, t1);
    const u256 alpha = fe_add(fe_dbl(t2), t2);
    const u256 beta4 = fe_dbl(fe_dbl(beta));
    jac r;
    r.X = fe_sub(fe_sqr(alpha), fe_dbl(beta4));
    const u256 yz = fe_add(p.Y, p.Z);
    r.Z = fe_sub(fe_sub(fe_sqr(yz), gamma), delta);
    const u256 g2 = fe_sqr(gamma);
    const u256 g8 = fe_dbl(fe_dbl(fe_dbl(g2)));
    r.Y = fe_sub(fe_mul(alpha, fe_sub(beta4, r.X)), g8);
    return r;
}

// add-2007-bl: 11M + 5S, with the exceptional cases resolved explicitly.
FAB_HD jac jac_add(const jac& p, const jac& q)
{
    if (jac_is_infinity(p)) return q;
    if (jac_is_infinity(q)) return p;
    const u256 z1z1 = fe_sqr(p.Z);
    const u256 z2z2 = fe_sqr(q.Z);
    const u256 u1 = fe_mul(p.X, z2z2);
    const u256 u2 = fe_mul(q.X, z1z1);
    const u256 s1 = fe_mul(fe_mul(p.Y, q.Z), z2z2);
    const u256 s2 = fe_mul(fe_mul(q.Y, p.Z), z1z1);
    const u256 h = fe_sub(u2, u1);
    const u256 rr0 = fe_sub(s2, s1);
    if (u256_is_zero(h)) {
        if (u256_is_zero(rr0)) return jac_double(p);   // same point
        return jac_infinity();                          // opposite points
    }
    const u256 h2 = fe_dbl(h);
    const u256 i = fe_sqr(h2);
    const u256 j = fe_mul(h, i);
    const u256 rr = fe_dbl(rr0);
    const u256 v = fe_mul(u1, i);
    jac r;
    r.X = fe_sub(fe_sub(fe_sqr(rr), j), fe_dbl(v));
    r.Y = fe_sub(fe_mul(rr, fe_sub(v, r.X)), fe_dbl(fe_mul(s1, j)));
    const u256 zs = fe_add(p.Z, q.Z);
    r.Z = fe_mul(fe_sub(fe_sub(fe_sqr(zs), z1z1), z2z2), h);
    return r;
}

// madd-2007-bl (Z2 = 1): 7M + 4S.
FAB_HD jac jac_add_aff(const jac& p, const aff& q)
{
    if (jac_is_infinity(p)) return jac_from_aff(q);
    const u256 z1z1 = fe_sqr(p.Z);
    const u256 u2 = fe_mul(q.x, z1z1);
    const u256 s2 = fe_mul(fe_mul(q.y, p.Z), z1z1);
    const u256 h = fe_sub(u2, p.X);
    const u256 rr0 = fe_sub(s2, p.Y);
    if (u256_is_zero(h)) {
        if (u256_is_zero(rr0)) return jac_double(p);
        return jac_infinity();
    }
    const u256 hh = fe_sqr(h);
    const u256 i = fe_dbl(fe_dbl(hh));
    const u256 j = fe_mul(h, i);
    const u256 rr = fe_dbl(rr0);
    const u256 v = fe_mul(p.X, i);
    jac r;
    r.X = fe_sub(fe_sub(fe_sqr(rr), j), fe_dbl(v));
    r.Y = fe_sub(fe_mul(rr, fe_sub(v, r.X)), fe_dbl(fe_mul(p.Y, j)));
    const u256 zh = fe_add(p.Z, h);
    r.Z = fe_sub(fe_sub(fe_sqr(zh), z1z1), hh);
    return r;
}

// y^2 == x^3 - 3x + b  (x, y already in Montgomery form and < p)
FAB_HD bool aff_on_curve(const aff& a)
{
    const u256 x2 = fe_sqr(a.x);
    const u256 x3 = fe_mul(x2, a.x);
    const u256 x_3 = fe_add(fe_dbl(a.x), a.x);
    const u256 rhs = fe_add(fe_sub(x3, x_3), fe_b_mont());
    return u256_eq(fe_sqr(a.y), rhs);
}

// Jacobian -> affine (one field inversion); only used when building tables. p must not be infinity.
FAB_HD aff jac_to_aff(const jac& p)
{
    const u256 zi = fe_inv(p.Z);
    const u256 zi2 = fe_sqr(zi);
    aff r;
    r.x = fe_mul(p.X, zi2);
    r.y = fe_mul(p.Y, fe_mul(zi2, zi));
    return r;
}

}  // namespace fabgpu
