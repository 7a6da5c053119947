// bn254 G1 point arithmetic for the MSM kernels (curve y^2 = x^3 + 3 over Fq, a = 0).
//
// Replaces the reference's Jacobian group layer on the MSM path
//   groups/group.hpp:153-217 (dbl), :219-322 (mixed_add), :324-448 (add), :450-534 (normalize)
// Values, not representations, are what the reference's callers observe (SURVEY.md §8 note 1:
// batched_scalar_multiplications normalises its outputs, pippenger's Jacobian result is only ever
// normalised or added by its callers), so the accumulators here use extended Jacobian "XYZZ"
// coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2): a mixed add is 8M + 2S = 10 field products
// against 11 for the reference's madd-2007-bl shape, and no z-doubling bookkeeping.  The two products of
// y3 = R (Q - x3) - y1 PPP share ONE Montgomery reduction (Fq::mul2: 200 wide multiply-adds instead of 2 x 136).
//
// Conventions: affine inputs follow the reference (x, y Montgomery limbs; infinity <=> bit 63 of
// the top y limb, group.hpp:133-151).  An XYZZ accumulator is infinity <=> ZZ == 0.
// All coordinates are kept lazily reduced in [0, 2p) (see bbg_field.cuh).
#pragma once
#include "bbg_field.cuh"

namespace bbg
{

struct affine_pt
{
    fe x, y;
};
struct xyzz_pt
{
    fe x, y, zz, zzz;
};

struct G1
{
    static BBG_HD bool affine_is_infinity(const affine_pt& p) { return (p.y.v[7] >> 31) != 0; }
    static BBG_HD void affine_set_infinity(affine_pt& p)
    {
        p.x = Fq::zero();
        p.y = Fq::zero();
        p.y.v[7] = 0x80000000u;
    }
    static BBG_HD bool is_infinity(const xyzz_pt& p) { return Fq::is_zero_raw(p.zz); }
    static BBG_HD xyzz_pt infinity()
    {
        xyzz_pt r;
        r.x = Fq::zero();
        r.y = Fq::zero();
        r.zz = Fq::zero();
        r.zzz = Fq::zero();
        return r;
    }
    static BBG_HD xyzz_pt from_affine(const affine_pt& p)
    {
        xyzz_pt r;
        r.x = p.x;
        r.y = p.y;
        r.zz = Fq::one();
        r.zzz = Fq::one();
        return r;
    }
    // (x, y) -> (x, -y) when negate != 0   (reference: conditional_negate_affine, group_impl_asm.tcc:70-153)
    static BBG_HD affine_pt cond_negate(const affine_pt& p, uint32_t negate)
    {
        affine_pt r = p;
        if (negate) r.y = Fq::neg(p.y);
        return r;
    }
    // phi(P) = (beta x, -y) = -lambda P: the odd entries of the reference's point table
    // (generate_pippenger_point_table, scalar_multiplication.cpp:131-140)
    static BBG_HD affine_pt endo_table_entry(const affine_pt& p)
    {
        affine_pt r;
        r.x = Fq::mul_full(p.x, Fq::constant([](int i) { return FqParams::CUBE(i); }));
        r.y = Fq::reduce(Fq::neg(p.y));
        // reference computes p - y, which maps y = 0 to p; y = 0 is not on the curve so this cannot occur
        return r;
    }

    // 2 * (affine p) in XYZZ  (EFD mdbl-2008-s-1, a = 0): 2M + 3S-ish
    static BBG_HD xyzz_pt dbl_affine(const affine_pt& p)
    {
        xyzz_pt r;
        fe U = Fq::dbl(p.y);
        fe V = Fq::sqr(U);
        fe W = Fq::mul(U, V);
        fe S = Fq::mul(p.x, V);
        fe xx = Fq::sqr(p.x);
        fe M = Fq::add(Fq::dbl(xx), xx);
        fe X3 = Fq::sub(Fq::sqr(M), Fq::dbl(S));
        r.x = X3;
        r.y = Fq::mul2(M, Fq::sub(S, X3), Fq::neg(W), p.y); // M (S - x3) - W y1 under one reduction
        r.zz = V;
        r.zzz = W;
        return r;
    }
    // 2 * p  (EFD dbl-2008-s-1, a = 0)
    static BBG_HD xyzz_pt dbl(const xyzz_pt& p)
    {
        if (is_infinity(p)) return p;
        xyzz_pt r;
        fe U = Fq::dbl(p.y);
        fe V = Fq::sqr(U);
        fe W = Fq::mul(U, V);
        fe S = Fq::mul(p.x, V);
        fe xx = Fq::sqr(p.x);
        fe M = Fq::add(Fq::dbl(xx), xx);
        fe X3 = Fq::sub(Fq::sqr(M), Fq::dbl(S));
        r.x = X3;
        r.y = Fq::mul2(M, Fq::sub(S, X3), Fq::neg(W), p.y); // M (S - x3) - W y1 under one reduction
        r.zz = Fq::mul(V, p.zz);
        r.zzz = Fq::mul(W, p.zzz);
        return r;
    }
    // acc + (affine q), q != infinity  (EFD madd-2008-s): 8M + 2S, two of the M under one reduction.
    // Exception paths mirror the reference's mixed_add (group.hpp:241-254, :311-320):
    // acc = infinity -> q;  q == acc -> double;  q == -acc -> infinity.
    static BBG_HD xyzz_pt madd(const xyzz_pt& acc, const affine_pt& q)
    {
        if (is_infinity(acc)) return from_affine(q);
        fe U2 = Fq::mul(q.x, acc.zz);
        fe S2 = Fq::mul(q.y, acc.zzz);
        fe P = Fq::sub(U2, acc.x);
        fe R = Fq::sub(S2, acc.y);
        if (Fq::is_zero(P))
        {
            if (Fq::is_zero(R)) return dbl_affine(q);
            return infinity();
        }
        fe PP = Fq::sqr(P);
        fe PPP = Fq::mul(P, PP);
        fe Q = Fq::mul(acc.x, PP);
        xyzz_pt r;
        fe X3 = Fq::sub(Fq::sub(Fq::sqr(R), PPP), Fq::dbl(Q));
        r.x = X3;
        r.y = Fq::mul2(R, Fq::sub(Q, X3), Fq::neg(acc.y), PPP); // R (Q - x3) - y1 PPP under one reduction
        r.zz = Fq::mul(acc.zz, PP);
        r.zzz = Fq::mul(acc.zzz, PPP);
        return r;
    }
    // a + b  (EFD add-2008-s): 12M + 2S, all exception paths (reference add: group.hpp:324-448)
    static BBG_HD xyzz_pt add(const xyzz_pt& a, const xyzz_pt& b)
    {
        if (is_infinity(a)) return b;
        if (is_infinity(b)) return a;
        fe U1 = Fq::mul(a.x, b.zz);
        fe U2 = Fq::mul(b.x, a.zz);
        fe S1 = Fq::mul(a.y, b.zzz);
        fe S2 = Fq::mul(b.y, a.zzz);
        fe P = Fq::sub(U2, U1);
        fe R = Fq::sub(S2, S1);
        if (Fq::is_zero(P))
        {
            if (Fq::is_zero(R)) return dbl(a);
            return infinity();
        }
        fe PP = Fq::sqr(P);
        fe PPP = Fq::mul(P, PP);
        fe Q = Fq::mul(U1, PP);
        xyzz_pt r;
        fe X3 = Fq::sub(Fq::sub(Fq::sqr(R), PPP), Fq::dbl(Q));
        r.x = X3;
        r.y = Fq::mul2(R, Fq::sub(Q, X3), Fq::neg(S1), PPP);
        r.zz = Fq::mul(Fq::mul(a.zz, b.zz), PP);
        r.zzz = Fq::mul(Fq::mul(a.zzz, b.zzz), PPP);
        return r;
    }
    // XYZZ -> the reference's normalised form: canonical affine x, y (z = fq::one implied);
    // infinity -> x = y = 0 with the infinity flag (group.hpp:450-469, :787-799)
    static BBG_HD affine_pt to_affine(const xyzz_pt& p)
    {
        affine_pt r;
        if (is_infinity(p))
        {
            affine_set_infinity(r);
            return r;
        }
        fe inv = Fq::invert(Fq::mul(p.zz, p.zzz));
        r.x = Fq::mul_full(p.x, Fq::mul(inv, p.zzz));
        r.y = Fq::mul_full(p.y, Fq::mul(inv, p.zz));
        return r;
    }
};

BBG_HD affine_pt load_affine(const void* p)
{
    affine_pt r;
    r.x = load_fe(p);
    r.y = load_fe((const char*)p + 32);
    return r;
}
BBG_HD void store_affine(void* p, const affine_pt& a)
{
    store_fe(p, a.x);
    store_fe((char*)p + 32, a.y);
}
BBG_HD xyzz_pt load_xyzz(const void* p)
{
    xyzz_pt r;
    r.x = load_fe(p);
    r.y = load_fe((const char*)p + 32);
    r.zz = load_fe((const char*)p + 64);
    r.zzz = load_fe((const char*)p + 96);
    return r;
}
BBG_HD void store_xyzz(void* p, const xyzz_pt& a)
{
    store_fe(p, a.x);
    store_fe((char*)p + 32, a.y);
    store_fe((char*)p + 64, a.zz);
    store_fe((char*)p + 96, a.zzz);
}
// the same for pointers known to be GLOBAL memory (32-byte aligned): one 256-bit access per coordinate (bbg_field.cuh)
BBG_HD affine_pt load_affine_const(const void* p) // read-only tables (the SRS)
{
    affine_pt r;
    r.x = load_fe_const(p);
    r.y = load_fe_const((const char*)p + 32);
    return r;
}
BBG_HD xyzz_pt load_xyzz_global(const void* p)
{
    xyzz_pt r;
    r.x = load_fe_wide(p);
    r.y = load_fe_wide((const char*)p + 32);
    r.zz = load_fe_wide((const char*)p + 64);
    r.zzz = load_fe_wide((const char*)p + 96);
    return r;
}
BBG_HD void store_xyzz_global(void* p, const xyzz_pt& a)
{
    store_fe_global(p, a.x);
    store_fe_global((char*)p + 32, a.y);
    store_fe_global((char*)p + 64, a.zz);
    store_fe_global((char*)p + 96, a.zzz);
}


} // namespace bbg
