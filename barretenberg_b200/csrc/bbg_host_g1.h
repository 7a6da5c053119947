// Host-side finish of an MSM: combine the per-window bucket reductions the GPU produced, fold the
// windows (the reference's "acc = 2^(c+1) acc + window sum", scalar_multiplication.cpp:619-639) and
// normalise (group.hpp:450-469).  This is the "final host-side fold" of the design: O(windows * c)
// point operations (~350 for a 2^20-point MSM, < 0.1 ms), a strictly serial dependency chain that a GPU
// thread would take 4-5x longer to walk.  It is NOT a CPU fallback for the MSM: every bucket
// accumulation and reduction runs in the CUDA kernels of bbg_msm.cu, and nothing here can run
// without their output.
//
// Fq on the host uses 4 x 64-bit limbs with unsigned __int128 products; same lazy [0,2p) value
// conventions as the device code (bbg_field.cuh).
#pragma once
#include <stdint.h>
#include <string.h>

namespace bbg
{
namespace hostg1
{
typedef unsigned __int128 u128;
struct hfq
{
    uint64_t v[4];
};
struct hxyzz
{
    hfq x, y, zz, zzz;
};

static const hfq HP = { { 0x3C208C16D87CFD47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL } };  // fq.hpp:12-15
static const hfq HP2 = { { 0x7841182db0f9fa8eULL, 0x2f02d522d0e3951aULL, 0x70a08b6d0302b0bbULL, 0x60c89ce5c2634053ULL } }; // fq.hpp:23-26
static const hfq HONE = { { 0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL } }; // fq.hpp:33-36
static const uint64_t HNINV = 0x87d20782e4866389ULL;                                                                       // fq.hpp:64

inline bool is_zero_raw(const hfq& a) { return (a.v[0] | a.v[1] | a.v[2] | a.v[3]) == 0; }
inline uint64_t sub4(hfq& r, const hfq& a, const hfq& b)
{
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i)
    {
        u128 d = (u128)a.v[i] - b.v[i] - borrow;
        r.v[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    return borrow;
}
inline void add4(hfq& r, const hfq& a, const hfq& b)
{
    u128 c = 0;
    for (int i = 0; i < 4; ++i)
    {
        c += (u128)a.v[i] + b.v[i];
        r.v[i] = (uint64_t)c;
        c >>= 64;
    }
}
inline hfq reduce(const hfq& a)
{
    hfq t;
    return sub4(t, a, HP) ? a : t;
}
inline hfq add(const hfq& a, const hfq& b)
{
    hfq s, t;
    add4(s, a, b);
    return sub4(t, s, HP2) ? s : t;
}
inline hfq sub(const hfq& a, const hfq& b)
{
    hfq d, t;
    if (!sub4(d, a, b)) return d;
    add4(t, d, HP2);
    return t;
}
inline hfq dbl(const hfq& a) { return add(a, a); }
// Montgomery product, coarse result in [0,2p) for inputs in [0,2p) (4-limb CIOS)
inline hfq mul(const hfq& a, const hfq& b)
{
    uint64_t t[6] = { 0, 0, 0, 0, 0, 0 };
    for (int i = 0; i < 4; ++i)
    {
        uint64_t carry = 0;
        for (int j = 0; j < 4; ++j)
        {
            u128 x = (u128)a.v[j] * b.v[i] + t[j] + carry;
            t[j] = (uint64_t)x;
            carry = (uint64_t)(x >> 64);
        }
        u128 x = (u128)t[4] + carry;
        t[4] = (uint64_t)x;
        t[5] = (uint64_t)(x >> 64);
        const uint64_t m = t[0] * HNINV;
        x = (u128)m * HP.v[0] + t[0];
        carry = (uint64_t)(x >> 64);
        for (int j = 1; j < 4; ++j)
        {
            x = (u128)m * HP.v[j] + t[j] + carry;
            t[j - 1] = (uint64_t)x;
            carry = (uint64_t)(x >> 64);
        }
        x = (u128)t[4] + carry;
        t[3] = (uint64_t)x;
        t[4] = t[5] + (uint64_t)(x >> 64);
    }
    hfq r = { { t[0], t[1], t[2], t[3] } };
    return r;
}
inline hfq sqr(const hfq& a) { return mul(a, a); }
inline bool is_zero(const hfq& a) { return is_zero_raw(reduce(a)); }
inline hfq invert(const hfq& a) // a^(p-2), field.hpp:345-348
{
    uint64_t e[4] = { HP.v[0] - 2, HP.v[1], HP.v[2], HP.v[3] };
    hfq acc = HONE;
    bool started = false;
    for (int i = 255; i >= 0; --i)
    {
        if (started) acc = sqr(acc);
        if ((e[i >> 6] >> (i & 63)) & 1)
        {
            acc = started ? mul(acc, a) : a;
            started = true;
        }
    }
    return reduce(acc);
}

inline hxyzz infinity()
{
    hxyzz r;
    memset(&r, 0, sizeof r);
    return r;
}
inline bool is_infinity(const hxyzz& p) { return is_zero_raw(p.zz); }
// EFD dbl-2008-s-1 (a = 0), same formulas as G1::dbl in bbg_g1.cuh
inline hxyzz dbl(const hxyzz& p)
{
    if (is_infinity(p)) return p;
    hxyzz r;
    hfq U = dbl(p.y), V = sqr(U), W = mul(U, V), S = mul(p.x, V), xx = sqr(p.x);
    hfq M = add(dbl(xx), xx);
    hfq X3 = sub(sqr(M), dbl(S));
    r.x = X3;
    r.y = sub(mul(M, sub(S, X3)), mul(W, p.y));
    r.zz = mul(V, p.zz);
    r.zzz = mul(W, p.zzz);
    return r;
}
// EFD add-2008-s, same formulas and exception paths as G1::add
inline hxyzz add(const hxyzz& a, const hxyzz& b)
{
    if (is_infinity(a)) return b;
    if (is_infinity(b)) return a;
    hfq U1 = mul(a.x, b.zz), U2 = mul(b.x, a.zz), S1 = mul(a.y, b.zzz), S2 = mul(b.y, a.zzz);
    hfq P = sub(U2, U1), R = sub(S2, S1);
    if (is_zero(P))
    {
        if (is_zero(R)) return dbl(a);
        return infinity();
    }
    hfq PP = sqr(P), PPP = mul(P, PP), Q = mul(U1, PP);
    hxyzz r;
    hfq X3 = sub(sub(sqr(R), PPP), dbl(Q));
    r.x = X3;
    r.y = sub(mul(R, sub(Q, X3)), mul(S1, PPP));
    r.zz = mul(mul(a.zz, b.zz), PP);
    r.zzz = mul(mul(a.zzz, b.zzz), PPP);
    return r;
}
// -> reference normalised Jacobian (x, y canonical, z = fq::one); infinity -> x = y = 0, flag set, z = one
inline void to_normalized_jacobian(const hxyzz& p, uint64_t out[12])
{
    memset(out, 0, 96);
    memcpy(out + 8, HONE.v, 32);
    if (is_infinity(p))
    {
        out[7] = 1ULL << 63;
        return;
    }
    hfq inv = invert(mul(p.zz, p.zzz));
    hfq x = reduce(mul(p.x, mul(inv, p.zzz)));
    hfq y = reduce(mul(p.y, mul(inv, p.zz)));
    memcpy(out, x.v, 32);
    memcpy(out + 4, y.v, 32);
}
inline hxyzz from_normalized_jacobian(const uint64_t in[12])
{
    hxyzz r;
    if (in[7] >> 63) return infinity();
    memcpy(r.x.v, in, 32);
    memcpy(r.y.v, in + 4, 32);
    r.zz = HONE;
    r.zzz = HONE;
    return r;
}
} // namespace hostg1
} // namespace bbg
