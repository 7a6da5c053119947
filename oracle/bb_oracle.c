/* TEST INFRASTRUCTURE — NOT PRODUCT CODE.  See bb_oracle.h for scope, pinning status and
 * conventions.  Plain C restatement (gcc, unsigned __int128) of the reference algorithms; every
 * function cites the reference file:line it follows (paths relative to
 * /root/reference/src/barretenberg/).  Written to be obviously-correct rather than fast. */
#include "bb_oracle.h"

#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef uint64_t fe[4];

/* ------------------------------------------------------------------------------------------ */
/* Field parameters — values of curves/bn254/fq.hpp:12-64 and curves/bn254/fr.hpp:12-81.       */
/* ------------------------------------------------------------------------------------------ */
typedef struct
{
    fe p;        /* modulus */
    fe p2;       /* 2p */
    fe r2;       /* R^2 mod p */
    fe one;      /* R mod p */
    fe cube;     /* cube root of unity in Montgomery form (fq::beta / fr "beta" = lambda) */
    uint64_t ninv; /* -p^{-1} mod 2^64 */
} field_params;

static const field_params FP[2] = {
    { /* Fq */
      { 0x3C208C16D87CFD47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL },
      { 0x7841182db0f9fa8eULL, 0x2f02d522d0e3951aULL, 0x70a08b6d0302b0bbULL, 0x60c89ce5c2634053ULL },
      { 0xF32CFC5B538AFA89ULL, 0xB5E71911D44501FBULL, 0x47AB1EFF0A417FF6ULL, 0x06D89F71CAB8351FULL },
      { 0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL, 0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL },
      { 0x71930c11d782e155ULL, 0xa6bb947cffbe3323ULL, 0xaa303344d4741444ULL, 0x2c3b3f0d26594943ULL },
      0x87d20782e4866389ULL },
    { /* Fr */
      { 0x43E1F593F0000001ULL, 0x2833E84879B97091ULL, 0xB85045B68181585DULL, 0x30644E72E131A029ULL },
      { 0x87c3eb27e0000002ULL, 0x5067d090f372e122ULL, 0x70a08b6d0302b0baULL, 0x60c89ce5c2634053ULL },
      { 0x1BB8E645AE216DA7ULL, 0x53FE3AB1E35C59E3ULL, 0x8C49833D53BB8085ULL, 0x0216D0B17F4E44A5ULL },
      { 0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL, 0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL },
      { 0x93e7cede4a0329b3ULL, 0x7d4fdca77a96c167ULL, 0x8be4ba08b19a750aULL, 0x1cbd5653a5661c25ULL },
      0xc2e1f593efffffffULL },
};
/* fr.hpp:59-63 (2^28-th root of unity), :66-74 (generator 5 and inverse) */
static const fe FR_ROOT_OF_UNITY = { 0x636e735580d13d9cULL, 0xa22bf3742445ffd6ULL, 0x56452ac01eb203d8ULL, 0x1860ef942963f9e7ULL };
static const fe FR_GENERATOR = { 0x1b0d0ef99fffffe6ULL, 0xeaba68a3a32a913fULL, 0x47d8eb76d8dd0689ULL, 0x15d0085520f5bbc3ULL };
static const fe FR_GENERATOR_INV = { 0xd745397409999999ULL, 0xb4ada7d483c3efa8ULL, 0xc49ca2f8e57f3161ULL, 0x162a3754ac156cb3ULL };
/* g1.hpp:13-15: generator (x = fq::one, y) and curve constant b = 3, Montgomery form */
static const fe G1_ONE_Y = { 0xa6ba871b8b1e1b3aULL, 0x14f1d651eb8e167bULL, 0xccdd46def0f28c58ULL, 0x1c14ef83340fbe5eULL };
static const fe G1_B = { 0x7a17caa950ad28d7ULL, 0x1f6ac17ae15521b9ULL, 0x334bea4e696bd284ULL, 0x2a1f6744ce179d8eULL };

static void fe_copy(uint64_t* r, const uint64_t* a) { memmove(r, a, 32); }
static int fe_is_zero(const uint64_t* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static int fe_eq(const uint64_t* a, const uint64_t* b) { return memcmp(a, b, 32) == 0; }
/* field.hpp:179-188 */
static int fe_gt(const uint64_t* a, const uint64_t* b)
{
    for (int i = 3; i >= 0; --i)
    {
        if (a[i] != b[i]) return a[i] > b[i];
    }
    return 0;
}

/* a + b over 256 bits, carry dropped: field_impl_int128.tcc:165-173 */
static void raw_add(const uint64_t* a, const uint64_t* b, uint64_t* r)
{
    u128 c = 0;
    for (int i = 0; i < 4; ++i)
    {
        c += (u128)a[i] + b[i];
        r[i] = (uint64_t)c;
        c >>= 64;
    }
}
/* r = a - b; if that borrows, add `fix` back (mod 2^256): field_impl_int128.tcc:40-70 */
static void sub_fixup(const uint64_t* a, const uint64_t* b, const uint64_t* fix, uint64_t* r)
{
    uint64_t t[4];
    uint64_t borrow = 0;
    for (int i = 0; i < 4; ++i)
    {
        u128 d = (u128)a[i] - b[i] - borrow;
        t[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    if (borrow) raw_add(t, fix, t);
    fe_copy(r, t);
}

/* 256x256 -> 512 schoolbook: field_impl_int128.tcc:114-138 */
static void mul_512(const uint64_t* a, const uint64_t* b, uint64_t w[8])
{
    memset(w, 0, 64);
    for (int i = 0; i < 4; ++i)
    {
        uint64_t carry = 0;
        for (int j = 0; j < 4; ++j)
        {
            u128 t = (u128)a[i] * b[j] + w[i + j] + carry;
            w[i + j] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        w[i + 4] = carry;
    }
}
/* word-serial Montgomery reduction of a 512-bit value, no final subtraction:
 * field_impl_int128.tcc:72-110.  Result = (w + M*p) / 2^256 with M = -w/p mod 2^256. */
static void mont_reduce(const field_params* f, uint64_t w[8], uint64_t* out)
{
    uint64_t top = 0; /* carry into limb i+4 from previous rounds */
    for (int i = 0; i < 4; ++i)
    {
        uint64_t k = w[i] * f->ninv;
        uint64_t carry = 0;
        for (int j = 0; j < 4; ++j)
        {
            u128 t = (u128)k * f->p[j] + w[i + j] + carry;
            w[i + j] = (uint64_t)t;
            carry = (uint64_t)(t >> 64);
        }
        u128 t = (u128)w[i + 4] + carry + top;
        w[i + 4] = (uint64_t)t;
        top = (uint64_t)(t >> 64);
    }
    memcpy(out, w + 4, 32);
}

static void f_mul_coarse(const field_params* f, const uint64_t* a, const uint64_t* b, uint64_t* r)
{
    uint64_t w[8];
    mul_512(a, b, w);
    mont_reduce(f, w, r);
}
static void f_reduce_once(const field_params* f, const uint64_t* a, uint64_t* r) { sub_fixup(a, f->p, f->p, r); }
static void f_mul(const field_params* f, const uint64_t* a, const uint64_t* b, uint64_t* r)
{
    f_mul_coarse(f, a, b, r);
    f_reduce_once(f, r, r);
}
static void f_sqr(const field_params* f, const uint64_t* a, uint64_t* r) { f_mul(f, a, a, r); }
static void f_sqr_coarse(const field_params* f, const uint64_t* a, uint64_t* r) { f_mul_coarse(f, a, a, r); }
static void f_add(const field_params* f, const uint64_t* a, const uint64_t* b, uint64_t* r)
{
    raw_add(a, b, r);
    sub_fixup(r, f->p, f->p, r);
}
static void f_add_coarse(const field_params* f, const uint64_t* a, const uint64_t* b, uint64_t* r)
{
    raw_add(a, b, r);
    sub_fixup(r, f->p2, f->p2, r);
}
static void f_sub(const field_params* f, const uint64_t* a, const uint64_t* b, uint64_t* r) { sub_fixup(a, b, f->p, r); }
static void f_sub_coarse(const field_params* f, const uint64_t* a, const uint64_t* b, uint64_t* r) { sub_fixup(a, b, f->p2, r); }
static void f_neg(const field_params* f, const uint64_t* a, uint64_t* r) { f_sub(f, f->p, a, r); }
/* field_impl_int128.tcc:175-193: repeated coarse doubling */
static void f_quad_coarse(const field_params* f, const uint64_t* a, uint64_t* r)
{
    f_add_coarse(f, a, a, r);
    f_add_coarse(f, r, r, r);
}
static void f_oct_coarse(const field_params* f, const uint64_t* a, uint64_t* r)
{
    f_quad_coarse(f, a, r);
    f_add_coarse(f, r, r, r);
}
/* field.hpp:224-236 */
static void f_to_mont(const field_params* f, const uint64_t* a, uint64_t* r)
{
    uint64_t t[4], pp1[4] = { f->p[0] + 1, f->p[1], f->p[2], f->p[3] };
    fe_copy(t, a);
    while (fe_gt(t, pp1)) f_sub(f, t, f->p, t);
    f_mul(f, t, f->r2, r);
}
static void f_from_mont(const field_params* f, const uint64_t* a, uint64_t* r)
{
    static const fe raw_one = { 1, 0, 0, 0 };
    f_mul(f, a, raw_one, r);
}
/* field.hpp:246-287 pow (square-and-multiply from the top set bit, final loop-reduce) */
static void f_pow(const field_params* f, const uint64_t* a, const uint64_t* e, uint64_t* r)
{
    if (fe_is_zero(a))
    {
        memset(r, 0, 32);
        return;
    }
    int top = 255;
    while (top >= 0 && !((e[top >> 6] >> (top & 63)) & 1)) --top;
    uint64_t acc[4], base[4], pp1[4] = { f->p[0] + 1, f->p[1], f->p[2], f->p[3] };
    fe_copy(base, a);
    fe_copy(acc, a);
    for (int i = top - 1; i >= 0; --i)
    {
        f_sqr(f, acc, acc);
        if ((e[i >> 6] >> (i & 63)) & 1) f_mul(f, acc, base, acc);
    }
    while (fe_gt(acc, pp1)) f_sub(f, acc, f->p, acc);
    fe_copy(r, acc);
}
static void f_pow_small(const field_params* f, const uint64_t* a, uint64_t e, uint64_t* r)
{
    /* field.hpp:290-332: exponent 0 -> one, 1 -> a, 2 -> sqr, else generic ladder */
    if (e == 0)
    {
        fe_copy(r, f->one);
        return;
    }
    if (e == 1)
    {
        fe_copy(r, a);
        return;
    }
    uint64_t ee[4] = { e, 0, 0, 0 };
    if (fe_is_zero(a))
    {
        /* the reference's __pow_small has no zero guard, but 0^e = 0 either way */
        memset(r, 0, 32);
        return;
    }
    f_pow(f, a, ee, r);
}
static void f_invert(const field_params* f, const uint64_t* a, uint64_t* r)
{
    uint64_t e[4] = { f->p[0] - 2, f->p[1], f->p[2], f->p[3] };
    f_pow(f, a, e, r);
}

/* exported field API */
void orc_mul(int w, const uint64_t* a, const uint64_t* b, uint64_t* r) { f_mul(&FP[w], a, b, r); }
void orc_mul_coarse(int w, const uint64_t* a, const uint64_t* b, uint64_t* r) { f_mul_coarse(&FP[w], a, b, r); }
void orc_sqr(int w, const uint64_t* a, uint64_t* r) { f_sqr(&FP[w], a, r); }
void orc_add(int w, const uint64_t* a, const uint64_t* b, uint64_t* r) { f_add(&FP[w], a, b, r); }
void orc_sub(int w, const uint64_t* a, const uint64_t* b, uint64_t* r) { f_sub(&FP[w], a, b, r); }
void orc_add_coarse(int w, const uint64_t* a, const uint64_t* b, uint64_t* r) { f_add_coarse(&FP[w], a, b, r); }
void orc_sub_coarse(int w, const uint64_t* a, const uint64_t* b, uint64_t* r) { f_sub_coarse(&FP[w], a, b, r); }
void orc_reduce_once(int w, const uint64_t* a, uint64_t* r) { f_reduce_once(&FP[w], a, r); }
void orc_neg(int w, const uint64_t* a, uint64_t* r) { f_neg(&FP[w], a, r); }
void orc_to_mont(int w, const uint64_t* a, uint64_t* r) { f_to_mont(&FP[w], a, r); }
void orc_from_mont(int w, const uint64_t* a, uint64_t* r) { f_from_mont(&FP[w], a, r); }
void orc_invert(int w, const uint64_t* a, uint64_t* r) { f_invert(&FP[w], a, r); }
void orc_pow_small(int w, const uint64_t* a, uint64_t e, uint64_t* r) { f_pow_small(&FP[w], a, e, r); }
void orc_mul_n(int w, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n)
{
    for (size_t i = 0; i < n; ++i) f_mul(&FP[w], a + 4 * i, b + 4 * i, r + 4 * i);
}
/* Element-wise batches for the device self test (tests/test_field_selftest.py); op codes = bbg_field_selftest's
 * (include/bbgpu.h).  Ops whose device result is only defined as a residue (2, 6, 12) are returned canonical here. */
void orc_field_op_n(int w, int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n)
{
    const field_params* f = &FP[w];
    for (size_t i = 0; i < n; ++i)
    {
        const uint64_t* x = a + 4 * i;
        const uint64_t* y = b ? b + 4 * i : NULL;
        uint64_t* o = r + 4 * i;
        uint64_t t[4];
        switch (op)
        {
        case 0: f_mul_coarse(f, x, y, o); break;
        case 1: f_sqr_coarse(f, x, o); break;
        case 2: case 11: case 12: f_mul(f, x, y, o); break;
        case 3: f_add_coarse(f, x, y, o); break;
        case 4: f_sub_coarse(f, x, y, o); break;
        case 5: f_reduce_once(f, x, o); break;
        case 6: f_reduce_once(f, x, t); memset(o, 0, 32); f_sub(f, o, t, o); break; /* -x mod p, canonical, 0 -> 0 */
        case 7: f_to_mont(f, x, o); break;
        case 8: f_from_mont(f, x, o); break;
        case 9: f_invert(f, x, o); break;
        case 10: raw_add(x, f->p2, t); { /* x + 2p - y as a 256-bit integer, no correction */
            unsigned __int128 bw = 0;
            for (int k = 0; k < 4; ++k)
            {
                unsigned __int128 d = (unsigned __int128)t[k] - y[k] - (uint64_t)bw;
                o[k] = (uint64_t)d;
                bw = (d >> 64) & 1;
            }
        } break;
        default: memset(o, 0, 32);
        }
    }
}

void orc_constant(int which, uint64_t* r)
{
    const uint64_t* src = NULL;
    switch (which)
    {
    case 0: src = FP[0].p; break;
    case 1: src = FP[0].r2; break;
    case 2: src = FP[0].one; break;
    case 3: src = FP[0].cube; break;
    case 4: src = FP[1].p; break;
    case 5: src = FP[1].r2; break;
    case 6: src = FP[1].one; break;
    case 7: src = FP[1].cube; break;
    case 8: src = FR_ROOT_OF_UNITY; break;
    case 9: src = FR_GENERATOR; break;
    case 10: src = FR_GENERATOR_INV; break;
    case 11: src = FP[0].one; break; /* generator x = fq::one (g1.hpp:13) */
    case 12: src = G1_ONE_Y; break;
    case 13: src = G1_B; break;
    }
    if (src) memcpy(r, src, 32); else memset(r, 0, 32);
}

/* ------------------------------------------------------------------------------------------ */
/* Endomorphism split: fields/field.hpp:413-485.  k non-Montgomery; writes low 128 bits of     */
/* k1 to out[0..1] and of k2 to out[2..3].                                                     */
/* ------------------------------------------------------------------------------------------ */
void orc_split_endo(const uint64_t* k_in, uint64_t* out)
{
    static const fe g1c = { 0x7a7bd9d4391eb18dULL, 0x4ccef014a773d2cfULL, 0x0000000000000002ULL, 0 };
    static const fe g2c = { 0xd91d232ec7e0b3d7ULL, 0x0000000000000002ULL, 0, 0 };
    static const fe minus_b1 = { 0x8211bbeb7d4f1128ULL, 0x6f4d8248eeb859fcULL, 0, 0 };
    static const fe b2 = { 0x89d3256894d213e3ULL, 0, 0, 0 };
    const field_params* f = &FP[ORC_FR];
    uint64_t k[4], c1[8], c2[8], q1[8], q2[8], t1[4], t2[4];
    fe_copy(k, k_in);
    mul_512(g2c, k, c1);        /* c1 = (g2*k) >> 256 : high half */
    mul_512(g1c, k, c2);        /* c2 = (g1*k) >> 256 */
    mul_512(c1 + 4, minus_b1, q1);
    mul_512(c2 + 4, b2, q2);
    f_sub(f, q2, q1, t1);       /* low 256 bits only (:467) */
    f_mul(f, t1, f->cube, t2);  /* t1 * lambda (lambda is stored in Montgomery form) (:472) */
    f_add(f, k, t2, t2);        /* (:474) */
    out[0] = t2[0];
    out[1] = t2[1];
    out[2] = t1[0];
    out[3] = t1[1];
}

/* groups/wnaf.hpp:15-36 */
static uint32_t wnaf_bits_at(const uint64_t* s, size_t bits, size_t pos)
{
    size_t lo_idx = pos >> 6, hi_idx = (pos + bits - 1) >> 6;
    uint32_t lo = (uint32_t)(s[lo_idx] >> (pos & 63));
    uint32_t hi = 0;
    if (hi_idx != lo_idx) hi = (uint32_t)(s[hi_idx] << (64 - (pos & 63)));
    return (lo | hi) & ((1U << bits) - 1U);
}
/* groups/wnaf.hpp:38-55.  Entry = (|d|-1)/2 | (d<0)<<31, most significant window at wnaf[0],
 * stride num_points between windows.  Returns skew (1 if the scalar was even). */
int orc_fixed_wnaf(const uint64_t* scalar, uint32_t* wnaf, size_t num_points, size_t wnaf_bits)
{
    uint64_t s[2] = { scalar[0], scalar[1] };
    size_t entries = (127 + wnaf_bits - 1) / wnaf_bits;
    int skew = (s[0] & 1) == 0;
    uint32_t previous = wnaf_bits_at(s, wnaf_bits, 0) + (uint32_t)skew;
    for (size_t i = 1; i + 1 < entries; ++i)
    {
        uint32_t slice = wnaf_bits_at(s, wnaf_bits, i * wnaf_bits);
        uint32_t even = (slice & 1U) == 0U;
        /* if the next slice is even it borrows 2^w from us: our digit becomes previous - 2^w (negative) */
        wnaf[(entries - i) * num_points] = (((previous - (even << wnaf_bits)) ^ (0U - even)) >> 1) | (even << 31);
        previous = slice + even;
    }
    size_t final_bits = 127 - (127 / wnaf_bits) * wnaf_bits;
    uint32_t slice = wnaf_bits_at(s, final_bits, (entries - 1) * wnaf_bits);
    uint32_t even = (slice & 1U) == 0U;
    wnaf[num_points] = (((previous - (even << wnaf_bits)) ^ (0U - even)) >> 1) | (even << 31);
    wnaf[0] = (slice + even) >> 1;
    return skew;
}

/* ------------------------------------------------------------------------------------------ */
/* G1: groups/group.hpp.  Point layout x[0..3] y[4..7] z[8..11].                               */
/* ------------------------------------------------------------------------------------------ */
#define FQ (&FP[ORC_FQ])
static int pt_is_inf(const uint64_t* p) { return (int)(p[7] >> 63); }        /* :133-141 */
static void pt_set_inf(uint64_t* p) { p[7] = 1ULL << 63; }                    /* :143-151, field.hpp:199-202 */

/* group.hpp:153-217 */
void orc_g1_dbl(const uint64_t* p1, uint64_t* p2)
{
    if (pt_is_inf(p1))
    {
        pt_set_inf(p2);
        return;
    }
    uint64_t x[4], y[4], z[4], T0[4], T1[4], T2[4], T3[4], ox[4], oy[4], oz[4];
    fe_copy(x, p1); fe_copy(y, p1 + 4); fe_copy(z, p1 + 8);
    raw_add(z, z, oz);
    f_mul_coarse(FQ, oz, y, oz);
    f_reduce_once(FQ, oz, oz);
    f_sqr_coarse(FQ, x, T0);
    f_sqr_coarse(FQ, y, T1);
    f_sqr_coarse(FQ, T1, T2);
    f_add_coarse(FQ, T1, x, T1);
    f_sqr_coarse(FQ, T1, T1);
    f_add_coarse(FQ, T0, T2, T3);
    f_sub_coarse(FQ, T1, T3, T1);
    f_add_coarse(FQ, T1, T1, T1);
    f_add_coarse(FQ, T0, T0, T3);
    f_add_coarse(FQ, T3, T0, T3);
    f_add_coarse(FQ, T1, T1, T0);
    f_sqr_coarse(FQ, T3, ox);
    f_sub_coarse(FQ, ox, T0, ox);
    f_reduce_once(FQ, ox, ox);
    f_oct_coarse(FQ, T2, T2);
    f_sub_coarse(FQ, T1, ox, oy);
    f_mul_coarse(FQ, oy, T3, oy);
    f_sub_coarse(FQ, oy, T2, oy);
    f_reduce_once(FQ, oy, oy);
    fe_copy(p2, ox); fe_copy(p2 + 4, oy); fe_copy(p2 + 8, oz);
}

/* group.hpp:219-322 */
void orc_g1_mixed_add(const uint64_t* p1, const uint64_t* p2, uint64_t* p3)
{
    if (pt_is_inf(p1))
    {
        uint64_t t[12];
        memcpy(t, p2, 64);
        memcpy(t + 8, FQ->one, 32);
        memcpy(p3, t, 96);
        return;
    }
    uint64_t x1[4], y1[4], z1[4], T0[4], T1[4], T2[4], T3[4], ox[4], oy[4], oz[4];
    fe_copy(x1, p1); fe_copy(y1, p1 + 4); fe_copy(z1, p1 + 8);
    f_sqr_coarse(FQ, z1, T0);
    f_mul(FQ, p2, T0, T1);
    f_sub(FQ, T1, x1, T1);
    f_mul_coarse(FQ, z1, T0, T2);
    f_mul(FQ, T2, p2 + 4, T2);
    f_sub(FQ, T2, y1, T2);
    if (fe_is_zero(T1))
    {
        if (fe_is_zero(T2))
        {
            orc_g1_dbl(p1, p3);
        }
        else
        {
            uint64_t t[12];
            memcpy(t, p1, 96); /* the reference only flips the flag; other limbs are unspecified */
            pt_set_inf(t);
            memcpy(p3, t, 96);
        }
        return;
    }
    raw_add(T2, T2, T2);        /* __paralell_double_and_add_without_reduction: T2 *= 2, z3 = z1 + T1 */
    raw_add(z1, T1, oz);
    f_sqr_coarse(FQ, T1, T3);
    f_add_coarse(FQ, T0, T3, T0);
    f_sqr_coarse(FQ, oz, oz);
    f_sub_coarse(FQ, oz, T0, oz);
    f_reduce_once(FQ, oz, oz);
    f_quad_coarse(FQ, T3, T3);
    f_mul_coarse(FQ, T1, T3, T1);
    f_mul_coarse(FQ, T3, x1, T3);
    f_add_coarse(FQ, T3, T3, T0);
    f_add_coarse(FQ, T0, T1, T0);
    f_sqr_coarse(FQ, T2, ox);
    f_sub_coarse(FQ, ox, T0, ox);
    f_sub_coarse(FQ, T3, ox, T3);
    f_reduce_once(FQ, ox, ox);
    f_mul_coarse(FQ, T1, y1, T1);
    f_add_coarse(FQ, T1, T1, T1);
    f_mul_coarse(FQ, T3, T2, T3);
    f_sub_coarse(FQ, T3, T1, oy);
    f_reduce_once(FQ, oy, oy);
    fe_copy(p3, ox); fe_copy(p3 + 4, oy); fe_copy(p3 + 8, oz);
}

/* group.hpp:324-448 */
void orc_g1_add(const uint64_t* p1, const uint64_t* p2, uint64_t* p3)
{
    int z1 = pt_is_inf(p1), z2 = pt_is_inf(p2);
    if (z1 || z2)
    {
        if (z1 && !z2) memmove(p3, p2, 96);
        else if (z2 && !z1) memmove(p3, p1, 96);
        else pt_set_inf(p3);
        return;
    }
    uint64_t X1[4], Y1[4], Z1[4], X2[4], Y2[4], Z2[4];
    fe_copy(X1, p1); fe_copy(Y1, p1 + 4); fe_copy(Z1, p1 + 8);
    fe_copy(X2, p2); fe_copy(Y2, p2 + 4); fe_copy(Z2, p2 + 8);
    uint64_t Z1Z1[4], Z2Z2[4], U1[4], U2[4], S1[4], S2[4], F[4], H[4], I[4], J[4], ox[4], oy[4], oz[4];
    f_sqr_coarse(FQ, Z1, Z1Z1);
    f_sqr_coarse(FQ, Z2, Z2Z2);
    f_mul_coarse(FQ, X1, Z2Z2, U1);
    f_mul_coarse(FQ, X2, Z1Z1, U2);
    f_mul_coarse(FQ, Z2, Z2Z2, S1);
    f_mul_coarse(FQ, Z1, Z1Z1, S2);
    f_mul_coarse(FQ, S1, Y1, S1);
    f_mul_coarse(FQ, S2, Y2, S2);
    f_sub_coarse(FQ, U2, U1, H);
    f_reduce_once(FQ, H, H);
    f_sub_coarse(FQ, S2, S1, F);
    f_reduce_once(FQ, F, F);
    if (fe_is_zero(H))
    {
        if (fe_is_zero(F))
        {
            orc_g1_dbl(p1, p3);
        }
        else
        {
            uint64_t t[12];
            memcpy(t, p1, 96);
            pt_set_inf(t);
            memcpy(p3, t, 96);
        }
        return;
    }
    raw_add(F, F, F);
    raw_add(H, H, I);
    f_sqr_coarse(FQ, I, I);
    f_mul_coarse(FQ, H, I, J);
    f_mul_coarse(FQ, U1, I, U1);
    f_add_coarse(FQ, U1, U1, U2);
    f_add_coarse(FQ, U2, J, U2);
    f_sqr_coarse(FQ, F, ox);
    f_sub_coarse(FQ, ox, U2, ox);
    f_reduce_once(FQ, ox, ox);
    f_mul_coarse(FQ, J, S1, J);
    f_add_coarse(FQ, J, J, J);
    f_sub_coarse(FQ, U1, ox, oy);
    f_mul_coarse(FQ, oy, F, oy);
    f_sub_coarse(FQ, oy, J, oy);
    f_reduce_once(FQ, oy, oy);
    f_add_coarse(FQ, Z1, Z2, oz);
    f_add_coarse(FQ, Z1Z1, Z2Z2, Z1Z1);
    f_sqr_coarse(FQ, oz, oz);
    f_sub_coarse(FQ, oz, Z1Z1, oz);
    f_mul(FQ, oz, H, oz);
    fe_copy(p3, ox); fe_copy(p3 + 4, oy); fe_copy(p3 + 8, oz);
}

/* group.hpp:450-469 */
void orc_g1_normalize(const uint64_t* src, uint64_t* dst)
{
    uint64_t zi[4], zz[4], zzz[4], t[12];
    int inf = pt_is_inf(src);
    f_invert(FQ, src + 8, zi);
    f_sqr(FQ, zi, zz);
    f_mul(FQ, zi, zz, zzz);
    f_mul(FQ, src, zz, t);
    f_mul(FQ, src + 4, zzz, t + 4);
    memcpy(t + 8, FQ->one, 32);
    if (inf) pt_set_inf(t);
    memcpy(dst, t, 96);
}

/* group.hpp:474-534 (Montgomery's trick; infinity points are skipped but get z = one) */
void orc_g1_batch_normalize(uint64_t* pts, size_t n)
{
    if (n == 0) return;
    uint64_t(*tmp)[4] = malloc(sizeof(fe) * n);
    uint64_t acc[4], zi[4], zz[4], zzz[4];
    fe_copy(acc, FQ->one);
    for (size_t i = 0; i < n; ++i)
    {
        fe_copy(tmp[i], acc);
        if (!pt_is_inf(pts + 12 * i)) f_mul(FQ, acc, pts + 12 * i + 8, acc);
    }
    f_invert(FQ, acc, acc);
    for (size_t i = n; i-- > 0;)
    {
        uint64_t* p = pts + 12 * i;
        if (!pt_is_inf(p))
        {
            f_mul(FQ, acc, tmp[i], zi);
            f_sqr(FQ, zi, zz);
            f_mul(FQ, zi, zz, zzz);
            f_mul(FQ, p, zz, p);
            f_mul(FQ, p + 4, zzz, p + 4);
            f_mul(FQ, acc, p + 8, acc);
        }
        memcpy(p + 8, FQ->one, 32);
    }
    free(tmp);
}

/* Affine in, affine out (normalised; infinity = flag) for the device self test; op codes = bbg_g1_selftest's. */
static void aff_to_jac(const uint64_t* a, uint64_t* j)
{
    memcpy(j, a, 64);
    memcpy(j + 8, FQ->one, 32);
    if (a[7] >> 63) pt_set_inf(j);
}
void orc_g1_op_n(int op, const uint64_t* p, const uint64_t* q, uint64_t* out, size_t n)
{
    for (size_t i = 0; i < n; ++i)
    {
        uint64_t P[12], Q[12], t[12], u[12], r[12];
        aff_to_jac(p + 8 * i, P);
        aff_to_jac(q + 8 * i, Q);
        switch (op)
        {
        case 0: if (q[8 * i + 7] >> 63) memcpy(r, P, 96); else orc_g1_mixed_add(P, q + 8 * i, r); break;
        case 1: orc_g1_dbl(P, t); orc_g1_dbl(Q, u); orc_g1_add(t, u, r); break;
        case 2: orc_g1_dbl(P, t); orc_g1_dbl(t, r); break;
        case 3: orc_g1_add(P, Q, t); orc_g1_add(t, P, r); break;
        case 4: orc_g1_add(P, Q, r); break;
        case 5: {
            uint64_t tab[16];
            orc_generate_pippenger_point_table(p + 8 * i, tab, 1);
            aff_to_jac(tab + 8, r);
        } break;
        case 6: orc_g1_dbl(P, r); break;
        default: memcpy(r, P, 96); pt_set_inf(r);
        }
        orc_g1_normalize(r, t);
        memcpy(out + 8 * i, t, 64);
        if (pt_is_inf(t))
        {
            memset(out + 8 * i, 0, 64);
            out[8 * i + 7] = (uint64_t)1 << 63;
        }
    }
}

/* group.hpp:536-552: y^2 == x^3 + b, compared out of Montgomery form */
int orc_g1_on_curve(const uint64_t* a)
{
    if (pt_is_inf(a)) return 0;
    uint64_t xxx[4], yy[4];
    f_sqr(FQ, a, xxx);
    f_mul(FQ, a, xxx, xxx);
    f_add(FQ, xxx, G1_B, xxx);
    f_sqr(FQ, a + 4, yy);
    f_from_mont(FQ, xxx, xxx);
    f_from_mont(FQ, yy, yy);
    return fe_eq(xxx, yy);
}

/* Value-equivalent scalar multiplication (binary ladder, msb first).  The reference's
 * group_exponentiation (group.hpp:653-799) uses an endo-wNAF ladder; both return the affine
 * representation of the same group element, infinity flagged with x = y = 0 + msb. */
void orc_g1_scalar_mul(const uint64_t* affine, const uint64_t* scalar_mont, uint64_t* out)
{
    uint64_t k[4], acc[12];
    f_from_mont(&FP[ORC_FR], scalar_mont, k);
    memset(acc, 0, sizeof acc);
    pt_set_inf(acc);
    for (int i = 255; i >= 0; --i)
    {
        orc_g1_dbl(acc, acc);
        if ((k[i >> 6] >> (i & 63)) & 1) orc_g1_mixed_add(acc, affine, acc);
    }
    if (pt_is_inf(acc))
    {
        memset(out, 0, 64);
        pt_set_inf(out); /* affine y sits at limbs 4..7, the same offset as a Jacobian y */
        return;
    }
    orc_g1_batch_normalize(acc, 1);
    memcpy(out, acc, 64);
}

void orc_g1_arith_progression(const uint64_t* start_mont, const uint64_t* step_mont, uint64_t* out, size_t n)
{
    if (n == 0) return;
    uint64_t gen[8], base[8], step[8];
    memcpy(gen, FQ->one, 32);
    memcpy(gen + 4, G1_ONE_Y, 32);
    orc_g1_scalar_mul(gen, start_mont, base);
    orc_g1_scalar_mul(gen, step_mont, step);
    uint64_t* acc = malloc(96 * n);
    memcpy(acc, base, 64);
    memcpy(acc + 8, FQ->one, 32);
    for (size_t i = 1; i < n; ++i)
    {
        if (pt_is_inf(step)) memcpy(acc + 12 * i, acc + 12 * (i - 1), 96);
        else orc_g1_mixed_add(acc + 12 * (i - 1), step, acc + 12 * i);
    }
    orc_g1_batch_normalize(acc, n);
    for (size_t i = 0; i < n; ++i) memcpy(out + 8 * i, acc + 12 * i, 64);
    free(acc);
}

/* ------------------------------------------------------------------------------------------ */
/* MSM: curves/bn254/scalar_multiplication.cpp                                                 */
/* ------------------------------------------------------------------------------------------ */
size_t orc_get_optimal_bucket_width(size_t n) /* :21-81, thresholds restated as data */
{
    static const struct { size_t at_least; size_t width; } table[] = {
        { 14617149, 21 }, { 2139094, 18 }, { 100000, 15 }, { 144834, 14 }, { 25067, 12 }, { 13926, 11 },
        { 7659, 10 },     { 2436, 9 },     { 376, 7 },     { 231, 6 },     { 97, 5 },     { 35, 4 },
        { 10, 3 },        { 2, 2 },
    };
    for (size_t i = 0; i < sizeof table / sizeof table[0]; ++i)
    {
        if (n >= table[i].at_least) return table[i].width;
    }
    return 1;
}

/* :131-140.  table[2i] = P_i, table[2i+1] = (beta*x_i, -y_i) */
void orc_generate_pippenger_point_table(const uint64_t* points, uint64_t* table, size_t n)
{
    for (size_t i = n; i-- > 0;)
    {
        uint64_t px[4], py[4];
        fe_copy(px, points + 8 * i);
        fe_copy(py, points + 8 * i + 4);
        fe_copy(table + 16 * i, px);
        fe_copy(table + 16 * i + 4, py);
        f_mul(FQ, px, FQ->cube, table + 16 * i + 8);
        f_neg(FQ, py, table + 16 * i + 12);
    }
}

/* :576-648 with compute_wnaf_state (:265-308) inlined.  `k` = scalars already out of Montgomery form. */
static void pippenger_internal(const uint64_t* k, const uint64_t* table, size_t n, size_t forced_width, uint64_t* out)
{
    size_t c = forced_width ? forced_width : orc_get_optimal_bucket_width(n);
    size_t w = c + 1;
    size_t num_points = 2 * n;
    size_t rounds = (127 + w - 1) / w;
    size_t num_buckets = (size_t)1 << c;

    uint64_t* buckets = malloc(96 * num_buckets);
    uint32_t* wnaf = malloc(sizeof(uint32_t) * rounds * num_points);
    unsigned char* skew = malloc(num_points);
    for (size_t b = 0; b < num_buckets; ++b)
    {
        memset(buckets + 12 * b, 0, 96);
        pt_set_inf(buckets + 12 * b);
    }
    for (size_t i = 0; i < n; ++i)
    {
        uint64_t halves[4];
        orc_split_endo(k + 4 * i, halves);
        skew[2 * i] = (unsigned char)orc_fixed_wnaf(halves, wnaf + 2 * i, num_points, w);
        skew[2 * i + 1] = (unsigned char)orc_fixed_wnaf(halves + 2, wnaf + 2 * i + 1, num_points, w);
    }

    uint64_t acc[12], run[12], tmp[8];
    memset(acc, 0, sizeof acc);
    pt_set_inf(acc);
    for (size_t r = 0; r < rounds; ++r)
    {
        if (r == rounds - 1)
        {
            /* skew correction: subtract P_j once for every even half-scalar (:593-603) */
            for (size_t j = 0; j < num_points; ++j)
            {
                if (!skew[j]) continue;
                memcpy(tmp, table + 8 * j, 32);
                f_neg(FQ, table + 8 * j + 4, tmp + 4);
                orc_g1_mixed_add(buckets, tmp, buckets);
            }
        }
        for (size_t j = 0; j < num_points; ++j)
        {
            uint32_t entry = wnaf[r * num_points + j];
            size_t idx = entry & 0x0fffffffU; /* :83-88 */
            memcpy(tmp, table + 8 * j, 64);
            /* conditional_negate_affine (group_impl_int128.tcc): y -> p - y when the sign bit is set */
            if (entry >> 31) f_neg(FQ, table + 8 * j + 4, tmp + 4);
            orc_g1_mixed_add(buckets + 12 * idx, tmp, buckets + 12 * idx);
        }
        if (r > 0)
        {
            for (size_t j = 0; j < c; ++j) orc_g1_dbl(acc, acc);
        }
        memset(run, 0, sizeof run);
        pt_set_inf(run);
        for (size_t j = num_buckets - 1; j > 0; --j)
        {
            orc_g1_add(run, buckets + 12 * j, run);
            orc_g1_add(acc, run, acc);
            pt_set_inf(buckets + 12 * j);
        }
        orc_g1_add(run, buckets, run);
        orc_g1_dbl(acc, acc);
        orc_g1_add(acc, run, acc);
        pt_set_inf(buckets);
    }
    memcpy(out, acc, 96);
    free(buckets);
    free(wnaf);
    free(skew);
}

/* :457-476 */
void orc_pippenger(const uint64_t* scalars_mont, const uint64_t* table, size_t n, size_t forced_width, uint64_t* out)
{
    if (n == 0)
    {
        memcpy(out, FQ->one, 32);
        memcpy(out + 4, G1_ONE_Y, 32);
        memcpy(out + 8, FQ->one, 32);
        pt_set_inf(out);
        return;
    }
    uint64_t* k = malloc(32 * n);
    for (size_t i = 0; i < n; ++i) f_from_mont(&FP[ORC_FR], scalars_mont + 4 * i, k + 4 * i);
    pippenger_internal(k, table, n, forced_width, out);
    free(k);
}

void orc_msm_normalized(const uint64_t* scalars_mont, const uint64_t* table, size_t n, uint64_t* out)
{
    orc_pippenger(scalars_mont, table, n, 0, out);
    orc_g1_batch_normalize(out, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* NTT: polynomials/evaluation_domain.cpp + polynomial_arithmetic.cpp                          */
/* ------------------------------------------------------------------------------------------ */
#define FRP (&FP[ORC_FR])
struct orc_domain
{
    size_t size, log2_size;
    fe root, root_inverse, domain, domain_inverse, generator, generator_inverse;
    uint64_t* roots;          /* 2*size elements: forward rounds then inverse rounds */
    uint64_t** round_roots;   /* log2_size-1 pointers */
    uint64_t** inv_round_roots;
};

/* evaluation_domain.cpp:33-54: round i (m = 2^(i+1)) holds w_{2m}^j for j < m, built by a coarse-mul chain */
static void build_round_table(const uint64_t* root, size_t size, size_t log2_size, uint64_t* mem, uint64_t** rounds)
{
    size_t off = 0;
    for (size_t i = 0; i + 1 < log2_size; ++i)
    {
        size_t m = (size_t)1 << (i + 1);
        uint64_t round_root[4];
        rounds[i] = mem + 4 * off;
        f_pow_small(FRP, root, size / (2 * m), round_root);
        fe_copy(rounds[i], FRP->one);
        for (size_t j = 1; j < m; ++j) f_mul_coarse(FRP, rounds[i] + 4 * (j - 1), round_root, rounds[i] + 4 * j);
        off += m;
    }
}

orc_domain* orc_domain_new(size_t n)
{
    orc_domain* d = calloc(1, sizeof *d);
    d->size = n;
    while (((size_t)1 << d->log2_size) < n) d->log2_size++;
    /* field.hpp:487-494: square the 2^28-th root down to order n */
    fe_copy(d->root, FR_ROOT_OF_UNITY);
    for (size_t i = 28; i > d->log2_size; --i) f_sqr(FRP, d->root, d->root);
    f_invert(FRP, d->root, d->root_inverse);
    uint64_t raw_n[4] = { n, 0, 0, 0 };
    f_to_mont(FRP, raw_n, d->domain);
    f_invert(FRP, d->domain, d->domain_inverse);
    fe_copy(d->generator, FR_GENERATOR);
    fe_copy(d->generator_inverse, FR_GENERATOR_INV);
    if (d->log2_size >= 1 && n >= 2)
    {
        size_t rounds = d->log2_size - 1;
        d->roots = malloc(32 * 2 * n);
        d->round_roots = calloc(rounds + 1, sizeof(uint64_t*));
        d->inv_round_roots = calloc(rounds + 1, sizeof(uint64_t*));
        build_round_table(d->root, n, d->log2_size, d->roots, d->round_roots);
        build_round_table(d->root_inverse, n, d->log2_size, d->roots + 4 * n, d->inv_round_roots);
    }
    return d;
}
void orc_domain_free(orc_domain* d)
{
    if (!d) return;
    free(d->roots);
    free(d->round_roots);
    free(d->inv_round_roots);
    free(d);
}
void orc_domain_constant(const orc_domain* d, int which, uint64_t* r)
{
    const uint64_t* src[6] = { d->root, d->root_inverse, d->domain, d->domain_inverse, d->generator, d->generator_inverse };
    memcpy(r, src[which], 32);
}

/* polynomial_arithmetic.cpp:14-21 */
static uint32_t reverse_bits(uint32_t x, uint32_t bits)
{
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; ++i) r |= ((x >> i) & 1U) << (bits - 1 - i);
    return r;
}

/* polynomial_arithmetic.cpp:129-264, one thread.  Bit-reversed gather into scratch, round 0
 * without twiddles, then DIT rounds m = 2..n/2 with lazy [0,2p) values; the last round writes
 * reduce_once() results back to coeffs. */
static void fft_inner(uint64_t* coeffs, const orc_domain* d, uint64_t* const* table)
{
    size_t n = d->size;
    uint64_t* s = malloc(32 * n);
    for (size_t i = 0; i < n; ++i) fe_copy(s + 4 * i, coeffs + 4 * reverse_bits((uint32_t)i, (uint32_t)d->log2_size));
    for (size_t i = 0; i + 1 < n; i += 2)
    {
        uint64_t t[4];
        fe_copy(t, s + 4 * (i + 1));
        f_sub_coarse(FRP, s + 4 * i, s + 4 * (i + 1), s + 4 * (i + 1));
        f_add_coarse(FRP, t, s + 4 * i, s + 4 * i);
    }
    if (n <= 2)
    {
        for (size_t i = 0; i < n; ++i) f_reduce_once(FRP, s + 4 * i, coeffs + 4 * i);
        free(s);
        return;
    }
    size_t round = 0;
    for (size_t m = 2; m < n; m <<= 1, ++round)
    {
        const uint64_t* tw = table[round];
        int last = (m == (n >> 1));
        for (size_t i = 0; i < n / 2; ++i)
        {
            size_t k1 = (i & ~(m - 1)) << 1, j1 = i & (m - 1);
            uint64_t* lo = s + 4 * (k1 + j1);
            uint64_t* hi = s + 4 * (k1 + j1 + m);
            uint64_t t[4];
            f_mul_coarse(FRP, tw + 4 * j1, hi, t);
            f_sub_coarse(FRP, lo, t, hi);
            f_add_coarse(FRP, lo, t, lo);
            if (last)
            {
                f_reduce_once(FRP, hi, coeffs + 4 * (k1 + j1 + m));
                f_reduce_once(FRP, lo, coeffs + 4 * (k1 + j1));
            }
        }
    }
    free(s);
}

/* polynomial_arithmetic.cpp:81-102 with one chunk: coeffs[i] *= start * shift^i */
static void scale_by_generator(uint64_t* coeffs, size_t n, const uint64_t* start, const uint64_t* shift)
{
    uint64_t work[4];
    f_mul_coarse(FRP, start, FRP->one, work);
    for (size_t i = 0; i < n; ++i)
    {
        f_mul(FRP, coeffs + 4 * i, work, coeffs + 4 * i);
        f_mul_coarse(FRP, work, shift, work);
    }
}
static void scale_all(uint64_t* coeffs, size_t n, const uint64_t* k)
{
    for (size_t i = 0; i < n; ++i) f_mul(FRP, coeffs + 4 * i, k, coeffs + 4 * i);
}

/* polynomial_arithmetic.cpp:266-315 */
void orc_ntt(const orc_domain* d, int op, uint64_t* c, const uint64_t* constant)
{
    size_t n = d->size;
    uint64_t t[4];
    switch (op)
    {
    case 0: fft_inner(c, d, d->round_roots); break;
    case 1:
        fft_inner(c, d, d->inv_round_roots);
        scale_all(c, n, d->domain_inverse);
        break;
    case 2:
        scale_by_generator(c, n, FRP->one, d->generator);
        fft_inner(c, d, d->round_roots);
        break;
    case 3:
        fft_inner(c, d, d->inv_round_roots);
        scale_all(c, n, d->domain_inverse);
        scale_by_generator(c, n, FRP->one, d->generator_inverse);
        break;
    case 4:
        fft_inner(c, d, d->round_roots);
        scale_all(c, n, constant);
        break;
    case 5:
        fft_inner(c, d, d->inv_round_roots);
        f_mul(FRP, d->domain_inverse, constant, t);
        scale_all(c, n, t);
        break;
    case 6:
        f_mul(FRP, FRP->one, constant, t);
        scale_by_generator(c, n, t, d->generator);
        fft_inner(c, d, d->round_roots);
        break;
    }
}

/* polynomial_arithmetic.cpp:337-373 with one chunk: sum coeffs[i] * z^i */
void orc_poly_evaluate(const uint64_t* coeffs, const uint64_t* z, size_t n, uint64_t* out)
{
    uint64_t acc[4] = { 0, 0, 0, 0 }, zp[4], t[4];
    fe_copy(zp, FRP->one);
    for (size_t i = 0; i < n; ++i)
    {
        f_mul(FRP, coeffs + 4 * i, zp, t);
        f_add(FRP, acc, t, acc);
        f_mul_coarse(FRP, zp, z, zp);
    }
    fe_copy(out, acc);
}

/* polynomial_arithmetic.cpp:381-476: l_1[i] = numer[i mod S] / (g w_T^i - 1), numer[j] = (g^n w_S^j - 1) / n
 * (compute_multiplicative_subgroup :104-127 supplies g^n w_S^j).  Value-equivalent: canonical outputs. */
void orc_compute_lagrange_polynomial_fft(uint64_t* l_1, size_t log2_src, size_t log2_target)
{
    size_t T = (size_t)1 << log2_target, S = (size_t)1 << (log2_target - log2_src);
    orc_domain* src = orc_domain_new((size_t)1 << log2_src);
    orc_domain* tgt = orc_domain_new(T);
    uint64_t root_S[4], gn[4], numer[1024][4], x[4], den[4], inv[4];
    fe_copy(root_S, FR_ROOT_OF_UNITY);
    for (size_t i = 28; i > log2_target - log2_src; --i) f_sqr(FRP, root_S, root_S);
    fe_copy(gn, FR_GENERATOR);
    for (size_t i = 0; i < log2_src; ++i) f_sqr(FRP, gn, gn);
    for (size_t j = 0; j < S; ++j)
    {
        f_sub(FRP, gn, FRP->one, numer[j]);
        f_mul(FRP, numer[j], src->domain_inverse, numer[j]);
        f_mul(FRP, gn, root_S, gn);
    }
    fe_copy(x, FR_GENERATOR);
    for (size_t i = 0; i < T; ++i)
    {
        f_sub(FRP, x, FRP->one, den);
        f_invert(FRP, den, inv);
        f_mul(FRP, inv, numer[i & (S - 1)], l_1 + 4 * i);
        f_mul(FRP, x, tgt->root, x);
    }
    orc_domain_free(src);
    orc_domain_free(tgt);
}
