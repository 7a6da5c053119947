// bn254 Fq / Fr arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form with R = 2^256.
//
// Replaces the reference's x86-64 MULX/ADCX/ADOX field layer
//   fields/field_impl_asm.tcc:58-348, fields/asm_macros.hpp:34-413   (semantics: field_impl_int128.tcc)
// with IMAD / carry-chain code.  Same value conventions as the reference:
//   * "coarse" results live in [0, 2p) (reference: *_with_coarse_reduction), canonical ones in [0, p)
//   * every routine accepts inputs in [0, 2p); 4p < 2^256 for both moduli so sums never wrap
//   * limbs in memory are the reference's 4 x uint64 little-endian == our 8 x uint32 little-endian
//
// Montgomery product: operand-scanning CIOS over 32-bit words with two interleaved accumulators
// ("even"/"odd" columns) so every 32x32->64 product lands 64-bit aligned and each (mad.lo.cc,
// madc.hi.cc) pair can issue as one wide IMAD with carry.  Per product: 128 wide IMADs + 8 IMADs
// for the reduction factors (SURVEY.md §8a "136 MACs").
//
// The carry-chain primitives have two bodies: inline PTX under __CUDA_ARCH__, and a plain C++
// rendition used ONLY by the CPU-side kernel emulation build (tests/emul, -DBBG_EMULATE) so the
// limb logic can be checked against the oracle on machines without a GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BBG_HD __host__ __device__ __forceinline__
#define BBG_D __device__ __forceinline__
#else
#define BBG_HD inline
#define BBG_D inline
#endif

namespace bbg
{

struct fe
{
    uint32_t v[8];
};

// ---------------------------------------------------------------------------------------------
// Field parameters (values: reference curves/bn254/fq.hpp:12-64, curves/bn254/fr.hpp:12-81)
// ---------------------------------------------------------------------------------------------
// (accessor functions rather than static arrays: nvcc cannot ODR-use a host constexpr array in device code;
//  after unrolling every index is a literal and the words become instruction immediates)
#define BBG_CONST8(NAME, ...)                                                                      \
    static BBG_HD constexpr uint32_t NAME(int i)                                                   \
    {                                                                                              \
        constexpr uint32_t t[8] = { __VA_ARGS__ };                                                 \
        return t[i];                                                                               \
    }
struct FqParams
{
    BBG_CONST8(P, 0xD87CFD47u, 0x3C208C16u, 0x6871CA8Du, 0x97816A91u, 0x8181585Du, 0xB85045B6u, 0xE131A029u, 0x30644E72u)
    BBG_CONST8(P2, 0xB0F9FA8Eu, 0x7841182Du, 0xD0E3951Au, 0x2F02D522u, 0x0302B0BBu, 0x70A08B6Du, 0xC2634053u, 0x60C89CE5u)
    BBG_CONST8(R2, 0x538AFA89u, 0xF32CFC5Bu, 0xD44501FBu, 0xB5E71911u, 0x0A417FF6u, 0x47AB1EFFu, 0xCAB8351Fu, 0x06D89F71u)
    BBG_CONST8(ONE, 0xC58F0D9Du, 0xD35D438Du, 0xF5C70B3Du, 0x0A78EB28u, 0x7879462Cu, 0x666EA36Fu, 0x9A07DF2Fu, 0x0E0A77C1u)
    // cube root of unity beta (fq.hpp:53-56), Montgomery form
    BBG_CONST8(CUBE, 0xD782E155u, 0x71930C11u, 0xFFBE3323u, 0xA6BB947Cu, 0xD4741444u, 0xAA303344u, 0x26594943u, 0x2C3B3F0Du)
    static constexpr uint32_t NINV = 0xE4866389u; // low word of r_inv (fq.hpp:64): -p^-1 mod 2^32
    BBG_CONST8(NPRIME, 0xE4866389u, 0x87D20782u, 0x1ECA6AC9u, 0x9EDE7D65u, 0x1833DA80u, 0xD8AFCBD0u, 0x91888C6Bu, 0xF57A22B7u) // -p^-1 mod 2^256
};
struct FrParams
{
    BBG_CONST8(P, 0xF0000001u, 0x43E1F593u, 0x79B97091u, 0x2833E848u, 0x8181585Du, 0xB85045B6u, 0xE131A029u, 0x30644E72u)
    BBG_CONST8(P2, 0xE0000002u, 0x87C3EB27u, 0xF372E122u, 0x5067D090u, 0x0302B0BAu, 0x70A08B6Du, 0xC2634053u, 0x60C89CE5u)
    BBG_CONST8(R2, 0xAE216DA7u, 0x1BB8E645u, 0xE35C59E3u, 0x53FE3AB1u, 0x53BB8085u, 0x8C49833Du, 0x7F4E44A5u, 0x0216D0B1u)
    BBG_CONST8(ONE, 0x4FFFFFFBu, 0xAC96341Cu, 0x9F60CD29u, 0x36FC7695u, 0x7879462Eu, 0x666EA36Fu, 0x9A07DF2Fu, 0x0E0A77C1u)
    // lambda (fr.hpp:54-57, stored by the reference as fr::beta), Montgomery form
    BBG_CONST8(CUBE, 0x4A0329B3u, 0x93E7CEDEu, 0x7A96C167u, 0x7D4FDCA7u, 0xB19A750Au, 0x8BE4BA08u, 0xA5661C25u, 0x1CBD5653u)
    static constexpr uint32_t NINV = 0xEFFFFFFFu; // low word of r_inv (fr.hpp:81)
    BBG_CONST8(NPRIME, 0xEFFFFFFFu, 0xC2E1F593u, 0x4C6911B3u, 0x6586864Bu, 0x99062391u, 0xE39A9828u, 0x0D8341B2u, 0x73F82F1Du) // -p^-1 mod 2^256
};

// ---------------------------------------------------------------------------------------------
// Carry-chain primitives
// ---------------------------------------------------------------------------------------------
namespace cc
{
// r = a + b (256 bit); returns carry out
BBG_HD uint32_t add8(uint32_t* r, const uint32_t* a, const uint32_t* b)
{
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm volatile("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t t = 0;
    for (int i = 0; i < 8; ++i)
    {
        t += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)t;
        t >>= 32;
    }
    c = (uint32_t)t;
#endif
    return c;
}

// r = a - b (256 bit); returns borrow out as 0 / 0xffffffff
BBG_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b)
{
    uint32_t c;
#if defined(__CUDA_ARCH__)
    asm volatile("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(c)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]),
          "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t borrow = 0;
    for (int i = 0; i < 8; ++i)
    {
        uint64_t t = (uint64_t)a[i] - b[i] - borrow;
        r[i] = (uint32_t)t;
        borrow = (t >> 32) & 1;
    }
    c = borrow ? 0xffffffffu : 0u;
#endif
    return c;
}

// acc (four 64-bit columns acc[2j+1]:acc[2j]) = x[2j] * y        (no carries between columns)
BBG_HD void mul_row(uint32_t* acc, const uint32_t* x, uint32_t y)
{
#pragma unroll
    for (int j = 0; j < 8; j += 2)
    {
        uint64_t t = (uint64_t)x[j] * y;
        acc[j] = (uint32_t)t;
        acc[j + 1] = (uint32_t)(t >> 32);
    }
}

// acc += sum_j x[2j] * y * 2^(64 j) with one carry chain; the carry out of bit 256 is added to `top`
BBG_HD void mad_row_carry(uint32_t* acc, uint32_t& top, const uint32_t* x, uint32_t y)
{
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(y));
#else
    uint64_t carry = 0;
    for (int j = 0; j < 8; j += 2)
    {
        uint64_t prod = (uint64_t)x[j] * y;
        uint64_t lo = (uint64_t)acc[j] + (uint32_t)prod + carry;
        acc[j] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[j + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        acc[j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    top += (uint32_t)carry;
#endif
}

// same, carry out of bit 256 dropped (callers guarantee it is zero)
BBG_HD void mad_row(uint32_t* acc, const uint32_t* x, uint32_t y)
{
#if defined(__CUDA_ARCH__)
    asm volatile("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7])
        : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(y));
#else
    uint32_t dropped = 0;
    mad_row_carry(acc, dropped, x, y);
#endif
}

// Column shift fused with a row of products (the "T >>= 32" of CIOS):
//   e0  += sh[1]                                    (leftover low word joins the new even column 0)
//   sh   = (sh >> 64) + sum_j x[2j] * y * 2^(64 j) + carry of the line above
// `sh` holds the old even accumulator on entry and the new odd accumulator on exit.
BBG_HD void shift_mad_row(uint32_t* sh, uint32_t& e0, const uint32_t* x, uint32_t y)
{
#if defined(__CUDA_ARCH__)
    asm volatile("add.cc.u32 %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9, %13, %2;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
        "madc.hi.u32 %7, %12, %13, 0;"
        : "+r"(sh[0]), "+r"(sh[1]), "+r"(sh[2]), "+r"(sh[3]), "+r"(sh[4]), "+r"(sh[5]), "+r"(sh[6]), "+r"(sh[7]), "+r"(e0)
        : "r"(x[0]), "r"(x[2]), "r"(x[4]), "r"(x[6]), "r"(y));
#else
    uint64_t t = (uint64_t)e0 + sh[1];
    e0 = (uint32_t)t;
    uint64_t carry = t >> 32;
    for (int j = 0; j < 8; j += 2)
    {
        uint64_t prod = (uint64_t)x[j] * y;
        uint32_t in_lo = (j + 2 < 8) ? sh[j + 2] : 0u;
        uint32_t in_hi = (j + 3 < 8) ? sh[j + 3] : 0u;
        uint64_t lo = (uint64_t)in_lo + (uint32_t)prod + carry;
        uint64_t hi = (uint64_t)in_hi + (uint32_t)(prod >> 32) + (lo >> 32);
        sh[j] = (uint32_t)lo;
        sh[j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
#endif
}

// ---- single-instruction carry-chain pieces (device only; used by the truncated products of Field::mul_const) -------
#if defined(__CUDA_ARCH__)
BBG_D void mad_lo_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b)); }
BBG_D void madc_lo_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b)); }
BBG_D void madc_hi_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(r) : "r"(a), "r"(b)); }
BBG_D void addc_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
BBG_D void add_cc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }
BBG_D void addc(uint32_t& r, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); }

// acc[LIMB ..] += x[J] * y * 2^(32 LIMB) + x[J+2] * y * 2^(32 (LIMB+2)) + ...  as ONE carry chain over the products whose
// low word lies below limb END; the product that straddles END contributes its low word only.  FIRST starts the chain.
template <int J, int LIMB, int END, bool FIRST> BBG_D void mad_chain(uint32_t* acc, const uint32_t* x, uint32_t y)
{
    if constexpr (J < 8 && LIMB < END)
    {
        if constexpr (FIRST) mad_lo_cc(acc[LIMB], x[J], y);
        else madc_lo_cc(acc[LIMB], x[J], y);
        if constexpr (LIMB + 1 < END)
        {
            madc_hi_cc(acc[LIMB + 1], x[J], y);
            mad_chain<J + 2, LIMB + 2, END, false>(acc, x, y);
        }
    }
    else if constexpr (LIMB < END)
    {
        addc(acc[LIMB], acc[LIMB], 0u); // ran out of words of x below END: park the carry in the next (so far small) limb
    }
}
#endif

// ---- rows with the first SKIP products known to be zero (Field::sqr: row i only has the words j >= i) ---------------
// acc += sum_{j >= SKIP} x[2j] * y * 2^(64 j), carry out of bit 256 added to `top`
template <int SKIP> BBG_HD void mad_row_carry_skip(uint32_t* acc, uint32_t& top, const uint32_t* x, uint32_t y)
{
    if constexpr (SKIP >= 4) return;
#if defined(__CUDA_ARCH__)
    mad_lo_cc(acc[2 * SKIP], x[2 * SKIP], y);
    madc_hi_cc(acc[2 * SKIP + 1], x[2 * SKIP], y);
    if constexpr (SKIP < 3) { madc_lo_cc(acc[2 * SKIP + 2], x[2 * SKIP + 2], y); madc_hi_cc(acc[2 * SKIP + 3], x[2 * SKIP + 2], y); }
    if constexpr (SKIP < 2) { madc_lo_cc(acc[2 * SKIP + 4], x[2 * SKIP + 4], y); madc_hi_cc(acc[2 * SKIP + 5], x[2 * SKIP + 4], y); }
    if constexpr (SKIP < 1) { madc_lo_cc(acc[2 * SKIP + 6], x[2 * SKIP + 6], y); madc_hi_cc(acc[2 * SKIP + 7], x[2 * SKIP + 6], y); }
    addc(top, top, 0u);
#else
    uint64_t carry = 0;
    for (int j = 2 * SKIP; j < 8; j += 2)
    {
        uint64_t prod = (uint64_t)x[j] * y;
        uint64_t lo = (uint64_t)acc[j] + (uint32_t)prod + carry;
        acc[j] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[j + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        acc[j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    top += (uint32_t)carry;
#endif
}
// shift_mad_row with the first SKIP products zero: those columns are plain shifted adds
template <int SKIP> BBG_HD void shift_mad_row_skip(uint32_t* sh, uint32_t& e0, const uint32_t* x, uint32_t y)
{
#if defined(__CUDA_ARCH__)
    add_cc(e0, e0, sh[1]);
#pragma unroll
    for (int j = 0; j < 4; ++j)
    {
        const uint32_t lo_in = (2 * j + 2 < 8) ? sh[2 * j + 2] : 0u;
        const uint32_t hi_in = (2 * j + 3 < 8) ? sh[2 * j + 3] : 0u;
        if (j < SKIP)
        {
            addc_cc(sh[2 * j], lo_in, 0u);
            addc_cc(sh[2 * j + 1], hi_in, 0u);
        }
        else
        {
            sh[2 * j] = lo_in;
            madc_lo_cc(sh[2 * j], x[2 * j], y);
            sh[2 * j + 1] = hi_in;
            madc_hi_cc(sh[2 * j + 1], x[2 * j], y);
        }
    }
#else
    uint64_t t = (uint64_t)e0 + sh[1];
    e0 = (uint32_t)t;
    uint64_t carry = t >> 32;
    for (int j = 0; j < 8; j += 2)
    {
        uint64_t prod = (j >= 2 * SKIP) ? (uint64_t)x[j] * y : 0;
        uint32_t in_lo = (j + 2 < 8) ? sh[j + 2] : 0u;
        uint32_t in_hi = (j + 3 < 8) ? sh[j + 3] : 0u;
        uint64_t lo = (uint64_t)in_lo + (uint32_t)prod + carry;
        uint64_t hi = (uint64_t)in_hi + (uint32_t)(prod >> 32) + (lo >> 32);
        sh[j] = (uint32_t)lo;
        sh[j + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
#endif
}
} // namespace cc

// ---------------------------------------------------------------------------------------------
// Field operations
// ---------------------------------------------------------------------------------------------
template <typename FP> struct Field
{
    typedef FP params;

    static BBG_HD fe zero()
    {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = 0;
        return r;
    }
    template <typename Fn> static BBG_HD fe constant(Fn c)
    {
        fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = c(i);
        return r;
    }
    static BBG_HD fe one() { return constant([](int i) { return FP::ONE(i); }); }
    static BBG_HD fe modulus() { return constant([](int i) { return FP::P(i); }); }

    // Montgomery product, result in [0, 2p) for inputs in [0, 2p).
    // (reference: __mul_with_coarse_reduction, field_impl_asm.tcc:332-348 / field_impl_int128.tcc:254-263;
    //  same integer (ab + Mp)/2^256 as the 64-bit-limb reduction since M = -ab/p mod 2^256 either way)
    static BBG_HD fe mul(const fe& a, const fe& b)
    {
        uint32_t A[8], B[8];            // the two column accumulators; roles swap every word of b
        const uint32_t* ae = a.v;       // a[0], a[2], a[4], a[6]
        const uint32_t* ao = a.v + 1;   // a[1], a[3], a[5], a[7]
        uint32_t pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = FP::P(i);
        const uint32_t* pe = pl;     // p[0], p[2], p[4], p[6]
        const uint32_t* po = pl + 1; // p[1], p[3], p[5], p[7]

        // word 0: even = A, odd = B
        cc::mul_row(A, ae, b.v[0]);
        cc::mul_row(B, ao, b.v[0]);
        {
            uint32_t m = A[0] * FP::NINV;
            cc::mad_row(B, po, m);
            cc::mad_row_carry(A, B[7], pe, m);
        }
#pragma unroll
        for (int i = 1; i < 8; i += 2)
        {
            // odd word i: even = B, odd = A (A is shifted down two words on the way)
            {
                const uint32_t bi = b.v[i];
                cc::shift_mad_row(A, B[0], ao, bi);
                cc::mad_row_carry(B, A[7], ae, bi);
                uint32_t m = B[0] * FP::NINV;
                cc::mad_row(A, po, m);
                cc::mad_row_carry(B, A[7], pe, m);
            }
            if (i + 1 < 8)
            {
                // even word i+1: even = A, odd = B
                const uint32_t bi = b.v[i + 1];
                cc::shift_mad_row(B, A[0], ao, bi);
                cc::mad_row_carry(A, B[7], ae, bi);
                uint32_t m = A[0] * FP::NINV;
                cc::mad_row(B, po, m);
                cc::mad_row_carry(A, B[7], pe, m);
            }
        }
        // after word 7: even = B with B[0] == 0, odd = A.  result[k] = B[k+1] + A[k]
        fe r;
        uint32_t hi[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) hi[k] = B[k + 1];
        hi[7] = 0;
        cc::add8(r.v, A, hi);
        return r;
    }
    // (a b + c d) / 2^256 mod p in [0, 2p) for inputs in [0, 2p): two products under ONE interleaved reduction — word i adds
    // a b_i + c d_i and then a single multiple of p, 8 x (16 + 9) = 200 wide products where two Montgomery products take 272.
    // Bounds: the running sum stays below 5p (2^32 + 1) < 2^288, which is what the even / odd accumulator pair holds, and the
    // result below p (8p / 2^256 + 1) < 2.52p, so one conditional subtraction of 2p brings it back into the lazy range.
    // The mixed addition's y3 = R (Q - x3) - y1 PPP is the customer (bbg_g1.cuh).
    static BBG_HD fe mul2(const fe& a, const fe& b, const fe& c, const fe& d)
    {
        uint32_t A[8], B[8];
        const uint32_t *ae = a.v, *ao = a.v + 1, *ce = c.v, *co = c.v + 1;
        uint32_t pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = FP::P(i);
        const uint32_t *pe = pl, *po = pl + 1;
        // word 0: even = A, odd = B
        cc::mul_row(A, ae, b.v[0]);
        cc::mul_row(B, ao, b.v[0]);
        cc::mad_row(B, co, d.v[0]);
        cc::mad_row_carry(A, B[7], ce, d.v[0]);
        {
            uint32_t m = A[0] * FP::NINV;
            cc::mad_row(B, po, m);
            cc::mad_row_carry(A, B[7], pe, m);
        }
#pragma unroll
        for (int i = 1; i < 8; i += 2)
        {
            {
                // odd word i: even = B, odd = A (A is shifted down two words on the way)
                const uint32_t bi = b.v[i], di = d.v[i];
                cc::shift_mad_row(A, B[0], ao, bi);
                cc::mad_row_carry(B, A[7], ae, bi);
                cc::mad_row(A, co, di);
                cc::mad_row_carry(B, A[7], ce, di);
                uint32_t m = B[0] * FP::NINV;
                cc::mad_row(A, po, m);
                cc::mad_row_carry(B, A[7], pe, m);
            }
            if (i + 1 < 8)
            {
                // even word i+1: even = A, odd = B
                const uint32_t bi = b.v[i + 1], di = d.v[i + 1];
                cc::shift_mad_row(B, A[0], ao, bi);
                cc::mad_row_carry(A, B[7], ae, bi);
                cc::mad_row(B, co, di);
                cc::mad_row_carry(A, B[7], ce, di);
                uint32_t m = A[0] * FP::NINV;
                cc::mad_row(B, po, m);
                cc::mad_row_carry(A, B[7], pe, m);
            }
        }
        // after word 7: even = B with B[0] == 0, odd = A.  result[k] = B[k+1] + A[k], then back below 2p
        fe s, t;
        uint32_t hi[8], p2[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) hi[k] = B[k + 1];
        hi[7] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) p2[i] = FP::P2(i);
        cc::add8(s.v, A, hi);
        const uint32_t borrow = cc::sub8(t.v, s.v, p2);
        fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = borrow ? s.v[i] : t.v[i];
        return r;
    }
    // Montgomery square, result in [0, 2p) for a in [0, 2p): the same interleaved reduction as mul, but row i multiplies
    // a_i by  a_i B^i + 2 (a div B^(i+1)) B^(i+1)  (B = 2^32), i.e. by the words  a_i, a_(i+1) << 1, dd_(i+2) .. dd_7  with
    // dd_j = (a_j << 1) | (a_(j-1) >> 31) the words of 2a (2a < 2^256 since a < 2p < 2^255): every product a_i a_j is issued
    // once instead of twice - 36 + 64 wide products instead of 64 + 64; the skipped columns of the shifting accumulator
    // become plain carry adds.  Same integer (a^2 + M p) / 2^256 as mul(a, a).
    static BBG_HD fe sqr(const fe& a)
    {
        uint32_t A[8], B[8], X[9];
        uint32_t pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = FP::P(i);
        const uint32_t* pe = pl;
        const uint32_t* po = pl + 1;
        uint32_t dd[8];
        dd[0] = 0;
#pragma unroll
        for (int j = 1; j < 8; ++j) dd[j] = (a.v[j] << 1) | (a.v[j - 1] >> 31);
        X[8] = 0;
        // word 0: X = a_0, a_1 << 1, dd_2 .. dd_7
        X[0] = a.v[0];
        X[1] = a.v[1] << 1;
#pragma unroll
        for (int j = 2; j < 8; ++j) X[j] = dd[j];
        cc::mul_row(A, X, a.v[0]);
        cc::mul_row(B, X + 1, a.v[0]);
        {
            uint32_t m = A[0] * FP::NINV;
            cc::mad_row(B, po, m);
            cc::mad_row_carry(A, B[7], pe, m);
        }
        sqr_words<1>(A, B, X, a.v, pe, po);
        fe r;
        uint32_t hi[8];
#pragma unroll
        for (int k = 0; k < 7; ++k) hi[k] = B[k + 1];
        hi[7] = 0;
        cc::add8(r.v, A, hi);
        return r;
    }
    // words I (odd) and I + 1 (even) of the square; X[j] still holds dd_j for j >= I + 1 (a row only rewrites X[i], X[i+1])
    template <int I> static BBG_HD void sqr_words(uint32_t* A, uint32_t* B, uint32_t* X, const uint32_t* a, const uint32_t* pe, const uint32_t* po)
    {
        if constexpr (I < 8)
        {
            {
                // odd word I: even accumulator = B, odd accumulator = A (shifted down two words on the way)
                X[I] = a[I];
                if constexpr (I + 1 < 8) X[I + 1] = a[I + 1] << 1;
                const uint32_t y = a[I];
                cc::shift_mad_row_skip<(I - 1) / 2>(A, B[0], X + 1, y);
                cc::mad_row_carry_skip<(I + 1) / 2>(B, A[7], X, y);
                uint32_t m = B[0] * FP::NINV;
                cc::mad_row(A, po, m);
                cc::mad_row_carry(B, A[7], pe, m);
            }
            if constexpr (I + 1 < 8)
            {
                constexpr int E = I + 1;
                X[E] = a[E];
                if constexpr (E + 1 < 8) X[E + 1] = a[E + 1] << 1;
                const uint32_t y = a[E];
                cc::shift_mad_row_skip<E / 2>(B, A[0], X + 1, y);
                cc::mad_row_carry_skip<E / 2>(A, B[7], X, y);
                uint32_t m = A[0] * FP::NINV;
                cc::mad_row(B, po, m);
                cc::mad_row_carry(A, B[7], pe, m);
            }
            sqr_words<I + 2>(A, B, X, a, pe, po);
        }
    }

    // Product with a constant known in advance (NTT twiddles): r = a * w mod p as a residue in [0, 2p), for ANY a < 4p
    // and w < p given in PLAIN form together with wq = floor(w * 2^256 / p).  When a is a Montgomery-form value, so is
    // the result (a R * w = (a w) R), i.e. this replaces mul(a, to_mont(w)) and returns the same residue class.
    //   q = floor(a * wq / 2^256) from the product columns >= 6 only (the dropped columns are worth < 2^-28 of a unit),
    //   r = a * w - q * p  (mod 2^256);   a w / p - (a wq + dropped) / 2^256 < 4p / 2^256 + 2^-28 < 1  =>  r in [0, 2p).
    // 43 + 28 + 28 wide products and 16 low-word products instead of the Montgomery product's 128 + 8: the multiply pipe
    // has ~25% less to issue (the quotient needs no low half, the remainder no high half).
    static BBG_HD fe mul_const(const fe& a, const fe& w, const fe& wq)
    {
        fe r;
#if defined(__CUDA_ARCH__)
        // ---- q: limbs 8..15 of a * wq; E holds the products at even limb positions (limbs 6..15), O the odd (7..15)
        uint32_t E[16], O[16];
#pragma unroll
        for (int i = 6; i < 16; ++i) E[i] = O[i] = 0;
        mul_const_hi<0>(E, O, a.v, wq.v);
        uint32_t q[8], t7;
        cc::add_cc(t7, E[7], O[7]);
#pragma unroll
        for (int i = 0; i < 7; ++i) cc::addc_cc(q[i], E[8 + i], O[8 + i]);
        cc::addc(q[7], E[15], O[15]);
        (void)t7;
        // ---- r = low 256 bits of a * w + q * (2^256 - p)
        uint32_t L[8], M[8], np[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            L[i] = M[i] = 0;
            np[i] = i == 0 ? 0u - FP::P(0) : ~FP::P(i);
        }
        mul_const_lo<0>(L, M, a.v, w.v);
        mul_const_lo<0>(L, M, np, q);
        r.v[0] = L[0];
        cc::add_cc(r.v[1], L[1], M[1]);
#pragma unroll
        for (int i = 2; i < 7; ++i) cc::addc_cc(r.v[i], L[i], M[i]);
        cc::addc(r.v[7], L[7], M[7]);
#else
        // portable restatement with the same truncation (columns >= 6 of a * wq)
        uint64_t col[17];
        for (int k = 0; k < 17; ++k) col[k] = 0;
        for (int i = 0; i < 8; ++i)
        {
            for (int j = 0; j < 8; ++j)
            {
                if (i + j < 6) continue;
                const uint64_t t = (uint64_t)a.v[j] * wq.v[i];
                col[i + j] += (uint32_t)t;
                col[i + j + 1] += t >> 32;
            }
        }
        uint32_t q[8];
        uint64_t carry = 0;
        for (int k = 6; k < 16; ++k)
        {
            const uint64_t t = col[k] + carry;
            if (k >= 8) q[k - 8] = (uint32_t)t;
            carry = t >> 32;
        }
        uint64_t low[9];
        for (int k = 0; k < 9; ++k) low[k] = 0;
        for (int i = 0; i < 8; ++i)
        {
            const uint32_t np_i = i == 0 ? 0u - FP::P(0) : ~FP::P(i);
            for (int j = 0; i + j < 8; ++j)
            {
                const uint64_t t = (uint64_t)a.v[j] * w.v[i];
                const uint64_t u = (uint64_t)q[j] * np_i;
                low[i + j] += (uint64_t)(uint32_t)t + (uint32_t)u;
                low[i + j + 1] += (t >> 32) + (u >> 32);
            }
        }
        carry = 0;
        for (int k = 0; k < 8; ++k)
        {
            const uint64_t t = low[k] + carry;
            r.v[k] = (uint32_t)t;
            carry = t >> 32;
        }
#endif
        return r;
    }
    // wq = floor(w * 2^256 / p) for mul_const, from the canonical Montgomery form of w:  w 2^256 = wq p + w_mont, so
    // wq = -w_mont / p = w_mont * NPRIME (mod 2^256).  Table generation only - plain 64-bit arithmetic.
    static BBG_HD fe const_quotient(const fe& w_mont)
    {
        uint64_t low[9];
        for (int k = 0; k < 9; ++k) low[k] = 0;
        for (int i = 0; i < 8; ++i)
        {
            for (int j = 0; i + j < 8; ++j)
            {
                const uint64_t t = (uint64_t)w_mont.v[j] * FP::NPRIME(i);
                low[i + j] += (uint32_t)t;
                low[i + j + 1] += t >> 32;
            }
        }
        fe r;
        uint64_t carry = 0;
        for (int k = 0; k < 8; ++k)
        {
            const uint64_t t = low[k] + carry;
            r.v[k] = (uint32_t)t;
            carry = t >> 32;
        }
        return r;
    }
#if defined(__CUDA_ARCH__)
    // rows of the quotient estimate: word I of wq times the words of a that reach column 6 or above
    template <int I> static BBG_D void mul_const_hi(uint32_t* E, uint32_t* O, const uint32_t* a, const uint32_t* wq)
    {
        if constexpr (I < 8)
        {
            constexpr int JE = I <= 6 ? 6 - I : 1; // first word of a with i + j even and >= 6
            constexpr int JO = 7 - I;              // first word of a with i + j odd and >= 7
            cc::mad_chain<JE, I + JE, 16, true>(E, a, wq[I]);
            cc::mad_chain<JO, I + JO, 16, true>(O, a, wq[I]);
            mul_const_hi<I + 1>(E, O, a, wq);
        }
    }
    // rows of a low (mod 2^256) product: L += even-position products, M += odd-position products
    template <int I> static BBG_D void mul_const_lo(uint32_t* L, uint32_t* M, const uint32_t* x, const uint32_t* y)
    {
        if constexpr (I < 8)
        {
            cc::mad_chain<(I & 1), I + (I & 1), 8, true>(L, x, y[I]);
            cc::mad_chain<((I + 1) & 1), I + ((I + 1) & 1), 8, true>(M, x, y[I]);
            mul_const_lo<I + 1>(L, M, x, y);
        }
    }
#endif

    // [0,2p) -> [0,p)   (reference: reduce_once)
    static BBG_HD fe reduce(const fe& a)
    {
        fe t;
        uint32_t pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = FP::P(i);
        uint32_t borrow = cc::sub8(t.v, a.v, pl);
        fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = borrow ? a.v[i] : t.v[i];
        return r;
    }
    // a + b mod 2p, inputs/outputs in [0,2p)   (reference: __add_with_coarse_reduction)
    static BBG_HD fe add(const fe& a, const fe& b)
    {
        fe s, t;
        uint32_t p2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p2[i] = FP::P2(i);
        cc::add8(s.v, a.v, b.v);
        uint32_t borrow = cc::sub8(t.v, s.v, p2);
        fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = borrow ? s.v[i] : t.v[i];
        return r;
    }
    // a - b mod 2p, inputs/outputs in [0,2p)   (reference: __sub_with_coarse_reduction)
    static BBG_HD fe sub(const fe& a, const fe& b)
    {
        fe d, t;
        uint32_t p2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p2[i] = FP::P2(i);
        uint32_t borrow = cc::sub8(d.v, a.v, b.v);
        cc::add8(t.v, d.v, p2);
        fe r;
#pragma unroll
        for (int i = 0; i < 8; ++i) r.v[i] = borrow ? t.v[i] : d.v[i];
        return r;
    }
    // a - b + 2p without the conditional correction: result in (0, 4p) for a, b in [0, 2p).  Only valid as the
    // multiplicand of a product whose other factor is canonical (< p): then (4p * p) / 2^256 + p < 2p still holds.
    static BBG_HD fe sub_lazy(const fe& a, const fe& b)
    {
        fe t, r;
        uint32_t p2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) p2[i] = FP::P2(i);
        cc::add8(t.v, a.v, p2);
        cc::sub8(r.v, t.v, b.v);
        return r;
    }
    static BBG_HD fe dbl(const fe& a) { return add(a, a); }
    // -a in [0,2p): 2p - a  (a in [0,2p]);  maps 0 -> 2p?  no: 0 stays 0
    static BBG_HD fe neg(const fe& a)
    {
        fe z = zero();
        return sub(z, a);
    }
    // canonical results
    static BBG_HD fe mul_full(const fe& a, const fe& b) { return reduce(mul(a, b)); }     // reference __mul
    static BBG_HD fe to_mont(const fe& a) { return mul_full(a, constant([](int i) { return FP::R2(i); })); }       // field.hpp:224-232
    static BBG_HD fe from_mont(const fe& a)                                               // field.hpp:233-236
    {
        fe o = zero();
        o.v[0] = 1;
        return mul_full(a, o);
    }
    static BBG_HD bool is_zero_raw(const fe& a)
    {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= a.v[i];
        return o == 0;
    }
    // a == 0 (mod p) for a in [0,2p)
    static BBG_HD bool is_zero(const fe& a) { return is_zero_raw(reduce(a)); }
    static BBG_HD bool eq_raw(const fe& a, const fe& b)
    {
        uint32_t o = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) o |= a.v[i] ^ b.v[i];
        return o == 0;
    }
    // a^e for a small exponent, square-and-multiply msb first; coarse result
    static BBG_HD fe pow_u64(const fe& a, uint64_t e)
    {
        fe acc = one();
        bool started = false;
        for (int i = 63; i >= 0; --i)
        {
            if (started) acc = sqr(acc);
            if ((e >> i) & 1)
            {
                acc = started ? mul(acc, a) : a;
                started = true;
            }
        }
        return acc;
    }
    // Fermat inversion a^(p-2) (reference field.hpp:345-348); canonical result
    static BBG_HD fe invert(const fe& a)
    {
        uint32_t e[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) e[i] = FP::P(i);
        e[0] -= 2; // neither modulus has a low word < 2
        fe acc = one();
        bool started = false;
        for (int i = 255; i >= 0; --i)
        {
            if (started) acc = sqr(acc);
            if ((e[i >> 5] >> (i & 31)) & 1)
            {
                acc = started ? mul(acc, a) : a;
                started = true;
            }
        }
        return reduce(acc);
    }
    // The same inverse by Kaliski's almost-inverse (binary extended Euclid on shifts, additions and subtractions only: no
    // multiply-pipe work, ~1.4 * 254 short iterations) and a correcting power of two: phase one gives a^-1 * 2^k mod p with
    // 254 <= k <= 508, then * (1/2)^k and * R^3 as Montgomery products restore the Montgomery form.  About a seventh of the
    // Fermat chain's latency for a lone thread (what a batched inversion waits for); data-dependent time, which is fine here:
    // nothing on this path is secret.  a = 0 gives 0, as a^(p-2) does.  Canonical result.
    static BBG_HD fe invert_binary(const fe& a_in)
    {
        const fe a = reduce(a_in);
        uint32_t u[8], v[8], r[8], s[8], t[8];
#pragma unroll
        for (int i = 0; i < 8; ++i)
        {
            u[i] = FP::P(i);
            v[i] = a.v[i];
            r[i] = 0;
            s[i] = 0;
        }
        s[0] = 1;
        uint32_t k = 0;
        // invariants: u s + v r = p, so r, s <= p while the loop runs and r < 2p after it
        while ((v[0] | v[1] | v[2] | v[3] | v[4] | v[5] | v[6] | v[7]) != 0)
        {
            if ((u[0] & 1u) == 0)
            {
                shr1(u);
                shl1(s);
            }
            else if ((v[0] & 1u) == 0)
            {
                shr1(v);
                shl1(r);
            }
            else
            {
                const uint32_t borrow = cc::sub8(t, u, v);
                if (!borrow && (t[0] | t[1] | t[2] | t[3] | t[4] | t[5] | t[6] | t[7]) != 0) // u > v
                {
                    shr1(t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) u[i] = t[i];
                    cc::add8(t, r, s);
#pragma unroll
                    for (int i = 0; i < 8; ++i) r[i] = t[i];
                    shl1(s);
                }
                else
                {
                    cc::sub8(t, v, u);
                    shr1(t);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] = t[i];
                    cc::add8(t, s, r);
#pragma unroll
                    for (int i = 0; i < 8; ++i) s[i] = t[i];
                    shl1(r);
                }
            }
            ++k;
        }
        uint32_t pl[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) pl[i] = FP::P(i);
        if (!cc::sub8(t, r, pl)) // r >= p
        {
#pragma unroll
            for (int i = 0; i < 8; ++i) r[i] = t[i];
        }
        fe x;
        cc::sub8(x.v, pl, r); // p - r = a^-1 2^k (as plain residues); r = 0 only for a = 0
        // Montgomery form of 1/2: (R mod p) halved modulo p
        fe half = one();
        if (half.v[0] & 1u)
        {
            cc::add8(t, half.v, pl);
#pragma unroll
            for (int i = 0; i < 8; ++i) half.v[i] = t[i];
        }
        shr1(half.v);
        const fe r2 = constant([](int i) { return FP::R2(i); });
        // x = A^-1 2^k with A = a R the input:  x * (2^-k R) / R = a^-1 / R,  then * (R^2 * R^2 / R) / R = a^-1 R
        return reduce(mul(mul(x, pow_u64(half, k)), mul(r2, r2)));
    }
    static BBG_HD void shr1(uint32_t* x)
    {
#pragma unroll
        for (int i = 0; i < 7; ++i) x[i] = (x[i] >> 1) | (x[i + 1] << 31);
        x[7] >>= 1;
    }
    static BBG_HD void shl1(uint32_t* x)
    {
#pragma unroll
        for (int i = 7; i > 0; --i) x[i] = (x[i] << 1) | (x[i - 1] >> 31);
        x[0] <<= 1;
    }
};

typedef Field<FqParams> Fq;
typedef Field<FrParams> Fr;

// 16-byte vector load/store of a field element (the reference's field_t is 32-byte aligned)
BBG_HD fe load_fe(const void* p)
{
    const uint4* q = (const uint4*)p;
    uint4 lo = q[0], hi = q[1];
    fe r;
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}
BBG_HD void store_fe(void* p, const fe& a)
{
    uint4* q = (uint4*)p;
    uint4 lo, hi;
    lo.x = a.v[0]; lo.y = a.v[1]; lo.z = a.v[2]; lo.w = a.v[3];
    hi.x = a.v[4]; hi.y = a.v[5]; hi.z = a.v[6]; hi.w = a.v[7];
    q[0] = lo;
    q[1] = hi;
}

// The same for pointers KNOWN to be global memory, 32-byte aligned: one 256-bit access per element (sm_100 has
// ld / st.global.v8.b32 -> LDG.E.256 / STG.E.256): half the load / store instructions of the two-quadword form and a whole
// 32-byte sector per lane, which is what the strided accesses of the NTT passes and the MSM's point gathers are made of.
BBG_HD fe load_fe_global(const void* p)
{
#if defined(__CUDA_ARCH__)
    fe r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p)
                 : "memory");
    return r;
#else
    return load_fe(p);
#endif
}
// read-only data (tables, matrices, points) through the non-coherent path.  volatile on purpose: a plain asm counts as
// speculatable, and the compiler did move such a load above the null check that guards an optional table (illegal address)
BBG_HD fe load_fe_const(const void* p)
{
#if defined(__CUDA_ARCH__)
    fe r;
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
        : "l"(p));
    return r;
#else
    return load_fe(p);
#endif
}
// 256-bit load through the ordinary (coherent) path, no ordering against the surrounding stores
BBG_HD fe load_fe_wide(const void* p)
{
#if defined(__CUDA_ARCH__)
    fe r;
    asm volatile("ld.global.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]), "=r"(r.v[6]), "=r"(r.v[7])
                 : "l"(p));
    return r;
#else
    return load_fe(p);
#endif
}
BBG_HD void store_fe_global(void* p, const fe& a)
{
#if defined(__CUDA_ARCH__)
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]),
                 "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])
                 : "memory");
#else
    store_fe(p, a);
#endif
}

} // namespace bbg
