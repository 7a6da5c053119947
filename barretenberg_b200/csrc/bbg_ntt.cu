// Radix-2 Fr NTT family for sm_100a.
//
// Replaces polynomial_arithmetic::fft / ifft / coset_fft / coset_ifft / *_with_constant
//   (reference polynomials/polynomial_arithmetic.cpp:129-264 fft_inner_parallel, :81-102
//    scale_by_generator, :266-315 wrappers) on an evaluation_domain of size n = 2^log_n.
// Contract kept from the reference: natural order in and out, X_i = sum_j x_j w^(ij), inputs may be
// lazily reduced in [0,2p), outputs are canonical Montgomery limbs in [0,p).
//
// Algorithm (B200-first, not a port of the CPU loop nest):
//   n <= 2^11 : one CTA per polynomial, whole transform in shared memory.
//   n >= 2^12 : four-step split n = N1 * N2 in TWO HBM passes.
//       pass A  (column tiles, in -> scratch): for each column j2, N1-point DIF over j1, then multiply
//               by the inter-pass matrix M[i1][j2] = w_n^(i1 j2) (optionally carrying 1/n and the
//               coset factors) read coalesced in the same pattern as the data.
//       pass B  (row tiles, scratch -> out): for each row i1, N2-point DIF over j2, output transposed
//               to natural order X[i1 + N1 i2], canonical reduction fused into the store.
//     A tile is 2048 elements (64 KiB as eight 32-bit limb planes, XOR-swizzled against bank conflicts)
//     processed by 256 threads holding 8 elements each: radix-8 butterflies in registers, one shared
//     memory exchange per 3 stages.  Sub-transform twiddles live in shared memory as (plain value,
//     quotient) pairs for Fr::mul_const (66 KiB image: 19% fewer multiply-pipe slots per twiddle product
//     than a Montgomery product); one 512-thread CTA per SM whose two halves share the image and each
//     work on their own tile with their own named barrier (194 KiB smem, <=128 regs).  Work per element:
//     (log2 n)/2 products + 1 for the inter-pass twiddle (+1 per fused coset / constant scaling), i.e.
//     IMAD-bound, ~2 x 64 B of HBM traffic per element per pass.  Pass A starts its
//     tile's inter-pass matrix entries towards L2 when it loads the tile.
//   n >= 2^23 (up to the field's 2^28 two-adicity): one more outer split n = N0 * M, M <= 2^22: an outer pass A over
//       columns of length N0 (same kernel, matrix w_n^(i0 j)), then the two passes above on the N0 contiguous blocks of
//       M elements, whose last pass scatters block i0's output u to natural position i0 + N0 u.  THREE HBM passes;
//       coset / constant scalings run as one extra element-wise pass here (3% of the products, not worth more tables).
//   All twiddles, the inter-pass matrices and the coset power tables are generated ON THE DEVICE the
//   first time a domain size is used and cached (the reference's host round_roots tables are never
//   uploaded).
#include "bbg_internal.h"
#include "bbg_hostcopy.h"

#include <cstring>
#include <map>
#include <vector>

namespace bbg
{
namespace nttk
{
constexpr int TILE_LOG = 11;
constexpr int TILE = 1 << TILE_LOG;
constexpr int NT = 256;
constexpr int PLANE = TILE;                 // words per limb plane
constexpr int TWP = (TILE >> 1) + (TILE >> 6); // words per twiddle limb plane (padded)
// Tiles per CTA: the twiddle image holds every twiddle twice (plain value and precomputed quotient, Fr::mul_const), 66 KiB,
// so two 256-thread halves of one 512-thread CTA share it, each half working on its own tile with its own named barrier
// (2 x 64 KiB data + 66 KiB twiddles = 194 KiB, one CTA per SM; before: two 256-thread CTAs with a 33 KiB image each).
#ifdef BBG_EMULATE
constexpr int TPC = 1; // (the emulation has one barrier per block and no shared-memory limit)
#else
constexpr int TPC = 2;
#endif
constexpr size_t SMEM_BYTES = (size_t)(TPC * 8 * PLANE + 16 * TWP) * 4;
constexpr size_t SMEM_BYTES_SMALL = (size_t)(8 * PLANE + 16 * TWP) * 4;
constexpr int LO_TABLE_LOG = 11; // w^e = Thi[e >> 11] * Tlo[e & 2047] when generating matrices

#if defined(__CUDA_ARCH__) && defined(BBG_NTT_NOINLINE_MUL)
// one shared copy of the Montgomery product for the pass kernels (instruction-cache footprint)
__device__ __noinline__ fe fr_mul_shared(const fe a, const fe b) { return Fr::mul(a, b); }
#define NTT_MUL(a, b) fr_mul_shared((a), (b))
#else
#define NTT_MUL(a, b) Fr::mul((a), (b))
#endif
// BBG_NTT_ABLATE (development only, never set in the product build; results are WRONG): time the pass kernels with one
// ingredient removed to see what the multiply pipe is waiting for.  1: no CTA barriers  2: no shared-memory exchange
// 3: additions / subtractions without the modular correction  4: twiddles not loaded (constant)  5: no global traffic
#ifndef BBG_NTT_ABLATE
#define BBG_NTT_ABLATE 0
#endif
#ifndef BBG_NTT_PREFETCH
#define BBG_NTT_PREFETCH 1
#endif
// how the pass kernels read global memory: 0 two 128-bit loads, 1 one 256-bit load through the non-coherent path,
// 2 one 256-bit load through the ordinary path; per pass (A: column tiles, B: row tiles) and for pass A's matrix.
// Measured (r02, batch 8; fft 2^20 / coset_fft 2^21 / 2^22 in ms): all 128-bit 1.360 / 2.97 / 6.16; 256-bit stores and
// loads as set below 1.297 / 3.00 / 6.09 (pass B 2^22 2.70 -> 2.60, pass A 2^9-point columns 0.650 -> 0.625; pass A on
// single-column tiles does not gain, whichever way it loads)
#ifndef BBG_NTT_LOAD_A
#define BBG_NTT_LOAD_A 2
#endif
#ifndef BBG_NTT_LOAD_B
#define BBG_NTT_LOAD_B 1
#endif
#ifndef BBG_NTT_LOAD_MAT
#define BBG_NTT_LOAD_MAT 2
#endif
#ifndef BBG_NTT_STORE_A
#define BBG_NTT_STORE_A 1
#endif
#ifndef BBG_NTT_STORE_B
#define BBG_NTT_STORE_B 1
#endif
template <int MODE> BBG_D void ntt_store(fe* p, const fe& x)
{
    if constexpr (MODE == 1) store_fe_global(p, x);
    else store_fe(p, x);
}
template <int MODE> BBG_D fe ntt_load(const fe* p)
{
    if constexpr (MODE == 1) return load_fe_const(p);
    else if constexpr (MODE == 2) return load_fe_wide(p);
    else return load_fe(p);
}
#if BBG_NTT_ABLATE == 1 || BBG_NTT_ABLATE == 2
#define NTT_SYNC() ((void)0)
#elif defined(__CUDA_ARCH__)
// barrier over the 256 threads of this half of the CTA (ids 1 and 2; 0 is __syncthreads)
#define NTT_SYNC() asm volatile("bar.sync %0, %1;" ::"r"((int)(threadIdx.x / NT) + 1), "n"(NT) : "memory")
#else
#define NTT_SYNC() __syncthreads()
#endif
// Exchange between two radix steps: a thread of the next step (owned field at tile-slot bit QB2) reads what the threads
// that differ from it in thread bits [QB2, QB1) stored in this one (owned field at bit QB1 = QB2 + R).  Only those have to
// meet: a warp when QB1 <= 5, otherwise the 2^(QB1 - max(QB2, 5)) warps that share the remaining warp-index bits
// (named barriers 3..10; 1 and 2 are the whole-tile barriers of the two halves).
#ifndef BBG_NTT_FINE_SYNC
#define BBG_NTT_FINE_SYNC 1
#endif
template <int QB1, int QB2> BBG_D void exchange_sync()
{
#if BBG_NTT_ABLATE == 1 || BBG_NTT_ABLATE == 2
#elif defined(__CUDA_ARCH__) && BBG_NTT_FINE_SYNC
    if constexpr (QB1 <= 5) __syncwarp();
    else
    {
        constexpr int LO = (QB2 > 5 ? QB2 : 5) - 5, HI = QB1 - 5; // varying bits of the warp index (3 bits)
        static_assert(HI <= 3 && LO < HI, "tile of 256 threads");
        if constexpr (HI - LO == 3) NTT_SYNC();
        else
        {
            const int w = ((int)threadIdx.x >> 5) & 7;
            const int g = (w & ((1 << LO) - 1)) | ((w >> HI) << LO); // < 4
            asm volatile("bar.sync %0, %1;" ::"r"(3 + 4 * ((int)threadIdx.x / NT) + g), "n"(32 << (HI - LO)) : "memory");
        }
    }
#else
    NTT_SYNC();
#endif
}
BBG_HD int pad(int q) { return q + (q >> 5); }          // twiddle planes: power-of-two strides
// data planes: XOR swizzle of the bank bits with tile-slot bits 3..7.  A bijection on [0, TILE) that makes every
// access pattern of every radix step conflict-free (checked exhaustively for all sub-transform lengths and both
// tile layouts; the padded layout q + q/32 left 2- and 4-way conflicts in the low-bit steps, ncu r01).
BBG_HD int dswz(int q) { return q ^ ((q >> 3) & 31); }

// hint: bring the line holding p into L2 (no register, no scoreboard entry)
BBG_D void prefetch_l2(const void* p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
    (void)p;
#endif
}

// ---- next-tile staging ------------------------------------------------------------------------------------------------
// A tile's inputs travel global -> shared memory by cp.async (no registers, no scoreboard) into the tile's OWN data region
// while the previous tile's last radix step still multiplies and stores: the region is free from the moment that step has
// read it.  The staged image is raw: element of tile slot q as two 16-byte chunks 2q, 2q + 1, chunk c at chunk position
// c ^ ((c >> 3) & 1) (a warp's 128-bit loads of consecutive slots then cover all banks once per quarter warp).
// Measured (r02, tools/ntt_ablate.py, batch 8): SLOWER than loading into registers at the top of the tile - fft 2^20 1.386 ->
// 1.482 ms, coset_fft 2^22 6.19 -> 6.51 ms (two more whole-tile barriers, 100 more bytes of spills, and the tile waits as one
// for its copies where the warps used to trickle in as their own loads arrived).  Kept compiled out as the record of the
// experiment; the product build loads through registers.
#ifndef BBG_NTT_STAGE
#define BBG_NTT_STAGE 0
#endif
BBG_HD int stage_pos(int c) { return (c ^ ((c >> 3) & 1)) << 2; } // word offset of chunk c
BBG_D void cp_async16(uint32_t* smem_dst, const void* gmem_src)
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
#else
    memcpy(smem_dst, gmem_src, 16);
#endif
}
BBG_D void cp_async_commit()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
BBG_D void cp_async_wait_all()
{
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}
BBG_D fe staged_load(const uint32_t* data, int q)
{
    fe r;
    const uint4 lo = *reinterpret_cast<const uint4*>(data + stage_pos(2 * q));
    const uint4 hi = *reinterpret_cast<const uint4*>(data + stage_pos(2 * q + 1));
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
    return r;
}

BBG_D void sm_store(uint32_t* data, int q, const fe& x)
{
    const int p = dswz(q);
#pragma unroll
    for (int l = 0; l < 8; ++l) data[l * PLANE + p] = x.v[l];
}
BBG_D fe sm_load(const uint32_t* data, int q)
{
    const int p = dswz(q);
    fe r;
#pragma unroll
    for (int l = 0; l < 8; ++l) r.v[l] = data[l * PLANE + p];
    return r;
}
// twiddle e of the image: planes 0..7 = w^e in PLAIN form, planes 8..15 = floor(w^e 2^256 / p)  (Fr::mul_const)
BBG_D fe twq_load(const uint32_t* tw, int e)
{
    const int p = pad(e);
    fe r;
#pragma unroll
    for (int l = 0; l < 8; ++l) r.v[l] = tw[(8 + l) * TWP + p];
    return r;
}
BBG_D fe tw_load(const uint32_t* tw, int e)
{
#if BBG_NTT_ABLATE == 4
    fe r;
#pragma unroll
    for (int l = 0; l < 8; ++l) r.v[l] = 0x1234567u * (l + 1) + (e & 1);
    return r;
#else
    const int p = pad(e);
    fe r;
#pragma unroll
    for (int l = 0; l < 8; ++l) r.v[l] = tw[l * TWP + p];
    return r;
#endif
}
#if BBG_NTT_ABLATE == 3
BBG_D fe abl_add(const fe& a, const fe& b) { fe r; cc::add8(r.v, a.v, b.v); return r; }
BBG_D fe abl_sub(const fe& a, const fe& b) { fe r; cc::sub8(r.v, a.v, b.v); return r; }
#define NTT_ADD(a, b) abl_add((a), (b))
#define NTT_SUB(a, b) abl_sub((a), (b))
#define NTT_SUB_LAZY(a, b) abl_sub((a), (b))
#else
#define NTT_ADD(a, b) Fr::add((a), (b))
#define NTT_SUB(a, b) Fr::sub((a), (b))
#define NTT_SUB_LAZY(a, b) Fr::sub_lazy((a), (b))
#endif

struct PassParams
{
    const fe* src;
    fe* dst;
    size_t batch_stride;    // elements between polynomials
    int log_n;
    int num_tiles;          // tiles per polynomial
    int total_work;         // num_tiles * batch
    const uint32_t* sub_tw; // padded limb-plane image of w_N^e, e < N/2, N = 2^L
    const fe* mat;          // pass A: inter-pass matrix (n entries, row-major [i1][j2])
    const fe* vec;          // pass A: pre-scale P[j1] or null;  pass B: post-scale Q[i2] or null
    fe post_const;          // pass B: extra constant (fft/ifft_with_constant)
    int has_post_const;
    int scatter_shift;      // pass B as the last of three passes: block b's output o goes to dst[(o << scatter_shift) + b]
    int tile_begin = 0;     // this launch covers tiles tile_begin .. tile_begin + num_tiles - 1 (host-buffer transforms go in blocks)
};

// DIF butterflies on the 3-bit owned field x[m] <-> k = base | m << B of a 2^L-point transform:
// stages s = B+R-1 .. B, (u, v) -> (u + v, (u - v) * w_N^((k mod 2^s) << (L-1-s))).
template <int L, int B, int R> BBG_D void radix_step(fe (&x)[8], int base_low, const uint32_t* tw)
{
#pragma unroll
    for (int sl = R - 1; sl >= 0; --sl)
    {
        const int half = 1 << sl;
        const int s = B + sl;
#pragma unroll
        for (int m = 0; m < 8; ++m)
        {
            if (m & half) continue;
            const int jm = m & (half - 1);
            const fe u = x[m];
            const fe v = x[m | half];
            x[m] = NTT_ADD(u, v);
            if (B == 0 && jm == 0)
            {
                x[m | half] = NTT_SUB(u, v); // twiddle w^0
            }
            else
            {
                // the difference may stay in (0, 4p): mul_const takes any multiplicand below 4p
                const int e = (base_low | (jm << B)) << (L - 1 - s);
                x[m | half] = Fr::mul_const(NTT_SUB_LAZY(u, v), tw_load(tw, e), twq_load(tw, e));
            }
        }
    }
}

template <int L, bool COLS_LOW> struct TileMap
{
    static constexpr int CLOG = TILE_LOG - L; // log2(columns per tile)
    static constexpr int KSHIFT = COLS_LOW ? CLOG : 0;
    // tile slot q of element m for thread t when the owned field sits at k-bit B
    template <int B> static BBG_D int slot(int t, int m)
    {
        constexpr int QB = B + KSHIFT;
        return (t & ((1 << QB) - 1)) | (m << QB) | ((t >> QB) << (QB + 3));
    }
    static BBG_D int k_of(int q) { return COLS_LOW ? (q >> CLOG) : (q & ((1 << L) - 1)); }
    static BBG_D int c_of(int q) { return COLS_LOW ? (q & ((1 << CLOG) - 1)) : (q >> L); }
};

// queue the copies of tile `tile` of polynomial `src` into `data` (256 threads x 16 chunks of 16 bytes)
template <int L, bool COLS_LOW> BBG_D void stage_tile(const PassParams& p, const fe* src, int tile, uint32_t* data)
{
    typedef TileMap<L, COLS_LOW> TM;
    const int t = threadIdx.x & (NT - 1);
    const int rest = p.log_n - L;
#pragma unroll
    for (int j = 0; j < 2 * TILE / NT; ++j)
    {
        const int i = t + j * NT; // chunk: half (i & 1) of the element of tile slot i >> 1
        const int q = i >> 1;
        const int k = TM::k_of(q), c = TM::c_of(q);
        size_t g;
        if (COLS_LOW) g = ((size_t)k << rest) + ((size_t)tile << TM::CLOG) + c;
        else g = ((((size_t)tile << TM::CLOG) + c) << L) + k;
        cp_async16(data + stage_pos(i), reinterpret_cast<const char*>(src + g) + 16 * (i & 1));
#if BBG_NTT_PREFETCH
        // the matrix entries of a tile are the same index set as its inputs: start them towards L2 a whole tile ahead
        if (COLS_LOW && p.mat != nullptr && (i & 1) == 0) prefetch_l2(p.mat + g);
#endif
    }
    cp_async_commit();
}

// (nsrc, ntile): the tile this half of the CTA works on next (nsrc == nullptr: none), staged by the last step
template <int L, bool COLS_LOW, int B, int R, bool FIRST, bool LAST>
BBG_D void do_step(fe (&x)[8], const PassParams& p, const fe* src, fe* dst, int tile, uint32_t* data, const uint32_t* tw, const fe* nsrc, int ntile)
{
    typedef TileMap<L, COLS_LOW> TM;
    const int t = threadIdx.x & (NT - 1);
    const int rest = p.log_n - L; // log2 of the other dimension
    if (FIRST)
    {
#pragma unroll
        for (int m = 0; m < 8; ++m)
        {
            const int q = TM::template slot<B>(t, m);
            const int k = TM::k_of(q), c = TM::c_of(q);
#if BBG_NTT_STAGE && BBG_NTT_ABLATE != 5
            (void)c;
            x[m] = staged_load(data, q);
#else
            size_t g;
            if (COLS_LOW) g = ((size_t)k << rest) + ((size_t)tile << TM::CLOG) + c;
            else g = ((((size_t)tile << TM::CLOG) + c) << L) + k;
#if BBG_NTT_ABLATE == 5
            x[m] = Fr::zero();
            x[m].v[0] = (uint32_t)g;
#else
            x[m] = ntt_load<COLS_LOW ? BBG_NTT_LOAD_A : BBG_NTT_LOAD_B>(src + g); // (a pass never writes the buffer it reads)
#endif
#if BBG_NTT_PREFETCH
            if (COLS_LOW && p.mat != nullptr) prefetch_l2(p.mat + g);
#endif
#endif
        }
        // the coset pre-scale in a loop of its own: the element loads above are volatile asm statements, which the compiler keeps
        // in program order against the (volatile) products - interleaved, every load would wait for the product before it
        if (COLS_LOW && p.vec != nullptr)
        {
#pragma unroll
            for (int m = 0; m < 8; ++m) x[m] = NTT_MUL(x[m], load_fe(p.vec + TM::k_of(TM::template slot<B>(t, m))));
        }
    }
    else
    {
#if BBG_NTT_ABLATE != 2
#pragma unroll
        for (int m = 0; m < 8; ++m) x[m] = sm_load(data, TM::template slot<B>(t, m));
#endif
#if BBG_NTT_STAGE && BBG_NTT_ABLATE != 5
        if (LAST)
        {
            NTT_SYNC(); // every thread of the tile has its last inputs in registers: the region is free
            if (nsrc != nullptr) stage_tile<L, COLS_LOW>(p, nsrc, ntile, data);
        }
#endif
    }
    {
        const int q0 = TM::template slot<B>(t, 0);
        const int base_low = TM::k_of(q0) & ((1 << B) - 1);
        radix_step<L, B, R>(x, base_low, tw);
    }
    if (LAST)
    {
#pragma unroll
        for (int m = 0; m < 8; ++m)
        {
            const int q = TM::template slot<B>(t, m);
            const int k = TM::k_of(q), c = TM::c_of(q);
            const unsigned isub = __brev((unsigned)k) >> (32 - L);
            if (COLS_LOW)
            {
                const size_t o = ((size_t)isub << rest) + ((size_t)tile << TM::CLOG) + c;
#if BBG_NTT_ABLATE == 5
                const fe y5 = NTT_MUL(x[m], x[(m + 1) & 7]);
                if (y5.v[0] == 0x12345u && y5.v[7] == 0x777u) store_fe_global(dst + o, y5);
#else
                ntt_store<BBG_NTT_STORE_A>(dst + o, NTT_MUL(x[m], ntt_load<BBG_NTT_LOAD_MAT>(p.mat + o)));
#endif
            }
            else
            {
                const size_t row = ((size_t)tile << TM::CLOG) + c;
                size_t o = row + ((size_t)isub << rest);
                if (p.scatter_shift) o <<= p.scatter_shift; // (dst already points at this block's column)
                fe y = x[m];
                if (p.vec != nullptr) y = NTT_MUL(y, load_fe_const(p.vec + isub));
                if (p.has_post_const) y = NTT_MUL(y, p.post_const);
#if BBG_NTT_ABLATE == 5
                if (y.v[0] == 0x12345u && y.v[7] == 0x777u)
#endif
                ntt_store<BBG_NTT_STORE_B>(dst + o, Fr::reduce(y));
            }
        }
    }
    else
    {
        // (a thread stores to the very slots it loaded in this step: nobody else reads them before the exchange below)
#if !BBG_NTT_FINE_SYNC
        if (!FIRST) NTT_SYNC();
#endif
#if BBG_NTT_STAGE && BBG_NTT_ABLATE != 5
        if (FIRST) NTT_SYNC(); // the staged image has been read by everyone before the limb planes overwrite it
#endif
#if BBG_NTT_ABLATE != 2
#pragma unroll
        for (int m = 0; m < 8; ++m) sm_store(data, TM::template slot<B>(t, m), x[m]);
#endif
        exchange_sync<B + TM::KSHIFT, (B >= 3 ? B - 3 : 0) + TM::KSHIFT>();
    }
}

template <int L, bool COLS_LOW, int B, bool FIRST>
BBG_D void run_from(fe (&x)[8], const PassParams& p, const fe* src, fe* dst, int tile, uint32_t* data, const uint32_t* tw, const fe* nsrc, int ntile)
{
    if constexpr (B == 0)
    {
        do_step<L, COLS_LOW, 0, 3, FIRST, true>(x, p, src, dst, tile, data, tw, nsrc, ntile);
    }
    else
    {
        do_step<L, COLS_LOW, B, 3, FIRST, false>(x, p, src, dst, tile, data, tw, nsrc, ntile);
        if constexpr (B >= 3) run_from<L, COLS_LOW, B - 3, false>(x, p, src, dst, tile, data, tw, nsrc, ntile);
        else do_step<L, COLS_LOW, 0, B, false, true>(x, p, src, dst, tile, data, tw, nsrc, ntile);
    }
}

template <int L, bool COLS_LOW> __global__ void __launch_bounds__(NT * TPC, 1) ntt_pass_kernel(PassParams p)
{
    BBG_DYN_SMEM(smem_raw);
    const int half = (int)threadIdx.x / NT; // which of the CTA's tiles this thread works on
    uint32_t* data = (uint32_t*)smem_raw + half * 8 * PLANE;
    uint32_t* tw = (uint32_t*)smem_raw + TPC * 8 * PLANE;
    const int batch = p.total_work / p.num_tiles;
#if BBG_NTT_STAGE && BBG_NTT_ABLATE != 5
    {
        // the first tile's copies fly while the twiddle image is loaded
        const int work0 = blockIdx.x * TPC + half;
        if (work0 < p.total_work) stage_tile<L, COLS_LOW>(p, p.src + (size_t)(work0 % batch) * p.batch_stride, work0 / batch + p.tile_begin, data);
    }
#endif
    // twiddle image: 66 KiB per CTA by 16-byte cp.async (no register round trips: the plain copy loop was 33 dependent
    // load-store iterations, ~14 us of prologue per kernel - 8% of a single 2^20 transform), with the first tile's inputs
    // started towards L2 meanwhile
    for (int i = threadIdx.x; i < 4 * TWP; i += NT * TPC) cp_async16(tw + 4 * i, p.sub_tw + 4 * i);
    cp_async_commit();
    {
        const int work0 = blockIdx.x * TPC + half;
        if (work0 < p.total_work)
        {
            typedef TileMap<L, COLS_LOW> TM;
            const fe* src0 = p.src + (size_t)(work0 % batch) * p.batch_stride;
            const int tile0 = work0 / batch + p.tile_begin, t = threadIdx.x & (NT - 1), rest = p.log_n - L;
#pragma unroll
            for (int m = 0; m < 8; ++m)
            {
                const int q = TM::template slot<L - 3>(t, m);
                const int k = TM::k_of(q), c = TM::c_of(q);
                prefetch_l2(COLS_LOW ? src0 + ((size_t)k << rest) + ((size_t)tile0 << TM::CLOG) + c : src0 + ((((size_t)tile0 << TM::CLOG) + c) << L) + k);
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();
    for (int work = blockIdx.x * TPC + half; work < p.total_work; work += gridDim.x * TPC)
    {
        // polynomial-minor order: the CTAs in flight work on the same few tiles of all polynomials of the batch, so the
        // inter-pass matrix tile (pass A; 128 MiB per 2^22 transform, more than L2 holds) is read from HBM once per batch
        const int tile = work / batch + p.tile_begin;
        const size_t b = (size_t)(work % batch);
        const int nwork = work + (int)gridDim.x * TPC;
        const fe* nsrc = nwork < p.total_work ? p.src + (size_t)(nwork % batch) * p.batch_stride : nullptr;
#if BBG_NTT_STAGE && BBG_NTT_ABLATE != 5
        cp_async_wait_all(); // this thread's copies of the tile have landed ...
        NTT_SYNC();          // ... and everybody else's
#endif
        fe x[8];
        run_from<L, COLS_LOW, L - 3, true>(x, p, p.src + b * p.batch_stride, p.scatter_shift ? p.dst + b : p.dst + b * p.batch_stride, tile, data, tw, nsrc,
                                           nwork / batch + p.tile_begin);
#if !(BBG_NTT_STAGE && BBG_NTT_ABLATE != 5)
        NTT_SYNC(); // the last step's shared-memory reads finish before the next tile overwrites
#endif
    }
}

// ------------------------------------------------------------------------------------------------
// n <= 2048: whole transform in one CTA (tests, tiny circuits, the verifier's domains)
// ------------------------------------------------------------------------------------------------
struct SmallParams
{
    fe* coeffs;
    size_t batch_stride;
    int log_n;
    const uint32_t* sub_tw;
    const fe* pre;  // P[j] or null
    const fe* post; // Q[i] or null
    fe post_const;
    int has_post_const;
};

__global__ void __launch_bounds__(NT) ntt_small_kernel(SmallParams p)
{
    BBG_DYN_SMEM(smem_raw);
    uint32_t* data = (uint32_t*)smem_raw;
    uint32_t* tw = data + 8 * PLANE;
    const int L = p.log_n, n = 1 << L, t = threadIdx.x;
    fe* poly = p.coeffs + (size_t)blockIdx.x * p.batch_stride;
    {
        // only the part of the image this size reads (entries e < n / 2 of each of the 16 planes), by 16-byte cp.async
        const int per_plane = (pad((n >> 1) - 1) + 4) >> 2; // 16-byte chunks per plane covering words 0 .. pad(n / 2 - 1); <= TWP / 4
        for (int i = t; i < 16 * per_plane; i += NT)
        {
            const int off = (i / per_plane) * TWP + 4 * (i % per_plane);
            cp_async16(tw + off, p.sub_tw + off);
        }
        cp_async_commit();
    }
    for (int i = t; i < n; i += NT)
    {
        fe x = load_fe_global(poly + i);
        if (p.pre != nullptr) x = Fr::mul(x, load_fe_const(p.pre + i));
        sm_store(data, i, x);
    }
    cp_async_wait_all();
    for (int s = L - 1; s >= 0; --s)
    {
        __syncthreads();
        const int h = 1 << s;
        for (int i = t; i < (n >> 1); i += NT)
        {
            const int j = i & (h - 1);
            const int q0 = ((i >> s) << (s + 1)) | j;
            const fe u = sm_load(data, q0), v = sm_load(data, q0 | h);
            sm_store(data, q0, Fr::add(u, v));
            const int e = j << (L - 1 - s);
            const fe d = j != 0 ? Fr::mul_const(Fr::sub_lazy(u, v), tw_load(tw, e), twq_load(tw, e)) : Fr::sub(u, v);
            sm_store(data, q0 | h, d);
        }
    }
    __syncthreads();
    for (int i = t; i < n; i += NT)
    {
        const unsigned o = __brev((unsigned)i) >> (32 - L);
        fe y = sm_load(data, i);
        if (p.post != nullptr) y = Fr::mul(y, load_fe_const(p.post + o));
        if (p.has_post_const) y = Fr::mul(y, p.post_const);
        store_fe_global(poly + o, Fr::reduce(y));
    }
}

// ------------------------------------------------------------------------------------------------
// Table generation (device)
// ------------------------------------------------------------------------------------------------
// out[i] = scale * base^i
__global__ void gen_powers_kernel(fe* out, fe base, fe scale, unsigned count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    store_fe(out + i, Fr::mul(Fr::pow_u64(base, i), scale));
}
// padded limb-plane image of root^e, e < half: planes 0..7 the plain (non-Montgomery) value, planes 8..15 its quotient
// floor(w 2^256 / p) - the two constants Fr::mul_const multiplies by
__global__ void gen_subtw_image_kernel(uint32_t* img, fe root, unsigned half)
{
    const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= half) return;
    const fe w = Fr::reduce(Fr::pow_u64(root, e)); // canonical Montgomery form
    const fe plain = Fr::from_mont(w), quot = Fr::const_quotient(w);
    const int p = pad((int)e);
#pragma unroll
    for (int l = 0; l < 8; ++l)
    {
        img[l * TWP + p] = plain.v[l];
        img[(8 + l) * TWP + p] = quot.v[l];
    }
}
// M[i1][j2] = lo[e & 2047] * hi[e >> 11] * rf[i1] * cf[j2],  e = i1 * j2 < n
__global__ void gen_matrix_kernel(fe* M, const fe* lo, const fe* hi, const fe* rf, const fe* cf, int log_n, int log_cols)
{
    const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >> log_n) return;
    const size_t i1 = idx >> log_cols, j2 = idx & (((size_t)1 << log_cols) - 1);
    const size_t e = i1 * j2;
    fe w = Fr::mul(load_fe(lo + (e & ((1u << LO_TABLE_LOG) - 1))), load_fe(hi + (e >> LO_TABLE_LOG)));
    if (rf != nullptr) w = Fr::mul(w, load_fe(rf + i1));
    if (cf != nullptr) w = Fr::mul(w, load_fe(cf + j2));
    store_fe(M + idx, w);
}
// compute_lagrange_polynomial_fft (polynomial_arithmetic.cpp:381-476): the target-domain coset evaluations of
// L_1(X) = (X^n - 1) / (n (X - 1)):  out[i] = numer[i mod S] / (g * w_T^i - 1),  numer[j] = (g^n * w_S^j - 1) / n,
// S = T / n.  The reference inverts all T denominators with one serial Montgomery-trick sweep; here every thread
// owns a run of LAGRANGE_RUN consecutive i (denominators by repeated multiplication with w_T, prefix products, ONE
// Fermat inversion, back-substitution): ~15 products per element instead of a serial chain.
constexpr int LAGRANGE_RUN = 32;
__global__ void __launch_bounds__(64) lagrange_fft_kernel(fe* out, fe root_T, fe generator, const fe* numer, unsigned S_mask, unsigned T)
{
    const unsigned run = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned first = run * LAGRANGE_RUN;
    if (first >= T) return;
    const int count = (int)((T - first < (unsigned)LAGRANGE_RUN) ? T - first : LAGRANGE_RUN);
    const fe one = Fr::one();
    fe x = Fr::mul(generator, Fr::pow_u64(root_T, first)); // g * w_T^first
    fe den[LAGRANGE_RUN], prefix[LAGRANGE_RUN];
    fe acc = one;
    for (int i = 0; i < count; ++i)
    {
        den[i] = Fr::sub(x, one);
        prefix[i] = acc;
        acc = Fr::mul(acc, den[i]);
        x = Fr::mul(x, root_T);
    }
    fe inv = Fr::invert(acc);
    for (int i = count - 1; i >= 0; --i)
    {
        const fe d_inv = Fr::mul(inv, prefix[i]);
        inv = Fr::mul(inv, den[i]);
        store_fe(out + first + i, Fr::mul_full(d_inv, load_fe(numer + ((first + i) & S_mask))));
    }
}

// x[i] = canonical(x[i] * k * lo[i & mask] * hi[i >> lo_log]): the coset scalings of the three-pass sizes
__global__ void scale_powers_kernel(fe* x, const fe* lo, const fe* hi, int lo_log, fe k, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    {
        fe w = Fr::mul(load_fe(lo + (i & (((size_t)1 << lo_log) - 1))), load_fe(hi + (i >> lo_log)));
        store_fe(x + i, Fr::mul_full(Fr::mul(load_fe(x + i), w), k));
    }
}

// evaluation_domain::compute_lookup_table (evaluation_domain.cpp:33-54, :172-178): per direction, round i (m = 2^(i+1),
// i = 0 .. log2 size - 2) holds w_(2m)^j = root^(j size / 2m) for j < m at offset 2^(i+1) - 2; forward rounds fill
// roots[0, size), inverse rounds roots[size, 2 size).  The reference builds every round with a serial chain of coarse
// products; here each entry is one product of two table look-ups (canonical values: same field elements).
__global__ void domain_lookup_kernel(fe* roots, const fe* lo, const fe* hi, const fe* ilo, const fe* ihi, int lo_log, int log_size)
{
    const size_t size = (size_t)1 << log_size;
    for (size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x; t < 2 * size; t += (size_t)gridDim.x * blockDim.x)
    {
        const bool inverse = t >= size;
        const size_t u = inverse ? t - size : t;
        fe v = Fr::zero(); // the last two slots of each half are unused (the reference leaves them uninitialised)
        if (u + 2 < size)
        {
            const int i = 62 - __clzll((long long)(u + 2)); // round: 2^(i+1) <= u + 2 < 2^(i+2)
            const size_t j = u + 2 - ((size_t)1 << (i + 1));
            const size_t e = j << (log_size - i - 2); // < size / 2
            const fe* l = inverse ? ilo : lo;
            const fe* h = inverse ? ihi : hi;
            v = load_fe(l + (e & (((size_t)1 << lo_log) - 1)));
            if (h != nullptr) v = Fr::mul(v, load_fe(h + (e >> lo_log)));
            v = Fr::reduce(v);
        }
        store_fe(roots + t, v);
    }
}

// out[i] = in[i] * k
__global__ void scale_vector_kernel(fe* out, const fe* in, fe k, unsigned count)
{
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    store_fe(out + i, Fr::mul(load_fe(in + i), k));
}
} // namespace nttk

// ================================================================================================
// Host driver
// ================================================================================================
namespace
{
using namespace nttk;

size_t g_ntt_launches = 0;

// fr.hpp:59-63: 2^28-th primitive root of unity; :66-74 coset generator 5 and its inverse (Montgomery)
const fe ROOT_2_28 = { { 0x80D13D9Cu, 0x636E7355u, 0x2445FFD6u, 0xA22BF374u, 0x1EB203D8u, 0x56452AC0u, 0x2963F9E7u, 0x1860EF94u } };
const fe COSET_GEN = { { 0x9FFFFFE6u, 0x1B0D0EF9u, 0xA32A913Fu, 0xEABA68A3u, 0xD8DD0689u, 0x47D8EB76u, 0x20F5BBC3u, 0x15D00855u } };
const fe COSET_GEN_INV = { { 0x09999999u, 0xD7453974u, 0x83C3EFA8u, 0xB4ADA7D4u, 0xE57F3161u, 0xC49CA2F8u, 0xAC156CB3u, 0x162A3754u } };

// host-side evaluation of the handful of per-domain constants (the BBG_HD field code runs on the host too)
fe host_root_of_unity(unsigned log_n) // field.hpp:487-494
{
    fe r = ROOT_2_28;
    for (unsigned i = 28; i > log_n; --i) r = Fr::reduce(Fr::sqr(r));
    return r;
}
fe host_pow(fe base, uint64_t e) { return Fr::reduce(Fr::pow_u64(base, e)); }
fe host_domain_inverse(unsigned log_n) // evaluation_domain.cpp:65-66
{
    fe n = Fr::zero();
    n.v[0] = (uint32_t)(1u << log_n);
    return Fr::invert(Fr::to_mont(n));
}

struct DeviceBuf
{
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) bbg_rt::dev_free(p);
        p = nullptr;
        bytes = 0;
        int e = bbg_rt::dev_alloc(&p, need);
        if (e == 0) bytes = need;
        return e;
    }
    void release()
    {
        if (p) bbg_rt::dev_free(p);
        p = nullptr;
        bytes = 0;
    }
};

struct NttTables
{
    std::map<int, uint32_t*> sub_tw;    // key = L * 2 + inverse
    std::map<int, fe*> matrix;          // key = log_n * 8 + variant
    std::map<int, fe*> vec;             // key = log_n * 8 + kind
    DeviceBuf scratch;                  // pass A output
    DeviceBuf scratch_side;             // the same for transforms queued on a second stream (ntt_use_side_scratch)
    bool use_side = false;
    DeviceBuf tmp_vec;                  // per-call scaled pre-scale vector
    std::vector<void*> owned;
    bool smem_configured = false;
} g_tables;

DeviceBuf& active_scratch() { return g_tables.use_side ? g_tables.scratch_side : g_tables.scratch; }

int alloc_owned(void** p, size_t bytes)
{
    BBG_CHECK(bbg_rt::dev_alloc(p, bytes));
    g_tables.owned.push_back(*p);
    return 0;
}

int get_sub_tw(int L, bool inverse, cudaStream_t st, const uint32_t** out)
{
    const int key = L * 2 + (inverse ? 1 : 0);
    auto it = g_tables.sub_tw.find(key);
    if (it == g_tables.sub_tw.end())
    {
        uint32_t* img = nullptr;
        BBG_CHECK(alloc_owned((void**)&img, (size_t)16 * TWP * 4));
        BBG_CHECK(bbg_rt::dev_memset(img, 0, (size_t)16 * TWP * 4, st));
        fe root = host_root_of_unity((unsigned)L);
        if (inverse) root = Fr::invert(root);
        const unsigned half = L >= 1 ? (1u << (L - 1)) : 1u;
        BBG_LAUNCH_NOSYNC(gen_subtw_image_kernel, dim3((half + 127) / 128), dim3(128), st, img, root, half);
        ++g_ntt_launches;
        BBG_CHECK(bbg_rt::sync(st)); // (one-off; callers on other streams may use the table next)
        it = g_tables.sub_tw.emplace(key, img).first;
    }
    *out = it->second;
    return 0;
}

enum vec_kind
{
    VEC_PRE_COSET = 0,   // big: g^(N2 j1), j1 < N1        small: g^j, j < n
    VEC_POST_ICOSET = 1, // big: g^(-N1 i2), i2 < N2       small: g^(-i) / n, i < n
    VEC_G_LO = 2,        // three-pass sizes: g^j, j < 2^11           VEC_G_HI = 3: g^(2^11 j), j < n / 2^11
    VEC_G_HI = 3,
    VEC_GINV_LO = 4,     // the same for g^-1
    VEC_GINV_HI = 5,
};

// split used for n >= 2^12: N1 = 2^L1 rows (pass A transform length), N2 = 2^L2 columns (pass B length)
void split(unsigned log_n, int& L1, int& L2)
{
    if (log_n > 2 * (unsigned)TILE_LOG)
    {
        // three passes: outer length 2^L1, then the two-pass transform of size 2^L2 <= 2^22 on every block
        L1 = (int)(log_n + 2) / 3;
        L2 = (int)log_n - L1;
        return;
    }
    L1 = (int)(log_n + 1) / 2;
    L2 = (int)log_n - L1;
    // 2^20 = 2^9 x 2^11 rather than 2^10 x 2^10: a 10-stage sub-transform is 3 + 3 + 3 + 1 stages, its fourth radix step one
    // stage of trivial twiddles behind a whole exchange; 9 + 11 stages need five exchanges instead of six and pass A reads
    // 128-byte column pieces instead of 64-byte ones (measured, batch 8: 1.387 -> 1.358 ms; 2^11 x 2^9: 1.396; at 2^21 the
    // default 2^11 x 2^10 stays ahead of 2^10 x 2^11, 2.98 against 3.05 ms)
    if (log_n == 20) { L1 = 9; L2 = 11; }
    // development override: BBG_NTT_SPLIT="20:9,21:10" sets L1 per log_n (read once; both lengths must stay in 6..11)
    static int override_l1[32];
    static bool parsed = false;
    if (!parsed)
    {
        parsed = true;
        if (const char* e = getenv("BBG_NTT_SPLIT"))
        {
            while (*e)
            {
                const int lg = atoi(e);
                const char* colon = strchr(e, ':');
                if (colon == nullptr) break;
                const int l1 = atoi(colon + 1);
                if (lg >= 12 && lg <= 22 && l1 >= 6 && l1 <= TILE_LOG && lg - l1 >= 6 && lg - l1 <= TILE_LOG) override_l1[lg] = l1;
                const char* comma = strchr(colon, ',');
                if (comma == nullptr) break;
                e = comma + 1;
            }
        }
    }
    if (log_n < 32 && override_l1[log_n] != 0)
    {
        L1 = override_l1[log_n];
        L2 = (int)log_n - L1;
    }
}

int get_vec(unsigned log_n, int kind, cudaStream_t st, const fe** out)
{
    const int key = (int)log_n * 8 + kind;
    auto it = g_tables.vec.find(key);
    if (it == g_tables.vec.end())
    {
        fe* v = nullptr;
        unsigned count;
        fe base, scale = Fr::one();
        if (kind >= VEC_G_LO)
        {
            const fe g = (kind == VEC_G_LO || kind == VEC_G_HI) ? COSET_GEN : COSET_GEN_INV;
            const bool hi = (kind == VEC_G_HI || kind == VEC_GINV_HI);
            count = hi ? 1u << (log_n - LO_TABLE_LOG) : 1u << LO_TABLE_LOG;
            base = hi ? host_pow(g, (uint64_t)1 << LO_TABLE_LOG) : g;
        }
        else if (log_n <= (unsigned)TILE_LOG)
        {
            count = 1u << log_n;
            if (kind == VEC_PRE_COSET) base = COSET_GEN;
            else
            {
                base = COSET_GEN_INV;
                scale = host_domain_inverse(log_n);
            }
        }
        else
        {
            int L1, L2;
            split(log_n, L1, L2);
            if (kind == VEC_PRE_COSET)
            {
                count = 1u << L1;
                base = host_pow(COSET_GEN, (uint64_t)1 << L2);
            }
            else
            {
                count = 1u << L2;
                base = host_pow(COSET_GEN_INV, (uint64_t)1 << L1);
            }
        }
        BBG_CHECK(alloc_owned((void**)&v, (size_t)count * 32));
        BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((count + 127) / 128), dim3(128), st, v, base, scale, count);
        ++g_ntt_launches;
        BBG_CHECK(bbg_rt::sync(st));
        it = g_tables.vec.emplace(key, v).first;
    }
    *out = it->second;
    return 0;
}

// variant: 0 forward, 1 inverse (x 1/n), 2 coset forward (x g^j2), 3 coset inverse (x g^-i1 / n),
//          4 inverse roots without the 1/n (the inner transform of the three-pass sizes)
int get_matrix(unsigned log_n, int variant, cudaStream_t st, const fe** out)
{
    const int key = (int)log_n * 8 + variant;
    auto it = g_tables.matrix.find(key);
    if (it == g_tables.matrix.end())
    {
        int L1, L2;
        split(log_n, L1, L2);
        const size_t n = (size_t)1 << log_n;
        const bool inverse = (variant & 1) != 0 || variant == 4;
        fe w = host_root_of_unity(log_n);
        if (inverse) w = Fr::invert(w);
        const fe scale = (variant & 1) != 0 ? host_domain_inverse(log_n) : Fr::one();
        const unsigned lo_count = 1u << LO_TABLE_LOG, hi_count = (unsigned)(n >> LO_TABLE_LOG);
        fe *lo = nullptr, *hi = nullptr, *rf = nullptr, *cf = nullptr, *M = nullptr;
        BBG_CHECK(bbg_rt::dev_alloc((void**)&lo, (size_t)lo_count * 32));
        BBG_CHECK(bbg_rt::dev_alloc((void**)&hi, (size_t)hi_count * 32));
        BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((lo_count + 127) / 128), dim3(128), st, lo, w, scale, lo_count);
        BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((hi_count + 127) / 128), dim3(128), st, hi, host_pow(w, lo_count), Fr::one(), hi_count);
        g_ntt_launches += 2;
        if (variant == 2)
        {
            BBG_CHECK(bbg_rt::dev_alloc((void**)&cf, ((size_t)32) << L2));
            BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3(((1u << L2) + 127) / 128), dim3(128), st, cf, COSET_GEN, Fr::one(), 1u << L2);
            ++g_ntt_launches;
        }
        if (variant == 3)
        {
            BBG_CHECK(bbg_rt::dev_alloc((void**)&rf, ((size_t)32) << L1));
            BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3(((1u << L1) + 127) / 128), dim3(128), st, rf, COSET_GEN_INV, Fr::one(), 1u << L1);
            ++g_ntt_launches;
        }
        BBG_CHECK(alloc_owned((void**)&M, n * 32));
        BBG_LAUNCH_NOSYNC(gen_matrix_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), st, M, lo, hi, rf, cf, (int)log_n, L2);
        ++g_ntt_launches;
        BBG_CHECK(bbg_rt::sync(st));
        bbg_rt::dev_free(lo);
        bbg_rt::dev_free(hi);
        if (rf) bbg_rt::dev_free(rf);
        if (cf) bbg_rt::dev_free(cf);
        it = g_tables.matrix.emplace(key, M).first;
    }
    *out = it->second;
    return 0;
}

template <int L, bool COLS_LOW> int launch_pass_L(const PassParams& p, cudaStream_t st)
{
    static bool configured = false;
    if (!configured)
    {
        BBG_CHECK(bbg_rt::set_smem_limit((const void*)ntt_pass_kernel<L, COLS_LOW>, SMEM_BYTES));
        configured = true;
    }
    int grid = (2 / TPC) * bbg_rt::num_sms();
    if (grid > (p.total_work + TPC - 1) / TPC) grid = (p.total_work + TPC - 1) / TPC;
    auto kernel = ntt_pass_kernel<L, COLS_LOW>; // (alias: the template's comma would split the macro argument)
    bbg_prof::Scope prof(COLS_LOW ? bbg_prof::NTT_PASS_A : bbg_prof::NTT_PASS_B, st);
    BBG_LAUNCH(kernel, dim3((unsigned)grid), dim3(NT * TPC), SMEM_BYTES, st, p);
    ++g_ntt_launches;
    return bbg_rt::last_error();
}
template <bool COLS_LOW> int launch_pass(int L, const PassParams& p, cudaStream_t st)
{
    switch (L)
    {
    case 6: return launch_pass_L<6, COLS_LOW>(p, st);
    case 7: return launch_pass_L<7, COLS_LOW>(p, st);
    case 8: return launch_pass_L<8, COLS_LOW>(p, st);
    case 9: return launch_pass_L<9, COLS_LOW>(p, st);
    case 10: return launch_pass_L<10, COLS_LOW>(p, st);
    case 11: return launch_pass_L<11, COLS_LOW>(p, st);
    }
    return 1001;
}
} // namespace

size_t ntt_launch_count() { return g_ntt_launches; }
// Transforms queued while this is on use a second scratch buffer, so they may run on another stream concurrently with
// transforms queued while it is off (one host thread drives both streams).  Only for operations without a per-call scaled
// vector (plain / coset forward and inverse): those share g_tables.tmp_vec.
void ntt_use_side_scratch(bool on) { g_tables.use_side = on; }

// d_out: T = 2^log_target field elements on the device
int lagrange_fft_device(void* d_out, unsigned log_src, unsigned log_target, cudaStream_t st)
{
    if (log_target < log_src || log_target > 28 || log_target - log_src > 10) return 1002;
    const unsigned T = 1u << log_target, S = 1u << (log_target - log_src);
    // numer[j] = (g^n * w_S^j - 1) / n on the host: S is 2 or 4 in the prover (compute_multiplicative_subgroup, :104-127)
    fe gn = COSET_GEN;
    for (unsigned i = 0; i < log_src; ++i) gn = Fr::reduce(Fr::sqr(gn));
    const fe w_S = host_root_of_unity(log_target - log_src);
    const fe n_inv = host_domain_inverse(log_src);
    std::vector<fe> numer(S);
    fe cur = gn;
    for (unsigned j = 0; j < S; ++j)
    {
        numer[j] = Fr::mul_full(Fr::sub(cur, Fr::one()), n_inv);
        cur = Fr::mul_full(cur, w_S);
    }
    BBG_CHECK(g_tables.tmp_vec.ensure((size_t)S * 32));
    BBG_CHECK(bbg_rt::h2d(g_tables.tmp_vec.p, numer.data(), (size_t)S * 32, st));
    BBG_CHECK(bbg_rt::sync(st)); // numer is a stack-lifetime host buffer
    const unsigned runs = (T + LAGRANGE_RUN - 1) / LAGRANGE_RUN;
    BBG_LAUNCH_NOSYNC(lagrange_fft_kernel, dim3((runs + 63) / 64), dim3(64), st, (fe*)d_out, host_root_of_unity(log_target), COSET_GEN,
                      (const fe*)g_tables.tmp_vec.p, S - 1, T);
    ++g_ntt_launches;
    return bbg_rt::last_error();
}

int domain_lookup_table_device(void* d_roots, unsigned log_size, cudaStream_t st)
{
    if (log_size < 1 || log_size > 28) return 1002;
    const int lo_log = (int)(log_size < (unsigned)LO_TABLE_LOG ? log_size : (unsigned)LO_TABLE_LOG);
    const unsigned lo_count = 1u << lo_log, hi_count = log_size > (unsigned)lo_log ? 1u << (log_size - lo_log) : 0u;
    fe* tab = nullptr;
    BBG_CHECK(bbg_rt::dev_alloc((void**)&tab, (size_t)2 * (lo_count + hi_count + 1) * 32));
    fe *lo = tab, *hi = lo + lo_count, *ilo = hi + hi_count, *ihi = ilo + lo_count;
    const fe w = host_root_of_unity(log_size);
    const fe wi = Fr::invert(w);
    BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((lo_count + 127) / 128), dim3(128), st, lo, w, Fr::one(), lo_count);
    BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((lo_count + 127) / 128), dim3(128), st, ilo, wi, Fr::one(), lo_count);
    if (hi_count)
    {
        BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((hi_count + 127) / 128), dim3(128), st, hi, host_pow(w, lo_count), Fr::one(), hi_count);
        BBG_LAUNCH_NOSYNC(gen_powers_kernel, dim3((hi_count + 127) / 128), dim3(128), st, ihi, host_pow(wi, lo_count), Fr::one(), hi_count);
    }
    const int grid = 8 * bbg_rt::num_sms();
    BBG_LAUNCH_NOSYNC(domain_lookup_kernel, dim3((unsigned)grid), dim3(256), st, (fe*)d_roots, (const fe*)lo, hi_count ? (const fe*)hi : (const fe*)nullptr,
                      (const fe*)ilo, hi_count ? (const fe*)ihi : (const fe*)nullptr, lo_log, (int)log_size);
    g_ntt_launches += hi_count ? 5 : 3;
    int e = bbg_rt::last_error();
    if (e == 0) e = bbg_rt::sync(st);
    bbg_rt::dev_free(tab);
    return e;
}

#ifndef BBG_EMULATE
namespace
{
std::vector<cudaEvent_t> g_block_events; // ntt_host_blocks: one per block in flight, created at first use
}
#endif
int ntt_release_tables()
{
#ifndef BBG_EMULATE
    for (cudaEvent_t e : g_block_events) cudaEventDestroy(e);
    g_block_events.clear();
#endif
    for (void* p : g_tables.owned) bbg_rt::dev_free(p);
    g_tables.owned.clear();
    g_tables.sub_tw.clear();
    g_tables.matrix.clear();
    g_tables.vec.clear();
    g_tables.scratch.release();
    g_tables.scratch_side.release();
    g_tables.tmp_vec.release();
    return 0;
}

namespace
{
// n = 2^23 .. 2^28: outer pass over columns of length N0 = 2^L0, then the two-pass transform of size M = n / N0 on each of
// the N0 contiguous blocks, the last pass writing block i0's output u to i0 + N0 u.  coeffs -> scratch -> scratch2 -> coeffs.
int ntt_three_pass(void* d_coeffs, size_t stride, size_t batch, unsigned log_n, bool inverse, bool coset, bool with_constant, const fe& k,
                   cudaStream_t st)
{
    const size_t n = (size_t)1 << log_n;
    int L0, LM;
    split(log_n, L0, LM); // outer length 2^L0, inner size 2^LM (12 .. 22)
    int L1, L2;
    split((unsigned)LM, L1, L2);
    const size_t M = (size_t)1 << LM;
    BBG_CHECK(active_scratch().ensure(2 * n * 32));
    fe* s1 = (fe*)active_scratch().p;
    fe* s2 = s1 + n;
    const fe *g_lo = nullptr, *g_hi = nullptr;
    if (coset)
    {
        BBG_CHECK(get_vec(log_n, inverse ? VEC_GINV_LO : VEC_G_LO, st, &g_lo));
        BBG_CHECK(get_vec(log_n, inverse ? VEC_GINV_HI : VEC_G_HI, st, &g_hi));
    }
    const int scale_grid = 8 * bbg_rt::num_sms();
    for (size_t poly = 0; poly < batch; ++poly)
    {
        fe* x = (fe*)d_coeffs + poly * stride;
        if (coset && !inverse)
        {
            // scale_by_generator (polynomial_arithmetic.cpp:81-102): x[j] *= k g^j before the transform
            BBG_LAUNCH_NOSYNC(scale_powers_kernel, dim3((unsigned)scale_grid), dim3(256), st, x, g_lo, g_hi, LO_TABLE_LOG, with_constant ? k : Fr::one(), n);
            ++g_ntt_launches;
        }
        PassParams o;
        o.src = x;
        o.dst = s1;
        o.batch_stride = n;
        o.log_n = (int)log_n;
        o.num_tiles = (int)(n >> TILE_LOG);
        o.total_work = o.num_tiles;
        o.vec = nullptr;
        o.has_post_const = 0;
        o.post_const = Fr::one();
        o.scatter_shift = 0;
        BBG_CHECK(get_sub_tw(L0, inverse, st, &o.sub_tw));
        BBG_CHECK(get_matrix(log_n, inverse ? 1 : 0, st, &o.mat)); // w_n^(+-i0 j) (x 1/n for the inverse)
        BBG_CHECK(launch_pass<true>(L0, o, st));
        PassParams a = o;
        a.src = s1;
        a.dst = s2;
        a.batch_stride = M;
        a.log_n = LM;
        a.num_tiles = (int)(M >> TILE_LOG);
        a.total_work = a.num_tiles << L0;
        BBG_CHECK(get_sub_tw(L1, inverse, st, &a.sub_tw));
        BBG_CHECK(get_matrix((unsigned)LM, inverse ? 4 : 0, st, &a.mat));
        BBG_CHECK(launch_pass<true>(L1, a, st));
        PassParams b = a;
        b.src = s2;
        b.dst = x;
        b.mat = nullptr;
        b.scatter_shift = L0;
        BBG_CHECK(get_sub_tw(L2, inverse, st, &b.sub_tw));
        if (with_constant && !coset)
        {
            b.has_post_const = 1; // fft / ifft_with_constant (:279-285, :301-309)
            b.post_const = k;
        }
        BBG_CHECK(launch_pass<false>(L2, b, st));
        if (coset && inverse)
        {
            // coset_ifft (:311-315): x[i] *= g^-i after the inverse transform
            BBG_LAUNCH_NOSYNC(scale_powers_kernel, dim3((unsigned)scale_grid), dim3(256), st, x, g_lo, g_hi, LO_TABLE_LOG, Fr::one(), n);
            ++g_ntt_launches;
        }
    }
    return bbg_rt::last_error();
}
} // namespace

// The two passes of an n = 2^12 .. 2^22 transform: tables fetched (generated at first use), parameters filled in
int setup_two_pass(void* d_coeffs, size_t stride, size_t batch, unsigned log_n, bool inverse, bool coset, bool with_constant, const fe& k, cudaStream_t st,
                   PassParams& a, PassParams& b, int& L1, int& L2)
{
    const size_t n = (size_t)1 << log_n;
    split(log_n, L1, L2);
    BBG_CHECK(active_scratch().ensure(batch * n * 32));
    a.src = (const fe*)d_coeffs;
    a.dst = (fe*)active_scratch().p;
    a.batch_stride = stride;
    a.log_n = (int)log_n;
    a.num_tiles = (int)(n >> TILE_LOG);
    a.total_work = a.num_tiles * (int)batch;
    a.vec = nullptr;
    a.has_post_const = 0;
    a.post_const = Fr::one();
    a.scatter_shift = 0;
    b = a;
    // pass A writes polynomial i at scratch + i * n; pass B reads it back from there
    b.src = (const fe*)active_scratch().p;
    b.dst = (fe*)d_coeffs;
    BBG_CHECK(get_sub_tw(L1, inverse, st, &a.sub_tw));
    BBG_CHECK(get_sub_tw(L2, inverse, st, &b.sub_tw));
    const int variant = (inverse ? 1 : 0) | (coset ? 2 : 0);
    BBG_CHECK(get_matrix(log_n, variant, st, &a.mat));
    b.mat = nullptr;
    if (coset && !inverse)
    {
        const fe* pc = nullptr;
        BBG_CHECK(get_vec(log_n, VEC_PRE_COSET, st, &pc));
        if (with_constant)
        {
            const unsigned cnt = 1u << L1;
            BBG_CHECK(g_tables.tmp_vec.ensure((size_t)cnt * 32));
            BBG_LAUNCH_NOSYNC(scale_vector_kernel, dim3((cnt + 127) / 128), dim3(128), st, (fe*)g_tables.tmp_vec.p, pc, k, cnt);
            ++g_ntt_launches;
            pc = (const fe*)g_tables.tmp_vec.p;
        }
        a.vec = pc;
    }
    else if (coset && inverse)
    {
        BBG_CHECK(get_vec(log_n, VEC_POST_ICOSET, st, &b.vec));
    }
    else if (with_constant)
    {
        b.has_post_const = 1;
        b.post_const = k;
    }
    return 0;
}

int ntt_device(void* d_coeffs, size_t stride, size_t batch, unsigned log_n, int op, const uint64_t* constant, cudaStream_t st)
{
    if (log_n < 1 || log_n > 28) return 1002; // n = 2 .. 2^28 (the two-adicity of the field, fr.hpp:59-63)
    if (op < OP_FFT || op > OP_COSET_FFT_WITH_CONSTANT) return 1003;
    if (batch == 0) return 0;
    const bool inverse = (op == OP_IFFT || op == OP_COSET_IFFT || op == OP_IFFT_WITH_CONSTANT);
    const bool coset = (op == OP_COSET_FFT || op == OP_COSET_IFFT || op == OP_COSET_FFT_WITH_CONSTANT);
    const bool with_constant = (op >= OP_FFT_WITH_CONSTANT);
    if (with_constant && constant == nullptr) return 1004;
    fe k = Fr::one();
    if (with_constant) k = load_fe(constant);
    const size_t n = (size_t)1 << log_n;

    if (log_n <= (unsigned)TILE_LOG)
    {
        SmallParams sp;
        sp.coeffs = (fe*)d_coeffs;
        sp.batch_stride = stride;
        sp.log_n = (int)log_n;
        BBG_CHECK(get_sub_tw((int)log_n, inverse, st, &sp.sub_tw));
        sp.pre = nullptr;
        sp.post = nullptr;
        sp.has_post_const = 0;
        sp.post_const = Fr::one();
        if (coset && !inverse)
        {
            const fe* pc = nullptr;
            BBG_CHECK(get_vec(log_n, VEC_PRE_COSET, st, &pc));
            if (with_constant)
            {
                // coset_fft_with_constant: generator start value = constant (polynomial_arithmetic.cpp:293-299)
                BBG_CHECK(g_tables.tmp_vec.ensure(n * 32));
                BBG_LAUNCH_NOSYNC(scale_vector_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), st, (fe*)g_tables.tmp_vec.p, pc, k, (unsigned)n);
                ++g_ntt_launches;
                pc = (const fe*)g_tables.tmp_vec.p;
            }
            sp.pre = pc;
        }
        else if (coset && inverse)
        {
            BBG_CHECK(get_vec(log_n, VEC_POST_ICOSET, st, &sp.post)); // carries 1/n
        }
        else if (inverse)
        {
            sp.has_post_const = 1;
            sp.post_const = host_domain_inverse(log_n);
            if (with_constant) sp.post_const = Fr::reduce(Fr::mul(sp.post_const, k)); // :301-309
        }
        else if (with_constant)
        {
            sp.has_post_const = 1;
            sp.post_const = k;
        }
        if (!g_tables.smem_configured)
        {
            BBG_CHECK(bbg_rt::set_smem_limit((const void*)ntt_small_kernel, SMEM_BYTES_SMALL));
            g_tables.smem_configured = true;
        }
        bbg_prof::Scope prof(bbg_prof::NTT_SMALL, st);
        BBG_LAUNCH(ntt_small_kernel, dim3((unsigned)batch), dim3(NT), SMEM_BYTES_SMALL, st, sp);
        ++g_ntt_launches;
        return bbg_rt::last_error();
    }

    if (log_n > 2 * (unsigned)TILE_LOG) return ntt_three_pass(d_coeffs, stride, batch, log_n, inverse, coset, with_constant, k, st);

    int L1, L2;
    PassParams a, b;
    BBG_CHECK(setup_two_pass(d_coeffs, stride, batch, log_n, inverse, coset, with_constant, k, st, a, b, L1, L2));
    // pass A: d_coeffs (stride) -> scratch (dense);  pass B: scratch (dense) -> d_coeffs (stride)
    {
        PassParams pa = a;
        // separate strides for source and destination are expressed by running pass A per layout:
        // scratch is dense (stride n).  When the caller's stride differs, process polynomials one by one.
        if (stride == n || batch == 1)
        {
            pa.batch_stride = stride;
            // src stride == dst stride only when stride == n; batch == 1 makes the stride irrelevant
            BBG_CHECK(launch_pass<true>(L1, pa, st));
            PassParams pb = b;
            pb.batch_stride = stride;
            BBG_CHECK(launch_pass<false>(L2, pb, st));
        }
        else
        {
            for (size_t i = 0; i < batch; ++i)
            {
                PassParams p1 = a;
                p1.src = (const fe*)d_coeffs + i * stride;
                p1.total_work = a.num_tiles;
                BBG_CHECK(launch_pass<true>(L1, p1, st));
                PassParams p2 = b;
                p2.dst = (fe*)d_coeffs + i * stride;
                p2.total_work = a.num_tiles;
                BBG_CHECK(launch_pass<false>(L2, p2, st));
            }
        }
    }
    return 0;
}
#ifndef BBG_EMULATE
// ---- a host buffer's transform with the copies cut into blocks --------------------------------------------------------------
// The reference's signatures are blocking and in place on HOST memory: upload, two passes, download, 2 x 2.44 + 0.77 ms for
// a 2^22 polynomial, the link idle while the passes run and the SMs idle while it copies.  Pass A works on COLUMN tiles of the
// N1 x N2 input matrix and pass B's row tiles write COLUMNS of the N2 x N1 output matrix, so both copies are cut into column
// blocks (cudaMemcpy2DAsync: N1 or N2 row pieces of a few KB each): pass A of block k runs while block k + 1 is on the wire,
// pass B of block k + 1 while block k travels back — what stays exposed is one block's pass at either end.
// [safe_lo, safe_hi): the byte range of the buffer that is page-locked (the registration cache locks whole pages only).  Row
// pieces outside it (the first piece of the first row, the last of the last) go through bbg_hostcopy's mixed-memory path, the
// downloads among them at the very end because a copy into pageable memory blocks the calling thread.
namespace
{
int block_event(size_t i, cudaEvent_t* ev)
{
    while (g_block_events.size() <= i)
    {
        cudaEvent_t e;
        BBG_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        g_block_events.push_back(e);
    }
    *ev = g_block_events[i];
    return 0;
}
struct EdgePiece
{
    size_t off, len;
};
// columns [x0, x0 + w) (bytes) of `rows` rows of `pitch` bytes, between the host buffer h and its device image d
int copy_column_block(char* h, char* d, size_t x0, size_t w, size_t pitch, size_t rows, bool to_device, size_t safe_lo, size_t safe_hi, cudaStream_t st,
                      std::vector<EdgePiece>* deferred)
{
    const cudaMemcpyKind kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    // rows 1 .. rows - 2 lie inside the locked range whatever the block (a row is longer than a page)
    if (rows > 2)
    {
        const size_t off = pitch + x0;
        BBG_CHECK(to_device ? cudaMemcpy2DAsync(d + off, pitch, h + off, pitch, w, rows - 2, kind, st)
                            : cudaMemcpy2DAsync(h + off, pitch, d + off, pitch, w, rows - 2, kind, st));
    }
    for (int edge = 0; edge < (rows > 1 ? 2 : 1); ++edge)
    {
        const size_t off = (edge == 0 ? 0 : (rows - 1) * pitch) + x0;
        if (off >= safe_lo && off + w <= safe_hi)
            BBG_CHECK(to_device ? cudaMemcpyAsync(d + off, h + off, w, kind, st) : cudaMemcpyAsync(h + off, d + off, w, kind, st));
        else if (to_device)
            BBG_CHECK(bbg_hostcopy::h2d(d + off, h + off, w, st));
        else
            deferred->push_back({ off, w });
    }
    return 0;
}
} // namespace

bool ntt_host_blocks_applicable(unsigned log_n)
{
    static const unsigned lo = [] { const char* e = getenv("BBG_NTT_HOST_BLOCKS_MIN_LOG"); const int v = e ? atoi(e) : 0; return v >= 14 && v <= 22 ? (unsigned)v : 21u; }();
    return log_n >= lo && log_n <= 2 * (unsigned)TILE_LOG;
}

int ntt_host_blocks(void* h_coeffs, size_t safe_lo, size_t safe_hi, void* d_coeffs, unsigned log_n, int op, const uint64_t* constant, cudaStream_t st,
                    cudaStream_t copy_in, cudaStream_t copy_out)
{
    if (!ntt_host_blocks_applicable(log_n)) return 1002;
    if (op < OP_FFT || op > OP_COSET_FFT_WITH_CONSTANT) return 1003;
    const bool inverse = (op == OP_IFFT || op == OP_COSET_IFFT || op == OP_IFFT_WITH_CONSTANT);
    const bool coset = (op == OP_COSET_FFT || op == OP_COSET_IFFT || op == OP_COSET_FFT_WITH_CONSTANT);
    const bool with_constant = (op >= OP_FFT_WITH_CONSTANT);
    if (with_constant && constant == nullptr) return 1004;
    fe k = Fr::one();
    if (with_constant) k = load_fe(constant);
    const size_t n = (size_t)1 << log_n;
    int L1, L2;
    PassParams a, b;
    BBG_CHECK(setup_two_pass(d_coeffs, n, 1, log_n, inverse, coset, with_constant, k, st, a, b, L1, L2));
    const int tiles = (int)(n >> TILE_LOG);
    // Block width on the wire: 16 KiB row pieces.  Narrower pieces cost the copy engines more than the overlap returns
    // (coset_fft 2^22 from a page-locked buffer, one B200: whole-buffer copies 5.69 ms; 2 KiB pieces 6.65, 8 KiB 5.35,
    // 16 KiB 5.25 ms), which is also why sizes below 2^21 — whose rows leave no room for two such blocks — keep whole copies.
    size_t piece = 16384;
    if (const char* e = getenv("BBG_NTT_HOST_BLOCK_KB")) // development
    {
        const int v = atoi(e);
        if (v >= 1 && v <= 1024 && (v & (v - 1)) == 0) piece = (size_t)v << 10;
    }
    auto tiles_per_block = [&](int cl) {
        size_t t = (piece / 32) >> cl;
        if (t < 1) t = 1;
        if (t > (size_t)tiles) t = (size_t)tiles;
        return (int)t;
    };
    char* h = (char*)h_coeffs;
    char* d = (char*)d_coeffs;
    cudaEvent_t ev;
    // uploads may start once everything already queued on the work stream is done with the device image
    BBG_CHECK(block_event(0, &ev));
    BBG_CHECK(cudaEventRecord(ev, st));
    BBG_CHECK(cudaStreamWaitEvent(copy_in, ev, 0));
    size_t events_used = 1;
    {
        // pass A: tile t = columns [t << cl, (t + 1) << cl) of the 2^L1 x 2^L2 input matrix
        const int cl = TILE_LOG - L1;
        const size_t pitch = ((size_t)32) << L2, rows = (size_t)1 << L1;
        const int block_tiles = tiles_per_block(cl), blocks = tiles / block_tiles;
        for (int blk = 0; blk < blocks; ++blk)
        {
            const size_t x0 = ((size_t)(blk * block_tiles) << cl) * 32, w = ((size_t)block_tiles << cl) * 32;
            BBG_CHECK(copy_column_block(h, d, x0, w, pitch, rows, true, safe_lo, safe_hi, copy_in, nullptr));
            BBG_CHECK(block_event(events_used++, &ev));
            BBG_CHECK(cudaEventRecord(ev, copy_in));
            BBG_CHECK(cudaStreamWaitEvent(st, ev, 0));
            PassParams pa = a;
            pa.num_tiles = block_tiles;
            pa.total_work = block_tiles;
            pa.tile_begin = blk * block_tiles;
            BBG_CHECK(launch_pass<true>(L1, pa, st));
        }
    }
    std::vector<EdgePiece> deferred;
    {
        // pass B: tile t = rows [t << cl, ..) of the intermediate matrix = COLUMNS [t << cl, ..) of the 2^L2 x 2^L1 output matrix
        const int cl = TILE_LOG - L2;
        const size_t pitch = ((size_t)32) << L1, rows = (size_t)1 << L2;
        const int block_tiles = tiles_per_block(cl), blocks = tiles / block_tiles;
        for (int blk = 0; blk < blocks; ++blk)
        {
            PassParams pb = b;
            pb.num_tiles = block_tiles;
            pb.total_work = block_tiles;
            pb.tile_begin = blk * block_tiles;
            BBG_CHECK(launch_pass<false>(L2, pb, st));
            BBG_CHECK(block_event(events_used++, &ev));
            BBG_CHECK(cudaEventRecord(ev, st));
            BBG_CHECK(cudaStreamWaitEvent(copy_out, ev, 0));
            const size_t x0 = ((size_t)(blk * block_tiles) << cl) * 32, w = ((size_t)block_tiles << cl) * 32;
            BBG_CHECK(copy_column_block(h, d, x0, w, pitch, rows, false, safe_lo, safe_hi, copy_out, &deferred));
        }
    }
    for (const EdgePiece& e : deferred) BBG_CHECK(bbg_hostcopy::d2h(h + e.off, d + e.off, e.len, copy_out));
    BBG_CHECK(bbg_rt::sync(copy_out));
    return bbg_rt::sync(st);
}
#endif
} // namespace bbg
