// Pippenger multi-scalar multiplication over bn254 G1 for sm_100a.
//
// Replaces scalar_multiplication::pippenger / batched_scalar_multiplications
//   (reference curves/bn254/scalar_multiplication.cpp:457-476, :576-648, :650-772) for the value they
//   compute: sum_i k_i * P_i over the reference's interleaved 2n point table [P_i, phi(P_i)].
// The result is a group element, independent of bucket width, thread partition and summation order
// (SURVEY.md §8 note 3), so the pipeline below is free to be GPU-shaped:
//
//   1. digits   one thread per scalar: out of Montgomery form, the reference's endomorphism split
//               (field.hpp:413-485) restated exactly in 32-bit limbs -> two 128-bit half scalars for table
//               entries 2i and 2i+1, then signed c-bit window digits d in [-2^(c-1)+1, 2^(c-1)] (zero digits
//               are dropped: a zero scalar costs nothing, unlike the reference's always-odd wNAF which turns
//               it into a P + (-P) pair per round).  Histogram of (window, |d|) by global atomics.
//   2. scan     exclusive prefix sum of the W * 2^(c-1) bucket counts.
//   3. scatter  counting sort: (point index | sign) written at its bucket's cursor -> entries grouped by
//               (window, bucket).  [pippenger.md:49-76 is the authors' sketch of this bucket-ordered form.]
//   4. accumulate  the sorted entry array is cut into fixed slices of S entries, one thread per slice:
//               perfectly balanced whatever the digit distribution.  A thread walks its slice doing mixed
//               additions acc += +-P (XYZZ coordinates, 10 field products) and flushes at bucket boundaries;
//               runs cut by a slice boundary go to per-slice head/tail slots.
//   5. fix-up   one thread per bucket that spans slices: tail + heads.
//   6. reduce   window sum = sum_b (b+1) * B_b: threads take chunks of 8 buckets (local running sums, the
//               reference's :628-640 loop in miniature), then block-wide tree reductions produce per window the
//               plain sums and the log2(chunks) bit-sliced sums T_r = sum_{t: bit r of t} A_t.
//   7. finish   host: S_w = sum V + sum A + 8 * sum_r 2^r T_r, fold windows, normalise (bbg_host_g1.h).
//
// Work at n = 2^20 (c = 16, W = 8): 2^24 mixed adds (1.68e8 Fq products) + ~2^20 full adds; HBM: 64 MiB digits
// written + read, 64 MiB sorted entries, ~1 GiB of 64-byte point gathers mostly served from L2.
#include "bbg_internal.h"
#include "bbg_host_g1.h"
#include "bbg_hostcopy.h"

#include <atomic>
#include <chrono>
#include <stdlib.h>
#include <string.h>
#include <vector>
#ifndef BBG_EMULATE
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>
#endif

namespace bbg
{
namespace msmk
{
constexpr uint32_t NO_DIGIT = 0xffffffffu;
constexpr int CHUNK_LOG = 3; // buckets per running-sum thread in stage 6 (serial depth 2 * 2^CHUNK_LOG - 1 additions)

struct Plan
{
    size_t n;          // scalars
    size_t num_points; // 2n table entries
    int c;             // window bits
    int W;             // windows
    uint32_t NB;       // buckets per window = 2^(c-1)
    uint32_t total_buckets;
    uint32_t S;        // slice length
    size_t max_entries;
    size_t max_slices;
    int chunk_log;
    uint32_t chunks_per_window;
    int reduce_outputs; // per window: chunk bits + 2
    uint32_t red_splits; // blocks per (window, output) of the reduction, a power of two
    // fixed-base form (a table of pre-doubled windows, below): every window's digits fall into ONE bucket set and entry
    // (w, j) reads point j of window w's table, entry_stride entries further on
    int sets;              // bucket sets per MSM: W, or 1 in the fixed-base form
    uint32_t entry_stride; // 0, or the 2 * n_srs entries of one window's table
};

// little multi-word helpers for the endomorphism split (portable: host and device)
template <int NA, int NB_> BBG_HD void mul_wide(const uint32_t* a, const uint32_t* b, uint32_t* out)
{
#pragma unroll
    for (int i = 0; i < NA + NB_; ++i) out[i] = 0;
#pragma unroll
    for (int i = 0; i < NA; ++i)
    {
        uint32_t carry = 0;
#pragma unroll
        for (int j = 0; j < NB_; ++j)
        {
            const uint64_t t = (uint64_t)a[i] * b[j] + out[i + j] + carry;
            out[i + j] = (uint32_t)t;
            carry = (uint32_t)(t >> 32);
        }
        out[i + NB_] = carry;
    }
}

// field.hpp:413-485, restated on 32-bit words.  k canonical, non-Montgomery.  k1, k2: low 128 bits.
BBG_HD void split_endo(const fe& k, uint32_t k1[4], uint32_t k2[4])
{
    // g1 = {0x7a7bd9d4391eb18d, 0x4ccef014a773d2cf, 2}, g2 = {0xd91d232ec7e0b3d7, 2},
    // minus_b1 = {0x8211bbeb7d4f1128, 0x6f4d8248eeb859fc}, b2 = {0x89d3256894d213e3}   (field.hpp:420-426)
    const uint32_t g1c[5] = { 0x391eb18du, 0x7a7bd9d4u, 0xa773d2cfu, 0x4ccef014u, 0x2u };
    const uint32_t g2c[3] = { 0xc7e0b3d7u, 0xd91d232eu, 0x2u };
    const uint32_t mb1[4] = { 0x7d4f1128u, 0x8211bbebu, 0xeeb859fcu, 0x6f4d8248u };
    const uint32_t b2c[2] = { 0x94d213e3u, 0x89d32568u };
    uint32_t w1[11], w2[13];
    mul_wide<8, 3>(k.v, g2c, w1); // c1 = (g2 * k) >> 256 = w1[8..10]
    mul_wide<8, 5>(k.v, g1c, w2); // c2 = (g1 * k) >> 256 = w2[8..12]
    uint32_t q1w[7], q2w[7];
    mul_wide<3, 4>(w1 + 8, mb1, q1w); // q1 = c1 * (-b1)
    mul_wide<5, 2>(w2 + 8, b2c, q2w); // q2 = c2 * b2
    fe q1 = Fr::zero(), q2 = Fr::zero();
#pragma unroll
    for (int i = 0; i < 7; ++i)
    {
        q1.v[i] = q1w[i];
        q2.v[i] = q2w[i];
    }
    // t1 = q2 - q1 (+p on borrow)  : __sub, field_impl_int128.tcc:40-54
    fe t1, t1p;
    uint32_t pl[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) pl[i] = FrParams::P(i);
    const uint32_t borrow = cc::sub8(t1.v, q2.v, q1.v);
    cc::add8(t1p.v, t1.v, pl);
#pragma unroll
    for (int i = 0; i < 8; ++i) t1.v[i] = borrow ? t1p.v[i] : t1.v[i];
    // t2 = t1 * lambda (Montgomery product with lambda in Montgomery form = plain product), then k + t2 reduced once
    fe t2 = Fr::mul_full(t1, Fr::constant([](int i) { return FrParams::CUBE(i); }));
    fe s;
    cc::add8(s.v, k.v, t2.v);
    s = Fr::reduce(s);
#pragma unroll
    for (int i = 0; i < 4; ++i)
    {
        k1[i] = s.v[i];
        k2[i] = t1.v[i];
    }
}

// c bits of a 128-bit value starting at bit `pos` (bits beyond 127 read as zero)
BBG_HD uint32_t window_bits(const uint32_t k[4], int pos, int c)
{
    const int word = pos >> 5, shift = pos & 31;
    uint64_t lo = word < 4 ? k[word] : 0u;
    uint64_t hi = word + 1 < 4 ? k[word + 1] : 0u;
    const uint64_t v = (lo | (hi << 32)) >> shift;
    return (uint32_t)v & ((1u << c) - 1u);
}

// ---- 1. digits -------------------------------------------------------------------------------------
// (the counting atomic returns the entry's rank inside its bucket: the scatter then needs no atomics of its own)
// set_stride: NB when every window has its own bucket set, 0 when all windows share one (fixed-base form)
// A warp whose 32 lanes all count into the SAME bucket adds once (match.all, one instruction; lane 0 adds 32, each lane
// takes its place behind the returned base): a polynomial with long runs of equal coefficients — constants, selectors,
// 0 / 1 witnesses occur in the prover — otherwise sends 2^21 atomics per window to one address and the L2 atomic unit
// serialises them (constant scalars at 2^20: digit pass 0.57 -> 0.10 ms).  Grouping lanes by value (match.any) would also
// catch interleaved repetitions but iterates over the distinct values: +0.1 ms on uniform scalars, measured, so not that.
BBG_D uint32_t count_into_bucket(uint32_t* counts, uint32_t key, bool counted)
{
#if defined(__CUDA_ARCH__)
    int same = 0;
    __match_all_sync(0xffffffffu, counted ? key : 0xffffffffu, &same);
    if (same)
    {
        const unsigned lane = threadIdx.x & 31u;
        uint32_t base = 0;
        if (lane == 0 && counted) base = atomicAdd(&counts[key], 32u);
        return __shfl_sync(0xffffffffu, base, 0) + lane;
    }
#endif
    return counted ? atomicAdd(&counts[key], 1u) : 0u;
}
__global__ void msm_digits_kernel(const fe* scalars, size_t n, int c, int W, uint32_t NB, uint32_t set_stride, uint32_t* digits, uint32_t* ranks,
                                  uint32_t* counts)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool live = i < n; // (no early exit: every lane of a warp takes part in the aggregated counting)
    uint32_t half[2][4] = { { 0, 0, 0, 0 }, { 0, 0, 0, 0 } };
    if (live)
    {
        const fe k = Fr::from_mont(load_fe(scalars + i)); // scalar_multiplication.cpp:469-472
        split_endo(k, half[0], half[1]);
    }
    const size_t num_points = 2 * n;
    uint32_t carry[2] = { 0, 0 };
    for (int w = 0; w < W; ++w)
    {
        uint32_t packed[2], rank[2] = { 0, 0 };
#pragma unroll
        for (int h = 0; h < 2; ++h)
        {
            uint32_t d = window_bits(half[h], w * c, c) + carry[h];
            uint32_t neg = 0;
            carry[h] = 0;
            if (w + 1 < W && d > NB)
            {
                d = (1u << c) - d;
                neg = 0x80000000u;
                carry[h] = 1;
            }
            // a top-window digit above NB cannot occur for half scalars < 2^127 (the reference's wNAF makes the
            // same assumption, wnaf.hpp:11); clamp defensively so no bucket index is ever out of range
            if (d > NB) d = NB;
            const bool counted = live && d != 0;
            packed[h] = counted ? ((d - 1) | neg) : NO_DIGIT;
            rank[h] = count_into_bucket(counts, (uint32_t)((size_t)w * set_stride + (d - 1)), counted);
        }
        if (live)
        {
            uint2 pair;
            pair.x = packed[0];
            pair.y = packed[1];
            *(uint2*)(digits + (size_t)w * num_points + 2 * i) = pair;
            uint2 rpair;
            rpair.x = rank[0];
            rpair.y = rank[1];
            *(uint2*)(ranks + (size_t)w * num_points + 2 * i) = rpair;
        }
    }
}

// ---- 2. scan (three small kernels; counts -> exclusive offsets, offsets[total] = number of entries) ---
constexpr int SCAN_BLOCK = 256;
constexpr int SCAN_ITEMS = 8; // per thread
__global__ void scan_block_sums_kernel(const uint32_t* in, uint32_t count, uint32_t* block_sums)
{
    __shared__ uint32_t red[SCAN_BLOCK];
    const uint32_t base = blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + threadIdx.x * SCAN_ITEMS;
    uint32_t s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i)
    {
        if (base + i < count) s += in[base + i];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int off = SCAN_BLOCK / 2; off > 0; off >>= 1)
    {
        if ((int)threadIdx.x < off) red[threadIdx.x] += red[threadIdx.x + off];
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[blockIdx.x] = red[0];
}
// single block: exclusive scan of block_sums in place, total appended at block_sums[num_blocks]
__global__ void scan_spine_kernel(uint32_t* block_sums, uint32_t num_blocks)
{
    __shared__ uint32_t buf[SCAN_BLOCK];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < num_blocks; base += SCAN_BLOCK)
    {
        const uint32_t idx = base + threadIdx.x;
        const uint32_t v = idx < num_blocks ? block_sums[idx] : 0;
        buf[threadIdx.x] = v;
        __syncthreads();
        for (int off = 1; off < SCAN_BLOCK; off <<= 1) // Hillis-Steele inclusive scan
        {
            uint32_t add = (int)threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
            __syncthreads();
            buf[threadIdx.x] += add;
            __syncthreads();
        }
        const uint32_t incl = buf[threadIdx.x];
        const uint32_t start = running;
        if (idx < num_blocks) block_sums[idx] = start + incl - v;
        __syncthreads();
        if (threadIdx.x == SCAN_BLOCK - 1) running = start + incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_sums[num_blocks] = running;
}
__global__ void scan_apply_kernel(const uint32_t* in, uint32_t count, const uint32_t* block_sums, uint32_t* out)
{
    __shared__ uint32_t buf[SCAN_BLOCK];
    const uint32_t base = blockIdx.x * SCAN_BLOCK * SCAN_ITEMS + threadIdx.x * SCAN_ITEMS;
    uint32_t local[SCAN_ITEMS];
    uint32_t s = 0;
    for (int i = 0; i < SCAN_ITEMS; ++i)
    {
        local[i] = base + i < count ? in[base + i] : 0;
        s += local[i];
    }
    buf[threadIdx.x] = s;
    __syncthreads();
    for (int off = 1; off < SCAN_BLOCK; off <<= 1)
    {
        uint32_t add = (int)threadIdx.x >= off ? buf[threadIdx.x - off] : 0;
        __syncthreads();
        buf[threadIdx.x] += add;
        __syncthreads();
    }
    uint32_t run = block_sums[blockIdx.x] + buf[threadIdx.x] - s;
    for (int i = 0; i < SCAN_ITEMS; ++i)
    {
        if (base + i < count) out[base + i] = run;
        run += local[i];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == SCAN_BLOCK - 1) out[count] = block_sums[gridDim.x];
}

// ---- 3. scatter ------------------------------------------------------------------------------------
// W = windows of the whole (batched) pipeline, W1 = windows of one MSM.  Plain form: window w owns bucket set w and entry
// j is table entry j.  Fixed-base form (entry_stride != 0): MSM w / W1 owns one bucket set and entry (w mod W1, j) is entry
// j of that window's pre-doubled table.
__global__ void msm_scatter_kernel(const uint32_t* digits, const uint32_t* ranks, size_t num_points, int W, int W1, uint32_t NB, uint32_t entry_stride,
                                   const uint32_t* offsets, uint32_t* sorted)
{
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= num_points * (size_t)W) return;
    const uint32_t d = digits[e];
    if (d == NO_DIGIT) return;
    const size_t w = e / num_points;
    uint32_t j = (uint32_t)(e - w * num_points);
    size_t set = w;
    if (entry_stride != 0)
    {
        set = w / (size_t)W1;
        j += (uint32_t)(w - set * (size_t)W1) * entry_stride;
    }
    const size_t b = set * NB + (d & 0x7fffffffu);
    const uint32_t pos = offsets[b] + ranks[e];
    sorted[pos] = j | (d & 0x80000000u);
}

// ---- 4. accumulate ---------------------------------------------------------------------------------
// BBG_MSM_WIDE (default 1): the point gathers and the bucket stores of the accumulate pass as 256-bit accesses
#ifndef BBG_MSM_WIDE
#define BBG_MSM_WIDE 1
#endif
#if BBG_MSM_WIDE
#define MSM_LOAD_POINT(p) load_affine_const(p)
#define MSM_STORE_BUCKET(p, v) store_xyzz_global((p), (v))
#else
#define MSM_LOAD_POINT(p) load_affine(p)
#define MSM_STORE_BUCKET(p, v) store_xyzz((p), (v))
#endif
BBG_D affine_pt fetch_point(const fe* table, uint32_t entry)
{
    affine_pt p = MSM_LOAD_POINT(table + 2 * (size_t)(entry & 0x7fffffffu));
    if (entry >> 31) p.y = Fq::neg(p.y);
    return p;
}

// DIRECT: the entries ARE the points (what the pair-sum rounds below leave: affine points in bucket order, possibly the
// point at infinity), no index array in between.
template <bool DIRECT> BBG_D affine_pt accumulate_fetch(const fe* table, const uint32_t* sorted, uint32_t i)
{
    if (DIRECT)
    {
        affine_pt p;
        p.x = load_fe_wide(table + 2 * (size_t)i);
        p.y = load_fe_wide(table + 2 * (size_t)i + 1);
        return p;
    }
    return fetch_point(table, sorted[i]);
}
template <bool DIRECT>
__global__ void __launch_bounds__(128, 4) msm_accumulate_kernel(const uint32_t* sorted, const uint32_t* offsets, uint32_t total_buckets,
                                                            const fe* table, uint32_t S, fe* buckets, fe* head, fe* tail)
{
    const uint32_t E = offsets[total_buckets];
    const size_t slice = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t start64 = slice * S;
    if (start64 >= E) return;
    const uint32_t start = (uint32_t)start64;
    const uint32_t end = (E - start < S) ? E : start + S;
    // bucket containing `start`: offsets[b] <= start < offsets[b+1]
    uint32_t lo = 0, hi = total_buckets; // invariant: offsets[lo] <= start < offsets[hi]
    while (hi - lo > 1)
    {
        const uint32_t mid = (lo + hi) >> 1;
        if (offsets[mid] <= start) lo = mid; else hi = mid;
    }
    uint32_t b = lo;
    uint32_t bucket_end = offsets[b + 1];
    bool is_head = offsets[b] < start;
    xyzz_pt acc = G1::infinity();
    affine_pt next = accumulate_fetch<DIRECT>(table, sorted, start);
    for (uint32_t i = start; i < end; ++i)
    {
        if (i == bucket_end)
        {
            MSM_STORE_BUCKET(is_head ? head + 4 * slice : buckets + 4 * (size_t)b, acc);
            acc = G1::infinity();
            is_head = false;
            // next non-empty bucket: a few steps, then a binary search — behind a giant bucket (constant scalars: one per
            // window) tens of thousands of empty ones follow, and walking them one dependent load at a time made a single
            // thread the whole kernel's tail (ncu: the SMs idle for half of the 4.4 ms of that case)
            int steps = 0;
            do
            {
                ++b;
                bucket_end = offsets[b + 1];
            } while (bucket_end == i && ++steps < 4);
            if (bucket_end == i)
            {
                uint32_t lo_b = b, hi_b = total_buckets; // offsets[lo_b + 1] <= i < offsets[hi_b] ... find the first b with offsets[b + 1] > i
                while (hi_b - lo_b > 1)
                {
                    const uint32_t mid = lo_b + ((hi_b - lo_b) >> 1);
                    if (offsets[mid] <= i) lo_b = mid; else hi_b = mid;
                }
                b = lo_b;
                bucket_end = offsets[b + 1];
            }
        }
        const affine_pt cur = next;
        if (i + 1 < end) next = accumulate_fetch<DIRECT>(table, sorted, i + 1);
        if (!DIRECT || !G1::affine_is_infinity(cur)) acc = G1::madd(acc, cur);
    }
    fe* dst;
    if (is_head) dst = head + 4 * slice;              // bucket began before this slice
    else if (bucket_end <= end) dst = buckets + 4 * (size_t)b; // complete inside the slice
    else dst = tail + 4 * slice;                      // continues into the next slice
    MSM_STORE_BUCKET(dst, acc);
}

// ---- 3b. pair-sum rounds: batched AFFINE additions ahead of the accumulate pass -----------------------------------------
// A mixed addition into an XYZZ accumulator costs 10 field products.  Two AFFINE points add with one inversion, three
// products and a square — and inversions batch (Montgomery's trick: three more products per denominator plus one shared
// inversion).  The sorted entry array is therefore first folded by R rounds of pairwise affine additions inside the buckets
// (6 products per addition + the inversion's share), each round halving the entries, and only the remaining ~1/2^R of the
// additions run through the accumulate pass.  What the reference's callers observe is the sum, not the order (SURVEY §8
// note 3), so the result is unchanged.  NOT the default: on B200 the rounds turn a multiply-bound pass into an HBM-bound one
// and lose (pick_pair_rounds below; profiles/r02_msm_pair_rounds.md) — selectable with BBG_MSM_PAIR_ROUNDS, tested either way.
//
// One round: the input is an array of affine points partitioned into bucket regions [start[b], start[b+1]), the first
// count[b] positions of a region valid.  Positions are paired GLOBALLY, (2j, 2j+1) = pair slot j: two valid points of the
// same bucket are added, anything else (a lone first / last point of a bucket, which the alignment leaves over) is copied.
// The point at position p of bucket b goes to out_start[b] + (p >> 1) - (start[b] >> 1): with out_start[b] = (start[b] >> 1)
// + b the regions of the next round follow in closed form (msm_pair_plan_kernel), no scan between rounds; the last round
// writes to the dense offsets of a scan over the final counts, which is what the accumulate pass expects.
//
// A 256-thread CTA takes an item of 256 * B consecutive pair slots, warp w the slots [32 B w, 32 B (w + 1)), lane l every
// 32nd of them (coalesced point loads and stores).  Forward: each thread walks its B slots, classifies them and multiplies
// the denominators x2 - x1 (2 y for the rare P + P) into a running product, saving the prefix products to an L2-resident
// scratch.  Then ONE inversion per CTA: a product tree over the 256 totals in shared memory, Fermat inversion of the root
// by one thread, the tree walked back down.  Backward: each thread walks its slots in reverse, peels the inverse of each
// denominator off the running inverse and finishes the addition.  An intermediate point at infinity (P + (-P)) is marked by
// an all-ones top word of x and the reference's flag bit in y, values no reduced coordinate can take.
constexpr int PAIR_NT = 256;
constexpr int PAIR_CTAS_PER_SM = 2;
constexpr uint32_t PAIR_BMAX = 96;
constexpr uint32_t PAIR_BMIN = 8;
constexpr int PAIR_MAX_ROUNDS = 4;
constexpr uint32_t PAIR_BUCKET_MASK = (1u << 26) - 1u;
constexpr uint32_t PAIR_A_OUT = 1u << 26, PAIR_B_OUT = 1u << 27, PAIR_B_SAME = 1u << 28;
constexpr uint32_t PAIR_KIND_SHIFT = 29; // 0 nothing to compute, 1 add, 2 double, 3 the sum is the point at infinity
constexpr uint32_t PAIR_INF_WORD = 0xffffffffu;

struct PairParams
{
    const fe* src;             // points of this round (round 1: the point table, read through `sorted`)
    const uint32_t* sorted;    // round 1 only: entry = table index | sign << 31
    const uint32_t* in_start;  // [total_buckets + 1]
    const uint32_t* in_count;  // [total_buckets], nullptr: every position of a region is valid (round 1)
    const uint32_t* out_start; // [total_buckets]
    fe* dst;
    uint32_t total_buckets;
    uint32_t bmax; // pair slots per thread and item, <= PAIR_BMAX
    fe* prefix;    // scratch: gridDim.x * PAIR_BMAX * PAIR_NT field elements
    uint32_t* rec; // scratch: gridDim.x * PAIR_BMAX * PAIR_NT words
};

BBG_HD bool pair_is_infinity(const fe& x) { return x.v[7] == PAIR_INF_WORD; }
BBG_HD affine_pt pair_infinity()
{
    affine_pt r;
    r.x = Fq::zero();
    r.y = Fq::zero();
    r.x.v[7] = PAIR_INF_WORD;
    r.y.v[7] = 0x80000000u; // (group.hpp:133-151: what G1::affine_is_infinity tests)
    return r;
}
// smallest b' >= b with start[b' + 1] > p, given start[b] <= p < start[total_buckets]: the bucket region holding position p
BBG_D uint32_t pair_advance(const uint32_t* start, uint32_t total_buckets, uint32_t b, uint32_t p)
{
    for (int i = 0; i < 4; ++i)
    {
        if (start[b + 1] > p) return b;
        ++b;
    }
    uint32_t lo = b, hi = total_buckets; // start[lo] <= p < start[hi]  (runs of empty buckets: constant scalars leave 2^18 of them)
    while (hi - lo > 1)
    {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (start[mid] <= p) lo = mid; else hi = mid;
    }
    return lo;
}
BBG_D uint32_t pair_count(const PairParams& q, uint32_t b) { return q.in_count != nullptr ? q.in_count[b] : q.in_start[b + 1] - q.in_start[b]; }
template <bool FIRST> BBG_D fe pair_load_x(const PairParams& q, uint32_t p)
{
    if (FIRST) return load_fe_const(q.src + 2 * (size_t)(q.sorted[p] & 0x7fffffffu));
    return load_fe_wide(q.src + 2 * (size_t)p);
}
template <bool FIRST> BBG_D affine_pt pair_load(const PairParams& q, uint32_t p)
{
    if (FIRST) return fetch_point(q.src, q.sorted[p]);
    affine_pt r;
    r.x = load_fe_wide(q.src + 2 * (size_t)p);
    r.y = load_fe_wide(q.src + 2 * (size_t)p + 1);
    return r;
}
BBG_D void pair_store(const PairParams& q, uint32_t b, uint32_t j, const affine_pt& a)
{
    fe* d = q.dst + 2 * (size_t)(q.out_start[b] + j - (q.in_start[b] >> 1));
    store_fe_global(d, a.x);
    store_fe_global(d + 1, a.y);
}

BBG_D void pair_prefetch_l2(const void* p)
{
#if defined(__CUDA_ARCH__)
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); // (a hint: no register, no scoreboard entry)
#else
    (void)p;
#endif
}
// the point at position p, or (first round) the table entry e = sorted[p] already in a register
template <bool FIRST> BBG_D affine_pt pair_point(const PairParams& q, uint32_t p, uint32_t e)
{
    if (FIRST) return fetch_point(q.src, e);
    return pair_load<false>(q, p);
}

template <bool FIRST> __global__ void __launch_bounds__(PAIR_NT, PAIR_CTAS_PER_SM) msm_pair_round_kernel(PairParams q)
{
    __shared__ __align__(16) uint32_t tree[2 * PAIR_NT * 8]; // product tree over the threads' totals: node i = node 2i * node 2i+1, leaves 256..511
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    const uint32_t E = q.in_start[q.total_buckets];
    const uint32_t slots = (E >> 1) + (E & 1u);
    // slots per thread: whole waves of the grid at no more than bmax slots
    const uint32_t per_wave = (uint32_t)PAIR_NT * gridDim.x;
    const uint32_t waves = (slots + per_wave * q.bmax - 1) / (per_wave * q.bmax);
    uint32_t B = waves == 0 ? PAIR_BMIN : (slots + per_wave * waves - 1) / (per_wave * waves);
    if (B < PAIR_BMIN) B = PAIR_BMIN;
    if (B > q.bmax) B = q.bmax;
    const uint32_t item_slots = (uint32_t)PAIR_NT * B;
    const uint32_t items = (slots + item_slots - 1) / item_slots;
    fe* prefix = q.prefix + (size_t)blockIdx.x * PAIR_BMAX * PAIR_NT;
    uint32_t* rec = q.rec + (size_t)blockIdx.x * PAIR_BMAX * PAIR_NT;
    // second position of pair slot j (the last slot of an odd-length array has none: its first position stands in, unused)
    auto second = [&](uint32_t j) { return 2u * j + 1u < E ? 2u * j + 1u : 2u * j; };

    for (uint32_t item = blockIdx.x; item < items; item += gridDim.x)
    {
        const uint32_t wbase = item * item_slots + warp * 32u * B + lane;
        // ---- forward: classify, denominators, prefix products ----
        // The x coordinates travel one slot ahead of their use and, in the first round, their table indices two slots ahead:
        // a slot is one product, far too little to cover a gather from HBM behind an index load.
        fe run = Fq::one();
        uint32_t b = 0;
        fe nxa = Fq::zero(), nxb = Fq::zero();
        uint32_t ia = 0, ib = 0;
        if (wbase < slots)
        {
            b = pair_advance(q.in_start, q.total_buckets, 0, 2u * wbase);
            nxa = pair_load_x<FIRST>(q, 2u * wbase);
            nxb = pair_load_x<FIRST>(q, second(wbase));
            if (FIRST && B > 1 && wbase + 32u < slots)
            {
                ia = q.sorted[2u * (wbase + 32u)];
                ib = q.sorted[second(wbase + 32u)];
            }
        }
        for (uint32_t k = 0; k < B; ++k)
        {
            const uint32_t j = wbase + k * 32u;
            uint32_t r = 0;
            if (j < slots)
            {
                const fe xa = nxa, xb = nxb;
                const uint32_t jn = j + 32u;
                if (k + 1 < B && jn < slots)
                {
                    if (FIRST)
                    {
                        nxa = load_fe_const(q.src + 2 * (size_t)(ia & 0x7fffffffu));
                        nxb = load_fe_const(q.src + 2 * (size_t)(ib & 0x7fffffffu));
                        if (k + 2 < B && jn + 32u < slots)
                        {
                            ia = q.sorted[2u * (jn + 32u)];
                            ib = q.sorted[second(jn + 32u)];
                        }
                    }
                    else
                    {
                        nxa = load_fe_wide(q.src + 2 * (size_t)(2u * jn));
                        nxb = load_fe_wide(q.src + 2 * (size_t)second(jn));
                    }
                }
                const uint32_t p0 = 2u * j, p1 = p0 + 1u;
                b = pair_advance(q.in_start, q.total_buckets, b, p0);
                const uint32_t valid_end = q.in_start[b] + pair_count(q, b);
                const bool a_valid = p0 < valid_end;
                const bool same = p1 < q.in_start[b + 1];
                bool b_valid = false;
                if (same) b_valid = p1 < valid_end;
                else if (p1 < E)
                {
                    const uint32_t b1 = pair_advance(q.in_start, q.total_buckets, b + 1, p1);
                    b_valid = p1 < q.in_start[b1] + pair_count(q, b1);
                }
                r = b;
                if (a_valid && b_valid && same)
                {
                    if (pair_is_infinity(xa)) r |= PAIR_B_OUT | PAIR_B_SAME; // 0 + B
                    else if (pair_is_infinity(xb)) r |= PAIR_A_OUT;          // A + 0
                    else
                    {
                        fe d = Fq::sub(xb, xa);
                        uint32_t kind = 1;
                        if (Fq::is_zero(d))
                        {
                            const affine_pt A = pair_load<FIRST>(q, p0), Bp = pair_load<FIRST>(q, p1);
                            kind = 3; // A + (-A)
                            if (Fq::is_zero(Fq::sub(A.y, Bp.y)))
                            {
                                d = Fq::dbl(A.y); // A + A: slope 3 x^2 / 2 y
                                if (!Fq::is_zero(d)) kind = 2;
                            }
                        }
                        if (kind != 3)
                        {
                            store_fe_global(prefix + (size_t)k * PAIR_NT + tid, run);
                            run = Fq::mul(run, d);
                        }
                        r |= kind << PAIR_KIND_SHIFT;
                    }
                }
                else
                {
                    if (a_valid) r |= PAIR_A_OUT;
                    if (b_valid) r |= PAIR_B_OUT | (same ? PAIR_B_SAME : 0u);
                }
            }
            rec[(size_t)k * PAIR_NT + tid] = r;
        }
        // ---- one inversion for the CTA ----
        fe* node = (fe*)tree;
        node[PAIR_NT + tid] = run;
        for (uint32_t size = PAIR_NT / 2; size >= 1; size >>= 1)
        {
            __syncthreads();
            if (tid < size) node[size + tid] = Fq::mul(node[2 * (size + tid)], node[2 * (size + tid) + 1]);
        }
        __syncthreads();
#ifdef BBG_PAIR_FERMAT
        if (tid == 0) node[1] = Fq::invert(node[1]);
#else
        if (tid == 0) node[1] = Fq::invert_binary(node[1]); // (a lone thread: ~35 K cycles against ~200 K for the Fermat chain)
#endif
        for (uint32_t size = 1; size < PAIR_NT; size <<= 1)
        {
            __syncthreads();
            if (tid < size)
            {
                const uint32_t i = size + tid;
                const fe inv = node[i], left = node[2 * i], right = node[2 * i + 1];
                node[2 * i] = Fq::mul(inv, right);
                node[2 * i + 1] = Fq::mul(inv, left);
            }
        }
        __syncthreads();
        fe inv = node[PAIR_NT + tid]; // 1 / (this thread's product)
        __syncthreads();               // (the tree is rewritten by the next item)
        // ---- backward: finish the additions ----
        // The lines of the slot below are asked into L2 one slot ahead (in the first round behind indices read two ahead).
        uint32_t ca = 0, cb = 0, na = 0, nb = 0; // first round: table entries of this slot / of the one below
        auto entry_a = [&](uint32_t jj) { return jj < slots ? q.sorted[2u * jj] : 0u; };
        auto entry_b = [&](uint32_t jj) { return jj < slots ? q.sorted[second(jj)] : 0u; };
        if (FIRST)
        {
            const uint32_t jt = wbase + (B - 1) * 32u;
            ca = entry_a(jt);
            cb = entry_b(jt);
            if (B > 1)
            {
                na = entry_a(jt - 32u);
                nb = entry_b(jt - 32u);
            }
        }
        for (uint32_t k = B; k-- > 0;)
        {
            const uint32_t j = wbase + k * 32u, p0 = 2u * j, p1 = p0 + 1u;
            const uint32_t ea = ca, eb = cb;
            if (k > 0)
            {
                if (FIRST)
                {
                    if (j - 32u < slots)
                    {
                        pair_prefetch_l2(q.src + 2 * (size_t)(na & 0x7fffffffu));
                        pair_prefetch_l2(q.src + 2 * (size_t)(nb & 0x7fffffffu));
                    }
                    ca = na;
                    cb = nb;
                    if (k > 1)
                    {
                        na = entry_a(j - 64u);
                        nb = entry_b(j - 64u);
                    }
                }
                else if (j - 32u < slots)
                    pair_prefetch_l2(q.src + 4 * (size_t)(j - 32u)); // (both points of a slot share a 128-byte line)
            }
            const uint32_t r = rec[(size_t)k * PAIR_NT + tid];
            if ((r & ~PAIR_BUCKET_MASK) == 0) continue;
            const uint32_t b0 = r & PAIR_BUCKET_MASK, kind = r >> PAIR_KIND_SHIFT;
            if (kind == 1)
            {
                const affine_pt A = pair_point<FIRST>(q, p0, ea), Bp = pair_point<FIRST>(q, p1, eb);
                const fe inv_d = Fq::mul(inv, load_fe_global(prefix + (size_t)k * PAIR_NT + tid));
                inv = Fq::mul(inv, Fq::sub(Bp.x, A.x));
                const fe lam = Fq::mul(Fq::sub(Bp.y, A.y), inv_d);
                affine_pt S;
                S.x = Fq::sub(Fq::sub(Fq::sqr(lam), A.x), Bp.x);
                S.y = Fq::sub(Fq::mul(lam, Fq::sub(A.x, S.x)), A.y);
                pair_store(q, b0, j, S);
            }
            else if (kind == 2)
            {
                const affine_pt A = pair_point<FIRST>(q, p0, ea);
                const fe inv_d = Fq::mul(inv, load_fe_global(prefix + (size_t)k * PAIR_NT + tid));
                inv = Fq::mul(inv, Fq::dbl(A.y));
                const fe xx = Fq::sqr(A.x);
                const fe lam = Fq::mul(Fq::add(Fq::dbl(xx), xx), inv_d);
                affine_pt S;
                S.x = Fq::sub(Fq::sub(Fq::sqr(lam), A.x), A.x);
                S.y = Fq::sub(Fq::mul(lam, Fq::sub(A.x, S.x)), A.y);
                pair_store(q, b0, j, S);
            }
            else if (kind == 3)
                pair_store(q, b0, j, pair_infinity());
            if (r & PAIR_A_OUT) pair_store(q, b0, j, pair_point<FIRST>(q, p0, ea));
            if (r & PAIR_B_OUT)
            {
                const uint32_t b1 = (r & PAIR_B_SAME) ? b0 : pair_advance(q.in_start, q.total_buckets, b0 + 1, p1);
                pair_store(q, b1, j, pair_point<FIRST>(q, p1, eb));
            }
        }
    }
}

// Regions of every round from the sort's offsets, one thread per bucket (and one for the end marker b = total_buckets):
// start_{r+1}[b] = (start_r[b] >> 1) + b, count_{r+1}[b] = number of pair slots the count_r[b] valid positions touch.
// starts / counts: [rounds][total_buckets + 1], row r = the INPUT regions of round r (row 0 of starts is the offsets array
// itself and is not written); final_counts[b] = what is left after the last round (scanned into the dense offsets).
__global__ void msm_pair_plan_kernel(const uint32_t* offsets, uint32_t total_buckets, int rounds, uint32_t* starts, uint32_t* counts,
                                     uint32_t* final_counts)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b > total_buckets) return;
    uint32_t s = offsets[b];
    uint32_t n = b < total_buckets ? offsets[b + 1] - s : 0u;
    const size_t row = (size_t)total_buckets + 1;
    for (int r = 1; r <= rounds; ++r)
    {
        n = n == 0 ? 0u : ((s + n - 1u) >> 1) - (s >> 1) + 1u;
        s = (s >> 1) + b;
        if (r < rounds)
        {
            starts[(size_t)r * row + b] = s;
            counts[(size_t)r * row + b] = n;
        }
    }
    if (b < total_buckets) final_counts[b] = n;
}

// ---- 5. fix-up of buckets that span slices; empty buckets become infinity ----------------------------
// A bucket cut by slice edges = tail[s0] + head[s0+1] + ... + head[s1].  Short spans are summed by the bucket's own
// thread; long ones (one digit value shared by very many scalars: constant or highly repetitive polynomials do
// occur in the prover) are queued and reduced by a whole block each, so the cost of a giant bucket is
// O(span / 128 + log 128) additions instead of O(span).
constexpr uint32_t FIXUP_SERIAL_SPAN = 16;
constexpr uint32_t FIXUP_SUB = 2048; // a bucket spanning more slices than this is reduced in sub-spans of that length
constexpr int FIXUP_LARGE_GRID = 592; // 4 x 148: each block loops over queued buckets
constexpr int FIXUP_BLOCK = 128;
__global__ void msm_fixup_kernel(const uint32_t* offsets, uint32_t total_buckets, uint32_t S, fe* buckets, const fe* head, const fe* tail,
                                 uint32_t* work_count, uint32_t* work_list, uint32_t* giant_list)
{
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    const uint32_t o0 = offsets[b], o1 = offsets[b + 1];
    if (o0 == o1)
    {
        store_xyzz(buckets + 4 * (size_t)b, G1::infinity());
        return;
    }
    const uint32_t s0 = o0 / S, s1 = (o1 - 1) / S;
    if (s0 == s1) return;
    if (s1 - s0 > FIXUP_SUB)
    {
        giant_list[atomicAdd(work_count + 1, 1u)] = b; // reduced in sub-spans by many blocks (msm_fixup_giant_kernel)
        return;
    }
    if (s1 - s0 > FIXUP_SERIAL_SPAN)
    {
        work_list[atomicAdd(work_count, 1u)] = b;
        return;
    }
    xyzz_pt sum = load_xyzz(tail + 4 * (size_t)s0);
    for (uint32_t s = s0 + 1; s <= s1; ++s) sum = G1::add(sum, load_xyzz(head + 4 * (size_t)s));
    store_xyzz(buckets + 4 * (size_t)b, sum);
}
__global__ void __launch_bounds__(FIXUP_BLOCK) msm_fixup_large_kernel(const uint32_t* offsets, uint32_t S, fe* buckets, const fe* head, const fe* tail,
                                                                      const uint32_t* work_count, const uint32_t* work_list)
{
    __shared__ uint32_t sm[FIXUP_BLOCK * 32];
    const uint32_t count = *work_count;
    for (uint32_t item = blockIdx.x; item < count; item += gridDim.x)
    {
        const uint32_t b = work_list[item];
        const uint32_t o0 = offsets[b], o1 = offsets[b + 1];
        const uint32_t s0 = o0 / S, s1 = (o1 - 1) / S;
        xyzz_pt sum = G1::infinity();
        if (threadIdx.x == 0) sum = load_xyzz(tail + 4 * (size_t)s0);
        for (uint32_t s = s0 + 1 + threadIdx.x; s <= s1; s += FIXUP_BLOCK) sum = G1::add(sum, load_xyzz(head + 4 * (size_t)s));
        for (int off = FIXUP_BLOCK / 2; off > 0; off >>= 1)
        {
            __syncthreads();
            if ((int)threadIdx.x >= off && (int)threadIdx.x < 2 * off) store_xyzz(sm + 32 * (threadIdx.x - off), sum);
            __syncthreads();
            if ((int)threadIdx.x < off) sum = G1::add(sum, load_xyzz(sm + 32 * threadIdx.x));
        }
        if (threadIdx.x == 0) store_xyzz(buckets + 4 * (size_t)b, sum);
        __syncthreads();
    }
}
// Giant buckets (spans over FIXUP_SUB slices; constant scalars: a bucket of 2^20 entries spans ~19 000 slices and one block
// per BUCKET left all but a dozen SMs idle for a millisecond).  level 0: the heads s0 + 1 .. s1 are cut into sub-spans of
// FIXUP_SUB slices, one block each; the block's sum — for the first sub-span including the bucket's tail slot — replaces
// the first head of its sub-span.  level 1: one block per bucket adds those partial sums into the bucket.  The list is
// short (a giant bucket holds > 2048 slices of the entry array), so every block walks it to find its items.
__global__ void __launch_bounds__(FIXUP_BLOCK) msm_fixup_giant_kernel(const uint32_t* offsets, uint32_t S, fe* buckets, fe* head, const fe* tail,
                                                                      const uint32_t* work_count, const uint32_t* giant_list, int level)
{
    __shared__ uint32_t sm[FIXUP_BLOCK * 32];
    const uint32_t count = work_count[1];
    uint32_t item = 0; // running index over (bucket, sub-span) pairs at level 0, over buckets at level 1
    for (uint32_t w = 0; w < count; ++w)
    {
        const uint32_t b = giant_list[w];
        const uint32_t o0 = offsets[b], o1 = offsets[b + 1];
        const uint32_t s0 = o0 / S, s1 = (o1 - 1) / S;
        const uint32_t nsub = (s1 - s0 + FIXUP_SUB - 1) / FIXUP_SUB; // heads s0 + 1 .. s1
        const uint32_t first = level == 0 ? 0u : nsub, last = level == 0 ? nsub : nsub + 1; // level 1: one pseudo sub-span
        for (uint32_t sub = first; sub < last; ++sub, ++item)
        {
            if (item % gridDim.x != blockIdx.x) continue;
            xyzz_pt sum = G1::infinity();
            if (level == 0)
            {
                const uint32_t lo = s0 + 1 + sub * FIXUP_SUB;
                const uint32_t hi = lo + FIXUP_SUB - 1 < s1 ? lo + FIXUP_SUB - 1 : s1; // inclusive
                if (threadIdx.x == 0 && sub == 0) sum = load_xyzz(tail + 4 * (size_t)s0);
                for (uint32_t s = lo + threadIdx.x; s <= hi; s += FIXUP_BLOCK) sum = G1::add(sum, load_xyzz(head + 4 * (size_t)s));
            }
            else
            {
                for (uint32_t k = threadIdx.x; k < nsub; k += FIXUP_BLOCK) sum = G1::add(sum, load_xyzz(head + 4 * (size_t)(s0 + 1 + k * FIXUP_SUB)));
            }
            for (int off = FIXUP_BLOCK / 2; off > 0; off >>= 1)
            {
                __syncthreads();
                if ((int)threadIdx.x >= off && (int)threadIdx.x < 2 * off) store_xyzz(sm + 32 * (threadIdx.x - off), sum);
                __syncthreads();
                if ((int)threadIdx.x < off) sum = G1::add(sum, load_xyzz(sm + 32 * threadIdx.x));
            }
            if (threadIdx.x == 0)
            {
                if (level == 0) store_xyzz(head + 4 * (size_t)(s0 + 1 + sub * FIXUP_SUB), sum); // (every read of this sub-span is behind the barriers above)
                else store_xyzz(buckets + 4 * (size_t)b, sum);
            }
            __syncthreads();
        }
    }
}

// ---- 6a. chunk running sums: A_t = sum_v B, V_t = sum_v v * B  (v = 0 .. 2^chunk_log - 1) ---------------
__global__ void __launch_bounds__(128) msm_chunk_kernel(const fe* buckets, uint32_t total_chunks, int chunk_log, fe* A, fe* V)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= total_chunks) return;
    const fe* base = buckets + 4 * ((size_t)t << chunk_log);
    xyzz_pt run = G1::infinity(), acc = G1::infinity();
    for (int v = (1 << chunk_log) - 1; v >= 1; --v)
    {
        run = G1::add(run, load_xyzz(base + 4 * v));
        acc = G1::add(acc, run);
    }
    run = G1::add(run, load_xyzz(base));
    store_xyzz(A + 4 * (size_t)t, run);
    store_xyzz(V + 4 * (size_t)t, acc);
}

// ---- 6b. block tree reductions: out[w][r] -------------------------------------------------------------
// r < chunk_bits : sum of A_t over chunks t of window w with bit r set
// r = chunk_bits : sum of V_t;   r = chunk_bits + 1 : sum of A_t
// A window's chunks are cut into `splits` contiguous ranges of `len` chunks (both powers of two), one block each, so that a
// single large bucket set (the fixed-base form: 2^18 buckets in ONE set) is reduced by hundreds of blocks instead of a
// dozen; a block only visits the chunks that count for its output (bit r set), so all of its threads carry the same load.
// part[(w * outputs + r) * splits + s]; msm_reduce_final_kernel adds the `splits` partial sums.
// The blocks are latency-bound (a chain of full additions per thread, then eight tree levels), so what counts is that ALL of
// them are resident at once: two per SM at <= 128 registers, and `splits` chosen so that the grid fits that single wave.
constexpr int RED_BLOCK = 256;
constexpr int RED_BLOCKS_PER_SM = 2;
__global__ void __launch_bounds__(RED_BLOCK, RED_BLOCKS_PER_SM) msm_reduce_kernel(const fe* A, const fe* V, uint32_t chunks_per_window, int chunk_bits, uint32_t len, fe* part)
{
    __shared__ uint32_t sm[RED_BLOCK * 32];
    const int r = blockIdx.x, w = blockIdx.y;
    const uint32_t base = blockIdx.z * len;
    const fe* src = (r == chunk_bits ? V : A) + 4 * ((size_t)w * chunks_per_window + base);
    xyzz_pt sum = G1::infinity();
    if (r >= chunk_bits || (1u << r) >= len)
    {
        // every chunk of the range counts, or (bit r is a bit of the range's index) none does
        if (r >= chunk_bits || ((base >> r) & 1u))
            for (uint32_t t = threadIdx.x; t < len; t += RED_BLOCK) sum = G1::add(sum, load_xyzz(src + 4 * (size_t)t));
    }
    else
    {
        const uint32_t low = (1u << r) - 1;
        for (uint32_t u = threadIdx.x; u < (len >> 1); u += RED_BLOCK)
        {
            const uint32_t t = ((u & ~low) << 1) | (1u << r) | (u & low); // u with a one inserted at bit r
            sum = G1::add(sum, load_xyzz(src + 4 * (size_t)t));
        }
    }
    for (int off = RED_BLOCK / 2; off > 0; off >>= 1)
    {
        __syncthreads();
        if ((int)threadIdx.x >= off && (int)threadIdx.x < 2 * off) store_xyzz(sm + 32 * (threadIdx.x - off), sum);
        __syncthreads();
        if ((int)threadIdx.x < off) sum = G1::add(sum, load_xyzz(sm + 32 * threadIdx.x));
    }
    if (threadIdx.x == 0) store_xyzz(part + 4 * (((size_t)w * (chunk_bits + 2) + r) * gridDim.z + blockIdx.z), sum);
}
// out[o] = sum_s part[o * splits + s]; one block of blockDim.x (a power of two <= 64) threads per output
__global__ void __launch_bounds__(64) msm_reduce_final_kernel(const fe* part, uint32_t splits, fe* out)
{
    __shared__ uint32_t sm[32 * 32];
    const fe* src = part + 4 * (size_t)blockIdx.x * splits;
    xyzz_pt sum = G1::infinity();
    for (uint32_t s = threadIdx.x; s < splits; s += blockDim.x) sum = G1::add(sum, load_xyzz(src + 4 * (size_t)s));
    for (int off = (int)blockDim.x / 2; off > 0; off >>= 1)
    {
        __syncthreads();
        if ((int)threadIdx.x >= off && (int)threadIdx.x < 2 * off) store_xyzz(sm + 32 * (threadIdx.x - off), sum);
        __syncthreads();
        if ((int)threadIdx.x < off) sum = G1::add(sum, load_xyzz(sm + 32 * threadIdx.x));
    }
    if (threadIdx.x == 0) store_xyzz(out + 4 * (size_t)blockIdx.x, sum);
}

// ---- 6c. the same outputs with two additions per chunk instead of (chunk_bits + 2) / 2 ---------------------------------
// One halving tree over the chunk totals x_j yields EVERY bit-sliced sum: when the tree folds thread j + off into thread j
// (off = 128, 64, .., 1), the values held by threads [off, 2 off) are exactly the partial sums of the chunks whose index bit
// log2(off) is set; those threads keep them and halve among themselves in the remaining steps while the main tree carries
// on below them.  After eight steps thread 0 holds the block total and thread 2^h the sum over chunks with bit h set:
// 502 additions per 256 chunks, eight deep, every thread at most one addition per step.
// grid (blocks per set, sets, 2): z = 0 works on the A_t of chunks [256 B, 256 B + 256) of set w and writes T_0..T_7 and the
// block total, z = 1 is the plain sum of the V_t.  part[(w * nb + B) * 10 + o], o = 0..7: T_o, 8: total, 9: sum V.
constexpr int TREE_BLOCK = 256;
constexpr int TREE_OUT = 10;
BBG_D xyzz_pt bit_tree(xyzz_pt v, bool all_bits, uint32_t* sm)
{
    const int tid = (int)threadIdx.x;
    const int top = tid > 0 ? (1 << (31 - __clz(tid))) : 0; // highest set bit of tid: base of the group tid belongs to
    for (int off = TREE_BLOCK / 2; off > 0; off >>= 1)
    {
        __syncthreads();
        store_xyzz(sm + 32 * tid, v);
        __syncthreads();
        const bool active = tid < off || (all_bits && tid >= 2 * off && tid - top < off);
        if (active) v = G1::add(v, load_xyzz(sm + 32 * (tid + off)));
    }
    return v;
}
__global__ void __launch_bounds__(TREE_BLOCK, 2) msm_tree_reduce_kernel(const fe* A, const fe* V, uint32_t chunks_per_window, fe* part)
{
    __shared__ uint32_t sm[TREE_BLOCK * 32];
    const uint32_t B = blockIdx.x, w = blockIdx.y, nb = gridDim.x;
    const bool is_v = blockIdx.z != 0;
    const uint32_t t = B * TREE_BLOCK + threadIdx.x;
    xyzz_pt v = G1::infinity();
    if (t < chunks_per_window) v = load_xyzz((is_v ? V : A) + 4 * ((size_t)w * chunks_per_window + t));
    v = bit_tree(v, !is_v, sm);
    fe* dst = part + 4 * (((size_t)w * nb + B) * TREE_OUT);
    const int tid = (int)threadIdx.x;
    if (is_v)
    {
        if (tid == 0) store_xyzz(dst + 4 * 9, v);
    }
    else if (tid == 0) store_xyzz(dst + 4 * 8, v);
    else if ((tid & (tid - 1)) == 0) store_xyzz(dst + 4 * (31 - __clz(tid)), v);
}
// second level over the nb <= 256 blocks of a set, grid (10, sets): o < 8: T_o = sum_B part[B][o];  o = 8: sum V;
// o = 9: the bit tree over the block totals gives T_(8+h) and the sum of all A_t.  out[w][r] as msm_reduce_kernel writes it.
__global__ void __launch_bounds__(TREE_BLOCK, 2) msm_tree_final_kernel(const fe* part, uint32_t nb, int chunk_bits, fe* out)
{
    __shared__ uint32_t sm[TREE_BLOCK * 32];
    const int o = blockIdx.x, tid = (int)threadIdx.x;
    const uint32_t w = blockIdx.y;
    const int src_o = o < 8 ? o : (o == 8 ? 9 : 8);
    xyzz_pt v = G1::infinity();
    if ((uint32_t)tid < nb) v = load_xyzz(part + 4 * (((size_t)w * nb + tid) * TREE_OUT + src_o));
    v = bit_tree(v, o == 9, sm);
    fe* dst = out + 4 * ((size_t)w * (chunk_bits + 2));
    if (o < 8)
    {
        if (tid == 0 && o < chunk_bits) store_xyzz(dst + 4 * o, v);
    }
    else if (o == 8)
    {
        if (tid == 0) store_xyzz(dst + 4 * chunk_bits, v);
    }
    else if (tid == 0) store_xyzz(dst + 4 * (chunk_bits + 1), v);
    else if ((tid & (tid - 1)) == 0 && 8 + (31 - __clz(tid)) < chunk_bits) store_xyzz(dst + 4 * (8 + (31 - __clz(tid))), v);
}

// table[2i] = P_i, table[2i+1] = (beta x_i, -y_i)   (scalar_multiplication.cpp:131-140)
__global__ void endo_table_kernel(const fe* points, fe* table, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const affine_pt p = load_affine(points + 2 * i);
    store_affine(table + 4 * i, p);
    store_affine(table + 4 * i + 2, G1::endo_table_entry(p));
}
// ---- fixed-base tables (generate_pippenger_precompute_table, scalar_multiplication.cpp:90-129) -------------------------
// pre[w][j] = 2^(c w) * table[j] for w < W: with every window's point pre-doubled, the digits of ALL windows of a scalar
// can be added into one bucket set (pippenger_precomputed, :478-573, does exactly that on the CPU) — no doublings between
// windows, 1 / W of the buckets to reduce, and room for a wider window.  One thread per base point: c doublings per window,
// one inversion per stored point; the odd (endomorphism) entries follow from the even ones, phi(2^k P) = 2^k phi(P).
__global__ void __launch_bounds__(128) msm_precompute_kernel(const fe* table, fe* pre, size_t n, int c, int W)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const affine_pt p = load_affine(table + 4 * i);
    store_affine(pre + 4 * i, p);
    store_affine(pre + 4 * i + 2, load_affine(table + 4 * i + 2));
    xyzz_pt acc = G1::from_affine(p);
    for (int w = 1; w < W; ++w)
    {
        for (int k = 0; k < c; ++k) acc = G1::dbl(acc);
        const affine_pt a = G1::to_affine(acc);
        fe* dst = pre + (size_t)w * 4 * n + 4 * i;
        store_affine(dst, a);
        store_affine(dst + 2, G1::endo_table_entry(a));
        acc = G1::from_affine(a); // (keeps zz = 1: the next doublings start from the cheap form)
    }
}

// generate_pippenger_precompute_table's own output layout (scalar_multiplication.cpp:90-129): n PLAIN points in,
// table[i * n + j] = 2^((bits + 1)(i + 1)) * P_j for i < rounds - 1, canonical affine coordinates (the reference's
// batch_normalize leaves x z^-2, y z^-3 fully reduced).
__global__ void __launch_bounds__(128) precompute_plain_kernel(const fe* points, fe* out, size_t n, int bits_per_window, int rounds)
{
    const size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    const affine_pt p = load_affine(points + 2 * j);
    xyzz_pt acc = G1::affine_is_infinity(p) ? G1::infinity() : G1::from_affine(p);
    for (int i = 0; i + 1 < rounds; ++i)
    {
        for (int k = 0; k < bits_per_window; ++k) acc = G1::dbl(acc);
        const affine_pt a = G1::to_affine(acc);
        store_affine(out + 2 * ((size_t)i * n + j), a);
        if (!G1::affine_is_infinity(a)) acc = G1::from_affine(a);
    }
}

// ---- SRS loader ------------------------------------------------------------------------------
// io::read_transcript + read_g1_elements_from_buffer (io/io.hpp:76-98, :157-182) followed by
// generate_pippenger_point_table (scalar_multiplication.cpp:131-140) in one pass over the raw transcript bytes:
// monomials[0] = the G1 generator, monomials[i] = file point i - 1 (each coordinate: four 64-bit limbs, least significant
// limb first, big-endian bytes inside a limb, plain value) -> Montgomery form -> table[2i] = P_i, table[2i+1] = (beta x_i, -y_i).
BBG_D fe load_transcript_fq(const uint8_t* p)
{
    const uint4* q = (const uint4*)p;
    const uint4 lo = q[0], hi = q[1];
    const uint32_t w[8] = { lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w };
    fe r;
#pragma unroll
    for (int k = 0; k < 4; ++k)
    {
        // bswap64 of the limb (w[2k] low word, w[2k+1] high word as loaded on a little-endian machine)
        r.v[2 * k] = __byte_perm(w[2 * k + 1], 0, 0x0123);
        r.v[2 * k + 1] = __byte_perm(w[2 * k], 0, 0x0123);
    }
    return Fq::to_mont(r); // fq::__to_montgomery_form (field.hpp:224-232), canonical
}
__global__ void srs_from_transcript_kernel(const uint8_t* g1_bytes, fe* table, size_t n)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_pt p;
    if (i == 0)
    {
        p.x = Fq::one(); // g1::affine_one() (g1.hpp:13-15), io.hpp:177
        p.y = Fq::reduce(Fq::dbl(Fq::one()));
    }
    else
    {
        const uint8_t* src = g1_bytes + (i - 1) * 64;
        p.x = load_transcript_fq(src);
        p.y = load_transcript_fq(src + 32);
    }
    store_affine(table + 4 * i, p);
    store_affine(table + 4 * i + 2, G1::endo_table_entry(p));
}

// ---- synthetic point sets: (start + i * step) * G  (BASELINE configs[3]: "random multiples of the G1 generator") ----
constexpr int GEN_RUN = 32;
BBG_HD xyzz_pt scalar_mul_generator(const fe& k_mont)
{
    affine_pt g;
    g.x = Fq::one(); // generator (1, 2): g1.hpp:13-14
    g.y = Fq::dbl(Fq::one());
    const fe k = Fr::from_mont(k_mont);
    xyzz_pt acc = G1::infinity();
    for (int i = 253; i >= 0; --i)
    {
        acc = G1::dbl(acc);
        if ((k.v[i >> 5] >> (i & 31)) & 1u) acc = G1::madd(acc, g);
    }
    return acc;
}
__global__ void __launch_bounds__(64) g1_progression_kernel(fe a0, fe d, fe* points, size_t n)
{
    const size_t run = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t first = run * GEN_RUN;
    if (first >= n) return;
    const int count = (int)((n - first < (size_t)GEN_RUN) ? n - first : GEN_RUN);
    // scalar of the run's first point: a0 + first * d
    fe idx = Fr::zero();
    idx.v[0] = (uint32_t)first;
    idx.v[1] = (uint32_t)(first >> 32);
    const fe s = Fr::reduce(Fr::add(a0, Fr::mul(Fr::to_mont(idx), d)));
    const affine_pt step = G1::to_affine(scalar_mul_generator(d));
    xyzz_pt cur = scalar_mul_generator(s);
    // walk the run, remembering prefix products of zz * zzz for one shared inversion
    xyzz_pt pts[GEN_RUN];
    fe prefix[GEN_RUN];
    fe acc = Fq::one();
    for (int i = 0; i < count; ++i)
    {
        pts[i] = cur;
        prefix[i] = acc;
        if (!G1::is_infinity(cur)) acc = Fq::mul(acc, Fq::mul(cur.zz, cur.zzz));
        if (i + 1 < count) cur = G1::affine_is_infinity(step) ? cur : G1::madd(cur, step);
    }
    fe inv = Fq::invert(acc);
    for (int i = count - 1; i >= 0; --i)
    {
        affine_pt out;
        if (G1::is_infinity(pts[i]))
        {
            G1::affine_set_infinity(out);
        }
        else
        {
            const fe zi = Fq::mul(inv, prefix[i]); // 1 / (zz * zzz)
            out.x = Fq::mul_full(pts[i].x, Fq::mul(zi, pts[i].zzz));
            out.y = Fq::mul_full(pts[i].y, Fq::mul(zi, pts[i].zz));
            inv = Fq::mul(inv, Fq::mul(pts[i].zz, pts[i].zzz));
        }
        store_affine(points + 2 * (first + i), out);
    }
}
} // namespace msmk

// ================================================================================================
// Host driver
// ================================================================================================
namespace
{
using namespace msmk;
std::atomic<size_t> g_msm_launches{ 0 };

struct Workspace
{
    void* p = nullptr;
    size_t bytes = 0;
    int ensure(size_t need)
    {
        if (need <= bytes) return 0;
        if (p) bbg_rt::dev_free(p);
        p = nullptr;
        bytes = 0;
        const int e = bbg_rt::dev_alloc(&p, need);
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

// An MSM in flight: everything the host finish needs once the kernels are done.  launch queues the kernels and the
// copy of the per-window reductions into a pinned host slot and returns a ticket; finish waits for that copy and folds
// the windows.  Two workspaces let MSMs on two streams overlap (the latency-bound tail kernels of one under the accumulate
// pass of the next); several tickets may be pending per workspace because the copy is ordered behind the kernels on the
// launching stream.
constexpr int MSM_TICKETS = 6;
constexpr int MAX_DEVICES = 16;
struct PeerJob;
struct MsmTicket
{
    bool pending = false;
    bool zero = false;   // n == 0: every sum is the point at infinity
    Plan single{};
    Plan pl{};
    size_t batch = 0;
    size_t red_count = 0;
    void* host_red = nullptr;
    size_t host_bytes = 0;
    // multi-GPU: the shards of this MSM that run on the other devices (primary context only)
    PeerJob* peer_jobs[MAX_DEVICES] = {};
    int peer_count = 0;
#ifndef BBG_EMULATE
    cudaEvent_t done = nullptr;
    cudaEvent_t scalars_ready = nullptr; // recorded on the launching stream before the peers read the scalars
#endif
};

// Everything one device needs to run MSMs: the primary device owns g_primary, every other device of a multi-GPU
// library instance has its own inside its worker (below).
// a point table of this device with its pre-doubled windows (fixed-base form)
struct FixedBase
{
    const char* base;  // the plain 2 * n_srs-entry table the MSM calls name
    size_t n_srs;
    int c, W;
    void* pre;         // W x 2 * n_srs entries
};
struct MsmContext
{
    Workspace ws[2];
    MsmTicket tickets[MSM_TICKETS];
    std::vector<FixedBase> fixed; // (touched by the thread that drives the device only)
};
MsmContext g_primary;

int ticket_host_buffer(MsmTicket& t, size_t bytes)
{
    if (bytes <= t.host_bytes) return 0;
#ifndef BBG_EMULATE
    if (t.host_red) cudaFreeHost(t.host_red);
    t.host_red = nullptr;
    t.host_bytes = 0;
    BBG_CHECK(cudaHostAlloc(&t.host_red, bytes, cudaHostAllocDefault));
#else
    free(t.host_red);
    t.host_red = malloc(bytes);
    if (t.host_red == nullptr) return 2;
#endif
    t.host_bytes = bytes;
    return 0;
}

void context_release(MsmContext& ctx)
{
    for (Workspace& w : ctx.ws) w.release();
    for (FixedBase& f : ctx.fixed) bbg_rt::dev_free(f.pre);
    ctx.fixed.clear();
    for (MsmTicket& t : ctx.tickets)
    {
#ifndef BBG_EMULATE
        if (t.host_red) cudaFreeHost(t.host_red);
        if (t.done) cudaEventDestroy(t.done);
        if (t.scalars_ready) cudaEventDestroy(t.scalars_ready);
        t.done = nullptr;
        t.scalars_ready = nullptr;
#else
        free(t.host_red);
#endif
        t.host_red = nullptr;
        t.host_bytes = 0;
        t.pending = false;
    }
}

// Window plan.  Cost model in units of one mixed addition, fitted to B200 measurements (r01, 2^17 .. 2^26 points):
//   per entry  1.13  (accumulate 0.16 ns + histogram / scatter atomics), + 0.15 when there are fewer than 2^16 buckets
//              in total (few distinct counters: the L2 atomic units serialise);
//   per bucket 11 (chunk + reduce + fix-up, ~1.8 ns) + 0.3 * avg_bucket_size / 64 (buckets cut by slice edges).
// A narrow top window (few real bits left over) funnels all 2n of its entries into 2^top buckets — heavy atomic
// contention and giant buckets — so it is penalised; e.g. c = 21 at 2^25 points left a 1-bit top window and cost
// 158 ms against 88 ms for c = 19.
void pick_windows(size_t n, int& c_out, int& W_out)
{
    int lg = 0;
    while (((size_t)1 << lg) < 2 * n) ++lg;
    const int target = lg - 5;
    int c_lo = target - 4 < 5 ? 5 : target - 4;
    if (c_lo > 22) c_lo = 22;
    int c_hi = target + 3 > 22 ? 22 : target + 3;
    if (c_hi < c_lo) c_hi = c_lo; // tiny n: the range collapses to the smallest window
    if (const char* e = getenv("BBG_MSM_WINDOW")) // development override
    {
        const int v = atoi(e);
        if (v >= 2 && v <= 22) c_lo = c_hi = v;
    }
    double best = -1;
    c_out = c_lo;
    W_out = (128 + c_lo - 1) / c_lo;
    for (int c = c_lo; c <= c_hi; ++c)
    {
        int W = (128 + c - 1) / c;
        while ((W - 1) * c >= 127) --W;
        const int top_bits = 127 - (W - 1) * c;
        const double buckets = (double)((size_t)1 << (c - 1));
        const double avg = 2.0 * (double)n / buckets;
        const double per_entry = 1.13 + ((double)W * buckets < 65536.0 ? 0.15 : 0.0);
        double cost = (double)W * 2.0 * (double)n * per_entry + (double)W * buckets * (11.0 + 0.3 * avg / 64.0);
        if (top_bits < 10) cost += 2.0 * (double)n * (10 - top_bits) / 4.0;
        if (best < 0 || cost < best)
        {
            best = cost;
            c_out = c;
            W_out = W;
        }
    }
}

// Slice length of the accumulate pass.  Every thread performs exactly S mixed additions, so the pass runs in waves of
// 128 x 4 x #SM threads and a partly filled last wave is pure loss: 2^24 entries at S = 64 are 3.46 waves, i.e. 87%
// (what ncu showed for the multiply pipe).  S is chosen so that the slices fill a whole number of waves, near 64 entries
// per thread for large inputs and down to 16 for small ones (which otherwise occupy a fraction of the machine).
uint32_t pick_slice(size_t max_entries)
{
    if (const char* e = getenv("BBG_MSM_SLICE")) // development override
    {
        const int v = atoi(e);
        if (v >= 8 && v <= 512) return (uint32_t)v;
    }
    const double wave = 128.0 * 4.0 * (double)bbg_rt::num_sms();
    const double per_wave_thread = (double)max_entries / wave; // entries per thread if everything ran in ONE wave
    if (per_wave_thread <= 16.0) return 16;
    const double target = max_entries >= ((size_t)1 << 26) ? 128.0 : 64.0;
    size_t k = (size_t)(per_wave_thread / target + 0.999); // waves at ~target entries per thread
    if (k < 1) k = 1;
    uint32_t S = (uint32_t)(per_wave_thread / (double)k + 0.999);
    if (S < 16) S = 16;
    return S;
}

// Window width of a fixed-base table for MSMs of about n points: one bucket set of 2^(c-1) buckets whatever the number of
// windows, so the per-bucket cost is paid once and c can grow.  Costs in mixed-addition equivalents, fitted to B200
// measurements (tools/msm_fixed_base.py, 2^17..2^20): 1.13 per sorted entry; ~6 per bucket (fix-up, chunk sums and the split
// tree reduction, which all scale with the ONE bucket set); and the top window, whose 127 - (W-1) c bits put all of its 2n
// digits into the lowest 2^(top-1) buckets of the shared set: beyond ~2048 entries per bucket the counting atomics of the
// digit pass contend and the fix-up sees giant buckets (measured at 2^20: c = 17, top 8 bits, digits 0.16 -> 0.57 ms,
// fix-up 0.11 -> 0.20 ms; c = 18, top 1 bit: +0.7 / +1.1 ms).
void pick_windows_fixed_base(size_t n, int& c_out, int& W_out)
{
    double best = -1;
    c_out = 8;
    W_out = 16;
    for (int c = 8; c <= 22; ++c)
    {
        int W = (128 + c - 1) / c;
        while ((W - 1) * c >= 127) --W;
        if (W > 16) continue;
        const int top_bits = 127 - (W - 1) * c;
        const double buckets = (double)((size_t)1 << (c - 1));
        double cost = (double)W * 2.0 * (double)n * 1.13 + buckets * 6.0;
        const double per_top_bucket = 2.0 * (double)n / (double)((size_t)1 << (top_bits - 1));
        if (per_top_bucket > 2048.0) cost += 2.0 * (double)n * (per_top_bucket / 2048.0 < 8.0 ? per_top_bucket / 2048.0 : 8.0) * 0.25;
        if (best < 0 || cost < best)
        {
            best = cost;
            c_out = c;
            W_out = W;
        }
    }
}

// fixed_c != 0: the fixed-base form over a table pre-doubled for that window width
Plan make_plan(size_t n, int fixed_c = 0, uint32_t entry_stride = 0)
{
    Plan pl;
    pl.n = n;
    pl.num_points = 2 * n;
    if (fixed_c != 0)
    {
        pl.c = fixed_c;
        pl.W = (128 + fixed_c - 1) / fixed_c;
        while ((pl.W - 1) * pl.c >= 127) --pl.W;
    }
    else
        pick_windows(n, pl.c, pl.W);
    pl.sets = fixed_c != 0 ? 1 : pl.W;
    pl.entry_stride = fixed_c != 0 ? entry_stride : 0;
    pl.NB = 1u << (pl.c - 1);
    pl.total_buckets = pl.NB * (uint32_t)pl.sets;
    pl.max_entries = pl.num_points * (size_t)pl.W;
    pl.S = pick_slice(pl.max_entries);
    pl.max_slices = (pl.max_entries + pl.S - 1) / pl.S;
    // buckets per running-sum thread: 8 for large bucket sets; a small MSM (a device's share of a multi-GPU commitment: 2^15
    // buckets) has too few chunks of 8 to occupy the machine and each thread's 15 dependent additions are its whole run time:
    // measured at 2^17 points (fixed-base, 2^15 buckets) 0.748 ms with chunks of 8, 0.701 with 4, 0.677 with 2; at 2^20 points
    // (2^18 buckets) 2.95 / 3.05 / 3.60 ms (profiles/r02_msm_chunk_log.jsonl)
    const size_t set_buckets = ((size_t)1 << (pl.c - 1)) * (size_t)pl.sets;
    int chunk_log = set_buckets >= ((size_t)1 << 17) ? CHUNK_LOG : set_buckets >= ((size_t)1 << 16) ? 2 : 1;
    if (const char* e = getenv("BBG_MSM_CHUNK_LOG")) // development
    {
        const int v = atoi(e);
        if (v >= 1 && v <= 5) chunk_log = v;
    }
    pl.chunk_log = chunk_log < pl.c - 1 ? chunk_log : pl.c - 1;
    pl.chunks_per_window = pl.NB >> pl.chunk_log;
    int bits = 0;
    while ((1u << bits) < pl.chunks_per_window) ++bits;
    pl.reduce_outputs = bits + 2;
    pl.red_splits = 1; // (set per launch: depends on the batch and on the device's SM count)
    return pl;
}

// Rounds of pairwise affine additions ahead of the accumulate pass (3b).  OFF unless BBG_MSM_PAIR_ROUNDS asks for them:
// measured on B200 (r02, profiles/r02_msm_pair_rounds.md) the rounds need 30% fewer multiply-pipe slots than the accumulate
// pass they replace but are bound by HBM instead: every entry's point is gathered twice (x for the denominators, x and y
// again once the inverses exist), a 64-byte point costs a 128-byte line from a table far larger than L2, and round one
// moves 4.8 GB (3.2 TB/s, the multiply pipe 42% busy) where the accumulate pass moves 1.9 GB at 88% — 2^20 points over
// fixed-base windows: 3.33 ms with one round, 3.50 ms with three, against 3.04 ms without.
int pick_pair_rounds(size_t max_entries, uint32_t total_buckets)
{
    (void)max_entries;
    int rounds = 0;
    if (const char* e = getenv("BBG_MSM_PAIR_ROUNDS")) // development / tests
    {
        const int v = atoi(e);
        if (v >= 0 && v <= PAIR_MAX_ROUNDS && total_buckets <= PAIR_BUCKET_MASK) rounds = v;
    }
    return rounds;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// `batch` MSMs of the same size over the same point table in ONE pipeline on the current device: MSM b's windows become
// the virtual windows b * W .. b * W + W - 1 of a single sort / accumulate / reduce pass, so the latency-bound tail
// kernels (scan, fix-up, chunk, reduce) and the host round trip are paid once per batch (the prover commits 3 + 1 + 3 + 2
// polynomials per proof against the same SRS, prover.cpp:65-124, :640-652).
int context_launch(MsmContext& ctx, int id, int workspace, const void* const* d_scalars, size_t batch, const void* d_table, size_t n, cudaStream_t st)
{
    MsmTicket& tk = ctx.tickets[id];
    tk.batch = batch;
    tk.zero = (n == 0);
    if (n == 0)
    {
        tk.pending = true;
        return 0;
    }
    Workspace& g_ws = ctx.ws[workspace];
    if (2 * n > ((size_t)1 << 28)) return 1008;
    // a table registered with pre-doubled windows (msm_fixed_base_build) switches to the fixed-base form
    int fixed_c = 0;
    uint32_t entry_stride = 0;
    for (const FixedBase& f : ctx.fixed)
    {
        if ((const char*)d_table >= f.base && (const char*)d_table + n * 128 <= f.base + f.n_srs * 128)
        {
            fixed_c = f.c;
            entry_stride = (uint32_t)(2 * f.n_srs);
            d_table = (const char*)f.pre + ((const char*)d_table - f.base);
            break;
        }
    }
    const Plan single = make_plan(n, fixed_c, entry_stride);
    if (single.c < 2 || single.c > 22 || single.W < 1 || single.W > 64 || (single.W - 1) * single.c >= 127 || single.W * single.c < 128)
        return 1009; // planner invariant
    if (single.max_entries * batch >= ((size_t)1 << 32) || (size_t)single.total_buckets * batch >= ((size_t)1 << 31)) return 1008;
    Plan pl = single; // the batch as one MSM with batch * W windows
    pl.W = single.W * (int)batch;
    pl.sets = single.sets * (int)batch;
    pl.total_buckets = single.total_buckets * (uint32_t)batch;
    pl.max_entries = single.max_entries * batch;
    pl.max_slices = (pl.max_entries + pl.S - 1) / pl.S;

    // pair-sum rounds (3b) fold the sorted entries before the accumulate pass, which then runs on what is left
    const int pair_rounds = pick_pair_rounds(pl.max_entries, pl.total_buckets);
    const size_t pair_row = (size_t)pl.total_buckets + 1;
    const unsigned pair_grid = (unsigned)(PAIR_CTAS_PER_SM * bbg_rt::num_sms());
    size_t pair_cap[PAIR_MAX_ROUNDS + 1] = { pl.max_entries }; // entries (with the holes between regions) after round r
    for (int r = 1; r <= pair_rounds; ++r) pair_cap[r] = pair_cap[r - 1] / 2 + pl.total_buckets + 1;
    uint32_t S_acc = pl.S;
    size_t acc_slices = pl.max_slices;
    if (pair_rounds > 0)
    {
        if (pair_cap[1] >= ((size_t)1 << 32)) return 1008;
        S_acc = pick_slice(pair_cap[pair_rounds]);
        acc_slices = (pair_cap[pair_rounds] + S_acc - 1) / S_acc;
    }

    // carve the workspace
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes); return o; };
    const size_t o_digits = carve(pl.max_entries * 4);
    const size_t o_ranks = carve(pl.max_entries * 4);
    const size_t o_sorted = carve(pl.max_entries * 4 + 16);
    const size_t o_counts = carve((size_t)pl.total_buckets * 4);
    const size_t o_work_count = carve(256); // directly after counts: zeroed by the same memset
    const size_t o_offsets = carve(((size_t)pl.total_buckets + 1) * 4);
    const uint32_t scan_blocks = (pl.total_buckets + SCAN_BLOCK * SCAN_ITEMS - 1) / (SCAN_BLOCK * SCAN_ITEMS);
    const size_t o_spine = carve(((size_t)scan_blocks + 1) * 4);
    const size_t o_buckets = carve((size_t)pl.total_buckets * 128);
    const size_t o_head = carve(acc_slices * 128);
    const size_t o_tail = carve(acc_slices * 128);
    const size_t o_work_list = carve((acc_slices / FIXUP_SERIAL_SPAN + 2) * 4);
    const size_t o_giant_list = carve((acc_slices / FIXUP_SUB + 2) * 4);
    const uint32_t total_chunks = pl.chunks_per_window * (uint32_t)pl.sets;
    const size_t o_A = carve((size_t)total_chunks * 128);
    const size_t o_V = carve((size_t)total_chunks * 128);
    const size_t red_count = (size_t)pl.sets * pl.reduce_outputs;
    const size_t o_red = carve(red_count * 128);
    // Which reduction: both are latency-bound chains of full additions (~3.7 us each for a lone warp, measured).  Bucket sets
    // of 2^11 .. 2^16 chunks take the two-additions-per-chunk tree (6c: two launches, 2 x 9 additions deep whatever the
    // size: 0.125 ms against 0.185 / 0.33 ms for 2^12 / 2^15 chunks per set at 2^20 points); smaller sets are quicker through
    // the single launch of the bit-sliced reduction (6b: chunks / 256 + 8 deep; 0.068 against 0.105 ms at 2^17 points), which
    // is also kept beyond 2^16 chunks (c >= 21: sizes where this tail no longer counts), cut into the largest power of two
    // of ranges per (set, output) that still fits one resident wave.
    uint32_t tree_blocks = (pl.chunks_per_window > 1024u && pl.chunks_per_window <= 256u * TREE_BLOCK) ? pl.chunks_per_window / TREE_BLOCK : 0;
    if (const char* e = getenv("BBG_MSM_TREE_REDUCE")) // development / tests: 0 / 1 force the bit-sliced reduction / the tree
    {
        if (e[0] == '0') tree_blocks = 0;
        else if (pl.chunks_per_window <= 256u * TREE_BLOCK) tree_blocks = (pl.chunks_per_window + TREE_BLOCK - 1) / TREE_BLOCK;
    }
    while (tree_blocks == 0 && red_count * (pl.red_splits * 2) <= (size_t)(RED_BLOCKS_PER_SM * bbg_rt::num_sms()) &&
           pl.chunks_per_window / (pl.red_splits * 2) >= (uint32_t)RED_BLOCK)
        pl.red_splits *= 2;
    if (const char* e = getenv("BBG_MSM_RED_SPLITS")) // tests: force the split path where the SM count would not choose it
    {
        const uint32_t v = (uint32_t)atoi(e);
        if (v >= 1 && (v & (v - 1)) == 0 && pl.chunks_per_window / v >= 1) pl.red_splits = v;
    }
    const size_t o_red_part = carve(tree_blocks != 0 ? (size_t)pl.sets * tree_blocks * TREE_OUT * 128 : red_count * pl.red_splits * 128);
    // pair-sum rounds: two point buffers (the rounds alternate between them), the regions of the rounds, the dense offsets of
    // what is left, and the CTAs' scratch
    size_t o_pair_x = 0, o_pair_y = 0, o_pair_starts = 0, o_pair_counts = 0, o_pair_offsets = 0, o_pair_prefix = 0, o_pair_rec = 0;
    if (pair_rounds > 0)
    {
        o_pair_x = carve(pair_cap[1] * 64);
        o_pair_y = carve(pair_rounds > 1 ? pair_cap[2] * 64 : 0);
        o_pair_starts = carve((size_t)pair_rounds * pair_row * 4);
        o_pair_counts = carve((size_t)pair_rounds * pair_row * 4);
        o_pair_offsets = carve(pair_row * 4);
        o_pair_prefix = carve((size_t)pair_grid * PAIR_BMAX * PAIR_NT * 32);
        o_pair_rec = carve((size_t)pair_grid * PAIR_BMAX * PAIR_NT * 4);
    }
    BBG_CHECK(g_ws.ensure(off));
    char* ws = (char*)g_ws.p;
    uint32_t* digits = (uint32_t*)(ws + o_digits);
    uint32_t* ranks = (uint32_t*)(ws + o_ranks);
    uint32_t* sorted = (uint32_t*)(ws + o_sorted);
    uint32_t* counts = (uint32_t*)(ws + o_counts);
    uint32_t* offsets = (uint32_t*)(ws + o_offsets);
    uint32_t* spine = (uint32_t*)(ws + o_spine);
    fe* buckets = (fe*)(ws + o_buckets);
    fe* head = (fe*)(ws + o_head);
    fe* tail = (fe*)(ws + o_tail);
    uint32_t* work_count = (uint32_t*)(ws + o_work_count);
    uint32_t* work_list = (uint32_t*)(ws + o_work_list);
    uint32_t* giant_list = (uint32_t*)(ws + o_giant_list);
    fe* A = (fe*)(ws + o_A);
    fe* V = (fe*)(ws + o_V);
    fe* red = (fe*)(ws + o_red);

    // counts and the fix-up work counter are adjacent: one memset
    BBG_CHECK(bbg_rt::dev_memset(counts, 0, (o_work_count - o_counts) + 256, st));
    {
        bbg_prof::Scope prof(bbg_prof::MSM_DIGITS, st);
        for (size_t b = 0; b < batch; ++b)
            BBG_LAUNCH_NOSYNC(msm_digits_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), st, (const fe*)d_scalars[b], n, pl.c, single.W, pl.NB,
                              pl.entry_stride != 0 ? 0u : pl.NB, digits + b * single.max_entries, ranks + b * single.max_entries, counts + b * single.total_buckets);
    }
    bbg_prof::Scope* prof_scan = new bbg_prof::Scope(bbg_prof::MSM_SCAN, st);
    BBG_LAUNCH(scan_block_sums_kernel, dim3(scan_blocks), dim3(SCAN_BLOCK), 0, st, (const uint32_t*)counts, pl.total_buckets, spine);
    BBG_LAUNCH(scan_spine_kernel, dim3(1), dim3(SCAN_BLOCK), 0, st, spine, scan_blocks);
    BBG_LAUNCH(scan_apply_kernel, dim3(scan_blocks), dim3(SCAN_BLOCK), 0, st, (const uint32_t*)counts, pl.total_buckets, (const uint32_t*)spine, offsets);
    delete prof_scan;
    {
        bbg_prof::Scope prof(bbg_prof::MSM_SCATTER, st);
        BBG_LAUNCH_NOSYNC(msm_scatter_kernel, dim3((unsigned)((pl.max_entries + 255) / 256)), dim3(256), st, (const uint32_t*)digits, (const uint32_t*)ranks,
                          pl.num_points, pl.W, single.W, pl.NB, pl.entry_stride, (const uint32_t*)offsets, sorted);
    }
    const uint32_t* acc_offsets = offsets; // the regions the accumulate pass and the fix-up work on
    if (pair_rounds > 0)
    {
        bbg_prof::Scope prof(bbg_prof::MSM_PAIR, st);
        uint32_t* starts = (uint32_t*)(ws + o_pair_starts);
        uint32_t* pcounts = (uint32_t*)(ws + o_pair_counts);
        uint32_t* dense = (uint32_t*)(ws + o_pair_offsets);
        fe* buf[2] = { (fe*)(ws + o_pair_x), (fe*)(ws + o_pair_y) };
        // (`counts` has served the sort: it now takes what every bucket holds after the last round)
        BBG_LAUNCH_NOSYNC(msm_pair_plan_kernel, dim3((unsigned)((pair_row + 255) / 256)), dim3(256), st, (const uint32_t*)offsets, pl.total_buckets, pair_rounds,
                          starts, pcounts, counts);
        BBG_LAUNCH(scan_block_sums_kernel, dim3(scan_blocks), dim3(SCAN_BLOCK), 0, st, (const uint32_t*)counts, pl.total_buckets, spine);
        BBG_LAUNCH(scan_spine_kernel, dim3(1), dim3(SCAN_BLOCK), 0, st, spine, scan_blocks);
        BBG_LAUNCH(scan_apply_kernel, dim3(scan_blocks), dim3(SCAN_BLOCK), 0, st, (const uint32_t*)counts, pl.total_buckets, (const uint32_t*)spine, dense);
        for (int r = 1; r <= pair_rounds; ++r)
        {
            PairParams q;
            q.src = r == 1 ? (const fe*)d_table : buf[r & 1]; // round r writes buf[(r - 1) & 1]
            q.sorted = r == 1 ? sorted : nullptr;
            q.in_start = r == 1 ? offsets : starts + (size_t)(r - 1) * pair_row;
            q.in_count = r == 1 ? nullptr : pcounts + (size_t)(r - 1) * pair_row;
            q.out_start = r == pair_rounds ? dense : starts + (size_t)r * pair_row;
            q.dst = buf[(r - 1) & 1];
            q.total_buckets = pl.total_buckets;
            q.bmax = PAIR_BMAX;
            if (const char* e = getenv("BBG_MSM_PAIR_BMAX")) // development
            {
                const int v = atoi(e);
                if (v >= (int)PAIR_BMIN && v <= (int)PAIR_BMAX) q.bmax = (uint32_t)v;
            }
            q.prefix = (fe*)(ws + o_pair_prefix);
            q.rec = (uint32_t*)(ws + o_pair_rec);
            if (r == 1) BBG_LAUNCH(msm_pair_round_kernel<true>, dim3(pair_grid), dim3(PAIR_NT), 0, st, q);
            else BBG_LAUNCH(msm_pair_round_kernel<false>, dim3(pair_grid), dim3(PAIR_NT), 0, st, q);
        }
        acc_offsets = dense;
        g_msm_launches += 4 + pair_rounds;
    }
    {
        bbg_prof::Scope prof(bbg_prof::MSM_ACCUMULATE, st);
        if (pair_rounds > 0)
            BBG_LAUNCH_NOSYNC(msm_accumulate_kernel<true>, dim3((unsigned)((acc_slices + 127) / 128)), dim3(128), st, (const uint32_t*)nullptr, acc_offsets,
                              pl.total_buckets, (const fe*)(ws + (((pair_rounds - 1) & 1) ? o_pair_y : o_pair_x)), S_acc, buckets, head, tail);
        else
            BBG_LAUNCH_NOSYNC(msm_accumulate_kernel<false>, dim3((unsigned)((acc_slices + 127) / 128)), dim3(128), st, (const uint32_t*)sorted, acc_offsets,
                              pl.total_buckets, (const fe*)d_table, S_acc, buckets, head, tail);
    }
    {
        bbg_prof::Scope prof(bbg_prof::MSM_FIXUP, st);
        BBG_LAUNCH_NOSYNC(msm_fixup_kernel, dim3((pl.total_buckets + 127) / 128), dim3(128), st, acc_offsets, pl.total_buckets, S_acc, buckets,
                          (const fe*)head, (const fe*)tail, work_count, work_list, giant_list);
        BBG_LAUNCH(msm_fixup_large_kernel, dim3(FIXUP_LARGE_GRID), dim3(FIXUP_BLOCK), 0, st, acc_offsets, S_acc, buckets, (const fe*)head, (const fe*)tail,
                   (const uint32_t*)work_count, (const uint32_t*)work_list);
        BBG_LAUNCH(msm_fixup_giant_kernel, dim3(FIXUP_LARGE_GRID), dim3(FIXUP_BLOCK), 0, st, acc_offsets, S_acc, buckets, head, (const fe*)tail,
                   (const uint32_t*)work_count, (const uint32_t*)giant_list, 0);
        BBG_LAUNCH(msm_fixup_giant_kernel, dim3(FIXUP_LARGE_GRID), dim3(FIXUP_BLOCK), 0, st, acc_offsets, S_acc, buckets, head, (const fe*)tail,
                   (const uint32_t*)work_count, (const uint32_t*)giant_list, 1);
    }
    {
        bbg_prof::Scope prof(bbg_prof::MSM_CHUNK, st);
        BBG_LAUNCH_NOSYNC(msm_chunk_kernel, dim3((total_chunks + 127) / 128), dim3(128), st, (const fe*)buckets, total_chunks, pl.chunk_log, A, V);
    }
    {
        bbg_prof::Scope prof(bbg_prof::MSM_REDUCE, st);
        fe* red_part = (fe*)(ws + o_red_part);
        if (tree_blocks != 0)
        {
            BBG_LAUNCH(msm_tree_reduce_kernel, dim3(tree_blocks, (unsigned)pl.sets, 2), dim3(TREE_BLOCK), 0, st, (const fe*)A, (const fe*)V,
                       pl.chunks_per_window, red_part);
            BBG_LAUNCH(msm_tree_final_kernel, dim3(TREE_OUT, (unsigned)pl.sets), dim3(TREE_BLOCK), 0, st, (const fe*)red_part, tree_blocks,
                       pl.reduce_outputs - 2, red);
        }
        else
        {
            BBG_LAUNCH(msm_reduce_kernel, dim3((unsigned)pl.reduce_outputs, (unsigned)pl.sets, pl.red_splits), dim3(RED_BLOCK), 0, st, (const fe*)A,
                       (const fe*)V, pl.chunks_per_window, pl.reduce_outputs - 2, pl.chunks_per_window / pl.red_splits, pl.red_splits > 1 ? red_part : red);
            if (pl.red_splits > 1)
                BBG_LAUNCH(msm_reduce_final_kernel, dim3((unsigned)red_count), dim3(pl.red_splits < 64 ? pl.red_splits : 64), 0, st, (const fe*)red_part,
                           pl.red_splits, red);
        }
    }
    g_msm_launches += 11 + batch + ((tree_blocks != 0 || pl.red_splits > 1) ? 1 : 0);
    BBG_CHECK(bbg_rt::last_error());

    // the per-window reductions travel to the ticket's pinned slot behind the kernels
    BBG_CHECK(ticket_host_buffer(tk, red_count * 128));
    BBG_CHECK(bbg_rt::d2h(tk.host_red, red, red_count * 128, st));
#ifndef BBG_EMULATE
    if (tk.done == nullptr) BBG_CHECK(cudaEventCreateWithFlags(&tk.done, cudaEventDisableTiming));
    BBG_CHECK(cudaEventRecord(tk.done, st));
#endif
    tk.single = single;
    tk.pl = pl;
    tk.red_count = red_count;
    tk.pending = true;
    return 0;
}

// Build the pre-doubled windows of a table on the current device and remember them in `ctx` (fixed-base form).
int context_fixed_base_build(MsmContext& ctx, const void* d_table, size_t n_srs, int c, int W, cudaStream_t st)
{
    for (const FixedBase& f : ctx.fixed)
        if (f.base == (const char*)d_table) return 0;
    FixedBase f;
    f.base = (const char*)d_table;
    f.n_srs = n_srs;
    f.c = c;
    f.W = W;
    f.pre = nullptr;
    BBG_CHECK(bbg_rt::dev_alloc(&f.pre, (size_t)W * n_srs * 128));
    BBG_LAUNCH_NOSYNC(msm_precompute_kernel, dim3((unsigned)((n_srs + 127) / 128)), dim3(128), st, (const fe*)d_table, (fe*)f.pre, n_srs, c, W);
    g_msm_launches += 1;
    int e = bbg_rt::last_error();
    if (e == 0) e = bbg_rt::sync(st);
    if (e != 0)
    {
        bbg_rt::dev_free(f.pre);
        return e;
    }
    ctx.fixed.push_back(f);
    return 0;
}
void context_fixed_base_drop(MsmContext& ctx, const void* d_table)
{
    for (size_t i = 0; i < ctx.fixed.size(); ++i)
    {
        if (ctx.fixed[i].base != (const char*)d_table) continue;
        bbg_rt::dev_free(ctx.fixed[i].pre);
        ctx.fixed.erase(ctx.fixed.begin() + (long)i);
        return;
    }
}

// ---- 7. host finish: waits for the ticket's kernels, folds the windows of every MSM of the batch ---------------------
int context_finish(MsmContext& ctx, int ticket, hostg1::hxyzz* out)
{
    MsmTicket& tk = ctx.tickets[ticket];
    tk.pending = false;
    const size_t batch = tk.batch;
    if (tk.zero)
    {
        for (size_t b = 0; b < batch; ++b) out[b] = hostg1::infinity();
        return 0;
    }
#ifndef BBG_EMULATE
    BBG_CHECK(cudaEventSynchronize(tk.done));
#endif
    const Plan& single = tk.single;
    const Plan& pl = tk.pl;
    const hostg1::hxyzz* r_data = (const hostg1::hxyzz*)tk.host_red;
    const auto host_t0 = std::chrono::steady_clock::now();
    const int bits = pl.reduce_outputs - 2;
    for (size_t b = 0; b < batch; ++b)
    {
        hostg1::hxyzz result = hostg1::infinity();
        for (int w = single.sets - 1; w >= 0; --w) // (fixed-base form: one set, nothing to double)
        {
            const hostg1::hxyzz* rw = r_data + (b * (size_t)single.sets + (size_t)w) * pl.reduce_outputs;
            // sum_t t * A_t = sum_r 2^r T_r   (Horner from the top bit)
            hostg1::hxyzz tsum = hostg1::infinity();
            for (int r = bits - 1; r >= 0; --r)
            {
                tsum = hostg1::dbl(tsum);
                tsum = hostg1::add(tsum, rw[r]);
            }
            for (int i = 0; i < pl.chunk_log; ++i) tsum = hostg1::dbl(tsum); // * chunk size
            // bucket value = t * chunk + v + 1
            hostg1::hxyzz sw = hostg1::add(hostg1::add(tsum, rw[bits]), rw[bits + 1]);
            // result = 2^c * result + S_w   (reference :619-639, here with plain c-bit windows)
            if (w != single.sets - 1)
                for (int i = 0; i < pl.c; ++i) result = hostg1::dbl(result);
            result = hostg1::add(result, sw);
        }
        out[b] = result;
    }
    bbg_prof::add_host_ms(bbg_prof::MSM_HOST_FINISH, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count());
    return 0;
}

// ================================================================================================
// Multi-GPU: point ranges fanned out over the devices of one box
// ================================================================================================
// The reference splits an MSM into contiguous point ranges inside batched_scalar_multiplications, one per OpenMP thread
// (scalar_multiplication.cpp:703-728), and folds the partial sums on the calling thread (:750-765).  The same split here,
// one range per GPU: the primary device (the one every other entry point of the library runs on) keeps range 0, each other
// device has a worker thread with its own stream, MSM context and replica of the registered point tables; a worker pulls
// its slice of the scalars (peer copy over NVLink from the primary's HBM, or its own PCIe upload from the caller's host
// buffer), runs the same single-device pipeline and folds its windows; the launching thread adds the per-device sums.
// No collective: the only exchange is 128 bytes per device and MSM, handed over in host memory.
#ifndef BBG_EMULATE
constexpr int PEER_INFLIGHT = 4;
struct PeerJob
{
    int kind = 0; // 0: an MSM shard;  1: build the fixed-base windows of d_table (n_srs = count, c = batch, W = lo);  2: drop them
    const void* src[4] = {};
    size_t batch = 0;
    size_t lo = 0, count = 0; // scalar index range of this shard
    bool host_src = false;
    int src_device = 0;
    cudaEvent_t ready = nullptr; // scalars valid on the source device (device sources only)
    const void* d_table = nullptr; // this device's replica, already offset to the shard
    hostg1::hxyzz out[4];
    int err = 0;
    std::atomic<int> done{ 0 };
};

struct Replica
{
    const char* base0;
    size_t bytes;
    void* peer[MAX_DEVICES];
};

class PeerWorker
{
  public:
    PeerWorker(int device, int primary) : device_(device), primary_(primary), th_([this]() { run(); }) {}
    ~PeerWorker()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        th_.join();
    }
    int device() const { return device_; }
    void submit(PeerJob* j)
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            q_.push_back(j);
            ++outstanding_;
        }
        cv_.notify_one();
    }
    void quiesce()
    {
        std::unique_lock<std::mutex> lock(m_);
        idle_.wait(lock, [this]() { return outstanding_ == 0; });
    }
    int init_error()
    {
        std::unique_lock<std::mutex> lock(m_);
        idle_.wait(lock, [this]() { return started_; });
        return init_err_;
    }

  private:
    struct Inflight
    {
        PeerJob* job;
        int ticket;
    };
    int launch(PeerJob* j, unsigned seq, int* ticket)
    {
        const size_t bytes = j->count * 32;
        Workspace& buf = scalars_[seq % PEER_INFLIGHT];
        BBG_CHECK(buf.ensure(j->batch * bytes));
        const void* ptrs[4];
        if (!j->host_src) BBG_CHECK(cudaStreamWaitEvent(stream_, j->ready, 0));
        for (size_t b = 0; b < j->batch; ++b)
        {
            char* dst = (char*)buf.p + b * bytes;
            const char* src = (const char*)j->src[b] + j->lo * 32;
            if (j->host_src) BBG_CHECK(bbg_hostcopy::h2d_ring(ring_, dst, src, bytes, stream_));
            else BBG_CHECK(cudaMemcpyPeerAsync(dst, device_, src, j->src_device, bytes, stream_));
            ptrs[b] = dst;
        }
        int id = -1;
        for (int i = 0; i < MSM_TICKETS && id < 0; ++i)
            if (!ctx_.tickets[i].pending) id = i;
        if (id < 0) return 1007;
        BBG_CHECK(context_launch(ctx_, id, (int)(seq & 1), ptrs, j->batch, j->d_table, j->count, stream_));
        *ticket = id;
        return 0;
    }
    void complete(PeerJob* j, int err)
    {
        j->err = err;
        j->done.store(1, std::memory_order_release);
        std::lock_guard<std::mutex> lock(m_);
        if (--outstanding_ == 0) idle_.notify_all();
    }
    void run()
    {
        bbg_prof::thread_muted() = true;
        bbg_hostcopy::RegCache::lookup_only_thread() = true; // a worker never waits for the library to go idle: it IS its work
        int e = (int)cudaSetDevice(device_);
        if (e == 0) e = (int)cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking);
        if (e == 0)
        {
            // peer copies go straight over NVLink when the devices can address each other (staged by the driver otherwise)
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, device_, primary_) == cudaSuccess && can) cudaDeviceEnablePeerAccess(primary_, 0);
            cudaGetLastError();
        }
        {
            std::lock_guard<std::mutex> lock(m_);
            init_err_ = e;
            started_ = true;
        }
        idle_.notify_all();
        if (e != 0) return drain();
        std::vector<Inflight> inflight;
        unsigned seq = 0;
        for (;;)
        {
            PeerJob* j = nullptr;
            {
                std::unique_lock<std::mutex> lock(m_);
                if (inflight.empty()) cv_.wait(lock, [this]() { return stop_ || !q_.empty(); });
                if (!q_.empty() && inflight.size() < (size_t)PEER_INFLIGHT)
                {
                    j = q_.front();
                    q_.pop_front();
                }
                else if (stop_ && inflight.empty())
                    break;
            }
            if (j != nullptr && j->kind != 0)
            {
                // table maintenance: nothing of ours may be in flight on this device while tables come and go
                while (!inflight.empty())
                {
                    Inflight f = inflight.front();
                    inflight.erase(inflight.begin());
                    complete(f.job, context_finish(ctx_, f.ticket, f.job->out));
                }
                int err = 0;
                if (j->kind == 1) err = context_fixed_base_build(ctx_, j->d_table, j->count, (int)j->batch, (int)j->lo, stream_);
                else context_fixed_base_drop(ctx_, j->d_table);
                if (err != 0) cudaGetLastError();
                complete(j, err);
                continue;
            }
            if (j != nullptr)
            {
                int ticket = -1;
                const int err = launch(j, seq++, &ticket);
                if (err != 0)
                {
                    cudaGetLastError();
                    complete(j, err);
                }
                else
                    inflight.push_back({ j, ticket });
                // keep queueing while the launching thread keeps submitting (a prover queues its three wire
                // commitments back to back); finish the oldest once nothing is waiting
                std::lock_guard<std::mutex> lock(m_);
                if (!q_.empty() && inflight.size() < (size_t)PEER_INFLIGHT) continue;
            }
            if (!inflight.empty())
            {
                Inflight f = inflight.front();
                inflight.erase(inflight.begin());
                const int err = context_finish(ctx_, f.ticket, f.job->out);
                complete(f.job, err);
            }
        }
        cudaStreamSynchronize(stream_);
        context_release(ctx_);
        for (Workspace& w : scalars_) w.release();
        ring_.release();
        cudaStreamDestroy(stream_);
    }
    void drain() // initialisation failed: fail every job handed to us
    {
        for (;;)
        {
            PeerJob* j = nullptr;
            {
                std::unique_lock<std::mutex> lock(m_);
                cv_.wait(lock, [this]() { return stop_ || !q_.empty(); });
                if (q_.empty()) return;
                j = q_.front();
                q_.pop_front();
            }
            complete(j, init_err_);
        }
    }
    int device_, primary_;
    std::mutex m_;
    std::condition_variable cv_, idle_;
    std::deque<PeerJob*> q_;
    bool stop_ = false, started_ = false;
    int init_err_ = 0;
    int outstanding_ = 0;
    cudaStream_t stream_ = nullptr;
    MsmContext ctx_;
    Workspace scalars_[PEER_INFLIGHT];
    bbg_hostcopy::Ring ring_;
    std::thread th_; // last: started once everything above exists
};

struct MultiState
{
    int primary = 0;
    std::vector<std::unique_ptr<PeerWorker>> peers; // devices 1 .. G-1 of the instance
    std::vector<Replica> replicas;
    size_t min_points = (size_t)1 << 15;       // below this one device is faster than the hand-over
    size_t min_shard = (size_t)1 << 13;
    double primary_share = 1.0;                // the primary's range relative to an equal split (it also runs the NTTs)
    std::vector<PeerJob*> free_jobs;
} g_multi;

PeerJob* job_alloc()
{
    if (!g_multi.free_jobs.empty())
    {
        PeerJob* j = g_multi.free_jobs.back();
        g_multi.free_jobs.pop_back();
        j->done.store(0, std::memory_order_relaxed);
        j->err = 0;
        j->kind = 0;
        return j;
    }
    return new PeerJob();
}
void job_free(PeerJob* j) { g_multi.free_jobs.push_back(j); }

const Replica* find_replica(const void* d_table, size_t bytes)
{
    for (const Replica& r : g_multi.replicas)
        if ((const char*)d_table >= r.base0 && (const char*)d_table + bytes <= r.base0 + r.bytes) return &r;
    return nullptr;
}
#endif // !BBG_EMULATE
} // namespace

size_t msm_launch_count() { return g_msm_launches.load(); }

int msm_multi_device_count()
{
#ifndef BBG_EMULATE
    return 1 + (int)g_multi.peers.size();
#else
    return 1;
#endif
}

int msm_multi_quiesce()
{
#ifndef BBG_EMULATE
    for (auto& p : g_multi.peers) p->quiesce();
#endif
    return 0;
}

// devices[0] must be the device the library was initialised on (current on the calling thread)
int msm_multi_init(const int* devices, int count)
{
#ifndef BBG_EMULATE
    if (count < 1 || count > MAX_DEVICES || devices == nullptr) return 1007;
    if (!g_multi.peers.empty())
    {
        // idempotent for the same device list
        if ((int)g_multi.peers.size() + 1 != count || g_multi.primary != devices[0]) return 1007;
        for (int i = 1; i < count; ++i)
            if (g_multi.peers[(size_t)i - 1]->device() != devices[i]) return 1007;
        return 0;
    }
    g_multi.primary = devices[0];
    if (const char* e = getenv("BBG_MULTI_MIN_POINTS")) g_multi.min_points = (size_t)atol(e) > 0 ? (size_t)atol(e) : 1;
    if (const char* e = getenv("BBG_MULTI_MIN_SHARD")) g_multi.min_shard = (size_t)atol(e) > 0 ? (size_t)atol(e) : 1;
    if (const char* e = getenv("BBG_MULTI_PRIMARY_SHARE"))
    {
        const double v = atof(e);
        if (v > 0.0 && v <= 4.0) g_multi.primary_share = v;
    }
    for (int i = 1; i < count; ++i)
    {
        if (devices[i] == devices[0]) return 1007;
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, devices[0], devices[i]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(devices[i], 0);
        cudaGetLastError();
        g_multi.peers.emplace_back(new PeerWorker(devices[i], devices[0]));
    }
    for (auto& p : g_multi.peers)
    {
        const int e = p->init_error();
        if (e != 0)
        {
            g_multi.peers.clear();
            return e;
        }
    }
    return 0;
#else
    (void)devices;
    return count == 1 ? 0 : 1007;
#endif
}

// Every device of the instance gets its own copy of a registered point table (8 GB at 2^26 points: HBM is 180 GB), so any
// sub-range of it can be sharded; copied device to device from the primary's, behind `st`.
int msm_multi_replicate(const void* d_base0, size_t bytes, cudaStream_t st)
{
#ifndef BBG_EMULATE
    if (g_multi.peers.empty()) return 0;
    Replica r;
    r.base0 = (const char*)d_base0;
    r.bytes = bytes;
    for (void*& p : r.peer) p = nullptr;
    int e = 0;
    for (size_t i = 0; i < g_multi.peers.size() && e == 0; ++i)
    {
        const int dev = g_multi.peers[i]->device();
        e = (int)cudaSetDevice(dev);
        if (e == 0) e = (int)cudaMalloc(&r.peer[i], bytes ? bytes : 256);
        cudaSetDevice(g_multi.primary);
        if (e == 0) e = (int)cudaMemcpyPeerAsync(r.peer[i], dev, d_base0, g_multi.primary, bytes, st);
    }
    if (e == 0) e = bbg_rt::sync(st);
    if (e != 0)
    {
        for (size_t i = 0; i < g_multi.peers.size(); ++i)
        {
            if (r.peer[i] == nullptr) continue;
            cudaSetDevice(g_multi.peers[i]->device());
            cudaFree(r.peer[i]);
        }
        cudaSetDevice(g_multi.primary);
        cudaGetLastError();
        return e;
    }
    g_multi.replicas.push_back(r);
#else
    (void)d_base0;
    (void)bytes;
    (void)st;
#endif
    return 0;
}

int msm_multi_drop_replica(const void* d_base0)
{
#ifndef BBG_EMULATE
    for (size_t k = 0; k < g_multi.replicas.size(); ++k)
    {
        if (g_multi.replicas[k].base0 != (const char*)d_base0) continue;
        msm_multi_quiesce();
        for (size_t i = 0; i < g_multi.peers.size(); ++i)
        {
            if (g_multi.replicas[k].peer[i] == nullptr) continue;
            cudaSetDevice(g_multi.peers[i]->device());
            cudaFree(g_multi.replicas[k].peer[i]);
        }
        cudaSetDevice(g_multi.primary);
        g_multi.replicas.erase(g_multi.replicas.begin() + (long)k);
        return 0;
    }
#else
    (void)d_base0;
#endif
    return 0;
}

// Fixed-base tables for a registered point table: pre-doubled windows on every device of the instance (each device works
// on about n_srs / devices points per MSM, which sets the window width).  A no-op when the table would not fit the memory
// budget (W x the table; BBG_SRS_PRECOMPUTE_MAX_MB, default a quarter of the free device memory).
int msm_fixed_base_build(const void* d_table_primary, size_t n_srs, cudaStream_t st)
{
    if (n_srs < 1024) return 0;
    size_t devices = 1;
#ifndef BBG_EMULATE
    devices += g_multi.peers.size();
    if (n_srs < g_multi.min_points) devices = 1;
    while (devices > 1 && n_srs / devices < g_multi.min_shard) --devices;
#endif
    int c = 0, W = 0;
    pick_windows_fixed_base(n_srs / devices, c, W);
    if (const char* e = getenv("BBG_MSM_FIXED_WINDOW")) // development override
    {
        const int v = atoi(e);
        if (v >= 8 && v <= 22)
        {
            c = v;
            W = (128 + c - 1) / c;
            while ((W - 1) * c >= 127) --W;
        }
    }
    const size_t bytes = (size_t)W * n_srs * 128;
    if ((size_t)(W - 1) * 2 * n_srs + 2 * n_srs >= ((size_t)1 << 31)) return 0; // entry indices carry the sign in bit 31
    size_t budget = 0;
    if (const char* e = getenv("BBG_SRS_PRECOMPUTE_MAX_MB")) budget = (size_t)atol(e) << 20;
#ifndef BBG_EMULATE
    if (budget == 0)
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess) budget = free_b / 4;
    }
#else
    if (budget == 0) budget = (size_t)1 << 30;
#endif
    if (bytes > budget) return 0;
    BBG_CHECK(context_fixed_base_build(g_primary, d_table_primary, n_srs, c, W, st));
#ifndef BBG_EMULATE
    if (const Replica* rep = find_replica(d_table_primary, n_srs * 128))
    {
        if (rep->base0 == (const char*)d_table_primary)
        {
            std::vector<PeerJob*> jobs;
            for (size_t i = 0; i < g_multi.peers.size(); ++i)
            {
                PeerJob* j = job_alloc();
                j->kind = 1;
                j->d_table = rep->peer[i];
                j->count = n_srs;
                j->batch = (size_t)c;
                j->lo = (size_t)W;
                jobs.push_back(j);
                g_multi.peers[i]->submit(j);
            }
            int e = 0;
            for (PeerJob* j : jobs)
            {
                while (!j->done.load(std::memory_order_acquire)) std::this_thread::yield();
                if (e == 0) e = j->err;
                job_free(j);
            }
            if (e != 0)
            {
                // all or nothing: a device without the windows would compute the same sums the slow way, which is fine, but
                // an allocation failure is worth reporting
                return e;
            }
        }
    }
#endif
    return 0;
}

int msm_fixed_base_drop(const void* d_table_primary)
{
#ifndef BBG_EMULATE
    if (const Replica* rep = find_replica(d_table_primary, 1))
    {
        if (rep->base0 == (const char*)d_table_primary)
        {
            std::vector<PeerJob*> jobs;
            for (size_t i = 0; i < g_multi.peers.size(); ++i)
            {
                PeerJob* j = job_alloc();
                j->kind = 2;
                j->d_table = rep->peer[i];
                jobs.push_back(j);
                g_multi.peers[i]->submit(j);
            }
            for (PeerJob* j : jobs)
            {
                while (!j->done.load(std::memory_order_acquire)) std::this_thread::yield();
                job_free(j);
            }
        }
    }
#endif
    context_fixed_base_drop(g_primary, d_table_primary);
    return 0;
}

// (c, W) of the fixed-base windows behind a primary-device table pointer; 0 when there are none
int msm_fixed_base_info(const void* d_table_primary, int* c, int* W)
{
    for (const FixedBase& f : g_primary.fixed)
    {
        if ((const char*)d_table_primary >= f.base && (const char*)d_table_primary < f.base + f.n_srs * 128)
        {
            if (c) *c = f.c;
            if (W) *W = f.W;
            return 1;
        }
    }
    return 0;
}

int msm_release_workspace()
{
#ifndef BBG_EMULATE
    msm_multi_quiesce();
    for (size_t k = 0; k < g_multi.replicas.size(); ++k)
    {
        for (size_t i = 0; i < g_multi.peers.size(); ++i)
        {
            if (g_multi.replicas[k].peer[i] == nullptr) continue;
            cudaSetDevice(g_multi.peers[i]->device());
            cudaFree(g_multi.replicas[k].peer[i]);
        }
    }
    if (!g_multi.peers.empty()) cudaSetDevice(g_multi.primary);
    g_multi.replicas.clear();
    g_multi.peers.clear(); // joins the workers, which free their own device memory
    for (PeerJob* j : g_multi.free_jobs) delete j;
    g_multi.free_jobs.clear();
#endif
    context_release(g_primary);
    return 0;
}

bool msm_ticket_pending(int ticket) { return ticket >= 0 && ticket < MSM_TICKETS && g_primary.tickets[ticket].pending; }

// Queue `batch` same-size MSMs on the primary device's stream `st` (and, when the library drives several GPUs and the table
// is a replicated one, the other point ranges on the other devices); hands back a ticket for msm_finish.
//   d_scalars   device pointers on the primary device, or — host_scalars — the caller's host buffers: then each device
//               uploads its own range over its own PCIe link (the primary's copy goes through `staging`, batch * n * 32
//               bytes of primary-device memory owned by the caller)
int msm_launch_any(int workspace, const void* const* scalars, bool host_scalars, void* staging, size_t batch, const void* d_table, size_t n,
                   cudaStream_t st, int* ticket_out)
{
    if (workspace < 0 || workspace > 1 || ticket_out == nullptr || batch == 0 || batch > 4) return 1007;
    int id = -1;
    for (int i = 0; i < MSM_TICKETS && id < 0; ++i)
        if (!g_primary.tickets[i].pending) id = i;
    if (id < 0) return 1007; // too many MSMs in flight
    MsmTicket& tk = g_primary.tickets[id];
    tk.peer_count = 0;
    size_t n0 = n;
#ifndef BBG_EMULATE
    const Replica* rep = nullptr;
    size_t devices = 1;
    if (!g_multi.peers.empty() && n >= g_multi.min_points && (rep = find_replica(d_table, n * 128)) != nullptr)
    {
        devices = 1 + g_multi.peers.size();
        while (devices > 1 && n / devices < g_multi.min_shard) --devices;
    }
    if (devices > 1)
    {
        // contiguous ranges; the primary's is scaled by primary_share, the others share the rest equally
        double share0 = g_multi.primary_share / ((double)(devices - 1) + g_multi.primary_share);
        n0 = (size_t)((double)n * share0);
        if (n0 > n) n0 = n;
        const size_t rest = n - n0;
        if (!host_scalars)
        {
            if (tk.scalars_ready == nullptr) BBG_CHECK(cudaEventCreateWithFlags(&tk.scalars_ready, cudaEventDisableTiming));
            BBG_CHECK(cudaEventRecord(tk.scalars_ready, st));
        }
        const size_t table_off = (size_t)((const char*)d_table - rep->base0);
        for (size_t g = 1; g < devices; ++g)
        {
            const size_t lo = n0 + rest * (g - 1) / (devices - 1), hi = n0 + rest * g / (devices - 1);
            if (hi == lo) continue;
            PeerJob* j = job_alloc();
            for (size_t b = 0; b < batch; ++b) j->src[b] = scalars[b];
            j->batch = batch;
            j->lo = lo;
            j->count = hi - lo;
            j->host_src = host_scalars;
            j->src_device = g_multi.primary;
            j->ready = tk.scalars_ready;
            j->d_table = (const char*)rep->peer[g - 1] + table_off + lo * 128;
            tk.peer_jobs[tk.peer_count++] = j;
            g_multi.peers[g - 1]->submit(j);
        }
    }
#endif
    const void* dev_ptrs[4];
    for (size_t b = 0; b < batch; ++b)
    {
        if (host_scalars)
        {
            char* dst = (char*)staging + b * n0 * 32;
            if (n0 > 0) BBG_CHECK(bbg_hostcopy::h2d(dst, scalars[b], n0 * 32, st));
            dev_ptrs[b] = dst;
        }
        else
            dev_ptrs[b] = scalars[b];
    }
    const int e = context_launch(g_primary, id, workspace, dev_ptrs, batch, d_table, n0, st);
    if (e != 0)
    {
#ifndef BBG_EMULATE
        // the peers' shards are already queued: let them finish before reporting the failure
        for (int k = 0; k < tk.peer_count; ++k)
        {
            while (!tk.peer_jobs[k]->done.load(std::memory_order_acquire)) std::this_thread::yield();
            job_free(tk.peer_jobs[k]);
        }
        tk.peer_count = 0;
#endif
        return e;
    }
    *ticket_out = id;
    return 0;
}

int msm_launch(int workspace, const void* const* d_scalars, size_t batch, const void* d_table, size_t n, cudaStream_t st, int* ticket_out)
{
    return msm_launch_any(workspace, d_scalars, false, nullptr, batch, d_table, n, st, ticket_out);
}

// out_xyzz_host: HOST buffer of batch x 16 uint64 (X, Y, ZZ, ZZZ), un-normalised sums
int msm_finish(int ticket, void* out_xyzz_host)
{
    if (ticket < 0 || ticket >= MSM_TICKETS || !g_primary.tickets[ticket].pending) return 1007;
    MsmTicket& tk = g_primary.tickets[ticket];
    hostg1::hxyzz sums[4];
    int e = context_finish(g_primary, ticket, sums);
#ifndef BBG_EMULATE
    for (int k = 0; k < tk.peer_count; ++k)
    {
        PeerJob* j = tk.peer_jobs[k];
        while (!j->done.load(std::memory_order_acquire)) std::this_thread::yield();
        if (e == 0) e = j->err;
        if (e == 0)
            for (size_t b = 0; b < tk.batch; ++b) sums[b] = hostg1::add(sums[b], j->out[b]); // scalar_multiplication.cpp:750-765
        job_free(j);
    }
    tk.peer_count = 0;
#endif
    if (e != 0) return e;
    for (size_t b = 0; b < tk.batch; ++b) memcpy((char*)out_xyzz_host + b * sizeof(hostg1::hxyzz), &sums[b], sizeof(hostg1::hxyzz));
    return 0;
}

int msm_device_batched(const void* const* d_scalars, size_t batch, const void* d_table, size_t n, void* out_xyzz_host, cudaStream_t st)
{
    if (batch == 0) return 0;
    int ticket = -1;
    BBG_CHECK(msm_launch(0, d_scalars, batch, d_table, n, st, &ticket));
    return msm_finish(ticket, out_xyzz_host);
}

// d_out_xyzz: HOST buffer of 16 uint64 (X, Y, ZZ, ZZZ), un-normalised sum
int msm_device(const void* d_scalars, const void* d_table, size_t n, void* out_xyzz_host, cudaStream_t st)
{
    const void* one[1] = { d_scalars };
    return msm_device_batched(one, 1, d_table, n, out_xyzz_host, st);
}

// points[i] = (start + i * step) * G, i < n, affine with canonical coordinates.  One thread per run of GEN_RUN
// consecutive points: a 254-step double-and-add for the run's first point, mixed additions of step * G after
// that, and Montgomery's trick to leave XYZZ with one inversion per run.
int g1_generate_progression_device(const uint64_t* start_mont, const uint64_t* step_mont, void* d_points, size_t n, cudaStream_t st)
{
    if (n == 0) return 0;
    bbg_prof::Scope prof(bbg_prof::G1_GENERATE, st);
    const fe a0 = load_fe(start_mont), d = load_fe(step_mont);
    const size_t runs = (n + GEN_RUN - 1) / GEN_RUN;
    BBG_LAUNCH_NOSYNC(g1_progression_kernel, dim3((unsigned)((runs + 63) / 64)), dim3(64), st, a0, d, (fe*)d_points, n);
    g_msm_launches += 1;
    return bbg_rt::last_error();
}

int g1_table_from_transcript_device(const void* d_g1_bytes, void* d_table, size_t n, cudaStream_t st)
{
    if (n == 0) return 0;
    BBG_LAUNCH_NOSYNC(srs_from_transcript_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), st, (const uint8_t*)d_g1_bytes, (fe*)d_table, n);
    g_msm_launches += 1;
    return bbg_rt::last_error();
}

int g1_precompute_plain_device(const void* d_points, void* d_out, size_t n, int bits_per_window, int rounds, cudaStream_t st)
{
    if (n == 0 || rounds < 2) return 0;
    BBG_LAUNCH_NOSYNC(precompute_plain_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), st, (const fe*)d_points, (fe*)d_out, n, bits_per_window, rounds);
    g_msm_launches += 1;
    return bbg_rt::last_error();
}

int g1_build_endo_table_device(const void* d_points, void* d_table, size_t n, cudaStream_t st)
{
    if (n == 0) return 0;
    BBG_LAUNCH_NOSYNC(endo_table_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), st, (const fe*)d_points, (fe*)d_table, n);
    g_msm_launches += 1;
    return bbg_rt::last_error();
}
} // namespace bbg
