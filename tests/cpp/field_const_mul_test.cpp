// Host check of Field::mul_const (portable path) against the Montgomery product: for random a < 4p and w < p,
// reduce(mul_const(a, w, wq)) must equal reduce(mul(a mod 2p..., to_mont(w))) as residues.  Exit code 0 = all equal.
#include <cstdint>
#include <cstdio>
#include <cstdlib>
struct uint4 { uint32_t x, y, z, w; }; // (the header's vector load helpers; CUDA type otherwise)
#include "../../barretenberg_b200/csrc/bbg_field.cuh"
using namespace bbg;
static uint64_t s = 0x9e3779b97f4a7c15ull;
static uint64_t next() { uint64_t z = (s += 0x9e3779b97f4a7c15ull); z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull; z = (z ^ (z >> 27)) * 0x94d049bb133111ebull; return z ^ (z >> 31); }
template <typename F> static fe rnd_below_p()
{
    fe r;
    for (int i = 0; i < 8; i += 2) { uint64_t v = next(); r.v[i] = (uint32_t)v; r.v[i + 1] = (uint32_t)(v >> 32); }
    r.v[7] &= 0x3fffffffu; // < 2^254 < 2p
    return F::reduce(r);   // [0, p) for inputs below 2p
}
template <typename F> static int run(const char* name, int count)
{
    int bad = 0;
    for (int it = 0; it < count; ++it)
    {
        fe w = rnd_below_p<F>();
        if (it == 0) w = F::zero();
        if (it == 1) { w = F::modulus(); w.v[0] -= 1; }
        fe a = rnd_below_p<F>();
        // spread a over [0, 4p): add 0..3 times p
        const int k = it & 3;
        fe pm = F::modulus();
        for (int t = 0; t < k; ++t) { fe n; cc::add8(n.v, a.v, pm.v); a = n; }
        if (it == 2) { a = F::modulus(); fe n; cc::add8(n.v, a.v, a.v); cc::add8(a.v, n.v, n.v); a.v[0] -= 1; } // 4p - 1
        const fe w_mont = F::to_mont(w);
        const fe wq = F::const_quotient(w_mont);
        const fe got = F::mul_const(a, w, wq);
        // range check: got < 2p
        fe p2 = F::modulus(); { fe n; cc::add8(n.v, p2.v, p2.v); p2 = n; }
        fe d; const uint32_t borrow = cc::sub8(d.v, got.v, p2.v);
        if (!borrow) { ++bad; if (bad < 5) printf("%s: result >= 2p at %d\n", name, it); continue; }
        // expected: a * w as Montgomery product with w_mont (inputs of mul must be < 2p: reduce a first, twice)
        fe ar = a;
        for (int t = 0; t < 3; ++t) ar = F::reduce(ar);
        const fe want = F::reduce(F::mul(ar, w_mont));
        if (!F::eq_raw(F::reduce(got), want)) { ++bad; if (bad < 5) printf("%s: mismatch at %d\n", name, it); }
    }
    printf("%s: %d cases, %d bad\n", name, count, bad);
    return bad;
}
// sqr(a) must be the same integer as mul(a, a) (not just the same residue) for every a < 2p
template <typename F> static int run_sqr(const char* name, int count)
{
    int bad = 0;
    fe p2 = F::modulus(); { fe n; cc::add8(n.v, p2.v, p2.v); p2 = n; }
    for (int it = 0; it < count; ++it)
    {
        fe a = rnd_below_p<F>();
        if (it & 1) { fe n; fe pm = F::modulus(); cc::add8(n.v, a.v, pm.v); a = n; } // [p, 2p)
        if (it == 0) a = F::zero();
        if (it == 1) { a = p2; a.v[0] -= 1; }                         // 2p - 1
        if (it == 2) { a = F::modulus(); }
        if (it == 3) { for (int l = 0; l < 8; ++l) a.v[l] = 0xffffffffu; a.v[7] = F::params::P2(7) - 1; } // all-ones words below 2p
        if (it == 4) { for (int l = 0; l < 8; ++l) a.v[l] = 0x80000000u; a.v[7] = 0x40000000u; }        // every shifted-out bit set
        if (it == 5) { for (int l = 0; l < 8; ++l) a.v[l] = 0xffffffffu; a.v[7] = 0x3fffffffu; }
        if (it == 6) { a = F::zero(); a.v[0] = 0xffffffffu; }
        if (it == 7) { a = F::zero(); a.v[7] = F::params::P2(7) - 1; }
        const fe want = F::mul(a, a), got = F::sqr(a);
        if (!F::eq_raw(want, got)) { ++bad; if (bad < 5) printf("%s sqr: mismatch at %d\n", name, it); }
    }
    printf("%s sqr: %d cases, %d bad\n", name, count, bad);
    return bad;
}
int main(int argc, char** argv)
{
    const int count = argc > 1 ? atoi(argv[1]) : 200000;
    int bad = run<Fr>("Fr", count) + run<Fq>("Fq", count) + run_sqr<Fr>("Fr", count) + run_sqr<Fq>("Fq", count);
    return bad ? 1 : 0;
}
