// TEST / BENCH INFRASTRUCTURE.  Writes a synthetic trusted-setup transcript in the on-disk format the reference's
// io::read_transcript expects (io/io.hpp:36-45, :76-98, :157-182), because srs_db/transcript.dat is
// absent from the reference checkout.  Secret x is fixed by a seed: monomials x^i * G1 for i = 1 .. N-1 and x * G2,
// so ReferenceString, the prover AND verify_proof's pairing check all work unchanged.
// Built against the reference's own headers; uses only the reference's group arithmetic.
//   usage: make_srs <num_points> <out_path>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <barretenberg/curves/bn254/fq.hpp>
#include <barretenberg/curves/bn254/fr.hpp>
#include <barretenberg/curves/bn254/g1.hpp>
#include <barretenberg/curves/bn254/g2.hpp>

using namespace barretenberg;

static void put_be32(FILE* f, uint32_t v)
{
    unsigned char b[4] = { (unsigned char)(v >> 24), (unsigned char)(v >> 16), (unsigned char)(v >> 8), (unsigned char)v };
    fwrite(b, 1, 4, f);
}
// one field element: raw (non-Montgomery) value, limbs in little-endian order, bytes big-endian inside each limb
static void put_fq(FILE* f, const fq::field_t& mont)
{
    fq::field_t raw;
    fq::__from_montgomery_form(mont, raw);
    for (int i = 0; i < 4; ++i)
    {
        uint64_t be = __builtin_bswap64(raw.data[i]);
        fwrite(&be, 8, 1, f);
    }
}

int main(int argc, char** argv)
{
    if (argc < 3)
    {
        fprintf(stderr, "usage: %s <num_points> <out_path>\n", argv[0]);
        return 2;
    }
    const size_t n = strtoull(argv[1], nullptr, 10);
    fr::field_t x = { { 0x9d2c5680a5a5a5a5ULL, 0x0123456789abcdefULL, 0xfedcba9876543210ULL, 0x0badc0de0badc0deULL & 0x0fffffffffffffffULL } };
    fr::__to_montgomery_form(x, x);

    // powers x^1 .. x^(n-1)
    std::vector<fr::field_t> powers(n > 1 ? n - 1 : 0);
    fr::field_t acc = x;
    for (size_t i = 0; i + 1 < n; ++i)
    {
        powers[i] = acc;
        fr::__mul(acc, x, acc);
    }
    std::vector<g1::affine_element> pts(powers.size());
#pragma omp parallel for schedule(dynamic, 256)
    for (size_t i = 0; i < powers.size(); ++i)
    {
        pts[i] = g1::group_exponentiation(g1::affine_one(), powers[i]);
    }
    g2::affine_element g2_one;
    g2::element g2_gen = g2::one();
    g2::jacobian_to_affine(g2_gen, g2_one);
    g2::affine_element g2_x = g2::group_exponentiation(g2_one, x);

    FILE* f = fopen(argv[2], "wb");
    if (!f)
    {
        perror("fopen");
        return 1;
    }
    const uint32_t num_g1 = (uint32_t)powers.size();
    put_be32(f, 0);      // transcript_number
    put_be32(f, 1);      // total_transcripts
    put_be32(f, num_g1); // total_g1_points
    put_be32(f, 2);      // total_g2_points
    put_be32(f, num_g1); // num_g1_points
    put_be32(f, 2);      // num_g2_points
    put_be32(f, 0);      // start_from
    for (size_t i = 0; i < pts.size(); ++i)
    {
        put_fq(f, pts[i].x);
        put_fq(f, pts[i].y);
    }
    const g2::affine_element g2s[2] = { g2_one, g2_x };
    for (int k = 0; k < 2; ++k)
    {
        put_fq(f, g2s[k].x.c0);
        put_fq(f, g2s[k].x.c1);
        put_fq(f, g2s[k].y.c0);
        put_fq(f, g2s[k].y.c1);
    }
    unsigned char checksum[64];
    memset(checksum, 0, sizeof checksum); // BLAKE2b checksum is not verified by the reader
    fwrite(checksum, 1, sizeof checksum, f);
    fclose(f);
    printf("wrote %s: %u G1 monomials + 2 G2 points\n", argv[2], num_g1);
    return 0;
}
