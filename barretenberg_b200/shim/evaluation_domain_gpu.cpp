// Drop-in replacement for evaluation_domain::compute_lookup_table of the reference's
//   src/barretenberg/polynomials/evaluation_domain.cpp:33-54, :172-178      
// Same class, same signature (evaluation_domain.hpp:35).  CircuitFFTState's constructor (circuit_state.hpp:11-19) calls it
// for the n, 2n and 4n domains: 14 n field elements built by serial product chains, 0.7 s of Prover(n) construction at
// n = 2^20 — and with the NTTs on the GPU only permutation.hpp:31 still reads them.  Here the tables are generated on the
// device (one product per entry) and copied into the same host layout: forward rounds in roots[0, size), inverse rounds in
// roots[size, 2 size), round i at offset 2^(i+1) - 2.  Entries are canonical Montgomery values (the reference's chain
// leaves them lazily reduced in [0, 2p)): the same field elements.
//
// How to link: build the reference's evaluation_domain.cpp with -Dcompute_lookup_table=cpu_reference_compute_lookup_table
// and add this file.
#include <cstdio>
#include <cstdlib>

#include <barretenberg/polynomials/evaluation_domain.hpp>

#include "bbgpu.h"
#include "shim_stats.h"

namespace barretenberg
{
void evaluation_domain::compute_lookup_table()
{
    int e = bbg_shim::ensure_library();
    bbg_shim::Timer timer("compute_lookup_table");
    roots = (fr::field_t*)(aligned_alloc(32, sizeof(fr::field_t) * size * 2));
    if (e == 0) e = bbg_fr_domain_lookup_table((uint64_t*)roots, (unsigned)log2_size);
    if (e != 0)
    {
        fprintf(stderr, "bbgpu shim: compute_lookup_table failed: %s (no CPU fallback)\n", bbg_error_string(e));
        abort();
    }
    // the per-round pointers, laid out as compute_lookup_table_single does (:35-41)
    const long num_rounds = (long)log2_size;
    fr::field_t* bases[2] = { &roots[0], &roots[size] };
    std::vector<fr::field_t*>* tables[2] = { &round_roots, &inverse_round_roots };
    for (int h = 0; h < 2; ++h)
    {
        tables[h]->emplace_back(bases[h]);
        for (long i = 1; i < num_rounds - 1; ++i) tables[h]->emplace_back(tables[h]->back() + (1UL << i));
    }
}
} // namespace barretenberg
