// Forwarders to the reference's own Prover::construct_proof / Prover::reset bodies.  Build this file AND the
// reference's prover.cpp with  -Dconstruct_proof=cpu_reference_construct_proof -Dreset=cpu_reference_reset : inside
// these two translation units the member functions then carry the renamed symbols, which leaves the original names
// free for prover_gpu.cpp.  (The reference's round structure still sends every MSM / NTT to the GPU through the
// other shims; it is kept for widget mixes the resident path does not cover.)
#include <barretenberg/waffle/proof_system/prover/prover.hpp>

namespace bbg_shim
{
waffle::plonk_proof reference_construct_proof(waffle::Prover& prover) { return prover.construct_proof(); }
void reference_reset(waffle::Prover& prover) { prover.reset(); }
} // namespace bbg_shim
