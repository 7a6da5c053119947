// TEST INFRASTRUCTURE — host rendition of the device field / curve code (see cuda_emul.h).
#include "cuda_emul.h"
#include "../../barretenberg_b200/csrc/bbg_g1.cuh"

namespace emul
{
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local pthread_barrier_t* t_barrier = nullptr;
thread_local unsigned char* t_dyn_smem = nullptr;
} // namespace emul

using namespace bbg;
template <typename F> static void binop_n(int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n)
{
    for (size_t i = 0; i < n; ++i)
    {
        fe x = load_fe(a + 4 * i), y = load_fe(b + 4 * i), z;
        switch (op)
        {
        case 0: z = F::mul(x, y); break;
        case 1: z = F::mul_full(x, y); break;
        case 2: z = F::add(x, y); break;
        case 3: z = F::sub(x, y); break;
        case 4: z = F::reduce(x); break;
        case 5: z = F::invert(x); break;
        case 6: z = F::to_mont(x); break;
        case 7: z = F::from_mont(x); break;
        default: z = F::zero();
        }
        store_fe(r + 4 * i, z);
    }
}
extern "C" {
void emu_field_op_n(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* r, size_t n)
{
    if (field == 0) binop_n<Fq>(op, a, b, r, n); else binop_n<Fr>(op, a, b, r, n);
}
// acc (xyzz 16 limbs u64) op affine/xyzz -> xyzz
void emu_g1_madd(const uint64_t* acc, const uint64_t* q, uint64_t* out) { store_xyzz(out, G1::madd(load_xyzz(acc), load_affine(q))); }
void emu_g1_add(const uint64_t* a, const uint64_t* b, uint64_t* out) { store_xyzz(out, G1::add(load_xyzz(a), load_xyzz(b))); }
void emu_g1_dbl(const uint64_t* a, uint64_t* out) { store_xyzz(out, G1::dbl(load_xyzz(a))); }
void emu_g1_from_affine(const uint64_t* a, uint64_t* out) { store_xyzz(out, G1::from_affine(load_affine(a))); }
void emu_g1_to_affine(const uint64_t* a, uint64_t* out) { store_affine(out, G1::to_affine(load_xyzz(a))); }
void emu_g1_endo(const uint64_t* a, uint64_t* out) { store_affine(out, G1::endo_table_entry(load_affine(a))); }
}
