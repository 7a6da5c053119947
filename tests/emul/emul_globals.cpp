// TEST INFRASTRUCTURE — storage for the CPU kernel-emulation thread-locals (see cuda_emul.h).
#include "cuda_emul.h"
namespace emul
{
thread_local dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
thread_local pthread_barrier_t* t_barrier = nullptr;
thread_local unsigned char* t_dyn_smem = nullptr;
} // namespace emul
