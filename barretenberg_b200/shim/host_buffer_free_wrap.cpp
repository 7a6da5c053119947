// Release hook for the host-buffer registration cache (include/bbgpu.h: bbg_set_host_register_cache).
//
// The shims page-lock long-lived caller buffers in place (barretenberg::polynomial coefficients, witness and selector
// polynomials, ReferenceString::monomials) the second time the library sees them, so that every later NTT / MSM call copies
// them at the pinned PCIe rate.  A page-locked range must be handed back to the library before its memory is returned to
// the allocator.  The reference frees through free() (types.hpp:25 `#define aligned_free free`, polynomial.cpp:70,
// reference_string.cpp:68) and through operator delete (std::vector), so the prover is linked with
//   -Wl,--wrap=free -Wl,--wrap=_ZdlPv -Wl,--wrap=_ZdlPvm -Wl,--wrap=_ZdaPv -Wl,--wrap=realloc
// and this file: every release made by the objects of that link first tells the library to forget the block.  The lookup
// is one atomic load while nothing is registered and a short table walk otherwise.
#include <cstddef>
#include <cstdlib>

#include "bbgpu.h"

extern "C" {
void __real_free(void* p);
void* __real_realloc(void* p, size_t bytes);
void __real__ZdlPv(void* p);
void __real__ZdlPvm(void* p, size_t bytes);
void __real__ZdaPv(void* p);

void __wrap_free(void* p)
{
    if (p != nullptr) bbg_host_buffer_forget(p);
    __real_free(p);
}
void* __wrap_realloc(void* p, size_t bytes)
{
    if (p != nullptr) bbg_host_buffer_forget(p);
    return __real_realloc(p, bytes);
}
void __wrap__ZdlPv(void* p)
{
    if (p != nullptr) bbg_host_buffer_forget(p);
    __real__ZdlPv(p);
}
void __wrap__ZdlPvm(void* p, size_t bytes)
{
    if (p != nullptr) bbg_host_buffer_forget(p);
    __real__ZdlPvm(p, bytes);
}
void __wrap__ZdaPv(void* p)
{
    if (p != nullptr) bbg_host_buffer_forget(p);
    __real__ZdaPv(p);
}
}
