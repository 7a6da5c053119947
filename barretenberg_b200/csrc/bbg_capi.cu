// extern "C" surface of libbbgpu.so — see include/bbgpu.h for the contract and the reference
// signatures each entry point stands behind.  No CPU fallback: every compute entry point needs the
// CUDA device selected by bbg_init().
#include "../../include/bbgpu.h"
#include "bbg_internal.h"
#include "bbg_host_g1.h"
#include "bbg_hostcopy.h"
#include "bbg_plonk.h"

#include <mutex>
#include <vector>

namespace bbg
{
int microbench_launch(int mode, int iters, uint32_t* d_out, int blocks, double* ops, cudaStream_t st);
int field_selftest_device(int field, int op, const void* d_a, const void* d_b, void* d_out, size_t count, cudaStream_t st);
int g1_selftest_device(int op, const void* d_p, const void* d_q, void* d_out, size_t count, cudaStream_t st);
size_t selftest_launch_count();
}

namespace
{
using namespace bbg;

std::mutex g_mutex;
bool g_ready = false;
cudaStream_t g_stream = nullptr;
bool g_own_stream = false;
uint64_t g_misc_launches = 0;

struct GrowBuf
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
GrowBuf g_stage_coeffs;  // NTT host path: polynomial staging
GrowBuf g_stage_scalars; // MSM host path: scalar staging
GrowBuf g_stage_table;   // MSM host path: unregistered point tables
// bbg_msm_g1_launch: each MSM queued on the second stream owns its scalar staging until its ticket is finished (in a
// multi-GPU instance the other devices pull their ranges out of it on their own streams)
struct AsyncSlot
{
    GrowBuf scalars;
    int ticket = -1;
} g_async_slots[6];

struct SrsEntry
{
    const uint64_t* host_base;
    size_t n; // original points; table has 2n entries of 8 uint64
    void* d_table;
    uint64_t fingerprint; // of sampled host entries, re-checked on every cache hit
    bool automatic;       // created by the auto cache (may be evicted)
    int pins;             // resident provers holding d_table (bbg_plonk_set_srs): never evicted while > 0
};
std::vector<std::pair<const void*, const uint64_t*>> g_prover_srs; // prover -> host_base of the entry it pins
std::vector<SrsEntry> g_srs;
// point sets handed over WITHOUT their endomorphism entries (pippenger_low_memory, pippenger_precomputed): the 2n-entry
// table is built on the device and kept under the caller's address like an auto-cached SRS (n x 64 host bytes behind it)
struct PlainEntry
{
    const uint64_t* host_points;
    size_t n;
    void* d_table;
    uint64_t fingerprint;
};
std::vector<PlainEntry> g_plain;
uint64_t plain_fingerprint(const uint64_t* pts, size_t n)
{
    uint64_t h = 1469598103934665603ULL ^ (uint64_t)n;
    for (size_t s = 0; s < 32; ++s)
    {
        const size_t idx = n <= 32 ? (s < n ? s : n - 1) : (s * (n - 1)) / 31;
        for (int w = 0; w < 8; ++w)
        {
            h ^= pts[8 * idx + w];
            h *= 1099511628211ULL;
        }
    }
    return h;
}
bool g_auto_srs = false;
bool g_srs_precompute = false; // build fixed-base windows for every registered / cached table (bbg_set_srs_precompute)
constexpr size_t AUTO_SRS_MIN_POINTS = 1024; // below this an upload per call is cheaper than bookkeeping
constexpr size_t AUTO_SRS_MAX_ENTRIES = 4;

// FNV-1a over 32 table entries spread across [0, 2n): cheap enough to recompute on every call, and enough to
// notice that a host buffer was freed / rewritten behind a cached address
uint64_t table_fingerprint(const uint64_t* table, size_t n)
{
    const size_t entries = 2 * n;
    uint64_t h = 1469598103934665603ULL ^ (uint64_t)n;
    for (size_t s = 0; s < 32; ++s)
    {
        const size_t idx = entries <= 32 ? (s < entries ? s : entries - 1) : (s * (entries - 1)) / 31;
        const uint64_t* e = table + 8 * idx;
        for (int w = 0; w < 8; ++w)
        {
            h ^= e[w];
            h *= 1099511628211ULL;
        }
    }
    return h;
}

#ifndef BBG_EMULATE
cudaEvent_t g_ev_start = nullptr, g_ev_stop = nullptr;
// copy engines for the host-buffer NTT path: uploads, kernels and downloads of different polynomials overlap
cudaStream_t g_copy_in = nullptr, g_copy_out = nullptr;
cudaStream_t g_msm_stream = nullptr; // bbg_msm_g1_partial_dev_launch: an MSM running beside the work stream
cudaEvent_t g_msm_fence = nullptr;
bbg_hostcopy::Ring g_msm_ring; // pageable scalars of bbg_msm_g1_launch: their chunks wait behind MSM kernels, not behind the work stream
std::vector<cudaEvent_t> g_pipe_events;
#endif

int ensure_ready()
{
    if (g_ready) return 0;
#ifdef BBG_EMULATE
    g_ready = true;
    return 0;
#else
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return BBG_E_NO_DEVICE;
    return BBG_E_NOT_INITIALISED;
#endif
}

// nothing queued by this library is still running on any device (before device memory is handed back)
void quiesce_all()
{
    bbg_rt::sync(g_stream);
#ifndef BBG_EMULATE
    if (g_msm_stream) cudaStreamSynchronize(g_msm_stream);
#endif
    msm_multi_quiesce();
}

void srs_drop(size_t i)
{
    quiesce_all();
    msm_fixed_base_drop(g_srs[i].d_table);
    msm_multi_drop_replica(g_srs[i].d_table);
    bbg_rt::dev_free(g_srs[i].d_table);
    for (size_t k = 0; k < g_prover_srs.size();)
    {
        if (g_prover_srs[k].second == g_srs[i].host_base) g_prover_srs.erase(g_prover_srs.begin() + (long)k);
        else ++k;
    }
    g_srs.erase(g_srs.begin() + (long)i);
}

// d_table: filled (or being filled on the work stream) by the caller; ownership moves to the cache
int srs_add(const uint64_t* host_base, size_t n, void* d_table, bool automatic)
{
    SrsEntry s;
    s.host_base = host_base;
    s.n = n;
    s.d_table = d_table;
    s.automatic = automatic;
    s.pins = 0;
    s.fingerprint = table_fingerprint(host_base, n);
    int e = msm_multi_replicate(d_table, n * 128, g_stream); // no-op with one device
    if (e == 0 && g_srs_precompute) e = msm_fixed_base_build(d_table, n, g_stream);
    if (e != 0)
    {
        bbg_rt::sync(g_stream);
        msm_fixed_base_drop(d_table);
        msm_multi_drop_replica(d_table);
        bbg_rt::dev_free(d_table);
        return e;
    }
    g_srs.push_back(s);
    return 0;
}

// device pointer for a host point-table pointer: registered SRS (sub-range allowed) or a fresh upload.
// keep: the caller will use the pointer beyond this call (a launched MSM, a resident prover), so an unknown table is
// cached whatever its size and whether or not the auto cache is on — never the shared staging buffer.
int resolve_table(const uint64_t* points, size_t n, const void** d_table, bool keep = false)
{
    for (size_t i = 0; i < g_srs.size(); ++i)
    {
        SrsEntry& s = g_srs[i];
        if (points >= s.host_base && points + 16 * n <= s.host_base + 16 * s.n)
        {
            if (table_fingerprint(s.host_base, s.n) != s.fingerprint)
            {
                // the host buffer changed under a cached address: drop the stale copy and fall through
                srs_drop(i);
                break;
            }
            *d_table = (const char*)s.d_table + ((const char*)points - (const char*)s.host_base);
            return 0;
        }
    }
    if (keep || (g_auto_srs && n >= AUTO_SRS_MIN_POINTS))
    {
        // first sight of a large table: keep it on the device (prover.cpp calls the MSM 9 times per proof on the same
        // ReferenceString::monomials buffer; copies of a ReferenceString get their own entry)
        size_t autos = 0;
        for (const SrsEntry& s : g_srs) autos += s.automatic ? 1 : 0;
        if (autos >= AUTO_SRS_MAX_ENTRIES)
        {
            for (size_t i = 0; i < g_srs.size(); ++i)
            {
                if (g_srs[i].automatic && g_srs[i].pins == 0)
                {
                    srs_drop(i);
                    break;
                }
            }
        }
        void* d = nullptr;
        BBG_CHECK(bbg_rt::dev_alloc(&d, n * 128));
        const int e = bbg_hostcopy::h2d(d, points, n * 128, g_stream);
        if (e != 0)
        {
            bbg_rt::dev_free(d);
            return e;
        }
        BBG_CHECK(srs_add(points, n, d, true));
        *d_table = d;
        return 0;
    }
    BBG_CHECK(g_stage_table.ensure(n * 128));
    BBG_CHECK(bbg_hostcopy::h2d(g_stage_table.p, points, n * 128, g_stream));
    *d_table = g_stage_table.p;
    return 0;
}

// `batch` same-size MSMs of host scalars over one host table; each device of the instance uploads its own point range
int msm_host_batched(const uint64_t* const* scalars, size_t batch, const uint64_t* points, size_t n, hostg1::hxyzz* out)
{
    if (n == 0)
    {
        for (size_t b = 0; b < batch; ++b) out[b] = hostg1::infinity();
        return 0;
    }
    if (points == nullptr) return BBG_E_BAD_ARGUMENT;
    for (size_t b = 0; b < batch; ++b)
        if (scalars[b] == nullptr) return BBG_E_BAD_ARGUMENT;
    const void* d_table = nullptr;
    BBG_CHECK(resolve_table(points, n, &d_table));
    BBG_CHECK(g_stage_scalars.ensure(batch * n * 32));
#ifndef BBG_EMULATE
    // the other devices' workers copy straight out of the caller's buffers: page-lock them here, once, as a whole
    if (msm_multi_device_count() > 1)
        for (size_t b = 0; b < batch; ++b) (void)bbg_hostcopy::reg_cache().classify(scalars[b], n * 32);
#endif
    int ticket = -1;
    BBG_CHECK(msm_launch_any(0, (const void* const*)scalars, true, g_stage_scalars.p, batch, d_table, n, g_stream, &ticket));
    return msm_finish(ticket, out);
}

int msm_host(const uint64_t* scalars, const uint64_t* points, size_t n, hostg1::hxyzz* out)
{
    const uint64_t* one[1] = { scalars };
    if (n > 0 && scalars == nullptr) return BBG_E_BAD_ARGUMENT;
    return msm_host_batched(one, 1, points, n, out);
}
} // namespace

extern "C" {

int bbg_init(int device)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (g_ready) return 0;
#ifndef BBG_EMULATE
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return BBG_E_NO_DEVICE;
    if (device < 0 || device >= count) return BBG_E_BAD_ARGUMENT;
    BBG_CHECK(cudaSetDevice(device));
    if (g_stream == nullptr)
    {
        BBG_CHECK(cudaStreamCreateWithFlags(&g_stream, cudaStreamNonBlocking));
        g_own_stream = true;
    }
    BBG_CHECK(cudaEventCreate(&g_ev_start));
    BBG_CHECK(cudaEventCreate(&g_ev_stop));
    BBG_CHECK(cudaStreamCreateWithFlags(&g_copy_in, cudaStreamNonBlocking));
    BBG_CHECK(cudaStreamCreateWithFlags(&g_copy_out, cudaStreamNonBlocking));
    BBG_CHECK(cudaStreamCreateWithFlags(&g_msm_stream, cudaStreamNonBlocking));
    BBG_CHECK(cudaEventCreateWithFlags(&g_msm_fence, cudaEventDisableTiming));
    bbg_hostcopy::reg_cache().set_quiesce([]() { quiesce_all(); });
#else
    (void)device;
#endif
    g_ready = true;
    return 0;
}

int bbg_init_multi(const int* devices, int count)
{
    if (devices == nullptr || count < 1) return BBG_E_BAD_ARGUMENT;
    BBG_CHECK(bbg_init(devices[0]));
    std::lock_guard<std::mutex> lock(g_mutex);
#ifndef BBG_EMULATE
    int current = -1, visible = 0;
    BBG_CHECK(cudaGetDevice(&current));
    BBG_CHECK(cudaGetDeviceCount(&visible));
    if (current != devices[0]) return BBG_E_BAD_ARGUMENT; // already initialised on another device
    for (int i = 0; i < count; ++i)
        if (devices[i] < 0 || devices[i] >= visible) return BBG_E_BAD_ARGUMENT;
#endif
    if (count == 1) return 0;
    BBG_CHECK(msm_multi_init(devices, count));
    // tables registered before the other devices joined
    for (const SrsEntry& s : g_srs)
    {
        msm_fixed_base_drop(s.d_table); // (its window width depends on the number of devices)
        BBG_CHECK(msm_multi_replicate(s.d_table, s.n * 128, g_stream));
        if (g_srs_precompute) BBG_CHECK(msm_fixed_base_build(s.d_table, s.n, g_stream));
    }
    return 0;
}

int bbg_device_count(void) { return msm_multi_device_count(); }

int bbg_shutdown(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (!g_ready) return 0;
    quiesce_all();
    ntt_release_tables();
    msm_release_workspace(); // also stops the other devices' workers and frees their table replicas
    plonk::release_helpers();
    g_stage_coeffs.release();
    g_stage_scalars.release();
    for (AsyncSlot& a : g_async_slots)
    {
        a.scalars.release();
        a.ticket = -1;
    }
    g_stage_table.release();
    for (SrsEntry& s : g_srs) bbg_rt::dev_free(s.d_table);
    g_srs.clear();
    for (PlainEntry& e : g_plain) bbg_rt::dev_free(e.d_table);
    g_plain.clear();
    g_prover_srs.clear();
    bbg_hostcopy::reg_cache().release();
#ifndef BBG_EMULATE
    if (g_ev_start) cudaEventDestroy(g_ev_start);
    if (g_ev_stop) cudaEventDestroy(g_ev_stop);
    g_ev_start = g_ev_stop = nullptr;
    for (cudaEvent_t ev : g_pipe_events) cudaEventDestroy(ev);
    g_pipe_events.clear();
    bbg_hostcopy::ring().release();
    g_msm_ring.release();
    if (g_copy_in) cudaStreamDestroy(g_copy_in);
    if (g_copy_out) cudaStreamDestroy(g_copy_out);
    g_copy_in = g_copy_out = nullptr;
    if (g_msm_stream)
    {
        cudaStreamSynchronize(g_msm_stream);
        cudaStreamDestroy(g_msm_stream);
    }
    if (g_msm_fence) cudaEventDestroy(g_msm_fence);
    g_msm_stream = nullptr;
    g_msm_fence = nullptr;
    if (g_own_stream && g_stream) cudaStreamDestroy(g_stream);
    g_stream = nullptr;
    g_own_stream = false;
#endif
    g_ready = false;
    return 0;
}

int bbg_set_stream(void* cuda_stream)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    bbg_rt::sync(g_stream);
#ifndef BBG_EMULATE
    if (g_own_stream && g_stream) cudaStreamDestroy(g_stream);
#endif
    g_stream = (cudaStream_t)cuda_stream;
    g_own_stream = false;
    return 0;
}

const char* bbg_error_string(int code)
{
    switch (code)
    {
    case 0: return "success";
    case BBG_E_BAD_SIZE: return "bbgpu: NTT size must be 2^1 .. 2^28";
    case BBG_E_BAD_OP: return "bbgpu: unknown NTT operation";
    case BBG_E_NULL_CONSTANT: return "bbgpu: this NTT operation needs a constant";
    case BBG_E_NOT_INITIALISED: return "bbgpu: bbg_init() has not been called";
    case BBG_E_NO_DEVICE: return "bbgpu: no CUDA device (there is no CPU fallback)";
    case BBG_E_BAD_ARGUMENT: return "bbgpu: bad argument";
    case BBG_E_TOO_LARGE: return "bbgpu: MSM too large";
    case 1001: return "bbgpu: internal: unsupported sub-transform size";
    case 1009: return "bbgpu: internal: MSM window planner produced an invalid plan";
    }
#ifndef BBG_EMULATE
    if (code > 0 && code < 1000) return cudaGetErrorString((cudaError_t)code);
#endif
    return "bbgpu: unknown error";
}

uint64_t bbg_launch_count(void)
{
    return (uint64_t)ntt_launch_count() + (uint64_t)msm_launch_count() + (uint64_t)plonk::launch_count() + (uint64_t)selftest_launch_count() + g_misc_launches;
}

// ---- NTT ----------------------------------------------------------------------------------------
int bbg_ntt_fr_dev(void* d_coeffs, size_t stride_elems, size_t batch, unsigned log2_n, int op, const uint64_t* constant)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (d_coeffs == nullptr && batch > 0) return BBG_E_BAD_ARGUMENT;
    return ntt_device(d_coeffs, stride_elems, batch, log2_n, op, constant, g_stream);
}

int bbg_ntt_fr_batched(uint64_t* const* coeffs, size_t batch, unsigned log2_n, int op, const uint64_t* constant)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (batch == 0) return 0;
    if (coeffs == nullptr) return BBG_E_BAD_ARGUMENT;
    if (log2_n < 1 || log2_n > 28) return BBG_E_BAD_SIZE;
    const size_t n = (size_t)1 << log2_n, bytes = n * 32;
    BBG_CHECK(g_stage_coeffs.ensure(batch * bytes));
    for (size_t i = 0; i < batch; ++i)
    {
        if (coeffs[i] == nullptr) return BBG_E_BAD_ARGUMENT;
    }
#ifndef BBG_EMULATE
    bool all_pinned = batch > 1;
    for (size_t i = 0; i < batch && all_pinned; ++i) all_pinned = bbg_hostcopy::caller_pinned_whole(coeffs[i], bytes);
    if (all_pinned)
    {
        // three-stage pipeline over polynomials: upload i+1 | transform i | download i-1 (PCIe is full duplex)
        while (g_pipe_events.size() < 2 * batch + 1)
        {
            cudaEvent_t ev;
            BBG_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            g_pipe_events.push_back(ev);
        }
        // uploads may start once everything already queued on the work stream is done with the staging buffer
        BBG_CHECK(cudaEventRecord(g_pipe_events[2 * batch], g_stream));
        BBG_CHECK(cudaStreamWaitEvent(g_copy_in, g_pipe_events[2 * batch], 0));
        for (size_t i = 0; i < batch; ++i)
        {
            BBG_CHECK(bbg_rt::h2d((char*)g_stage_coeffs.p + i * bytes, coeffs[i], bytes, g_copy_in));
            BBG_CHECK(cudaEventRecord(g_pipe_events[i], g_copy_in));
        }
        for (size_t i = 0; i < batch; ++i)
        {
            BBG_CHECK(cudaStreamWaitEvent(g_stream, g_pipe_events[i], 0));
            BBG_CHECK(ntt_device((char*)g_stage_coeffs.p + i * bytes, n, 1, log2_n, op, constant, g_stream));
            BBG_CHECK(cudaEventRecord(g_pipe_events[batch + i], g_stream));
            BBG_CHECK(cudaStreamWaitEvent(g_copy_out, g_pipe_events[batch + i], 0));
            BBG_CHECK(bbg_rt::d2h(coeffs[i], (char*)g_stage_coeffs.p + i * bytes, bytes, g_copy_out));
        }
        BBG_CHECK(bbg_rt::sync(g_copy_out));
        return bbg_rt::sync(g_stream);
    }
#endif
#ifndef BBG_EMULATE
    // one polynomial in a page-locked buffer (the caller's own, or one the registration cache locked in place): uploads,
    // passes and downloads overlapped block by block (bbg_ntt.cu, ntt_host_blocks)
    static const bool host_blocks = [] { const char* e = getenv("BBG_NTT_HOST_BLOCKS"); return e == nullptr || e[0] != '0'; }();
    if (batch == 1 && host_blocks && ntt_host_blocks_applicable(log2_n))
    {
        const bbg_hostcopy::PinnedSpan sp = bbg_hostcopy::reg_cache().peek(coeffs[0], bytes);
        if (sp.whole || (sp.mid > 0 && !sp.mixed))
            return ntt_host_blocks(coeffs[0], sp.whole ? 0 : sp.head, sp.whole ? bytes : sp.head + sp.mid, g_stage_coeffs.p, log2_n, op, constant, g_stream,
                                   g_copy_in, g_copy_out);
    }
#endif
    for (size_t i = 0; i < batch; ++i) BBG_CHECK(bbg_hostcopy::h2d((char*)g_stage_coeffs.p + i * bytes, coeffs[i], bytes, g_stream));
    BBG_CHECK(ntt_device(g_stage_coeffs.p, n, batch, log2_n, op, constant, g_stream));
    for (size_t i = 0; i < batch; ++i) BBG_CHECK(bbg_hostcopy::d2h(coeffs[i], (char*)g_stage_coeffs.p + i * bytes, bytes, g_stream));
    return bbg_rt::sync(g_stream);
}

int bbg_ntt_fr(uint64_t* coeffs, unsigned log2_n, int op, const uint64_t* constant)
{
    uint64_t* one[1] = { coeffs };
    return bbg_ntt_fr_batched(one, 1, log2_n, op, constant);
}

int bbg_compute_lagrange_polynomial_fft(uint64_t* l_1_coefficients, unsigned log2_src, unsigned log2_target)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (l_1_coefficients == nullptr) return BBG_E_BAD_ARGUMENT;
    if (log2_target < log2_src || log2_target > 24) return BBG_E_BAD_SIZE;
    const size_t bytes = ((size_t)32) << log2_target;
    BBG_CHECK(g_stage_coeffs.ensure(bytes));
    BBG_CHECK(lagrange_fft_device(g_stage_coeffs.p, log2_src, log2_target, g_stream));
    BBG_CHECK(bbg_hostcopy::d2h(l_1_coefficients, g_stage_coeffs.p, bytes, g_stream));
    return bbg_rt::sync(g_stream);
}

int bbg_compute_lagrange_polynomial_fft_dev(void* d_out, unsigned log2_src, unsigned log2_target)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (d_out == nullptr) return BBG_E_BAD_ARGUMENT;
    return lagrange_fft_device(d_out, log2_src, log2_target, g_stream);
}

// ---- MSM ----------------------------------------------------------------------------------------
int bbg_srs_register(const uint64_t* table_2n, size_t n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (table_2n == nullptr || n == 0) return BBG_E_BAD_ARGUMENT;
    for (size_t i = 0; i < g_srs.size(); ++i)
    {
        if (g_srs[i].host_base == table_2n)
        {
            srs_drop(i);
            break;
        }
    }
    void* d = nullptr;
    BBG_CHECK(bbg_rt::dev_alloc(&d, n * 128));
    int e = bbg_hostcopy::h2d(d, table_2n, n * 128, g_stream);
    if (e == 0) e = bbg_rt::sync(g_stream);
    if (e != 0)
    {
        bbg_rt::dev_free(d);
        return e;
    }
    return srs_add(table_2n, n, d, false);
}

int bbg_set_auto_srs_cache(int enable)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    g_auto_srs = enable != 0;
    return 0;
}

int bbg_set_srs_precompute(int enable)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    g_srs_precompute = enable != 0;
    quiesce_all();
    for (const SrsEntry& s : g_srs)
    {
        if (g_srs_precompute) BBG_CHECK(msm_fixed_base_build(s.d_table, s.n, g_stream));
        else msm_fixed_base_drop(s.d_table);
    }
    return 0;
}

int bbg_srs_device_table(const uint64_t* table_2n, void** d_table, int* window_bits, int* windows)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (table_2n == nullptr || d_table == nullptr) return BBG_E_BAD_ARGUMENT;
    for (const SrsEntry& s : g_srs)
    {
        if (table_2n >= s.host_base && table_2n < s.host_base + 16 * s.n)
        {
            *d_table = (char*)s.d_table + ((const char*)table_2n - (const char*)s.host_base);
            int c = 0, W = 0;
            msm_fixed_base_info(s.d_table, &c, &W);
            if (window_bits) *window_bits = c;
            if (windows) *windows = W;
            return 0;
        }
    }
    return BBG_E_BAD_ARGUMENT;
}

int bbg_srs_unregister(const uint64_t* table_2n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    for (size_t i = 0; i < g_srs.size(); ++i)
    {
        if (g_srs[i].host_base == table_2n)
        {
            srs_drop(i);
            return 0;
        }
    }
    return BBG_E_BAD_ARGUMENT;
}

int bbg_msm_g1(const uint64_t* scalars, const uint64_t* points_table, size_t n, uint64_t out_xyz[12])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    hostg1::hxyzz r;
    BBG_CHECK(msm_host(scalars, points_table, n, &r));
    hostg1::to_normalized_jacobian(r, out_xyz);
    return 0;
}

// pippenger_low_memory / pippenger_precomputed take the n plain points and apply the endomorphism on the fly
// (scalar_multiplication.cpp:142-263, :478-573); here the 2n-entry table is built on the device next to the upload, and —
// with the auto cache on — kept (with its fixed-base windows when bbg_set_srs_precompute is on) for the next call that
// names the same points
int bbg_msm_g1_points(const uint64_t* scalars, const uint64_t* points_n, size_t n, uint64_t out_xyz[12])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (out_xyz == nullptr) return BBG_E_BAD_ARGUMENT;
    hostg1::hxyzz r = hostg1::infinity();
    if (n > 0)
    {
        if (scalars == nullptr || points_n == nullptr) return BBG_E_BAD_ARGUMENT;
        const void* d_table = nullptr;
        for (size_t i = 0; i < g_plain.size() && d_table == nullptr; ++i)
        {
            PlainEntry& e = g_plain[i];
            if (e.host_points != points_n || e.n < n) continue;
            if (plain_fingerprint(e.host_points, e.n) == e.fingerprint)
            {
                d_table = e.d_table;
                break;
            }
            quiesce_all();
            msm_fixed_base_drop(e.d_table);
            msm_multi_drop_replica(e.d_table);
            bbg_rt::dev_free(e.d_table);
            g_plain.erase(g_plain.begin() + (long)i);
            break;
        }
        if (d_table == nullptr)
        {
            const bool keep = g_auto_srs && n >= AUTO_SRS_MIN_POINTS;
            void* d_tab = nullptr;
            if (keep) BBG_CHECK(bbg_rt::dev_alloc(&d_tab, n * 128));
            else
            {
                BBG_CHECK(g_stage_table.ensure(n * 192));
                d_tab = g_stage_table.p;
            }
            // the plain points are staged behind the staging table (or, for a kept table, in the scalar staging buffer)
            void* d_points = nullptr;
            if (keep)
            {
                BBG_CHECK(g_stage_coeffs.ensure(n * 64));
                d_points = g_stage_coeffs.p;
            }
            else
                d_points = (char*)d_tab + n * 128;
            int e = bbg_hostcopy::h2d(d_points, points_n, n * 64, g_stream);
            if (e == 0) e = g1_build_endo_table_device(d_points, d_tab, n, g_stream);
            if (e == 0 && keep)
            {
                if (g_plain.size() >= AUTO_SRS_MAX_ENTRIES)
                {
                    quiesce_all();
                    msm_fixed_base_drop(g_plain[0].d_table);
                    msm_multi_drop_replica(g_plain[0].d_table);
                    bbg_rt::dev_free(g_plain[0].d_table);
                    g_plain.erase(g_plain.begin());
                }
                e = bbg_rt::sync(g_stream);
                if (e == 0) e = msm_multi_replicate(d_tab, n * 128, g_stream);
                if (e == 0 && g_srs_precompute) e = msm_fixed_base_build(d_tab, n, g_stream);
                if (e == 0) g_plain.push_back({ points_n, n, d_tab, plain_fingerprint(points_n, n) });
            }
            if (e != 0)
            {
                if (keep)
                {
                    bbg_rt::sync(g_stream);
                    msm_fixed_base_drop(d_tab);
                    msm_multi_drop_replica(d_tab);
                    bbg_rt::dev_free(d_tab);
                }
                return e;
            }
            d_table = d_tab;
        }
        BBG_CHECK(g_stage_scalars.ensure(n * 32));
        int ticket = -1;
        const void* one[1] = { scalars };
        BBG_CHECK(msm_launch_any(0, one, true, g_stage_scalars.p, 1, d_table, n, g_stream, &ticket));
        BBG_CHECK(msm_finish(ticket, &r));
    }
    hostg1::to_normalized_jacobian(r, out_xyz);
    return 0;
}

// generate_pippenger_precompute_table (scalar_multiplication.cpp:90-129) on the device: table[i * n + j] =
// 2^((bits_per_bucket + 1)(i + 1)) P_j for i < rounds - 1, rounds = WNAF_SIZE(bits_per_bucket + 1), canonical affine
int bbg_generate_pippenger_precompute_table(const uint64_t* points_n, uint64_t* table, size_t n, unsigned bits_per_bucket)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (n == 0) return 0;
    if (points_n == nullptr || table == nullptr || bits_per_bucket < 1 || bits_per_bucket > 30) return BBG_E_BAD_ARGUMENT;
    const int w = (int)bits_per_bucket + 1;
    const int rounds = (127 + w) / w; // WNAF_SIZE (wnaf.hpp:5)
    if (rounds < 2) return 0;
    void *d_pts = nullptr, *d_out = nullptr;
    BBG_CHECK(bbg_rt::dev_alloc(&d_pts, n * 64));
    int e = bbg_rt::dev_alloc(&d_out, (size_t)(rounds - 1) * n * 64);
    if (e == 0) e = bbg_hostcopy::h2d(d_pts, points_n, n * 64, g_stream);
    if (e == 0) e = g1_precompute_plain_device(d_pts, d_out, n, w, rounds, g_stream);
    if (e == 0) e = bbg_hostcopy::d2h(table, d_out, (size_t)(rounds - 1) * n * 64, g_stream);
    if (e == 0) e = bbg_rt::sync(g_stream);
    bbg_rt::dev_free(d_pts);
    if (d_out) bbg_rt::dev_free(d_out);
    return e;
}

int bbg_msm_g1_batched(const uint64_t* const* scalars, const uint64_t* const* points_tables, size_t n, size_t batches, uint64_t* out_xyz)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (batches == 0) return 0;
    if (scalars == nullptr || points_tables == nullptr || out_xyz == nullptr) return BBG_E_BAD_ARGUMENT;
    // the prover's case (prover.cpp:65-124, :640-652): several polynomials against the same table -> one pipeline
    constexpr size_t GROUP = 4;
    size_t i = 0;
    while (i < batches)
    {
        size_t same = 1;
        while (n > 0 && i + same < batches && same < GROUP && points_tables[i + same] == points_tables[i]) ++same;
        if (same == 1)
        {
            hostg1::hxyzz r;
            BBG_CHECK(msm_host(scalars[i], points_tables[i], n, &r));
            hostg1::to_normalized_jacobian(r, out_xyz + 12 * i);
        }
        else
        {
            hostg1::hxyzz r[GROUP];
            BBG_CHECK(msm_host_batched(scalars + i, same, points_tables[i], n, r));
            for (size_t k = 0; k < same; ++k) hostg1::to_normalized_jacobian(r[k], out_xyz + 12 * (i + k));
        }
        i += same;
    }
    return 0;
}

int bbg_msm_g1_partial_dev(const void* d_scalars, const void* d_table, size_t n, uint64_t out_xyzz[16])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    return msm_device(d_scalars, d_table, n, out_xyzz, g_stream);
}

int bbg_msm_g1_partial_dev_launch(const void* d_scalars, const void* d_table, size_t n, int* ticket)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (ticket == nullptr || (n > 0 && (d_scalars == nullptr || d_table == nullptr))) return BBG_E_BAD_ARGUMENT;
    cudaStream_t st = g_stream;
#ifndef BBG_EMULATE
    // the MSM sees everything queued on the work stream so far, then runs beside whatever is queued next
    BBG_CHECK(cudaEventRecord(g_msm_fence, g_stream));
    BBG_CHECK(cudaStreamWaitEvent(g_msm_stream, g_msm_fence, 0));
    st = g_msm_stream;
#endif
    const void* one[1] = { d_scalars };
    return msm_launch(1, one, 1, d_table, n, st, ticket);
}

int bbg_msm_g1_partial_finish(int ticket, uint64_t out_xyzz[16])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (out_xyzz == nullptr) return BBG_E_BAD_ARGUMENT;
    return msm_finish(ticket, out_xyzz);
}

int bbg_msm_g1_launch(const uint64_t* scalars, const uint64_t* points_table, size_t n, int* ticket)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (ticket == nullptr) return BBG_E_BAD_ARGUMENT;
    *ticket = -1; // (n == 0: nothing to wait for, finish returns the point at infinity)
    if (n == 0) return 0;
    if (scalars == nullptr || points_table == nullptr) return BBG_E_BAD_ARGUMENT;
    const void* d_table = nullptr;
    // keep: the MSM outlives this call, so the table must not be the shared staging buffer
    BBG_CHECK(resolve_table(points_table, n, &d_table, /*keep=*/true));
    // own staging buffer per pending launch, own stream: uploads and kernels of successive launches are ordered by that
    // stream, and the calls made on the work stream in between (transforms with their own copies) run beside them
    AsyncSlot* slot = nullptr;
    for (AsyncSlot& a : g_async_slots)
    {
        if (a.ticket >= 0 && !msm_ticket_pending(a.ticket)) a.ticket = -1;
        if (slot == nullptr && a.ticket < 0) slot = &a;
    }
    if (slot == nullptr) return BBG_E_BAD_ARGUMENT; // too many MSMs in flight
    cudaStream_t st = g_stream;
#ifndef BBG_EMULATE
    st = g_msm_stream;
    // a table uploaded just now travels on the work stream: the MSM stream must see it
    BBG_CHECK(cudaEventRecord(g_msm_fence, g_stream));
    BBG_CHECK(cudaStreamWaitEvent(g_msm_stream, g_msm_fence, 0));
    if (slot->scalars.bytes < n * 32) BBG_CHECK(cudaStreamSynchronize(g_msm_stream)); // growing frees the old buffer
#endif
    BBG_CHECK(slot->scalars.ensure(n * 32));
#ifndef BBG_EMULATE
    // a buffer the registration cache page-locked is still "pageable" to its owner: theirs again when this call returns
    const bool wait_copy = bbg_hostcopy::copy_is_async(scalars, n * 32) && !bbg_hostcopy::caller_pinned_whole(scalars, n * 32);
    BBG_CHECK(bbg_hostcopy::h2d_ring(g_msm_ring, slot->scalars.p, scalars, n * 32, st));
    if (wait_copy)
    {
        BBG_CHECK(cudaEventRecord(g_msm_fence, st));
        BBG_CHECK(cudaEventSynchronize(g_msm_fence));
    }
#else
    BBG_CHECK(bbg_hostcopy::h2d(slot->scalars.p, scalars, n * 32, st));
#endif
    const void* one[1] = { slot->scalars.p };
    BBG_CHECK(msm_launch(1, one, 1, d_table, n, st, ticket));
    slot->ticket = *ticket;
    return 0;
}

int bbg_msm_g1_finish(int ticket, uint64_t out_xyz[12])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (out_xyz == nullptr) return BBG_E_BAD_ARGUMENT;
    hostg1::hxyzz r = hostg1::infinity();
    if (ticket != -1) BBG_CHECK(msm_finish(ticket, &r));
    hostg1::to_normalized_jacobian(r, out_xyz);
    return 0;
}

int bbg_msm_g1_dev(const void* d_scalars, const void* d_table, size_t n, uint64_t out_xyz[12])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    hostg1::hxyzz r;
    BBG_CHECK(msm_device(d_scalars, d_table, n, &r, g_stream));
    hostg1::to_normalized_jacobian(r, out_xyz);
    return 0;
}

int bbg_g1_fold_partials(const uint64_t* partials_xyzz, size_t count, uint64_t out_xyz[12])
{
    hostg1::hxyzz acc = hostg1::infinity();
    for (size_t i = 0; i < count; ++i)
    {
        hostg1::hxyzz p;
        memcpy(&p, partials_xyzz + 16 * i, sizeof p);
        acc = hostg1::add(acc, p);
    }
    hostg1::to_normalized_jacobian(acc, out_xyz);
    return 0;
}

int bbg_generate_pippenger_point_table(const uint64_t* points_n, uint64_t* table_2n, size_t n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (n == 0) return 0;
    if (points_n == nullptr || table_2n == nullptr) return BBG_E_BAD_ARGUMENT;
    void *d_pts = nullptr, *d_tab = nullptr;
    BBG_CHECK(bbg_rt::dev_alloc(&d_pts, n * 64));
    int e = bbg_rt::dev_alloc(&d_tab, n * 128);
    if (e == 0) e = bbg_hostcopy::h2d(d_pts, points_n, n * 64, g_stream);
    if (e == 0) e = g1_build_endo_table_device(d_pts, d_tab, n, g_stream);
    if (e == 0) e = bbg_hostcopy::d2h(table_2n, d_tab, n * 128, g_stream);
    if (e == 0) e = bbg_rt::sync(g_stream);
    bbg_rt::dev_free(d_pts);
    if (d_tab) bbg_rt::dev_free(d_tab);
    return e;
}

// ---- the reference's stand-alone polynomial helpers, host buffers ------------------
int bbg_fr_evaluate(const uint64_t* coeffs, size_t n, const uint64_t z[4], uint64_t out[4])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if ((coeffs == nullptr && n > 0) || z == nullptr || out == nullptr) return BBG_E_BAD_ARGUMENT;
    BBG_CHECK(g_stage_coeffs.ensure(n * 32 + 32));
    BBG_CHECK(bbg_hostcopy::h2d(g_stage_coeffs.p, coeffs, n * 32, g_stream));
    return plonk::evaluate_device(g_stage_coeffs.p, n, z, out, g_stream);
}

int bbg_fr_divide_by_pseudo_vanishing_polynomial(uint64_t* coeffs, unsigned log2_src, unsigned log2_target)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (coeffs == nullptr) return BBG_E_BAD_ARGUMENT;
    if (log2_target > 28) return BBG_E_BAD_SIZE;
    const size_t bytes = ((size_t)32) << log2_target;
    BBG_CHECK(g_stage_coeffs.ensure(bytes));
    BBG_CHECK(bbg_hostcopy::h2d(g_stage_coeffs.p, coeffs, bytes, g_stream));
    BBG_CHECK(plonk::divide_by_pseudo_vanishing_device(g_stage_coeffs.p, log2_src, log2_target, g_stream));
    BBG_CHECK(bbg_hostcopy::d2h(coeffs, g_stage_coeffs.p, bytes, g_stream));
    return bbg_rt::sync(g_stream);
}

int bbg_fr_compute_kate_opening_coefficients(const uint64_t* src, uint64_t* dest, const uint64_t z[4], size_t n, uint64_t f_out[4])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (src == nullptr || dest == nullptr || z == nullptr || f_out == nullptr || n == 0) return BBG_E_BAD_ARGUMENT;
    BBG_CHECK(g_stage_coeffs.ensure(2 * n * 32));
    char* d_src = (char*)g_stage_coeffs.p;
    char* d_dest = d_src + n * 32;
    BBG_CHECK(bbg_hostcopy::h2d(d_src, src, n * 32, g_stream));
    BBG_CHECK(plonk::kate_opening_device(d_src, d_dest, n, z, f_out, g_stream));
    BBG_CHECK(bbg_hostcopy::d2h(dest, d_dest, n * 32, g_stream));
    return bbg_rt::sync(g_stream);
}

// ---- prover construction helpers -----------------------------------------------
int bbg_fr_domain_lookup_table(uint64_t* roots, unsigned log2_size)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (roots == nullptr) return BBG_E_BAD_ARGUMENT;
    if (log2_size < 1 || log2_size > 28) return BBG_E_BAD_SIZE;
    const size_t bytes = ((size_t)64) << log2_size;
    BBG_CHECK(g_stage_coeffs.ensure(bytes));
    BBG_CHECK(domain_lookup_table_device(g_stage_coeffs.p, log2_size, g_stream));
    BBG_CHECK(bbg_hostcopy::d2h(roots, g_stage_coeffs.p, bytes, g_stream));
    return bbg_rt::sync(g_stream);
}

int bbg_srs_from_transcript(const uint8_t* g1_bytes, size_t n, uint64_t* table_2n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (table_2n == nullptr || n == 0 || (g1_bytes == nullptr && n > 1)) return BBG_E_BAD_ARGUMENT;
    // a table registered earlier behind the same host address is stale now
    for (size_t i = 0; i < g_srs.size(); ++i)
    {
        if (g_srs[i].host_base == table_2n)
        {
            srs_drop(i);
            break;
        }
    }
    struct
    {
        void* d_table;
    } s = { nullptr };
    BBG_CHECK(bbg_rt::dev_alloc(&s.d_table, n * 128));
    int e = 0;
    if (n > 1)
    {
        e = g_stage_table.ensure((n - 1) * 64);
        if (e == 0) e = bbg_hostcopy::h2d(g_stage_table.p, g1_bytes, (n - 1) * 64, g_stream);
    }
    if (e == 0) e = g1_table_from_transcript_device(g_stage_table.p, s.d_table, n, g_stream);
    if (e == 0) e = bbg_hostcopy::d2h(table_2n, s.d_table, n * 128, g_stream);
    if (e == 0) e = bbg_rt::sync(g_stream);
    if (e != 0)
    {
        bbg_rt::dev_free(s.d_table);
        return e;
    }
    return srs_add(table_2n, n, s.d_table, false); // the device copy IS the registered SRS: no second upload
}

// ---- HBM-resident PLONK prover rounds (bbg_plonk.cu) -------------------------------------------------
int bbg_plonk_create(unsigned log2_n, bbg_plonk_prover** out)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (out == nullptr) return BBG_E_BAD_ARGUMENT;
    plonk::Prover* p = nullptr;
    BBG_CHECK(plonk::create(log2_n, &p));
    *out = (bbg_plonk_prover*)p;
    return 0;
}
int bbg_plonk_destroy(bbg_plonk_prover* p)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (p == nullptr) return 0;
    quiesce_all();
    for (size_t k = 0; k < g_prover_srs.size();)
    {
        if (g_prover_srs[k].first != (const void*)p)
        {
            ++k;
            continue;
        }
        for (SrsEntry& s : g_srs)
            if (s.host_base == g_prover_srs[k].second && s.pins > 0) --s.pins;
        g_prover_srs.erase(g_prover_srs.begin() + (long)k);
    }
    plonk::destroy((plonk::Prover*)p);
    return 0;
}
int bbg_plonk_set_witness(bbg_plonk_prover* p, const uint64_t* w_l, const uint64_t* w_r, const uint64_t* w_o)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::set_witness((plonk::Prover*)p, w_l, w_r, w_o, g_stream);
}
int bbg_plonk_set_permutation(bbg_plonk_prover* p, const uint32_t* sigma_1_mapping, const uint32_t* sigma_2_mapping, const uint32_t* sigma_3_mapping)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::set_permutation((plonk::Prover*)p, sigma_1_mapping, sigma_2_mapping, sigma_3_mapping, g_stream);
}
int bbg_plonk_set_widgets(bbg_plonk_prover* p, const int* kinds, int count, const uint64_t* const* selectors)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::set_widgets((plonk::Prover*)p, kinds, count, selectors, g_stream);
}
int bbg_plonk_set_arithmetic_selectors(bbg_plonk_prover* p, const uint64_t* q_m, const uint64_t* q_l, const uint64_t* q_r, const uint64_t* q_o,
                                       const uint64_t* q_c)
{
    const int kind = BBG_WIDGET_ARITHMETIC;
    const uint64_t* q[5] = { q_m, q_l, q_r, q_o, q_c };
    return bbg_plonk_set_widgets(p, &kind, 1, q);
}
int bbg_plonk_set_srs(bbg_plonk_prover* p, const uint64_t* points_table, size_t n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || points_table == nullptr) return BBG_E_BAD_ARGUMENT;
    const void* d_table = nullptr;
    // the prover keeps the device pointer: the table is cached whatever its size and pinned against eviction until the
    // prover is destroyed or handed another table.  (A table whose host buffer is rewritten in place is dropped by the
    // fingerprint check of the next MSM that names it; call bbg_plonk_set_srs again after changing it.)
    BBG_CHECK(resolve_table(points_table, n, &d_table, /*keep=*/true));
    for (size_t k = 0; k < g_prover_srs.size();)
    {
        if (g_prover_srs[k].first != (const void*)p)
        {
            ++k;
            continue;
        }
        for (SrsEntry& s : g_srs)
            if (s.host_base == g_prover_srs[k].second && s.pins > 0) --s.pins;
        g_prover_srs.erase(g_prover_srs.begin() + (long)k);
    }
    for (SrsEntry& s : g_srs)
    {
        if (points_table >= s.host_base && points_table + 16 * n <= s.host_base + 16 * s.n)
        {
            ++s.pins;
            g_prover_srs.push_back({ (const void*)p, s.host_base });
            break;
        }
    }
    return plonk::set_srs((plonk::Prover*)p, d_table);
}
int bbg_plonk_round_wires(bbg_plonk_prover* p, uint64_t* out_xyz)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || out_xyz == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::round_wires((plonk::Prover*)p, out_xyz, g_stream);
}
int bbg_plonk_round_grand_product(bbg_plonk_prover* p, const uint64_t beta[4], const uint64_t gamma[4], uint64_t out_xyz[12])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || beta == nullptr || gamma == nullptr || out_xyz == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::round_grand_product((plonk::Prover*)p, beta, gamma, out_xyz, g_stream);
}
int bbg_plonk_round_quotient(bbg_plonk_prover* p, const uint64_t beta[4], const uint64_t gamma[4], const uint64_t alpha[4], const uint64_t alpha_base[4],
                             uint64_t* out_xyz)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || beta == nullptr || gamma == nullptr || alpha == nullptr || alpha_base == nullptr || out_xyz == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::round_quotient((plonk::Prover*)p, beta, gamma, alpha, alpha_base, out_xyz, g_stream);
}
int bbg_plonk_round_evaluations(bbg_plonk_prover* p, const uint64_t zeta[4], const uint64_t zeta_omega[4], uint64_t* out_evals)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || zeta == nullptr || zeta_omega == nullptr || out_evals == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::round_evaluations((plonk::Prover*)p, zeta, zeta_omega, out_evals, g_stream);
}
int bbg_plonk_round_linearise(bbg_plonk_prover* p, const uint64_t* scalars, const uint64_t zeta[4], uint64_t out_linear_eval[4])
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || scalars == nullptr || zeta == nullptr || out_linear_eval == nullptr) return BBG_E_BAD_ARGUMENT;
    return plonk::round_linearise((plonk::Prover*)p, scalars, zeta, out_linear_eval, g_stream);
}
int bbg_plonk_round_openings(bbg_plonk_prover* p, const uint64_t* nu_powers, const uint64_t beta_inv[4], const uint64_t zeta[4],
                             const uint64_t zeta_omega[4], const uint64_t* wire_shift_terms, const uint64_t* selector_terms, uint64_t* out_xyz)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (p == nullptr || nu_powers == nullptr || beta_inv == nullptr || zeta == nullptr || zeta_omega == nullptr || wire_shift_terms == nullptr ||
        selector_terms == nullptr || out_xyz == nullptr)
        return BBG_E_BAD_ARGUMENT;
    return plonk::round_openings((plonk::Prover*)p, nu_powers, beta_inv, zeta, zeta_omega, wire_shift_terms, selector_terms, out_xyz, g_stream);
}

// ---- device self test of the field / group primitives (bbg_selftest.cu) ---------------------------------
namespace
{
int selftest_run(bool group, int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count)
{
    const size_t elem = group ? 64 : 32;
    if (count == 0) return 0;
    if (a == nullptr || out == nullptr || (group && b == nullptr)) return BBG_E_BAD_ARGUMENT;
    BBG_CHECK(g_stage_coeffs.ensure(3 * count * elem));
    char* d_a = (char*)g_stage_coeffs.p;
    char* d_b = d_a + count * elem;
    char* d_o = d_b + count * elem;
    BBG_CHECK(bbg_hostcopy::h2d(d_a, a, count * elem, g_stream));
    if (b != nullptr) BBG_CHECK(bbg_hostcopy::h2d(d_b, b, count * elem, g_stream));
    if (group) BBG_CHECK(g1_selftest_device(op, d_a, d_b, d_o, count, g_stream));
    else BBG_CHECK(field_selftest_device(field, op, d_a, b != nullptr ? d_b : nullptr, d_o, count, g_stream));
    BBG_CHECK(bbg_hostcopy::d2h(out, d_o, count * elem, g_stream));
    return bbg_rt::sync(g_stream);
}
} // namespace
int bbg_field_selftest(int field, int op, const uint64_t* a, const uint64_t* b, uint64_t* out, size_t count)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (field < 0 || field > 1 || op < 0 || op > 14) return BBG_E_BAD_ARGUMENT;
    return selftest_run(false, field, op, a, b, out, count);
}
int bbg_g1_selftest(int op, const uint64_t* p, const uint64_t* q, uint64_t* out, size_t count)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (op < 0 || op > 6) return BBG_E_BAD_ARGUMENT;
    return selftest_run(true, 0, op, p, q, out, count);
}

// ---- caller-owned host buffers ----------------------------------------------------------------------
int bbg_set_host_register_cache(int enable)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (!enable) quiesce_all();
    bbg_hostcopy::reg_cache().enable(enable != 0);
    return 0;
}
int bbg_host_register_stats(double* register_ms, uint64_t* registered_bytes, uint64_t* registrations)
{
    if (register_ms == nullptr || registered_bytes == nullptr || registrations == nullptr) return BBG_E_BAD_ARGUMENT;
    unsigned long long b = 0, c = 0;
    bbg_hostcopy::reg_cache().stats(register_ms, &b, &c);
    *registered_bytes = b;
    *registrations = c;
    return 0;
}
int bbg_host_buffer_forget(const void* host_ptr)
{
    // no library lock: free() wrappers call this from any thread, usually for blocks the library never saw
    if (host_ptr != nullptr) bbg_hostcopy::reg_cache().forget(host_ptr);
    return 0;
}

// ---- device memory helpers ------------------------------------------------------------------------
int bbg_dev_alloc(void** d_ptr, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    return bbg_rt::dev_alloc(d_ptr, bytes);
}
int bbg_dev_free(void* d_ptr)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    quiesce_all();
    return bbg_rt::dev_free(d_ptr);
}
int bbg_copy_h2d(void* d_dst, const void* h_src, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    BBG_CHECK(bbg_rt::h2d(d_dst, h_src, bytes, g_stream));
    return bbg_rt::sync(g_stream);
}
int bbg_copy_d2h(void* h_dst, const void* d_src, size_t bytes)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    BBG_CHECK(bbg_rt::d2h(h_dst, d_src, bytes, g_stream));
    return bbg_rt::sync(g_stream);
}
int bbg_sync(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
#ifndef BBG_EMULATE
    if (g_msm_stream) BBG_CHECK(cudaStreamSynchronize(g_msm_stream));
#endif
    return bbg_rt::sync(g_stream);
}
int bbg_timer_start(void)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
#ifndef BBG_EMULATE
    return (int)cudaEventRecord(g_ev_start, g_stream);
#else
    return 0;
#endif
}
int bbg_timer_stop(float* elapsed_ms)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
#ifndef BBG_EMULATE
    BBG_CHECK(cudaEventRecord(g_ev_stop, g_stream));
    BBG_CHECK(cudaEventSynchronize(g_ev_stop));
    return (int)cudaEventElapsedTime(elapsed_ms, g_ev_start, g_ev_stop);
#else
    *elapsed_ms = 0.f;
    return 0;
#endif
}

int bbg_g1_generate_multiples_dev(const uint64_t start[4], const uint64_t step[4], void* d_points, size_t n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (start == nullptr || step == nullptr || (d_points == nullptr && n > 0)) return BBG_E_BAD_ARGUMENT;
    return g1_generate_progression_device(start, step, d_points, n, g_stream);
}

int bbg_generate_pippenger_point_table_dev(const void* d_points, void* d_table, size_t n)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (d_points == d_table) return BBG_E_BAD_ARGUMENT;
    return g1_build_endo_table_device(d_points, d_table, n, g_stream);
}

int bbg_profile_enable(int on)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    bbg_rt::sync(g_stream);
    bbg_prof::reset();
    bbg_prof::state().on = on != 0;
    return 0;
}
int bbg_profile_count(void) { return (int)bbg_prof::NUM_IDS; }
const char* bbg_profile_name(int id) { return bbg_prof::name(id); }
int bbg_profile_read(int id, double* total_ms, uint64_t* launches)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    if (id < 0 || id >= (int)bbg_prof::NUM_IDS || total_ms == nullptr || launches == nullptr) return BBG_E_BAD_ARGUMENT;
    bbg_prof::collect();
    *total_ms = bbg_prof::state().total_ms[id];
    *launches = bbg_prof::state().count[id];
    return 0;
}

int bbg_microbench(int mode, int iters, double* ops_per_second, float* elapsed_ms)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    BBG_CHECK(ensure_ready());
    if (iters <= 0 || ops_per_second == nullptr || elapsed_ms == nullptr) return BBG_E_BAD_ARGUMENT;
#ifndef BBG_EMULATE
    const int blocks = bbg_rt::num_sms() * 8;
    uint32_t* d_out = nullptr;
    BBG_CHECK(bbg_rt::dev_alloc((void**)&d_out, (size_t)blocks * 256 * 4));
    double ops = 0;
    int e = microbench_launch(mode, iters / 8 > 0 ? iters / 8 : 1, d_out, blocks, &ops, g_stream); // warm-up
    float best = 0.f;
    for (int rep = 0; rep < 5 && e == 0; ++rep)
    {
        cudaEventRecord(g_ev_start, g_stream);
        e = microbench_launch(mode, iters, d_out, blocks, &ops, g_stream);
        cudaEventRecord(g_ev_stop, g_stream);
        if (e == 0) e = (int)cudaEventSynchronize(g_ev_stop);
        float ms = 0.f;
        if (e == 0) e = (int)cudaEventElapsedTime(&ms, g_ev_start, g_ev_stop);
        if (rep == 0 || ms < best) best = ms;
        g_misc_launches += 1;
    }
    bbg_rt::dev_free(d_out);
    if (e != 0) return e;
    *elapsed_ms = best;
    *ops_per_second = ops / ((double)best * 1e-3);
    return 0;
#else
    (void)mode;
    *ops_per_second = 0;
    *elapsed_ms = 0;
    return 0;
#endif
}

} // extern "C"
