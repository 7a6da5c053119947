"""ctypes loader for libbbgpu.so (include/bbgpu.h).  Fails loudly when the CUDA library is missing."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
u64p = C.POINTER(C.c_uint64)


class BbgError(RuntimeError):
    def __init__(self, code, message):
        super().__init__("bbgpu error %d: %s" % (code, message))
        self.code = code


def library_path():
    return os.path.join(_HERE, "libbbgpu.so")


def _as_u64(a, shape_tail):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim < 1 or tuple(a.shape[-len(shape_tail):]) != tuple(shape_tail):
        raise ValueError("expected trailing shape %s, got %s" % (shape_tail, a.shape))
    return a


class Library:
    """One loaded copy of the C ABI.  `path` defaults to the in-tree CUDA build; the test-suite also points it
    at the CPU kernel-emulation build under tests/emul (test infrastructure, never the product path)."""

    OPS = {"fft": 0, "ifft": 1, "coset_fft": 2, "coset_ifft": 3, "fft_with_constant": 4,
           "ifft_with_constant": 5, "coset_fft_with_constant": 6}

    def __init__(self, path=None, device=0, devices=None):
        """device: the CUDA device of a single-GPU instance.  devices: a list — devices[0] primary, the rest take point
        ranges of large MSMs over registered tables (bbg_init_multi)."""
        path = path or library_path()
        if not os.path.exists(path):
            raise BbgError(-1, "%s not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                               "(there is no CPU fallback)" % path)
        self.path = path
        self.lib = C.CDLL(path)
        L = self.lib
        L.bbg_error_string.restype = C.c_char_p
        L.bbg_launch_count.restype = C.c_uint64
        L.bbg_init.argtypes = [C.c_int]
        L.bbg_set_stream.argtypes = [C.c_void_p]
        L.bbg_ntt_fr.argtypes = [C.c_void_p, C.c_uint, C.c_int, C.c_void_p]
        L.bbg_ntt_fr_batched.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint, C.c_int, C.c_void_p]
        L.bbg_ntt_fr_dev.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_uint, C.c_int, C.c_void_p]
        L.bbg_srs_register.argtypes = [C.c_void_p, C.c_size_t]
        L.bbg_srs_unregister.argtypes = [C.c_void_p]
        L.bbg_msm_g1.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.bbg_msm_g1_batched.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t, C.c_size_t, C.c_void_p]
        L.bbg_msm_g1_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.bbg_msm_g1_partial_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.bbg_g1_fold_partials.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.bbg_msm_g1_partial_dev_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        L.bbg_msm_g1_partial_finish.argtypes = [C.c_int, C.c_void_p]
        L.bbg_generate_pippenger_point_table.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbg_dev_alloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]
        L.bbg_dev_free.argtypes = [C.c_void_p]
        L.bbg_copy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbg_copy_d2h.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbg_timer_stop.argtypes = [C.POINTER(C.c_float)]
        L.bbg_microbench.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_float)]
        L.bbg_g1_generate_multiples_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbg_generate_pippenger_point_table_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbg_compute_lagrange_polynomial_fft.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        L.bbg_fr_domain_lookup_table.argtypes = [C.c_void_p, C.c_uint]
        L.bbg_fr_evaluate.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.bbg_fr_divide_by_pseudo_vanishing_polynomial.argtypes = [C.c_void_p, C.c_uint, C.c_uint]
        L.bbg_fr_compute_kate_opening_coefficients.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
        L.bbg_srs_from_transcript.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
        L.bbg_profile_name.restype = C.c_char_p
        L.bbg_profile_read.argtypes = [C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
        L.bbg_msm_g1_launch.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int)]
        L.bbg_msm_g1_finish.argtypes = [C.c_int, C.c_void_p]
        L.bbg_init_multi.argtypes = [C.POINTER(C.c_int), C.c_int]
        L.bbg_host_buffer_forget.argtypes = [C.c_void_p]
        L.bbg_field_selftest.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.bbg_g1_selftest.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        if devices is not None and len(devices) > 1:
            arr = (C.c_int * len(devices))(*[int(d) for d in devices])
            self.check(L.bbg_init_multi(arr, len(devices)))
        else:
            self.check(L.bbg_init(device if devices is None else int(devices[0])))

    FIELD_OPS = {"mul_coarse": 0, "sqr_coarse": 1, "mul_const": 2, "add_coarse": 3, "sub_coarse": 4, "reduce_once": 5, "neg": 6,
                 "to_mont": 7, "from_mont": 8, "invert": 9, "sub_lazy": 10, "mul": 11, "mul_const_raw": 12,
                 "invert_binary": 13, "mul2": 14}
    G1_OPS = {"mixed_add": 0, "add_doubled": 1, "dbl_dbl": 2, "accumulate": 3, "add": 4, "endo_entry": 5, "dbl_affine": 6}

    def set_srs_precompute(self, on=True):
        """Fixed-base windows for every registered / cached table from now on (and for those already registered)."""
        self.check(self.lib.bbg_set_srs_precompute(1 if on else 0))

    def srs_device_table(self, table):
        """(device pointer, window bits, windows) of a registered host table; windows == 0: no fixed-base tables."""
        d, c, w = C.c_void_p(), C.c_int(), C.c_int()
        self.lib.bbg_srs_device_table.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        self.check(self.lib.bbg_srs_device_table(table.ctypes.data_as(C.c_void_p), C.byref(d), C.byref(c), C.byref(w)))
        return d.value, c.value, w.value

    def device_count(self):
        return int(self.lib.bbg_device_count())

    def set_host_register_cache(self, on=True):
        self.check(self.lib.bbg_set_host_register_cache(1 if on else 0))

    def host_register_stats(self):
        ms, b, c = C.c_double(), C.c_uint64(), C.c_uint64()
        self.check(self.lib.bbg_host_register_stats(C.byref(ms), C.byref(b), C.byref(c)))
        return {"register_ms": ms.value, "registered_bytes": int(b.value), "registrations": int(c.value)}

    def host_buffer_forget(self, array_or_ptr):
        p = array_or_ptr.ctypes.data if isinstance(array_or_ptr, np.ndarray) else int(array_or_ptr)
        self.check(self.lib.bbg_host_buffer_forget(C.c_void_p(p)))

    def field_selftest(self, field, op, a, b=None):
        """Device field primitive `op` (FIELD_OPS) element-wise on (count, 4) uint64 arrays; field 0 = Fq, 1 = Fr."""
        a = _as_u64(a, (4,))
        bp = None
        if b is not None:
            b = _as_u64(b, (4,))
            assert b.shape == a.shape
            bp = b.ctypes.data_as(C.c_void_p)
        out = np.zeros_like(a)
        self.check(self.lib.bbg_field_selftest(field, self.FIELD_OPS[op], a.ctypes.data_as(C.c_void_p), bp, out.ctypes.data_as(C.c_void_p), a.shape[0]))
        return out

    def g1_selftest(self, op, p, q):
        p, q = _as_u64(p, (8,)), _as_u64(q, (8,))
        assert p.shape == q.shape
        out = np.zeros_like(p)
        self.check(self.lib.bbg_g1_selftest(self.G1_OPS[op], p.ctypes.data_as(C.c_void_p), q.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), p.shape[0]))
        return out

    # ------------------------------------------------------------------ plumbing
    def check(self, code):
        if code != 0:
            raise BbgError(code, self.lib.bbg_error_string(code).decode())

    def launch_count(self):
        return int(self.lib.bbg_launch_count())

    def set_stream(self, cuda_stream_handle):
        self.check(self.lib.bbg_set_stream(C.c_void_p(cuda_stream_handle)))

    def sync(self):
        self.check(self.lib.bbg_sync())

    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.lib.bbg_dev_alloc(C.byref(p), nbytes))
        return p.value

    def dev_free(self, p):
        self.check(self.lib.bbg_dev_free(C.c_void_p(p)))

    def h2d(self, d_ptr, array):
        a = np.ascontiguousarray(array)
        self.check(self.lib.bbg_copy_h2d(C.c_void_p(d_ptr), a.ctypes.data_as(C.c_void_p), a.nbytes))

    def d2h(self, array, d_ptr):
        assert array.flags["C_CONTIGUOUS"]
        self.check(self.lib.bbg_copy_d2h(array.ctypes.data_as(C.c_void_p), C.c_void_p(d_ptr), array.nbytes))

    def timer_start(self):
        self.check(self.lib.bbg_timer_start())

    def timer_stop(self):
        ms = C.c_float()
        self.check(self.lib.bbg_timer_stop(C.byref(ms)))
        return ms.value

    def microbench(self, mode, iters):
        ops, ms = C.c_double(), C.c_float()
        self.check(self.lib.bbg_microbench(mode, iters, C.byref(ops), C.byref(ms)))
        return ops.value, ms.value

    def profile_enable(self, on=True):
        self.check(self.lib.bbg_profile_enable(1 if on else 0))

    def profile_read(self):
        """{kernel name: (total_ms, launches)} accumulated since profile_enable()."""
        out = {}
        for i in range(self.lib.bbg_profile_count()):
            ms, cnt = C.c_double(), C.c_uint64()
            self.check(self.lib.bbg_profile_read(i, C.byref(ms), C.byref(cnt)))
            if cnt.value:
                out[self.lib.bbg_profile_name(i).decode()] = (ms.value, int(cnt.value))
        return out

    def generate_multiples_dev(self, start_mont, step_mont, d_points, n):
        a, b = _as_u64(start_mont, (4,)), _as_u64(step_mont, (4,))
        self.check(self.lib.bbg_g1_generate_multiples_dev(a.ctypes.data_as(C.c_void_p), b.ctypes.data_as(C.c_void_p), C.c_void_p(d_points), n))

    def generate_pippenger_point_table_dev(self, d_points, d_table, n):
        self.check(self.lib.bbg_generate_pippenger_point_table_dev(C.c_void_p(d_points), C.c_void_p(d_table), n))

    # ------------------------------------------------------------------ NTT
    def ntt(self, op, coeffs, constant=None):
        """In place on a C-contiguous uint64 array of shape (n, 4) (or (batch, n, 4)); returns it."""
        opc = self.OPS[op] if isinstance(op, str) else int(op)
        a = coeffs
        if not (isinstance(a, np.ndarray) and a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"] and a.shape[-1] == 4):
            raise ValueError("coeffs must be a C-contiguous uint64 array (..., n, 4)")
        n = a.shape[-2]
        log_n = n.bit_length() - 1
        if n < 2 or (1 << log_n) != n:
            raise ValueError("domain size must be a power of two >= 2")
        k = None
        if constant is not None:
            k = _as_u64(constant, (4,))
        kp = k.ctypes.data_as(C.c_void_p) if k is not None else None
        if a.ndim == 2:
            self.check(self.lib.bbg_ntt_fr(a.ctypes.data_as(C.c_void_p), log_n, opc, kp))
        else:
            batch = a.shape[0]
            ptrs = (C.c_void_p * batch)(*[a[i].ctypes.data for i in range(batch)])
            self.check(self.lib.bbg_ntt_fr_batched(ptrs, batch, log_n, opc, kp))
        return a

    def compute_lagrange_polynomial_fft(self, log2_src, log2_target):
        out = np.zeros((1 << log2_target, 4), dtype=np.uint64)
        self.check(self.lib.bbg_compute_lagrange_polynomial_fft(out.ctypes.data_as(C.c_void_p), log2_src, log2_target))
        return out

    def ntt_dev(self, op, d_ptr, log_n, batch=1, stride=None, constant=None):
        opc = self.OPS[op] if isinstance(op, str) else int(op)
        k = _as_u64(constant, (4,)) if constant is not None else None
        kp = k.ctypes.data_as(C.c_void_p) if k is not None else None
        self.check(self.lib.bbg_ntt_fr_dev(C.c_void_p(d_ptr), stride or (1 << log_n), batch, log_n, opc, kp))

    # ------------------------------------------------------------------ MSM
    def srs_register(self, table):
        t = _as_u64(table, (8,))
        assert t.shape[0] % 2 == 0
        self.check(self.lib.bbg_srs_register(t.ctypes.data_as(C.c_void_p), t.shape[0] // 2))
        return t  # keep alive: the cache is keyed on this buffer's address

    def srs_unregister(self, table):
        self.check(self.lib.bbg_srs_unregister(table.ctypes.data_as(C.c_void_p)))

    def msm(self, scalars, table, n=None):
        """Normalised Jacobian (12 uint64) of sum scalars[i] * P_i; table = 2n affine entries (n*2, 8)."""
        s = _as_u64(scalars, (4,))
        t = table if (isinstance(table, np.ndarray) and table.dtype == np.uint64 and table.flags["C_CONTIGUOUS"]) else _as_u64(table, (8,))
        n = s.shape[0] if n is None else n
        if t.shape[0] < 2 * n:
            raise ValueError("point table must hold 2n entries")
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.bbg_msm_g1(s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)))
        return out

    def msm_batched(self, scalar_list, table_list):
        b = len(scalar_list)
        ss = [_as_u64(s, (4,)) for s in scalar_list]
        n = ss[0].shape[0]
        if any(s.shape[0] != n for s in ss):
            # the reference prints and returns (scalar_multiplication.cpp:677-685); raise instead of silently doing nothing
            raise ValueError("batched_scalar_multiplications: each scalar mul must be same size")
        sp = (C.c_void_p * b)(*[s.ctypes.data for s in ss])
        tp = (C.c_void_p * b)(*[t.ctypes.data for t in table_list])
        out = np.zeros((b, 12), dtype=np.uint64)
        self.check(self.lib.bbg_msm_g1_batched(sp, tp, n, b, out.ctypes.data_as(C.c_void_p)))
        return out

    def msm_dev(self, d_scalars, d_table, n):
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.bbg_msm_g1_dev(C.c_void_p(d_scalars), C.c_void_p(d_table), n, out.ctypes.data_as(C.c_void_p)))
        return out

    def msm_partial_dev(self, d_scalars, d_table, n):
        out = np.zeros(16, dtype=np.uint64)
        self.check(self.lib.bbg_msm_g1_partial_dev(C.c_void_p(d_scalars), C.c_void_p(d_table), n, out.ctypes.data_as(C.c_void_p)))
        return out

    def msm_partial_dev_launch(self, d_scalars, d_table, n):
        """Queue the MSM on the library's second stream; returns a ticket for msm_partial_finish()."""
        t = C.c_int(-1)
        self.check(self.lib.bbg_msm_g1_partial_dev_launch(C.c_void_p(d_scalars), C.c_void_p(d_table), n, C.byref(t)))
        return t.value

    def msm_partial_finish(self, ticket):
        out = np.zeros(16, dtype=np.uint64)
        self.check(self.lib.bbg_msm_g1_partial_finish(ticket, out.ctypes.data_as(C.c_void_p)))
        return out

    def msm_launch(self, scalars, table, n):
        """bbg_msm_g1_launch: host buffers, MSM queued on the library's second stream; returns a ticket for msm_finish()."""
        s = _as_u64(scalars, (4,))
        t = table if (isinstance(table, np.ndarray) and table.dtype == np.uint64 and table.flags["C_CONTIGUOUS"]) else _as_u64(table, (8,))
        if t.shape[0] < 2 * n:
            raise ValueError("point table must hold 2n entries")
        ticket = C.c_int(-1)
        self.check(self.lib.bbg_msm_g1_launch(s.ctypes.data_as(C.c_void_p), t.ctypes.data_as(C.c_void_p), n, C.byref(ticket)))
        return ticket.value

    def msm_finish(self, ticket):
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.bbg_msm_g1_finish(ticket, out.ctypes.data_as(C.c_void_p)))
        return out

    def fold_partials(self, partials):
        p = _as_u64(partials, (16,))
        out = np.zeros(12, dtype=np.uint64)
        self.check(self.lib.bbg_g1_fold_partials(p.ctypes.data_as(C.c_void_p), p.shape[0], out.ctypes.data_as(C.c_void_p)))
        return out

    def evaluate(self, coeffs, z):
        """polynomial_arithmetic::evaluate: sum_i coeffs[i] z^i (canonical Montgomery limbs)."""
        c, zz = _as_u64(coeffs, (4,)), _as_u64(z, (4,))
        out = np.zeros(4, dtype=np.uint64)
        self.check(self.lib.bbg_fr_evaluate(c.ctypes.data_as(C.c_void_p), c.shape[0], zz.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
        return out

    def divide_by_pseudo_vanishing_polynomial(self, coeffs, log2_src):
        """In place on a C-contiguous (T, 4) uint64 array of coset evaluations; returns it."""
        assert coeffs.dtype == np.uint64 and coeffs.flags["C_CONTIGUOUS"] and coeffs.shape[-1] == 4
        log_t = coeffs.shape[0].bit_length() - 1
        self.check(self.lib.bbg_fr_divide_by_pseudo_vanishing_polynomial(coeffs.ctypes.data_as(C.c_void_p), log2_src, log_t))
        return coeffs

    def compute_kate_opening_coefficients(self, src, z):
        """Returns (quotient coefficients (n, 4), F(z))."""
        c, zz = _as_u64(src, (4,)), _as_u64(z, (4,))
        dest = np.zeros_like(c)
        f = np.zeros(4, dtype=np.uint64)
        self.check(self.lib.bbg_fr_compute_kate_opening_coefficients(c.ctypes.data_as(C.c_void_p), dest.ctypes.data_as(C.c_void_p),
                                                                    zz.ctypes.data_as(C.c_void_p), c.shape[0], f.ctypes.data_as(C.c_void_p)))
        return dest, f

    def domain_lookup_table(self, log2_size):
        """evaluation_domain::compute_lookup_table: (2 * size, 4) uint64, forward rounds then inverse rounds."""
        out = np.zeros((2 << log2_size, 4), dtype=np.uint64)
        self.check(self.lib.bbg_fr_domain_lookup_table(out.ctypes.data_as(C.c_void_p), log2_size))
        return out

    def srs_from_transcript(self, g1_bytes, n):
        """The 2n-entry point table from the raw transcript bytes of points 1 .. n-1 (generator first); the returned array
        is the host buffer the device copy is registered behind (keep it alive, srs_unregister() it when done)."""
        raw = np.frombuffer(bytes(g1_bytes), dtype=np.uint8)
        assert raw.size == 64 * (n - 1)
        table = np.zeros((2 * n, 8), dtype=np.uint64)
        self.check(self.lib.bbg_srs_from_transcript(raw.ctypes.data_as(C.c_void_p) if raw.size else None, n, table.ctypes.data_as(C.c_void_p)))
        return table

    def generate_pippenger_point_table(self, points):
        p = _as_u64(points, (8,))
        n = p.shape[0]
        table = np.zeros((2 * n, 8), dtype=np.uint64)
        self.check(self.lib.bbg_generate_pippenger_point_table(p.ctypes.data_as(C.c_void_p), table.ctypes.data_as(C.c_void_p), n))
        return table


_default = None


def default_library():
    """The in-tree CUDA build on device LOCAL_RANK (or 0); raises BbgError when it is not usable."""
    global _default
    if _default is None:
        _default = Library(device=int(os.environ.get("LOCAL_RANK", "0")))
    return _default
