"""Shared test infrastructure: ctypes loaders for the two CPU checkers and a seeded input generator.

TEST INFRASTRUCTURE ONLY.  `oracle/libbb_oracle.so` is the plain-C restatement, `oracle/_ref/libbb_ref.so`
is the unmodified reference compiled by oracle/Makefile (present in the build container and, as a
prebuilt file, on the GPU box).  Field elements are numpy uint64 arrays of shape (..., 4).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libbb_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libbb_ref.so")

FQ, FR = 0, 1
FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
MODULUS = {FQ: FQ_MODULUS, FR: FR_MODULUS}
R_MONT = 1 << 256

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)


def ptr(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u64p)


def ptr32(a):
    assert a.dtype == np.uint32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(u32p)


def build_oracle():
    if not os.path.exists(ORACLE_SO) or os.path.getmtime(ORACLE_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "bb_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "oracle"], stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        build_oracle()
        lib = C.CDLL(ORACLE_SO)
        lib.orc_domain_new.restype = C.c_void_p
        lib.orc_domain_new.argtypes = [C.c_size_t]
        lib.orc_domain_free.argtypes = [C.c_void_p]
        lib.orc_domain_constant.argtypes = [C.c_void_p, C.c_int, u64p]
        lib.orc_ntt.argtypes = [C.c_void_p, C.c_int, u64p, u64p]
        lib.orc_pippenger.argtypes = [u64p, u64p, C.c_size_t, C.c_size_t, u64p]
        lib.orc_msm_normalized.argtypes = [u64p, u64p, C.c_size_t, u64p]
        lib.orc_generate_pippenger_point_table.argtypes = [u64p, u64p, C.c_size_t]
        lib.orc_get_optimal_bucket_width.restype = C.c_size_t
        lib.orc_get_optimal_bucket_width.argtypes = [C.c_size_t]
        lib.orc_fixed_wnaf.argtypes = [u64p, u32p, C.c_size_t, C.c_size_t]
        lib.orc_mul_n.argtypes = [C.c_int, u64p, u64p, u64p, C.c_size_t]
        lib.orc_pow_small.argtypes = [C.c_int, u64p, C.c_uint64, u64p]
        lib.orc_g1_batch_normalize.argtypes = [u64p, C.c_size_t]
        lib.orc_poly_evaluate.argtypes = [u64p, u64p, C.c_size_t, u64p]
        lib.orc_g1_arith_progression.argtypes = [u64p, u64p, u64p, C.c_size_t]
        lib.orc_compute_lagrange_polynomial_fft.argtypes = [u64p, C.c_size_t, C.c_size_t]
        _oracle = lib
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def require_ref():
    """For `-m gpu` tests: the compiled reference is a prebuilt artefact that travels to the GPU box (oracle/_ref is
    git-ignored, not gpurun-ignored).  Its absence there is a broken checkout, not a reason to skip: fail loudly, so a
    green run can never mean "the reference comparisons were silently left out"."""
    if not have_ref():
        import pytest

        pytest.fail("oracle/_ref/libbb_ref.so is missing: build it where /root/reference exists "
                    "(python -c 'import __graft_entry__ as g; g.build()') — GPU parity tests do not skip")
    return ref()


def require_built(*names):
    """Same rule for the prebuilt harness binaries under build/ (tests/cpp/Makefile)."""
    missing = [n for n in names if not os.path.exists(os.path.join(ROOT, "build", n))]
    if missing:
        import pytest

        pytest.fail("build/%s missing: run make -C tests/cpp where /root/reference exists — GPU tests do not skip" % ", build/".join(missing))


def ref():
    """The compiled reference (None when oracle/_ref was never built)."""
    global _ref
    if _ref is None:
        if not have_ref():
            return None
        lib = C.CDLL(REF_SO)
        lib.ref_domain_new.restype = C.c_void_p
        lib.ref_domain_new.argtypes = [C.c_size_t]
        lib.ref_domain_free.argtypes = [C.c_void_p]
        lib.ref_domain_constant.argtypes = [C.c_void_p, C.c_int, u64p]
        lib.ref_domain_num_threads.restype = C.c_size_t
        lib.ref_domain_num_threads.argtypes = [C.c_void_p]
        lib.ref_ntt.argtypes = [C.c_void_p, C.c_int, C.c_void_p, u64p]
        lib.ref_pippenger.argtypes = [u64p, u64p, C.c_size_t, C.c_size_t, u64p]
        lib.ref_pippenger_inplace.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, u64p]
        lib.ref_batched_scalar_multiplications.argtypes = [C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_size_t, u64p]
        lib.ref_generate_pippenger_point_table.argtypes = [u64p, u64p, C.c_size_t]
        lib.ref_get_optimal_bucket_width.restype = C.c_size_t
        lib.ref_get_optimal_bucket_width.argtypes = [C.c_size_t]
        lib.ref_fixed_wnaf.argtypes = [u64p, u32p, C.c_size_t, C.c_size_t]
        lib.ref_g1_arith_progression.argtypes = [u64p, u64p, u64p, C.c_size_t]
        lib.ref_g1_batch_normalize.argtypes = [u64p, C.c_size_t]
        lib.ref_poly_evaluate.argtypes = [u64p, u64p, C.c_size_t, u64p]
        if hasattr(lib, "ref_divide_by_pseudo_vanishing_polynomial"):
            lib.ref_divide_by_pseudo_vanishing_polynomial.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
            lib.ref_compute_kate_opening_coefficients.argtypes = [C.c_void_p, C.c_void_p, u64p, C.c_size_t, u64p]
        lib.ref_r_inv.restype = C.c_uint64
        lib.ref_compute_lagrange_polynomial_fft.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t]
        lib.ref_aligned_alloc.restype = C.c_void_p
        lib.ref_aligned_alloc.argtypes = [C.c_size_t]
        lib.ref_aligned_free.argtypes = [C.c_void_p]
        for name in ("ref_fq_mul_n", "ref_fr_mul_n"):
            getattr(lib, name).argtypes = [u64p, u64p, u64p, C.c_size_t]
        for name in ("ref_fr_to_mont_n", "ref_fr_from_mont_n", "ref_split_endo_n"):
            getattr(lib, name).argtypes = [u64p, u64p, C.c_size_t]
        _ref = lib
    return _ref


# ---------------------------------------------------------------- integers <-> limbs, seeded generators
import sys  # noqa: E402

if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from barretenberg_b200.synthetic import from_limbs, splitmix64, to_limbs  # noqa: E402,F401
from barretenberg_b200.synthetic import random_field as _random_field  # noqa: E402


def limbs_array(ints):
    out = np.empty((len(ints), 4), dtype=np.uint64)
    for i, x in enumerate(ints):
        out[i] = to_limbs(x)
    return out


def mont(x, field=FR):
    return (x * R_MONT) % MODULUS[field]


def unmont(x, field=FR):
    return (x * pow(R_MONT, -1, MODULUS[field])) % MODULUS[field]


def random_field_raw(seed, n, field=FR):
    """n uniform-ish canonical values < p as (n,4) uint64 limbs."""
    return _random_field(seed, n, MODULUS[field])


def random_scalars_mont(seed, n):
    """Seeded Fr elements in canonical Montgomery form (uniform residues: Montgomery form of a uniform value)."""
    return random_field_raw(seed, n, FR)


def generator_multiples_table(seed, n):
    """Pippenger 2n table for points P_i = (a0 + i*d) * G (seeded a0, d), via the oracle.
    Returns (table (2n,8) uint64, a0, d) with a0, d python ints (non-Montgomery)."""
    lib = oracle()
    a0 = int(splitmix64(seed, 1)[0]) | 1
    d = int(splitmix64(seed + 1, 1)[0]) | 1
    pts = arithmetic_progression_points(a0, d, n)
    table = np.zeros((2 * n, 8), dtype=np.uint64)
    if n:
        lib.orc_generate_pippenger_point_table(ptr(pts), ptr(table), n)
    return table, a0, d


def arithmetic_progression_points(a0, d, n):
    """(n,8) affine points (a0 + i d) G using oracle mixed adds + one batch normalisation."""
    out = np.zeros((n, 8), dtype=np.uint64)
    if n:
        oracle().orc_g1_arith_progression(ptr(to_limbs(mont(a0))), ptr(to_limbs(mont(d))), ptr(out), n)
    return out


def closed_form_msm(scalars, a0, d):
    """Normalised Jacobian of sum_i k_i (a0 + i d) G = ((sum_i k_i (a0 + i d)) mod r) G  (SURVEY.md §8c-3):
    one Fr dot product and one oracle scalar multiplication — no CPU MSM needed at 2^20 / 2^26."""
    lib = oracle()
    from barretenberg_b200.synthetic import dot_mod_r

    s = dot_mod_r(np.ascontiguousarray(scalars), a0, d)
    gen = np.zeros(8, dtype=np.uint64)
    tmp = np.zeros(4, dtype=np.uint64)
    lib.orc_constant(11, ptr(tmp)); gen[:4] = tmp
    lib.orc_constant(12, ptr(tmp)); gen[4:] = tmp
    res = np.zeros(8, dtype=np.uint64)
    lib.orc_g1_scalar_mul(ptr(gen), ptr(to_limbs(mont(s))), ptr(res))
    out = np.zeros(12, dtype=np.uint64)
    out[:8] = res
    lib.orc_constant(2, ptr(tmp)); out[8:] = tmp
    return out


def is_infinity(pt):
    """pt: affine (8) or Jacobian (12) limbs."""
    return bool(int(pt[7]) >> 63)


def oracle_msm(scalars, table):
    """Normalised Jacobian (12 limbs) of sum scalars[i]*P_i via the C oracle."""
    out = np.zeros(12, dtype=np.uint64)
    n = scalars.shape[0]
    oracle().orc_msm_normalized(ptr(np.ascontiguousarray(scalars)), ptr(np.ascontiguousarray(table)), n, ptr(out))
    return out


class OracleDomain:
    def __init__(self, n):
        self.n = n
        self.h = oracle().orc_domain_new(n)

    def ntt(self, op, coeffs, constant=None):
        c = np.ascontiguousarray(coeffs).copy()
        k = ptr(np.ascontiguousarray(constant)) if constant is not None else None
        oracle().orc_ntt(self.h, op, ptr(c), k)
        return c

    def constant(self, which):
        r = np.zeros(4, dtype=np.uint64)
        oracle().orc_domain_constant(self.h, which, ptr(r))
        return r

    def __del__(self):
        try:
            oracle().orc_domain_free(self.h)
        except Exception:
            pass


NTT_OPS = {"fft": 0, "ifft": 1, "coset_fft": 2, "coset_ifft": 3, "fft_with_constant": 4,
           "ifft_with_constant": 5, "coset_fft_with_constant": 6}


def expected_domain_lookup_table(log2_size):
    """evaluation_domain.cpp:33-54, :172-178 as python integers: canonical Montgomery limbs, unused slots zero."""
    size = 1 << log2_size
    od = OracleDomain(size)
    p = FR_MODULUS
    out = np.zeros((2 * size, 4), dtype=np.uint64)
    for half, which in ((0, 0), (1, 1)):
        root = unmont(from_limbs(od.constant(which)))
        off = half * size
        for i in range(log2_size - 1):
            m = 1 << (i + 1)
            rr = pow(root, size // (2 * m), p)
            cur = 1
            for j in range(m):
                out[off + m - 2 + j] = to_limbs(mont(cur))
                cur = cur * rr % p
    return out


def transcript_g1_bytes(points_mont):
    """io.hpp:76-98 inverted: affine points (k, 8) uint64 Montgomery limbs -> raw transcript bytes (x then y, limbs in
    little-endian order, big-endian bytes inside a limb, plain values)."""
    out = bytearray()
    for pt in points_mont:
        for c in (pt[:4], pt[4:]):
            v = unmont(from_limbs(c), FQ)
            for k in range(4):
                out += ((v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF).to_bytes(8, "big")
    return bytes(out)
