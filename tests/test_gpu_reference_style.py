"""The reference's own hot-path unit tests, restated against the drop-in interface (same names, same checks):
test/test_polynomial_arithmetic.cpp:31-175 and test/test_scalar_multiplication.cpp:72-324.  The reference draws
its inputs from getentropy; here they are seeded."""
import numpy as np
import pytest

import barretenberg_b200 as bb
from barretenberg_b200 import polynomial_arithmetic as pa
from barretenberg_b200 import scalar_multiplication as sm
import helpers as H
from helpers import FR, ptr

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    return bb.default_library()


def fr_add(a, b):
    r = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_add(FR, ptr(np.ascontiguousarray(a)), ptr(np.ascontiguousarray(b)), ptr(r))
    return r


def fr_mul(a, b):
    r = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_mul(FR, ptr(np.ascontiguousarray(a)), ptr(np.ascontiguousarray(b)), ptr(r))
    return r


def test_fft_with_small_degree(lib):
    """polynomials.fft_with_small_degree (:31-56): fft output == evaluate(poly, w^i), raw limbs."""
    n = 16
    poly = H.random_scalars_mont(1, n)
    fft_transform = poly.copy()
    domain = bb.EvaluationDomain(n, library=lib)
    domain.compute_lookup_table()
    pa.fft(fft_transform, domain)
    root = H.OracleDomain(n).constant(0)
    work_root = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_constant(6, ptr(work_root))
    for i in range(n):
        expected = np.zeros(4, dtype=np.uint64)
        H.oracle().orc_poly_evaluate(ptr(poly), ptr(work_root), n, ptr(expected))
        assert (fft_transform[i] == expected).all()
        work_root = fr_mul(work_root, root)


@pytest.mark.parametrize("n", [256, 1 << 14])
def test_basic_fft_and_fft_coset_ifft_consistency(lib, n):
    """polynomials.basic_fft (:58-80), fft_ifft_consistency (:82-101), fft_coset_ifft_consistency (:104-128)."""
    domain = bb.EvaluationDomain(n, library=lib)
    expected = H.random_scalars_mont(2, n)
    result = expected.copy()
    pa.fft(result, domain)
    pa.ifft(result, domain)
    assert (result == expected).all()
    pa.coset_fft(result, domain)
    pa.coset_ifft(result, domain)
    assert (result == expected).all()


@pytest.mark.parametrize("n", [2, 4, 8, 1 << 11])
def test_fft_coset_ifft_cross_consistency(lib, n):
    """polynomials.fft_coset_ifft_cross_consistency (:130-175): the n, 2n and 4n coset evaluations of one
    polynomial agree on the shared points."""
    base = H.random_scalars_mont(3, n)
    expected = np.stack([fr_add(fr_add(base[i], base[i]), base[i]) for i in range(n)])
    poly_a = base.copy()
    poly_b = np.zeros((2 * n, 4), dtype=np.uint64)
    poly_c = np.zeros((4 * n, 4), dtype=np.uint64)
    poly_b[:n] = base
    poly_c[:n] = base
    small, mid, large = (bb.EvaluationDomain(m, library=lib) for m in (n, 2 * n, 4 * n))
    pa.coset_fft(poly_a, small)
    pa.coset_fft(poly_b, mid)
    pa.coset_fft(poly_c, large)
    for i in range(n):
        poly_a[i] = fr_add(fr_add(poly_a[i], poly_c[4 * i]), poly_b[2 * i])
    pa.coset_ifft(poly_a, small)
    assert (poly_a == expected).all()


def naive_sum(scalars, points):
    """sum of group_exponentiation(points[i], scalars[i]), normalised — the reference tests' expected value."""
    o = H.oracle()
    one = np.zeros(4, dtype=np.uint64)
    o.orc_constant(2, ptr(one))
    acc = np.zeros(12, dtype=np.uint64)
    acc[7] = np.uint64(1) << np.uint64(63)
    for s, p in zip(scalars, points):
        term = np.zeros(8, dtype=np.uint64)
        o.orc_g1_scalar_mul(ptr(np.ascontiguousarray(p)), ptr(np.ascontiguousarray(s)), ptr(term))
        if H.is_infinity(term):
            continue
        nxt = np.zeros(12, dtype=np.uint64)
        o.orc_g1_mixed_add(ptr(acc), ptr(term), ptr(nxt))
        acc = nxt
    out = np.zeros(12, dtype=np.uint64)
    if H.is_infinity(acc):
        return acc
    o.orc_g1_normalize(ptr(acc), ptr(out))
    return out


@pytest.mark.parametrize("num_points", [1, 2000])
def test_pippenger(lib, num_points):
    """scalar_multiplication.pippenger / pippenger_one (:72-138): vs the sum of naive scalar multiplications."""
    points = H.arithmetic_progression_points(0xABCDEF, 0x1357, num_points)
    scalars = H.random_scalars_mont(4, num_points)
    table = sm.generate_pippenger_point_table(points, library=lib)
    result = sm.pippenger(scalars, table, num_points, library=lib)
    assert (result == naive_sum(scalars, points)).all()


def test_pippenger_zero_points_and_mul_by_zero(lib):
    """:140-162: n = 0 and a single zero scalar both give the point at infinity."""
    points = H.arithmetic_progression_points(5, 7, 4)
    table = sm.generate_pippenger_point_table(points, library=lib)
    assert H.is_infinity(sm.pippenger(np.zeros((0, 4), dtype=np.uint64), table, 0, library=lib))
    assert H.is_infinity(sm.pippenger(np.zeros((1, 4), dtype=np.uint64), table, 1, library=lib))


def test_batched_scalar_multiplications(lib):
    """:286-324: 5 batches of 2000, outputs limb-equal (x, y, z) to normalize(pippenger()) per batch."""
    num_exponentiations, num_points = 5, 2000
    points = H.arithmetic_progression_points(0x777, 0x99, num_points)
    table = sm.generate_pippenger_point_table(points, library=lib)
    states = [sm.MultiplicationState(points=table, scalars=H.random_scalars_mont(10 + i, num_points), num_elements=num_points)
              for i in range(num_exponentiations)]
    sm.batched_scalar_multiplications(states, num_exponentiations, library=lib)
    for st in states:
        expected = H.oracle_msm(st.scalars, table)  # the oracle's literal pippenger + normalize
        assert (st.output == expected).all()
        assert (st.output[8:] == expected[8:]).all()  # z = fq::one
