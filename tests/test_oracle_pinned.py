"""Pins oracle/bb_oracle.c (the plain-C restatement) against
  (1) the reference's own known-answer vectors (tests/golden/reference_kats.json),
  (2) the unmodified reference compiled into oracle/_ref/libbb_ref.so (when present), limb for limb,
  (3) frozen outputs of that compiled reference (tests/golden/ref_vectors.npz, made by make_golden.py).
CPU only.
"""
import json
import os

import numpy as np
import pytest

import helpers as H
from helpers import FQ, FR, ptr, ptr32

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def L(hexes):
    return np.array([int(h, 16) for h in hexes], dtype=np.uint64)


@pytest.fixture(scope="module")
def kats():
    with open(os.path.join(GOLD, "reference_kats.json")) as f:
        return json.load(f)


def binop(name, field, a, b):
    r = np.zeros(4, dtype=np.uint64)
    getattr(H.oracle(), name)(field, ptr(a), ptr(b), ptr(r))
    return r


def unop(name, field, a):
    r = np.zeros(4, dtype=np.uint64)
    getattr(H.oracle(), name)(field, ptr(a), ptr(r))
    return r


@pytest.mark.parametrize("fname,field", [("fq", FQ), ("fr", FR)])
def test_field_kats(kats, fname, field):
    k = kats[fname]
    for v in k["mul"]:
        assert (binop("orc_mul", field, L(v["a"]), L(v["b"])) == L(v["r"])).all()
    for v in k["sqr"]:
        assert (unop("orc_sqr", field, L(v["a"])) == L(v["r"])).all()
    for v in k["add"]:
        assert (binop("orc_add", field, L(v["a"]), L(v["b"])) == L(v["r"])).all()
    for v in k["sub"]:
        assert (binop("orc_sub", field, L(v["a"]), L(v["b"])) == L(v["r"])).all()
    # to/from Montgomery of 1 (test_fq.cpp:135-149, test_fr.cpp:90-104)
    one = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_constant(2 if field == FQ else 6, ptr(one))
    raw1 = H.to_limbs(1)
    assert (unop("orc_to_mont", field, raw1) == one).all()
    assert (unop("orc_from_mont", field, one) == raw1).all()
    assert H.from_limbs(one) == H.R_MONT % H.MODULUS[field]


def to_mont_pt(coords):
    return np.concatenate([unop("orc_to_mont", FQ, L(c)) for c in coords])


def normalized(p):
    out = np.zeros(12, dtype=np.uint64)
    H.oracle().orc_g1_normalize(ptr(np.ascontiguousarray(p)), ptr(out))
    return out


def test_g1_kats(kats):
    lib = H.oracle()
    g = kats["g1"]
    a, b, r = to_mont_pt(g["mixed_add"]["a"]), to_mont_pt(g["mixed_add"]["b"]), to_mont_pt(g["mixed_add"]["r"])
    out = np.zeros(12, dtype=np.uint64)
    lib.orc_g1_mixed_add(ptr(a), ptr(b), ptr(out))
    assert (normalized(out) == normalized(r)).all()

    a, r = to_mont_pt(g["dbl_three_times"]["a"]), to_mont_pt(g["dbl_three_times"]["r"])
    out = a.copy()
    for _ in range(3):
        nxt = np.zeros(12, dtype=np.uint64)
        lib.orc_g1_dbl(ptr(out), ptr(nxt))
        out = nxt
    assert (normalized(out) == normalized(r)).all()

    a, b, r = to_mont_pt(g["add"]["a"]), to_mont_pt(g["add"]["b"]), to_mont_pt(g["add"]["r"])
    out = np.zeros(12, dtype=np.uint64)
    lib.orc_g1_add(ptr(a), ptr(b), ptr(out))
    assert (normalized(out) == normalized(r)).all()

    ge = g["group_exponentiation"]
    gen = np.zeros(8, dtype=np.uint64)
    t = np.zeros(4, dtype=np.uint64)
    lib.orc_constant(11, ptr(t)); gen[:4] = t
    lib.orc_constant(12, ptr(t)); gen[4:] = t
    assert lib.orc_g1_on_curve(ptr(gen)) == 1
    s = unop("orc_to_mont", FR, L(ge["scalar_raw"]))
    res = np.zeros(8, dtype=np.uint64)
    lib.orc_g1_scalar_mul(ptr(gen), ptr(s), ptr(res))
    assert (res == to_mont_pt(ge["r"])).all()
    # group_exponentiation_zero_and_one (test_g1.cpp:305-316)
    lib.orc_g1_scalar_mul(ptr(gen), ptr(np.zeros(4, dtype=np.uint64)), ptr(res))
    assert H.is_infinity(res)
    one = np.zeros(4, dtype=np.uint64)
    lib.orc_constant(6, ptr(one))
    lib.orc_g1_scalar_mul(ptr(gen), ptr(one), ptr(res))
    assert (res == gen).all()


def recover_wnaf(entries, skew, bits):
    n = len(entries)
    v = 0
    for i, e in enumerate(entries):
        d = ((int(e) & 0x0FFFFFFF) << 1) + 1
        if int(e) >> 31:
            d = -d
        v += d << (bits * (n - 1 - i))
    return v - skew


@pytest.mark.parametrize("bits", [2, 3, 5, 8, 13, 16, 19, 22])
def test_wnaf_roundtrip(kats, bits):
    """test_wnaf.cpp:35-91: fixed_wnaf -> recover for 0, 1, 2^64 and random 127-bit scalars."""
    lib = H.oracle()
    entries = (127 + bits - 1) // bits
    rnd = H.splitmix64(1234 + bits, 64).reshape(32, 2).copy()
    rnd[:, 1] &= np.uint64(0x7FFFFFFFFFFFFFFF)
    cases = [np.array([0, 0], dtype=np.uint64), np.array([1, 0], dtype=np.uint64), np.array([0, 1], dtype=np.uint64)] + list(rnd)
    for s in cases:
        w = np.zeros(entries, dtype=np.uint32)
        skew = lib.orc_fixed_wnaf(ptr(np.ascontiguousarray(s)), ptr32(w), 1, bits)
        assert recover_wnaf(w, skew, bits) == int(s[0]) + (int(s[1]) << 64)
    z = kats["wnaf"]["zero_w16"]
    if bits == z["bits"]:
        w = np.zeros(entries, dtype=np.uint32)
        skew = lib.orc_fixed_wnaf(ptr(np.zeros(2, dtype=np.uint64)), ptr32(w), 1, bits)
        assert list(map(int, w)) == z["entries"] and skew == z["skew"]


def test_endo_split_property():
    """test_fr.cpp:239-294: k == k1 - k2*lambda (mod r) from the low 128 bits; both < 2^127."""
    lib = H.oracle()
    lam = np.zeros(4, dtype=np.uint64)
    lib.orc_constant(7, ptr(lam))
    lam = H.unmont(H.from_limbs(lam))
    assert pow(lam, 3, H.FR_MODULUS) == 1 and lam != 1
    ks = H.random_field_raw(77, 2000, FR)
    ks[0] = H.to_limbs(1)
    ks[1] = H.to_limbs(0)
    ks[2] = H.to_limbs(H.FR_MODULUS - 1)
    for k in ks:
        out = np.zeros(4, dtype=np.uint64)
        lib.orc_split_endo(ptr(np.ascontiguousarray(k)), ptr(out))
        k1 = int(out[0]) | (int(out[1]) << 64)
        k2 = int(out[2]) | (int(out[3]) << 64)
        assert k1 < (1 << 127) and k2 < (1 << 127)
        assert (k1 - k2 * lam) % H.FR_MODULUS == H.from_limbs(k)


# ------------------------------------------------------------------ differential vs compiled reference
needs_ref = pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref/libbb_ref.so not built")


@needs_ref
def test_constants_match_reference():
    o, r = H.oracle(), H.ref()
    for which in range(14):
        a, b = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        o.orc_constant(which, ptr(a))
        r.ref_constant(which, ptr(b))
        assert (a == b).all(), which


@needs_ref
@pytest.mark.parametrize("field", [FQ, FR])
def test_field_ops_match_reference(field):
    o, r = H.oracle(), H.ref()
    n = 20000
    a = H.random_field_raw(1, n, field)
    b = H.random_field_raw(2, n, field)
    # also lazily-reduced inputs in [0, 2p): add p to a third of them (SURVEY §8 note 2)
    p = H.MODULUS[field]
    for i in range(0, n, 3):
        a[i] = H.to_limbs(H.from_limbs(a[i]) + p)
    got, exp = np.zeros_like(a), np.zeros_like(a)
    o.orc_mul_n(field, ptr(a), ptr(b), ptr(got), n)
    (r.ref_fq_mul_n if field == FQ else r.ref_fr_mul_n)(ptr(a), ptr(b), ptr(exp), n)
    assert (got == exp).all()
    pre = "ref_fq_" if field == FQ else "ref_fr_"
    for i in range(300):
        for oname, rname in (("orc_mul_coarse", "mul_coarse"), ("orc_add", "add"), ("orc_sub", "sub")):
            x, y = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
            getattr(o, oname)(field, ptr(a[i]), ptr(b[i]), ptr(x))
            getattr(r, pre + rname)(ptr(a[i]), ptr(b[i]), ptr(y))
            assert (x == y).all(), (oname, i)
        x, y = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        o.orc_sqr(field, ptr(a[i]), ptr(x)); getattr(r, pre + "sqr")(ptr(a[i]), ptr(y))
        assert (x == y).all()
    for i in range(5):
        x, y = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
        o.orc_invert(field, ptr(b[i]), ptr(x)); getattr(r, pre + "invert")(ptr(b[i]), ptr(y))
        assert (x == y).all()


@needs_ref
def test_split_and_wnaf_match_reference():
    o, r = H.oracle(), H.ref()
    n = 5000
    k = H.random_field_raw(5, n, FR)
    k[0] = 0
    k[1] = H.to_limbs(1)
    got, exp = np.zeros_like(k), np.zeros_like(k)
    for i in range(n):
        o.orc_split_endo(ptr(k[i]), ptr(got[i]))
    r.ref_split_endo_n(ptr(k), ptr(exp), n)
    assert (got == exp).all()
    for bits in (4, 13, 16, 19, 22):
        entries = (127 + bits - 1) // bits
        for i in range(200):
            s = np.ascontiguousarray(exp[i, :2])
            a, b = np.zeros(entries * 3, dtype=np.uint32), np.zeros(entries * 3, dtype=np.uint32)
            sa = o.orc_fixed_wnaf(ptr(s), ptr32(a), 3, bits)
            sb = r.ref_fixed_wnaf(ptr(s), ptr32(b), 3, bits)
            assert sa == sb and (a == b).all()


@needs_ref
def test_g1_ops_match_reference_limb_for_limb():
    o, r = H.oracle(), H.ref()
    pts = H.arithmetic_progression_points(12345, 777, 40)
    one = np.zeros(4, dtype=np.uint64)
    o.orc_constant(2, ptr(one))
    acc_o = np.concatenate([pts[0], one])
    acc_r = acc_o.copy()
    for i in range(1, 40):
        a, b = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
        o.orc_g1_mixed_add(ptr(acc_o), ptr(pts[i]), ptr(a))
        r.ref_g1_mixed_add(ptr(acc_r), ptr(pts[i]), ptr(b))
        assert (a == b).all()
        d1, d2 = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
        o.orc_g1_dbl(ptr(a), ptr(d1)); r.ref_g1_dbl(ptr(b), ptr(d2))
        assert (d1 == d2).all()
        s1, s2 = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
        o.orc_g1_add(ptr(a), ptr(d1), ptr(s1)); r.ref_g1_add(ptr(b), ptr(d2), ptr(s2))
        assert (s1 == s2).all()
        acc_o, acc_r = s1, s2
    # exception paths: P + P -> dbl, P + (-P) -> infinity, infinity + P (test_g1.cpp:124-241)
    p = np.concatenate([pts[3], one])
    a, b = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    o.orc_g1_mixed_add(ptr(p), ptr(pts[3]), ptr(a)); r.ref_g1_mixed_add(ptr(p), ptr(pts[3]), ptr(b))
    assert (a == b).all()
    neg = pts[3].copy()
    o.orc_neg(FQ, ptr(pts[3][4:].copy()), ptr(one))
    neg[4:] = one
    o.orc_g1_mixed_add(ptr(p), ptr(neg), ptr(a)); r.ref_g1_mixed_add(ptr(p), ptr(neg), ptr(b))
    assert H.is_infinity(a) and H.is_infinity(b)
    o.orc_g1_mixed_add(ptr(a), ptr(pts[5]), ptr(a)); r.ref_g1_mixed_add(ptr(b), ptr(pts[5]), ptr(b))
    assert (a == b).all() and (a[:8] == pts[5]).all()
    # generator progression agrees with the reference's own construction
    s0, st = H.to_limbs(H.mont(12345)), H.to_limbs(H.mont(777))
    rp = np.zeros((40, 8), dtype=np.uint64)
    r.ref_g1_arith_progression(ptr(s0), ptr(st), ptr(rp), 40)
    assert (rp == pts).all()


@needs_ref
@pytest.mark.parametrize("n,width", [(1, 0), (2, 0), (7, 0), (64, 0), (500, 0), (1000, 5), (3000, 0)])
def test_pippenger_matches_reference(n, width):
    """Same un-normalised Jacobian limbs as reference pippenger (same algorithm, same order)."""
    o, r = H.oracle(), H.ref()
    table, _, _ = H.generator_multiples_table(900 + n, n)
    rt = np.zeros_like(table)
    r.ref_generate_pippenger_point_table(ptr(np.ascontiguousarray(table[0::2])), ptr(rt), n)
    assert (rt == table).all()
    sc = H.random_scalars_mont(33 + n, n)
    if n > 4:
        sc[1] = 0                      # zero scalar
        sc[2] = sc[3]                  # repeated scalar
        sc[4] = H.to_limbs(H.from_limbs(sc[4]) + H.FR_MODULUS)  # lazily reduced input
    a, b = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    o.orc_pippenger(ptr(sc), ptr(table), n, width, ptr(a))
    r.ref_pippenger(ptr(sc), ptr(table), n, width, ptr(b))
    assert (a == b).all()
    assert o.orc_get_optimal_bucket_width(n) == r.ref_get_optimal_bucket_width(n)


@needs_ref
def test_pippenger_edge_cases_match_reference():
    o, r = H.oracle(), H.ref()
    table, _, _ = H.generator_multiples_table(5, 16)
    a, b = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64)
    o.orc_pippenger(ptr(np.zeros((0, 4), dtype=np.uint64)), ptr(table), 0, 0, ptr(a))
    r.ref_pippenger(ptr(np.zeros((1, 4), dtype=np.uint64)), ptr(table), 0, 0, ptr(b))
    assert H.is_infinity(a) and H.is_infinity(b)
    z = np.zeros((16, 4), dtype=np.uint64)  # all-zero scalars -> infinity (test_scalar_multiplication.cpp:140-162)
    o.orc_pippenger(ptr(z), ptr(table), 16, 0, ptr(a))
    r.ref_pippenger(ptr(z), ptr(table), 16, 0, ptr(b))
    assert H.is_infinity(a) and H.is_infinity(b)


@needs_ref
@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 8, 11])
def test_ntt_matches_reference(log_n):
    o, r = H.oracle(), H.ref()
    n = 1 << log_n
    od = H.OracleDomain(n)
    rd = r.ref_domain_new(n)
    for which in range(6):
        a = np.zeros(4, dtype=np.uint64)
        r.ref_domain_constant(rd, which, ptr(a))
        assert (a == od.constant(which)).all()
    x = H.random_scalars_mont(100 + log_n, n)
    x[0] = H.to_limbs(H.from_limbs(x[0]) + H.FR_MODULUS)  # [0,2p) input
    k = H.random_scalars_mont(7, 1)[0]
    buf = r.ref_aligned_alloc(32 * n)
    view = np.ctypeslib.as_array((H.C.c_uint64 * (4 * n)).from_address(buf)).reshape(n, 4)
    for name, op in H.NTT_OPS.items():
        view[:] = x
        r.ref_ntt(rd, op, buf, ptr(k))
        got = od.ntt(op, x, k)
        assert (got == view).all(), name
        assert all(H.from_limbs(v) < H.FR_MODULUS for v in got[:8])
    r.ref_aligned_free(buf)
    r.ref_domain_free(rd)


# ------------------------------------------------------------------ frozen reference outputs
def test_against_frozen_reference_vectors():
    path = os.path.join(GOLD, "ref_vectors.npz")
    g = np.load(path)
    o = H.oracle()
    n = int(g["msm_n"])
    a = np.zeros(12, dtype=np.uint64)
    o.orc_msm_normalized(ptr(np.ascontiguousarray(g["msm_scalars"])), ptr(np.ascontiguousarray(g["msm_table"])), n, ptr(a))
    assert (a == g["msm_out_normalized"]).all()
    x = g["ntt_in"]
    k = np.ascontiguousarray(g["ntt_constant"])
    od = H.OracleDomain(x.shape[0])
    for name, op in H.NTT_OPS.items():
        assert (od.ntt(op, x, k) == g["ntt_" + name]).all(), name
    got = np.zeros_like(g["fr_mul_a"])
    o.orc_mul_n(FR, ptr(np.ascontiguousarray(g["fr_mul_a"])), ptr(np.ascontiguousarray(g["fr_mul_b"])), ptr(got), got.shape[0])
    assert (got == g["fr_mul_r"]).all()
    got = np.zeros_like(g["fq_mul_a"])
    o.orc_mul_n(FQ, ptr(np.ascontiguousarray(g["fq_mul_a"])), ptr(np.ascontiguousarray(g["fq_mul_b"])), ptr(got), got.shape[0])
    assert (got == g["fq_mul_r"]).all()


def test_fft_against_horner_evaluation():
    """test_polynomial_arithmetic.cpp:31-56: n=16 fft equals evaluate() at w^i, raw limbs (canonical outputs)."""
    o = H.oracle()
    n = 16
    od = H.OracleDomain(n)
    x = H.random_scalars_mont(3, n)
    y = od.ntt(0, x)
    root = od.constant(0)
    w = np.zeros(4, dtype=np.uint64)
    o.orc_constant(6, ptr(w))  # one
    for i in range(n):
        e = np.zeros(4, dtype=np.uint64)
        o.orc_poly_evaluate(ptr(x), ptr(w), n, ptr(e))
        assert (e == y[i]).all()
        nw = np.zeros(4, dtype=np.uint64)
        o.orc_mul(FR, ptr(w), ptr(root), ptr(nw))
        w = nw


def test_ntt_roundtrips():
    """fft∘ifft and coset round trips (test_polynomial_arithmetic.cpp:58-128)."""
    for n in (2, 4, 256, 1 << 12):
        od = H.OracleDomain(n)
        x = H.random_scalars_mont(n, n)
        assert (od.ntt(1, od.ntt(0, x)) == x).all()
        assert (od.ntt(3, od.ntt(2, x)) == x).all()


@needs_ref
@pytest.mark.parametrize("log_src,log_tgt", [(2, 3), (4, 5), (8, 9), (8, 10)])
def test_lagrange_fft_matches_reference(log_src, log_tgt):
    o, r = H.oracle(), H.ref()
    t = 1 << log_tgt
    exp_ptr = r.ref_aligned_alloc(32 * t)
    exp = np.ctypeslib.as_array((H.C.c_uint64 * (4 * t)).from_address(exp_ptr)).reshape(t, 4)
    r.ref_compute_lagrange_polynomial_fft(exp_ptr, 1 << log_src, t)
    got = np.zeros((t, 4), dtype=np.uint64)
    o.orc_compute_lagrange_polynomial_fft(ptr(got), log_src, log_tgt)
    assert (got == exp).all()
    r.ref_aligned_free(exp_ptr)
