"""CPU tests of the kernel logic: the SAME kernel sources as libbbgpu.so compiled with g++ -DBBG_EMULATE
(tests/emul/cuda_emul.h runs each CUDA thread as an OS thread) and driven through the same C ABI, compared
with the oracle.  This is test infrastructure for a GPU-less box — the product never loads this library."""
import os
import subprocess

import numpy as np
import pytest

import barretenberg_b200 as bb
import helpers as H

ROOT = H.ROOT
EMUL_SO = os.path.join(ROOT, "tests", "emul", "libbbgpu_emul.so")


@pytest.fixture(scope="module")
def emu():
    src = os.path.join(ROOT, "barretenberg_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(src, f)) for f in os.listdir(src) if f.endswith((".cu", ".cuh", ".h")))
    if not os.path.exists(EMUL_SO) or os.path.getmtime(EMUL_SO) < newest:
        env = dict(os.environ)
        env.pop("CXX", None)
        subprocess.check_call(["make", "-C", src, "emul"], env=env, stdout=subprocess.DEVNULL)
    return bb.Library(EMUL_SO)


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 7, 11, 12, 13, 15])
def test_ntt_all_ops_vs_oracle(emu, log_n):
    n = 1 << log_n
    od = H.OracleDomain(n)
    x = H.random_scalars_mont(100 + log_n, n)
    x[0] = H.to_limbs(H.from_limbs(x[0]) + H.FR_MODULUS)  # lazily reduced input (SURVEY §8 note 2)
    k = H.random_scalars_mont(7, 1)[0]
    for name, op in H.NTT_OPS.items():
        got = emu.ntt(name, x.copy(), k)
        assert (got == od.ntt(op, x, k)).all(), name


def test_ntt_batched_and_golden(emu):
    g = np.load(os.path.join(ROOT, "tests", "golden", "ref_vectors.npz"))
    x = g["ntt_in"]
    for name in H.NTT_OPS:
        batch = np.stack([x, x[::-1].copy(), x]).copy()
        emu.ntt(name, batch, g["ntt_constant"])
        assert (batch[0] == g["ntt_" + name]).all() and (batch[2] == g["ntt_" + name]).all(), name


@pytest.mark.slow
def test_ntt_two_pass_every_subtransform_length(emu):
    """2^17 and 2^18 exercise sub-transform lengths 8 and 9 in both passes (2^19..2^22 run on the GPU only)."""
    for log_n in (14, 16, 17, 18):
        n = 1 << log_n
        od = H.OracleDomain(n)
        x = H.random_scalars_mont(log_n, n)
        for name in ("coset_fft", "coset_ifft"):
            assert (emu.ntt(name, x.copy()) == od.ntt(H.NTT_OPS[name], x)).all(), (log_n, name)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 300, 1000, 3000])
def test_msm_vs_oracle(emu, n):
    table, _, _ = H.generator_multiples_table(900 + n, max(n, 1))
    sc = H.random_scalars_mont(33 + n, n)
    if n > 4:
        sc[1] = 0
        sc[2] = sc[3]
        sc[4] = H.to_limbs(H.from_limbs(sc[4]) + H.FR_MODULUS)
    got = emu.msm(sc, table, n)
    exp = H.oracle_msm(sc, table)
    if H.is_infinity(exp):
        assert H.is_infinity(got)
    else:
        assert (got == exp).all()


def test_msm_edge_cases(emu):
    n = 200
    table, a0, d = H.generator_multiples_table(11, n)
    one = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_constant(6, H.ptr(one))
    # all-zero scalars -> infinity (test_scalar_multiplication.cpp:140-162)
    assert H.is_infinity(emu.msm(np.zeros((n, 4), dtype=np.uint64), table))
    # all scalars equal to one: a single giant bucket spanning many slices
    ones = np.tile(one, (n, 1))
    assert (emu.msm(ones, table) == H.oracle_msm(ones, table)).all()
    # P and -P pairs with equal scalars cancel: table of [P, -P, P, -P, ...]
    pts = np.ascontiguousarray(table[0::2]).copy()
    neg = np.zeros(4, dtype=np.uint64)
    for i in range(1, n, 2):
        pts[i, :4] = pts[i - 1, :4]
        H.oracle().orc_neg(H.FQ, H.ptr(pts[i - 1, 4:].copy()), H.ptr(neg))
        pts[i, 4:] = neg
    t2 = emu.generate_pippenger_point_table(pts)
    ref_t2 = np.zeros_like(t2)
    H.oracle().orc_generate_pippenger_point_table(H.ptr(pts), H.ptr(ref_t2), n)
    assert (t2 == ref_t2).all()
    sc = H.random_scalars_mont(5, n)
    sc[1::2] = sc[0::2]
    assert H.is_infinity(emu.msm(sc, t2))
    # repeated points with different scalars (P + P doubling path inside a bucket is possible)
    pts2 = np.tile(pts[0], (n, 1))
    t3 = emu.generate_pippenger_point_table(pts2)
    sc = H.random_scalars_mont(6, n)
    sc[:50] = sc[0]
    assert (emu.msm(sc, t3) == H.oracle_msm(sc, t3)).all()
    # closed form on generator multiples
    sc = H.random_scalars_mont(8, n)
    assert (emu.msm(sc, table) == H.closed_form_msm(sc, a0, d)).all()


def test_msm_giant_bucket_block_fixup(emu):
    """One digit value shared by every scalar -> a bucket spanning > 16 slices -> the block-reduction fix-up path."""
    n = 1500
    table, a0, d = H.generator_multiples_table(13, n)
    sc = np.tile(H.random_scalars_mont(99, 1)[0], (n, 1))
    assert (emu.msm(sc, table) == H.closed_form_msm(sc, a0, d)).all()
    # a few distinct values only: several large buckets per window
    vals = H.random_scalars_mont(98, 3)
    sc = vals[np.arange(n) % 3]
    assert (emu.msm(np.ascontiguousarray(sc), table) == H.closed_form_msm(np.ascontiguousarray(sc), a0, d)).all()


def test_msm_giant_bucket_sub_spans(emu, monkeypatch):
    """A bucket spanning more than FIXUP_SUB = 2048 slices (forced with a short slice length): the fix-up reduces it in
    sub-spans, one block each, then adds the partial sums; the empty buckets behind it are skipped by binary search."""
    monkeypatch.setenv("BBG_MSM_SLICE", "8")
    n = 20000  # 20000 entries per giant bucket / 8 per slice = 2500 slices: two sub-spans
    table, a0, d = H.generator_multiples_table(14, n)
    sc = np.tile(H.random_scalars_mont(96, 1)[0], (n, 1))
    assert (emu.msm(sc, table) == H.closed_form_msm(sc, a0, d)).all()
    vals = H.random_scalars_mont(95, 2)
    sc = np.ascontiguousarray(vals[np.arange(n) % 2])
    assert (emu.msm(sc, table) == H.closed_form_msm(sc, a0, d)).all()


def test_batched_msm_and_srs_cache(emu):
    n = 256
    table, _, _ = H.generator_multiples_table(21, n)
    keep = emu.srs_register(table)
    scs = [H.random_scalars_mont(40 + i, n) for i in range(3)]
    states = [bb.scalar_multiplication.MultiplicationState(points=keep, scalars=s, num_elements=n) for s in scs]
    bb.scalar_multiplication.batched_scalar_multiplications(states, library=emu)
    for st, s in zip(states, scs):
        assert (st.output == H.oracle_msm(s, table)).all()
    # sub-range call: &points[2 * offset] with its own scalar range (scalar_multiplication.cpp:720-723)
    off, m = 64, 100
    sub = keep[2 * off:]
    got = emu.msm(scs[0][off:off + m], sub, m)
    assert (got == H.oracle_msm(np.ascontiguousarray(scs[0][off:off + m]), np.ascontiguousarray(table[2 * off:2 * (off + m)]))).all()
    emu.srs_unregister(keep)
    with pytest.raises(ValueError):
        bad = [bb.scalar_multiplication.MultiplicationState(points=table, scalars=scs[0], num_elements=n),
               bb.scalar_multiplication.MultiplicationState(points=table, scalars=scs[1], num_elements=n - 1)]
        bb.scalar_multiplication.batched_scalar_multiplications(bad, library=emu)


def test_partials_fold(emu):
    """The multi-GPU plan on one process: point-range shards -> XYZZ partials -> fold == whole MSM."""
    n = 512
    table, _, _ = H.generator_multiples_table(31, n)
    sc = H.random_scalars_mont(32, n)
    parts = []
    for r in range(4):
        lo, hi = r * n // 4, (r + 1) * n // 4
        d_s = emu.dev_alloc((hi - lo) * 32)
        d_t = emu.dev_alloc((hi - lo) * 128)
        emu.h2d(d_s, sc[lo:hi])
        emu.h2d(d_t, table[2 * lo:2 * hi])
        parts.append(emu.msm_partial_dev(d_s, d_t, hi - lo))
        emu.dev_free(d_s)
        emu.dev_free(d_t)
    assert (emu.fold_partials(np.stack(parts)) == H.oracle_msm(sc, table)).all()


def test_device_point_generator(emu):
    """bbg_g1_generate_multiples_dev == the oracle's (a0 + i d) G progression, including a run boundary."""
    n, a0, d = 70, 12345, 777
    dp = emu.dev_alloc(n * 64)
    emu.generate_multiples_dev(H.to_limbs(H.mont(a0)), H.to_limbs(H.mont(d)), dp, n)
    out = np.zeros((n, 8), dtype=np.uint64)
    emu.d2h(out, dp)
    assert (out == H.arithmetic_progression_points(a0, d, n)).all()
    dt = emu.dev_alloc(n * 128)
    emu.generate_pippenger_point_table_dev(dp, dt, n)
    tab = np.zeros((2 * n, 8), dtype=np.uint64)
    emu.d2h(tab, dt)
    exp = np.zeros_like(tab)
    H.oracle().orc_generate_pippenger_point_table(H.ptr(out), H.ptr(exp), n)
    assert (tab == exp).all()
    emu.dev_free(dp)
    emu.dev_free(dt)


@pytest.mark.parametrize("log_src,log_tgt", [(2, 3), (3, 5), (6, 7), (10, 11), (10, 12)])
def test_compute_lagrange_polynomial_fft(emu, log_src, log_tgt):
    """polynomial_arithmetic.cpp:381-476 on the device vs the oracle restatement (canonical limbs)."""
    got = emu.compute_lagrange_polynomial_fft(log_src, log_tgt)
    exp = np.zeros_like(got)
    H.oracle().orc_compute_lagrange_polynomial_fft(H.ptr(exp), log_src, log_tgt)
    assert (got == exp).all()


@pytest.mark.parametrize("log2_size", [1, 2, 3, 5, 11, 12])
def test_domain_lookup_table(emu, log2_size):
    """evaluation_domain::compute_lookup_table on the device: every round of both directions as values"""
    assert (emu.domain_lookup_table(log2_size) == H.expected_domain_lookup_table(log2_size)).all()


@pytest.mark.parametrize("n", [1, 2, 5, 300])
def test_srs_from_transcript(emu, n):
    """raw transcript bytes -> [G, phi(G), P_1, phi(P_1), ...]; the device copy serves MSMs behind the returned buffer"""
    ref_table, _, _ = H.generator_multiples_table(77, max(n, 2))
    pts = np.ascontiguousarray(ref_table[0:2 * n:2])
    gen = np.zeros((1, 8), dtype=np.uint64)
    gen[0, :4] = H.to_limbs(H.mont(1, H.FQ))
    gen[0, 4:] = H.to_limbs(H.mont(2, H.FQ))
    pts[0] = gen[0]
    expect = np.zeros((2 * n, 8), dtype=np.uint64)
    H.oracle().orc_generate_pippenger_point_table(H.ptr(pts), H.ptr(expect), n)
    table = emu.srs_from_transcript(H.transcript_g1_bytes(pts[1:]), n)
    assert (table == expect).all()
    sc = H.random_scalars_mont(5, n)
    assert (emu.msm(sc, table, n) == H.oracle_msm(sc, expect)).all()
    emu.srs_unregister(table)


def test_plonk_rounds_reject_misuse(emu):
    """the round entry points return BBG_E_BAD_ARGUMENT (1007) instead of running on missing inputs or out of order"""
    import ctypes as C
    L = emu.lib
    L.bbg_plonk_create.argtypes = [C.c_uint, C.POINTER(C.c_void_p)]
    L.bbg_plonk_destroy.argtypes = [C.c_void_p]
    L.bbg_plonk_round_wires.argtypes = [C.c_void_p, C.c_void_p]
    L.bbg_plonk_round_grand_product.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.bbg_plonk_set_widgets.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.c_int, C.POINTER(C.c_void_p)]
    h = C.c_void_p()
    assert L.bbg_plonk_create(1, C.byref(h)) == 1002   # at least 4 gates
    assert L.bbg_plonk_create(30, C.byref(h)) == 1002  # beyond the supported size
    assert L.bbg_plonk_create(4, C.byref(h)) == 0
    out = np.zeros(36, dtype=np.uint64)
    k = np.zeros(4, dtype=np.uint64)
    assert L.bbg_plonk_round_wires(h, out.ctypes.data_as(C.c_void_p)) == 1007  # no witness / permutation / widgets / SRS yet
    assert L.bbg_plonk_round_grand_product(h, k.ctypes.data_as(C.c_void_p), k.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 1007
    kinds = (C.c_int * 2)(0, 0)  # the same widget kind twice
    sel = (C.c_void_p * 10)(*([k.ctypes.data] * 10))
    assert L.bbg_plonk_set_widgets(h, kinds, 2, sel) == 1007
    kinds = (C.c_int * 1)(9)     # unknown kind
    assert L.bbg_plonk_set_widgets(h, kinds, 1, sel) == 1007
    assert L.bbg_plonk_destroy(h) == 0


def test_msm_launch_finish_tickets(emu):
    """bbg_msm_g1_partial_dev_launch / _finish: several MSMs in flight (tickets finished out of order), an NTT queued in
    between, results equal to the one-call form; an empty MSM gives infinity; a stale ticket is rejected"""
    lib = emu
    n = 200
    table, _, _ = H.generator_multiples_table(41, n)
    d_t = lib.dev_alloc(n * 128)
    lib.h2d(d_t, table)
    scs = [H.random_scalars_mont(50 + i, n) for i in range(3)]
    d_s = []
    for s in scs:
        d = lib.dev_alloc(n * 32)
        lib.h2d(d, s)
        d_s.append(d)
    tickets = [lib.msm_partial_dev_launch(d, d_t, n) for d in d_s]
    x = H.random_scalars_mont(9, 1 << 10)
    assert (lib.ntt("ifft", lib.ntt("fft", x.copy())) == x).all()  # work-stream traffic while the MSMs are in flight
    empty = lib.msm_partial_dev_launch(d_s[0], d_t, 0)
    for i in (2, 0, 1):
        part = lib.msm_partial_finish(tickets[i])
        assert (lib.fold_partials(part.reshape(1, 16)) == H.oracle_msm(scs[i], table)).all()
    assert H.is_infinity(lib.fold_partials(lib.msm_partial_finish(empty).reshape(1, 16)))
    with pytest.raises(bb.BbgError):
        lib.msm_partial_finish(tickets[0])
    for d in d_s + [d_t]:
        lib.dev_free(d)


def test_msm_host_launch_finish(emu):
    """bbg_msm_g1_launch / _finish (host buffers): two MSMs queued, host-buffer transforms run in between, results equal
    to bbg_msm_g1; n = 0 gives infinity without a ticket; a ticket cannot be finished twice"""
    lib = emu
    n = 300
    table, _, _ = H.generator_multiples_table(43, n)
    scs = [H.random_scalars_mont(70 + i, n) for i in range(2)]
    tickets = [lib.msm_launch(s, table, n) for s in scs]
    x = H.random_scalars_mont(11, 1 << 12)
    assert (lib.ntt("coset_ifft", lib.ntt("coset_fft", x.copy())) == x).all()
    for i in (1, 0):
        got = lib.msm_finish(tickets[i])
        assert (got == H.oracle_msm(scs[i], table)).all()
        assert (got == lib.msm(scs[i], table, n)).all()
    assert H.is_infinity(lib.msm_finish(lib.msm_launch(scs[0], table, 0)))
    with pytest.raises(bb.BbgError):
        lib.msm_finish(tickets[0])


# ---- the reference's stand-alone polynomial helpers (polynomial_arithmetic.cpp:337-373, :478-591) ----------------------
def _canonical(a):
    """lazily reduced (k, 4) limbs -> canonical limbs"""
    out = a.copy()
    for i in range(a.shape[0]):
        out[i] = H.to_limbs(H.from_limbs(a[i]) % H.FR_MODULUS)
    return out


@pytest.mark.parametrize("n", [1, 2, 31, 4096, 4097, 12289])
def test_poly_evaluate(emu, n):
    lib = emu
    c = H.random_scalars_mont(70 + n % 97, n)
    z = H.random_scalars_mont(71, 1)[0]
    want = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_poly_evaluate(H.ptr(c), H.ptr(z), n, H.ptr(want))
    assert (lib.evaluate(c, z) == _canonical(want.reshape(1, 4))[0]).all()


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("log_src,log_target", [(3, 4), (3, 5), (8, 9), (8, 10), (10, 10)])
def test_divide_by_pseudo_vanishing_polynomial(emu, log_src, log_target):
    lib = emu
    T = 1 << log_target
    x = H.random_scalars_mont(80 + log_target, T)
    r = H.ref()
    buf = r.ref_aligned_alloc(T * 32)
    import ctypes as C
    C.memmove(buf, x.ctypes.data, T * 32)
    r.ref_divide_by_pseudo_vanishing_polynomial(buf, 1 << log_src, T)
    want = np.frombuffer((C.c_uint64 * (4 * T)).from_address(buf), dtype=np.uint64).reshape(T, 4).copy()
    r.ref_aligned_free(buf)
    got = lib.divide_by_pseudo_vanishing_polynomial(x.copy(), log_src)
    assert (got == want).all()


@pytest.mark.skipif(not H.have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("n", [1, 2, 32, 33, 1000, 5000])
def test_compute_kate_opening_coefficients(emu, n):
    """dest = (F(X) - F(z)) / (X - z): the reference's serial recurrence vs the suffix scan, as values; F(z) limb-equal"""
    lib = emu
    import ctypes as C
    src = H.random_scalars_mont(90 + n % 89, n)
    z = H.random_scalars_mont(91, 1)[0]
    r = H.ref()
    a, b = r.ref_aligned_alloc(n * 32), r.ref_aligned_alloc(n * 32)
    C.memmove(a, src.ctypes.data, n * 32)
    f_ref = np.zeros(4, dtype=np.uint64)
    r.ref_compute_kate_opening_coefficients(a, b, H.ptr(z), n, H.ptr(f_ref))
    want = np.frombuffer((C.c_uint64 * (4 * n)).from_address(b), dtype=np.uint64).reshape(n, 4).copy()
    r.ref_aligned_free(a)
    r.ref_aligned_free(b)
    got, f = lib.compute_kate_opening_coefficients(src, z)
    assert (f == f_ref).all()
    assert (got == _canonical(want)).all()


@pytest.mark.parametrize("window", [0, 9, 13, 15, 16])
def test_msm_fixed_base_windows(emu, window, monkeypatch):
    """Fixed-base form (generate_pippenger_precompute_table / pippenger_precomputed, scalar_multiplication.cpp:90-129,
    :478-573): tables registered while bbg_set_srs_precompute is on get pre-doubled windows, and every MSM over them — whole,
    sub-range, batched, degenerate scalars — must equal the plain Pippenger result and the oracle.  window = 0: the
    planner's own width; others force a width so that several window counts and top-window sizes are exercised."""
    if window:
        monkeypatch.setenv("BBG_MSM_FIXED_WINDOW", str(window))
    if window == 16:
        # the bit-sliced bucket reduction kept for very large bucket sets, in its split form (several blocks per output +
        # final sum); every other case takes the two-additions-per-chunk tree, with 1, 2 and 8 blocks per set
        monkeypatch.setenv("BBG_MSM_TREE_REDUCE", "0")
        monkeypatch.setenv("BBG_MSM_RED_SPLITS", "8")
    elif window in (9, 13):
        monkeypatch.setenv("BBG_MSM_TREE_REDUCE", "1")  # (small bucket sets would take the bit-sliced reduction)
    n = 1300
    table, a0, d = H.generator_multiples_table(55, n)
    emu.set_srs_precompute(True)
    try:
        keep = emu.srs_register(table)
        d_ptr, c, w = emu.srs_device_table(keep)
        assert w >= 6 and c * w >= 128 and (not window or c == window)
        sc = H.random_scalars_mont(56, n)
        sc[1] = 0
        sc[2] = sc[3]
        sc[4] = H.to_limbs(H.from_limbs(sc[4]) + H.FR_MODULUS)
        exp = H.oracle_msm(sc, table)
        assert (emu.msm(sc, keep) == exp).all()
        # sub-range of the registered table
        off, m = 301, 700
        assert (emu.msm(np.ascontiguousarray(sc[:m]), keep[2 * off:], m) == H.oracle_msm(np.ascontiguousarray(sc[:m]), np.ascontiguousarray(table[2 * off:2 * (off + m)]))).all()
        # batched: one bucket set per MSM of the batch
        scs = [H.random_scalars_mont(60 + i, n) for i in range(3)]
        got = emu.msm_batched(scs, [keep] * 3)
        for g_, s_ in zip(got, scs):
            assert (g_ == H.oracle_msm(s_, table)).all()
        # one digit value shared by every scalar, and all-zero scalars
        same = np.tile(H.random_scalars_mont(99, 1)[0], (n, 1))
        assert (emu.msm(same, keep) == H.closed_form_msm(same, a0, d)).all()
        assert H.is_infinity(emu.msm(np.zeros((n, 4), dtype=np.uint64), keep))
        emu.srs_unregister(keep)
        # the same table, unregistered: plain Pippenger windows, same point
        assert (emu.msm(sc, table) == exp).all()
    finally:
        emu.set_srs_precompute(False)


@pytest.mark.parametrize("rounds,bmax", [(1, 96), (2, 8), (3, 12), (4, 96)])
def test_msm_pair_sum_rounds(emu, rounds, bmax, monkeypatch):
    """Pair-sum rounds ahead of the accumulate pass (batched affine additions, bbg_msm.cu 3b), forced on for sizes the
    planner would leave to the accumulate pass alone: random scalars, zero / repeated / unreduced scalars, P + P and
    P + (-P) inside buckets (repeated points: doubling and cancellation), giant buckets, long runs of empty buckets, the
    fixed-base form and a batch — every result must equal the oracle's, whatever the number of rounds and the item size."""
    monkeypatch.setenv("BBG_MSM_PAIR_ROUNDS", str(rounds))
    monkeypatch.setenv("BBG_MSM_PAIR_BMAX", str(bmax))
    n = 1400
    table, a0, d = H.generator_multiples_table(71, n)
    sc = H.random_scalars_mont(72, n)
    sc[1] = 0
    sc[2] = sc[3]
    sc[4] = H.to_limbs(H.from_limbs(sc[4]) + H.FR_MODULUS)
    before = emu.launch_count()
    assert (emu.msm(sc, table) == H.oracle_msm(sc, table)).all()
    assert emu.launch_count() - before == 12 + 4 + rounds  # the rounds did run
    # one digit value shared by all scalars: one giant bucket per window, every other bucket empty
    same = np.tile(H.random_scalars_mont(73, 1)[0], (n, 1))
    assert (emu.msm(same, table) == H.closed_form_msm(same, a0, d)).all()
    vals = H.random_scalars_mont(74, 3)
    few = np.ascontiguousarray(vals[np.arange(n) % 3])
    assert (emu.msm(few, table) == H.closed_form_msm(few, a0, d)).all()
    assert H.is_infinity(emu.msm(np.zeros((n, 4), dtype=np.uint64), table))
    # P, -P, P, -P ... with pairwise equal scalars: every bucket cancels to the point at infinity
    m = 240
    pts = np.ascontiguousarray(table[0:2 * m:2]).copy()
    neg = np.zeros(4, dtype=np.uint64)
    for i in range(1, m, 2):
        pts[i, :4] = pts[i - 1, :4]
        H.oracle().orc_neg(H.FQ, H.ptr(pts[i - 1, 4:].copy()), H.ptr(neg))
        pts[i, 4:] = neg
    t2 = emu.generate_pippenger_point_table(pts)
    sc2 = H.random_scalars_mont(75, m)
    sc2[1::2] = sc2[0::2]
    assert H.is_infinity(emu.msm(sc2, t2))
    sc2 = H.random_scalars_mont(76, m)  # independent scalars: cancellations and doublings only where digits collide
    assert (emu.msm(sc2, t2) == H.oracle_msm(sc2, t2)).all()
    # one point repeated: every addition inside a bucket is P + P or P + (-P) or meets a point at infinity
    t3 = emu.generate_pippenger_point_table(np.tile(pts[0], (m, 1)))
    sc3 = H.random_scalars_mont(77, m)
    sc3[:40] = sc3[0]
    assert (emu.msm(sc3, t3) == H.oracle_msm(sc3, t3)).all()
    # fixed-base form (one bucket set, entries from every window's table) and a batch
    emu.set_srs_precompute(True)
    try:
        keep = emu.srs_register(table)
        assert (emu.msm(sc, keep) == H.oracle_msm(sc, table)).all()
        scs = [H.random_scalars_mont(80 + i, n) for i in range(2)]
        for g_, s_ in zip(emu.msm_batched(scs, [keep] * 2), scs):
            assert (g_ == H.oracle_msm(s_, table)).all()
        emu.srs_unregister(keep)
    finally:
        emu.set_srs_precompute(False)
