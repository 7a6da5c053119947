"""Parity tests proper: the CUDA path through the C ABI on a real B200 vs the oracle, the committed golden
vectors (frozen reference outputs), the compiled reference (oracle/_ref, prebuilt) at large sizes, and
size-independent properties at BASELINE.json's full sizes.  Bit-exact: integer arithmetic."""
import os

import numpy as np
import pytest

import barretenberg_b200 as bb
import helpers as H
from helpers import ptr

pytestmark = pytest.mark.gpu
GOLD = os.path.join(H.ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def lib():
    return bb.default_library()  # raises BbgError if libbbgpu.so or the GPU is missing: no fallback


def ref_ntt(op, x, k=None):
    """Compiled-reference NTT (multithreaded x86 asm) on a copy of x."""
    r = H.ref()
    n = x.shape[0]
    rd = r.ref_domain_new(n)
    buf = r.ref_aligned_alloc(32 * n)
    view = np.ctypeslib.as_array((H.C.c_uint64 * (4 * n)).from_address(buf)).reshape(n, 4)
    view[:] = x
    r.ref_ntt(rd, op, buf, ptr(np.ascontiguousarray(k)) if k is not None else None)
    out = view.copy()
    r.ref_aligned_free(buf)
    r.ref_domain_free(rd)
    return out


def ref_threads_pow2():
    r = H.ref()
    t = r.ref_omp_threads()
    p = 1
    while p * 2 <= t:
        p *= 2
    r.ref_set_omp_threads(p)  # evaluation_domain silently requires a power of two (SURVEY §5 hazard)
    return p


@pytest.mark.parametrize("log_n", list(range(1, 17)))
def test_ntt_all_ops_vs_oracle(lib, log_n):
    n = 1 << log_n
    od = H.OracleDomain(n)
    x = H.random_scalars_mont(100 + log_n, n)
    x[0] = H.to_limbs(H.from_limbs(x[0]) + H.FR_MODULUS)
    x[n // 2] = 0
    k = H.random_scalars_mont(7, 1)[0]
    for name, op in H.NTT_OPS.items():
        got = lib.ntt(name, x.copy(), k)
        assert (got == od.ntt(op, x, k)).all(), name


def test_ntt_golden_reference_vectors(lib):
    g = np.load(os.path.join(GOLD, "ref_vectors.npz"))
    for name in H.NTT_OPS:
        batch = np.stack([g["ntt_in"]] * 3).copy()
        lib.ntt(name, batch, g["ntt_constant"])
        for b in batch:
            assert (b == g["ntt_" + name]).all(), name


@pytest.mark.parametrize("log_n", [17, 18, 19, 20, 21, 22])
def test_ntt_large_vs_compiled_reference(lib, log_n):
    H.require_ref()  # -m gpu: a missing compiled reference fails, it does not skip
    ref_threads_pow2()
    n = 1 << log_n
    x = H.random_scalars_mont(log_n, n)
    k = H.random_scalars_mont(8, 1)[0]
    ops = ["fft", "ifft", "coset_fft", "coset_ifft"] if log_n >= 21 else list(H.NTT_OPS)
    for name in ops:
        got = lib.ntt(name, x.copy(), k)
        assert (got == ref_ntt(H.NTT_OPS[name], x, k)).all(), (log_n, name)


@pytest.mark.parametrize("log_n,ops", [(23, list(H.NTT_OPS)), (24, ["fft", "ifft", "coset_fft", "coset_ifft"]), (25, ["coset_fft", "ifft"])])
def test_ntt_three_pass_sizes_vs_compiled_reference(lib, log_n, ops):
    H.require_ref()  # -m gpu: a missing compiled reference fails, it does not skip
    """n > 2^22 (up to the field's 2^28 two-adicity): outer split + two-pass blocks, three HBM passes.  A circuit of
    2^21 .. 2^23 gates needs these for its 4n domain."""
    ref_threads_pow2()
    n = 1 << log_n
    x = H.random_scalars_mont(log_n, n)
    x[1] = H.to_limbs(H.from_limbs(x[1]) + H.FR_MODULUS)  # lazily reduced input
    k = H.random_scalars_mont(9, 1)[0]
    for name in ops:
        got = lib.ntt(name, x.copy(), k)
        assert (got == ref_ntt(H.NTT_OPS[name], x, k)).all(), (log_n, name)


def test_ntt_three_pass_round_trip_2p26(lib):
    """2^26 elements (2 GiB): round trips only (the CPU reference would take minutes)."""
    n = 1 << 26
    x = H.random_scalars_mont(26, n)
    fx = lib.ntt("coset_fft", x.copy())
    assert not (fx[:64] == x[:64]).all()
    assert (lib.ntt("coset_ifft", fx) == x).all()


@pytest.mark.parametrize("log_n", [20, 22])
def test_ntt_properties_full_size(lib, log_n):
    """Round trips, linearity and canonical outputs at the BASELINE sizes (no reference needed)."""
    n = 1 << log_n
    x = H.random_scalars_mont(1, n)
    y = H.random_scalars_mont(2, n)
    fx = lib.ntt("fft", x.copy())
    assert (lib.ntt("ifft", fx.copy()) == x).all()
    assert (lib.ntt("coset_ifft", lib.ntt("coset_fft", x.copy())) == x).all()
    # canonical: every output < p  (top limb first)
    p = H.to_limbs(H.FR_MODULUS)
    top = fx[:, 3]
    assert (top <= p[3]).all()
    # linearity on a sample: fft(x + y) == fft(x) + fft(y)
    s = np.zeros_like(x[:4096])
    o = H.oracle()
    xs, ys = x.copy(), y.copy()
    sum_xy = np.zeros_like(x)
    # add mod p with python ints on a strided sample is too slow at full size: use the oracle's batched mul trick
    # (x + y) computed limb-wise in numpy with carry
    carry = np.zeros(n, dtype=np.uint64)
    with np.errstate(over="ignore"):
        for i in range(4):
            t = xs[:, i] + ys[:, i]
            c1 = (t < xs[:, i]).astype(np.uint64)
            t2 = t + carry
            c2 = (t2 < t).astype(np.uint64)
            sum_xy[:, i] = t2
            carry = c1 | c2
    fs = lib.ntt("fft", sum_xy)  # inputs in [0,2p) are allowed
    fy = lib.ntt("fft", y.copy())
    idx = np.arange(0, n, n // 512)
    for i in idx:
        a = (H.from_limbs(fx[i]) + H.from_limbs(fy[i])) % H.FR_MODULUS
        assert H.from_limbs(fs[i]) == a
    del s, o


def test_ntt_prover_pattern_coset_4n(lib):
    """prover.cpp:418-425: low n coefficients non-zero, rest zero, coset_fft on the 4n domain; checked against
    the oracle at n = 2^12 (4n = 2^14)."""
    n = 1 << 12
    x = np.zeros((4 * n, 4), dtype=np.uint64)
    x[:n] = H.random_scalars_mont(3, n)
    od = H.OracleDomain(4 * n)
    assert (lib.ntt("coset_fft", x.copy()) == od.ntt(H.NTT_OPS["coset_fft"], x)).all()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 17, 64, 300, 1000, 4096, 10000])
def test_msm_vs_oracle(lib, n):
    table, _, _ = H.generator_multiples_table(900 + n, max(n, 1))
    sc = H.random_scalars_mont(33 + n, n)
    if n > 4:
        sc[1] = 0
        sc[2] = sc[3]
        sc[4] = H.to_limbs(H.from_limbs(sc[4]) + H.FR_MODULUS)
    got = lib.msm(sc, table, n)
    exp = H.oracle_msm(sc, table)
    if H.is_infinity(exp):
        assert H.is_infinity(got)
    else:
        assert (got == exp).all()


def test_msm_golden_reference_vector(lib):
    g = np.load(os.path.join(GOLD, "ref_vectors.npz"))
    got = lib.msm(g["msm_scalars"], g["msm_table"])
    assert (got == g["msm_out_normalized"]).all()


def test_msm_edge_cases(lib):
    n = 2000
    table, a0, d = H.generator_multiples_table(11, n)
    one = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_constant(6, ptr(one))
    assert H.is_infinity(lib.msm(np.zeros((n, 4), dtype=np.uint64), table))
    ones = np.tile(one, (n, 1))
    assert (lib.msm(ones, table) == H.oracle_msm(ones, table)).all()
    pts = np.ascontiguousarray(table[0::2]).copy()
    neg = np.zeros(4, dtype=np.uint64)
    for i in range(1, n, 2):
        pts[i, :4] = pts[i - 1, :4]
        H.oracle().orc_neg(H.FQ, ptr(pts[i - 1, 4:].copy()), ptr(neg))
        pts[i, 4:] = neg
    t2 = lib.generate_pippenger_point_table(pts)
    ref_t2 = np.zeros_like(t2)
    H.oracle().orc_generate_pippenger_point_table(ptr(pts), ptr(ref_t2), n)
    assert (t2 == ref_t2).all()
    sc = H.random_scalars_mont(5, n)
    sc[1::2] = sc[0::2]
    assert H.is_infinity(lib.msm(sc, t2))
    pts2 = np.tile(pts[0], (n, 1))
    t3 = lib.generate_pippenger_point_table(pts2)
    sc = H.random_scalars_mont(6, n)
    sc[:50] = sc[0]
    assert (lib.msm(sc, t3) == H.oracle_msm(sc, t3)).all()
    sc = H.random_scalars_mont(8, n)
    assert (lib.msm(sc, table) == H.closed_form_msm(sc, a0, d)).all()


@pytest.mark.parametrize("distinct", [1, 3, 1000])
def test_msm_repetitive_scalars_giant_buckets(lib, distinct):
    """Constant / highly repetitive scalar vectors (they occur in the prover) give buckets spanning thousands of
    slices: exercises the block-reduction fix-up path (its cost is a bench.py figure, `degenerate_msm_ms`, not an
    assertion here)."""
    n = 1 << 18
    table, a0, d = H.generator_multiples_table(13, n)
    vals = H.random_scalars_mont(97, distinct)
    sc = np.ascontiguousarray(vals[np.arange(n) % distinct])
    got = lib.msm(sc, table)
    assert (got == H.closed_form_msm(sc, a0, d)).all()


@pytest.mark.parametrize("log_n", [20, 21, 22])
def test_ntt_host_buffer_blocks(lib, log_n):
    """A single polynomial in a page-locked HOST buffer takes the block-pipelined path (ntt_host_blocks: column blocks
    of the upload under pass A, pass B under the column blocks of the download).  Every entry point, on a buffer the
    caller pinned and on a pageable, deliberately misaligned buffer that the registration cache locks in place (whole
    pages only: the first and last row pieces straddle unlocked pages), must give the device-resident transform's limbs
    — which test_ntt_vs_reference pins to the compiled reference at these sizes."""
    import torch

    n = 1 << log_n
    x = H.random_scalars_mont(500 + log_n, n)
    k = H.random_scalars_mont(7, 1)[0]
    d = lib.dev_alloc(n * 32)
    want = {}
    for name in H.NTT_OPS:
        lib.h2d(d, x)
        lib.ntt_dev(name, d, log_n, constant=k if name.endswith("with_constant") else None)
        out = np.zeros_like(x)
        lib.d2h(out, d)
        want[name] = out
    lib.dev_free(d)
    pinned = torch.zeros((n, 4), dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    for name in H.NTT_OPS:
        pinned[:] = x
        lib.ntt(name, pinned, k)
        assert (pinned == want[name]).all(), ("pinned", name)
    raw = np.zeros(n * 4 + 512 + 8, dtype=np.uint64)
    off = (-(raw.ctypes.data // 8) % 512) + 6  # 48 bytes past a page boundary: head and tail pages stay unlocked
    buf = raw[off:off + n * 4].reshape(n, 4)
    assert buf.ctypes.data % 4096 == 48
    lib.set_host_register_cache(True)
    try:
        before = lib.host_register_stats()["registrations"]
        for i in range(4):  # the cache page-locks a buffer that keeps coming back (6th copy)
            buf[:] = x
            lib.ntt("fft", buf)
            assert (buf == want["fft"]).all(), ("sighting", i)
        assert lib.host_register_stats()["registrations"] == before + 1
        for name in H.NTT_OPS:
            buf[:] = x
            lib.ntt(name, buf, k)
            assert (buf == want[name]).all(), ("registered", name)
        lib.host_buffer_forget(buf)
    finally:
        lib.set_host_register_cache(False)


@pytest.mark.parametrize("rounds", [0, 1, 2, 3, 4])
def test_msm_pair_sum_rounds(lib, rounds, monkeypatch):
    """Pair-sum rounds ahead of the accumulate pass (batched affine additions: bbg_msm.cu 3b), 0 .. 4 of them forced: the
    commitment does not depend on how the additions inside the buckets are arranged (SURVEY §8 note 3).  Uniform scalars
    against the closed form, constant scalars (one giant bucket per window, 2^15 - 1 empty ones), one point repeated
    (every addition a doubling, a cancellation or against the point at infinity) against the oracle, plain and fixed-base."""
    monkeypatch.setenv("BBG_MSM_PAIR_ROUNDS", str(rounds))
    n = 1 << 18
    table, a0, d = H.generator_multiples_table(17, n)
    sc = H.random_scalars_mont(41, n)
    sc[1] = 0
    sc[2] = sc[3]
    want = H.closed_form_msm(sc, a0, d)
    assert (lib.msm(sc, table) == want).all()
    const = np.tile(H.random_scalars_mont(42, 1)[0], (n, 1))
    assert (lib.msm(const, table) == H.closed_form_msm(const, a0, d)).all()
    m = 3000
    pts = np.tile(np.ascontiguousarray(table[0::2])[5], (m, 1))
    t3 = lib.generate_pippenger_point_table(pts)
    sc3 = H.random_scalars_mont(43, m)
    sc3[:64] = sc3[0]
    assert (lib.msm(sc3, t3) == H.oracle_msm(sc3, t3)).all()
    lib.set_srs_precompute(True)
    try:
        keep = lib.srs_register(table)
        assert (lib.msm(sc, keep) == want).all()
        lib.srs_unregister(keep)
    finally:
        lib.set_srs_precompute(False)


def test_batched_msm_and_srs_cache(lib):
    n = 4096
    table, _, _ = H.generator_multiples_table(21, n)
    keep = lib.srs_register(table)
    scs = [H.random_scalars_mont(40 + i, n) for i in range(3)]
    states = [bb.scalar_multiplication.MultiplicationState(points=keep, scalars=s, num_elements=n) for s in scs]
    bb.scalar_multiplication.batched_scalar_multiplications(states, library=lib)
    for st, s in zip(states, scs):
        assert (st.output == H.oracle_msm(s, table)).all()
    off, m = 640, 1000
    got = lib.msm(scs[0][off:off + m], keep[2 * off:], m)
    exp = H.oracle_msm(np.ascontiguousarray(scs[0][off:off + m]), np.ascontiguousarray(table[2 * off:2 * (off + m)]))
    assert (got == exp).all()
    lib.srs_unregister(keep)


@pytest.mark.parametrize("log_n", [16, 20])
def test_msm_full_size_closed_form_and_reference(lib, log_n):
    """BASELINE configs[0]: 2^20 points.  Points (a0 + i d) G, seeded scalars; expected value from the closed form
    (one Fr dot product + one scalar multiplication) and, when oracle/_ref is present, from the compiled reference's
    batched_scalar_multiplications (limb-equal normalised x, y — test_scalar_multiplication.cpp:315-323)."""
    n = 1 << log_n
    table, a0, d = H.generator_multiples_table(77, n)
    sc = H.random_scalars_mont(78, n)
    sc[5] = 0
    got = lib.msm(sc, table)
    assert (got == H.closed_form_msm(sc, a0, d)).all()
    # bench recipe scalars: powers of one element (bench_barretenberg.cpp:196-206)
    if H.require_ref() is not None:  # fails (never skips) when the compiled reference is missing
        r = H.ref()
        out = np.zeros((1, 12), dtype=np.uint64)
        sbuf = r.ref_aligned_alloc(32 * n)
        tbuf = r.ref_aligned_alloc(128 * n)
        sv = np.ctypeslib.as_array((H.C.c_uint64 * (4 * n)).from_address(sbuf)).reshape(n, 4)
        tv = np.ctypeslib.as_array((H.C.c_uint64 * (16 * n)).from_address(tbuf)).reshape(2 * n, 8)
        sv[:] = sc
        tv[:] = table
        ptrs = (H.C.c_void_p * 1)(sbuf)
        r.ref_batched_scalar_multiplications(ptrs, tbuf, n, 1, ptr(out))
        assert (out[0] == got).all()
        r.ref_aligned_free(sbuf)
        r.ref_aligned_free(tbuf)


@pytest.mark.parametrize("log_n,parts", [(24, 1), (26, 1), (26, 8)])
def test_msm_configs3_large_synthetic_closed_form(lib, log_n, parts):
    """BASELINE configs[3]: 2^26 synthetic points (a0 + i d) G generated on the device, uniform scalars.  No CPU oracle
    finishes this size in seconds, so the check is the size-independent closed form sum k_i (a0 + i d) G =
    [(sum k_i (a0 + i d)) mod r] G (one exact Fr dot product on the host, one 1-point MSM); parts = 8: the same points as 8
    point-range shards (the multi-GPU cut, scalar_multiplication.cpp:703-728) run one after another on this device, partials
    folded on the host.  Same kernels and planner as the bench's msm_2p26 leg."""
    import sys

    if H.ROOT not in sys.path:
        sys.path.insert(0, H.ROOT)
    from barretenberg_b200 import parallel
    from barretenberg_b200 import synthetic as S

    n = 1 << log_n
    a0, d = 0x7654321, 0xABCDE
    BLK = 1 << 20
    partials, total = [], 0
    for part in range(parts):
        lo, hi = parallel.shard_range(n, part, parts)
        n_loc = hi - lo
        d_points = lib.dev_alloc(n_loc * 64)
        d_table = lib.dev_alloc(n_loc * 128)
        lib.generate_multiples_dev(S.to_limbs(S.mont(a0 + lo * d)), S.to_limbs(S.mont(d)), d_points, n_loc)
        lib.generate_pippenger_point_table_dev(d_points, d_table, n_loc)
        lib.sync()
        lib.dev_free(d_points)
        d_scalars = lib.dev_alloc(n_loc * 32)
        for pos in range(lo, hi, BLK):
            piece = S.random_field(9000 + pos // BLK, min(BLK, hi - pos))
            lib.h2d(d_scalars + (pos - lo) * 32, piece)
            total = (total + S.dot_mod_r(piece, a0 + pos * d, d)) % S.FR_MODULUS
        partials.append(lib.msm_partial_dev(d_scalars, d_table, n_loc))
        lib.dev_free(d_scalars)
        lib.dev_free(d_table)
    got = lib.fold_partials(np.stack(partials))
    d_g, d_gt, d_s1 = lib.dev_alloc(64), lib.dev_alloc(128), lib.dev_alloc(32)
    lib.generate_multiples_dev(S.to_limbs(S.mont(1)), S.to_limbs(0), d_g, 1)
    lib.generate_pippenger_point_table_dev(d_g, d_gt, 1)
    lib.h2d(d_s1, S.to_limbs(S.mont(total)).reshape(1, 4))
    expect = lib.msm_dev(d_s1, d_gt, 1)
    for p_ in (d_g, d_gt, d_s1):
        lib.dev_free(p_)
    assert not H.is_infinity(got)
    assert (got == expect).all()


def test_msm_sharded_partials_fold(lib):
    """configs[2] on one GPU: 8 point-range shards -> XYZZ partials -> host fold == whole MSM."""
    n = 1 << 14
    table, a0, d = H.generator_multiples_table(31, n)
    sc = H.random_scalars_mont(32, n)
    parts = []
    for r in range(8):
        lo, hi = r * n // 8, (r + 1) * n // 8
        d_s = lib.dev_alloc((hi - lo) * 32)
        d_t = lib.dev_alloc((hi - lo) * 128)
        lib.h2d(d_s, sc[lo:hi])
        lib.h2d(d_t, table[2 * lo:2 * hi])
        parts.append(lib.msm_partial_dev(d_s, d_t, hi - lo))
        lib.dev_free(d_s)
        lib.dev_free(d_t)
    assert (lib.fold_partials(np.stack(parts)) == H.closed_form_msm(sc, a0, d)).all()


def test_device_resident_paths(lib):
    log_n = 14
    n = 1 << log_n
    x = H.random_scalars_mont(9, 3 * n).reshape(3, n, 4)
    d = lib.dev_alloc(x.nbytes)
    lib.h2d(d, x)
    lib.ntt_dev("coset_fft", d, log_n, batch=3)
    out = np.zeros_like(x)
    lib.d2h(out, d)
    lib.dev_free(d)
    od = H.OracleDomain(n)
    for i in range(3):
        assert (out[i] == od.ntt(H.NTT_OPS["coset_fft"], x[i])).all()


def test_device_point_generator(lib):
    n, a0, d = 5000, 12345, 777
    dp = lib.dev_alloc(n * 64)
    lib.generate_multiples_dev(H.to_limbs(H.mont(a0)), H.to_limbs(H.mont(d)), dp, n)
    out = np.zeros((n, 8), dtype=np.uint64)
    lib.d2h(out, dp)
    assert (out == H.arithmetic_progression_points(a0, d, n)).all()
    lib.dev_free(dp)


def test_profile_counters(lib):
    lib.profile_enable(True)
    n = 1 << 13
    x = H.random_scalars_mont(1, n)
    lib.ntt("fft", x)
    prof = lib.profile_read()
    lib.profile_enable(False)
    assert prof["ntt_pass_a"][1] == 1 and prof["ntt_pass_b"][1] == 1


@pytest.mark.parametrize("log_src,log_tgt", [(2, 3), (6, 7), (10, 11), (12, 14), (20, 21)])
def test_compute_lagrange_polynomial_fft(lib, log_src, log_tgt):
    """§8f widening: compute_lagrange_polynomial_fft (polynomial_arithmetic.cpp:381-476) vs the oracle (small sizes)
    and the compiled reference (prover size 2^20 -> 2^21)."""
    got = lib.compute_lagrange_polynomial_fft(log_src, log_tgt)
    t = 1 << log_tgt
    if log_tgt <= 14:
        exp = np.zeros_like(got)
        H.oracle().orc_compute_lagrange_polynomial_fft(ptr(exp), log_src, log_tgt)
        assert (got == exp).all()
    if H.require_ref() is not None:  # fails (never skips) when the compiled reference is missing
        r = H.ref()
        ref_threads_pow2()
        p = r.ref_aligned_alloc(32 * t)
        exp = np.ctypeslib.as_array((H.C.c_uint64 * (4 * t)).from_address(p)).reshape(t, 4)
        r.ref_compute_lagrange_polynomial_fft(p, 1 << log_src, t)
        assert (got == exp).all()
        r.ref_aligned_free(p)


# ---- prover construction helpers ---------------------------------------------------------------
@pytest.mark.parametrize("log2_size", [4, 11, 12, 16])
def test_domain_lookup_table_full(lib, log2_size):
    assert (lib.domain_lookup_table(log2_size) == H.expected_domain_lookup_table(log2_size)).all()


def test_domain_lookup_table_2p22_sampled(lib):
    """the 4n domain of a 2^20-gate circuit (256 MiB of tables): 2000 seeded entries of both directions against python
    integers, and the structure: round i starts with one, its second entry is the round's root"""
    lg = 22
    size = 1 << lg
    t = lib.domain_lookup_table(lg)
    od = H.OracleDomain(size)
    p = H.FR_MODULUS
    rng = np.random.default_rng(22)
    for half, which in ((0, 0), (1, 1)):
        root = H.unmont(H.from_limbs(od.constant(which)))
        for u in rng.integers(0, size - 2, size=1000):
            u = int(u)
            i = (u + 2).bit_length() - 2
            j = u + 2 - (1 << (i + 1))
            want = H.mont(pow(root, j * (size >> (i + 2)), p))
            assert H.from_limbs(t[half * size + u]) == want, (half, u)
    assert (t[size - 2:size] == 0).all() and (t[2 * size - 2:] == 0).all()


@pytest.mark.parametrize("n", [1, 2, 1000, 1 << 16])
def test_srs_from_transcript(lib, n):
    ref_table, _, _ = H.generator_multiples_table(78, max(n, 2))
    pts = np.ascontiguousarray(ref_table[0:2 * n:2])
    pts[0, :4] = H.to_limbs(H.mont(1, H.FQ))
    pts[0, 4:] = H.to_limbs(H.mont(2, H.FQ))
    expect = np.zeros((2 * n, 8), dtype=np.uint64)
    H.oracle().orc_generate_pippenger_point_table(H.ptr(pts), H.ptr(expect), n)
    table = lib.srs_from_transcript(H.transcript_g1_bytes(pts[1:]), n)
    assert (table == expect).all()
    sc = H.random_scalars_mont(6, n)
    assert (lib.msm(sc, table, n) == H.oracle_msm(sc, expect)).all()  # served by the device copy kept at load time
    lib.srs_unregister(table)


def test_msm_launch_finish_tickets(lib):
    """bbg_msm_g1_partial_dev_launch / _finish: several MSMs in flight (tickets finished out of order), an NTT queued in
    between, results equal to the one-call form; an empty MSM gives infinity; a stale ticket is rejected"""
    lib = lib
    n = 5000
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


def test_msm_host_launch_finish(lib):
    """bbg_msm_g1_launch / _finish (host buffers): two MSMs queued, host-buffer transforms run in between, results equal
    to bbg_msm_g1; n = 0 gives infinity without a ticket; a ticket cannot be finished twice"""
    n = 6000
    table, _, _ = H.generator_multiples_table(43, n)
    scs = [H.random_scalars_mont(70 + i, n) for i in range(2)]
    tickets = [lib.msm_launch(s, table, n) for s in scs]
    x = H.random_scalars_mont(11, 1 << 13)
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


@pytest.mark.parametrize("n", [1, 31, 4097, 1 << 16, 3 << 18, (1 << 20) + 5])
def test_poly_evaluate(lib, n):
    lib = lib
    c = H.random_scalars_mont(70 + n % 97, n)
    z = H.random_scalars_mont(71, 1)[0]
    want = np.zeros(4, dtype=np.uint64)
    H.oracle().orc_poly_evaluate(H.ptr(c), H.ptr(z), n, H.ptr(want))
    assert (lib.evaluate(c, z) == _canonical(want.reshape(1, 4))[0]).all()


@pytest.mark.parametrize("log_src,log_target", [(3, 5), (12, 13), (16, 18), (20, 21), (20, 22)])
def test_divide_by_pseudo_vanishing_polynomial(lib, log_src, log_target):
    H.require_ref()  # -m gpu: a missing compiled reference fails, it does not skip
    lib = lib
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


@pytest.mark.parametrize("n", [1, 33, 5000, 1 << 16, (1 << 20) - 3, 1 << 20])
def test_compute_kate_opening_coefficients(lib, n):
    H.require_ref()  # -m gpu: a missing compiled reference fails, it does not skip
    """dest = (F(X) - F(z)) / (X - z): the reference's serial recurrence vs the suffix scan, as values; F(z) limb-equal"""
    lib = lib
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


# ---- round 2: host-buffer hazards named by the round-1 review ---------------------------------------------------------
def test_msm_launched_beside_large_pageable_transforms(lib):
    """bbg_msm_g1_launch with >= 1 MiB of pageable scalars (the staging-ring path) followed by pageable transforms of
    >= 1 MiB on the work stream: the download ring must not reuse a slot whose MSM upload has not run yet, and an
    unregistered table must reach the MSM stream before its kernels and must not be overwritten by the next launch."""
    n = 1 << 16
    tables = [H.generator_multiples_table(61 + 2 * i, n) for i in range(2)]
    scs = [H.random_scalars_mont(80 + i, n) for i in range(2)]  # 2 MiB each, pageable numpy memory
    x = H.random_scalars_mont(12, 1 << 18)  # 8 MiB pageable
    for _ in range(2):  # second round: the same buffers again (registration cache hit when it is on)
        tickets = [lib.msm_launch(scs[i], tables[i][0], n) for i in range(2)]  # two different unregistered tables in flight
        y = lib.ntt("fft", x.copy())
        assert (lib.ntt("ifft", y) == x).all()
        for i in (1, 0):
            got = lib.msm_finish(tickets[i])
            assert (got == H.closed_form_msm(scs[i], tables[i][1], tables[i][2])).all(), i


def test_host_register_cache(lib):
    """Pageable buffers seen twice are page-locked in place: results unchanged on first, second and later sightings, after
    the caller rewrites the buffer, and after it tells the library to forget it."""
    lib.set_host_register_cache(True)
    try:
        n = 1 << 17  # 4 MiB
        x = H.random_scalars_mont(15, n)
        od = None
        buf = x.copy()
        expect = None
        for it in range(6):  # page-locked in place at the 6th copy (two per in-place transform)
            buf[:] = x
            lib.ntt("coset_fft", buf)
            if expect is None:
                expect = buf.copy()
                if H.have_ref():
                    assert (expect == ref_ntt(H.NTT_OPS["coset_fft"], x)).all()
            assert (buf == expect).all(), it
        # new contents behind the same (now page-locked) address
        x2 = H.random_scalars_mont(16, n)
        buf[:] = x2
        lib.ntt("coset_ifft", lib.ntt("coset_fft", buf))
        assert (buf == x2).all()
        # sub-ranges of the page-locked buffer, small and large, starting in its unregistered first page and in the middle:
        # a copy must never span page-locked and pageable memory in one piece
        for start, log_m in ((0, 12), (0, 15), (3, 12), (1000, 15), (n - (1 << 12), 12)):
            m = 1 << log_m
            buf[:] = x
            view = buf[start:start + m]
            lib.ntt("fft", view)
            ref_piece = x[start:start + m].copy()
            assert (view == H.OracleDomain(m).ntt(H.NTT_OPS["fft"], ref_piece)).all(), (start, log_m)
            assert (buf[:start] == x[:start]).all() and (buf[start + m:] == x[start + m:]).all()
        lib.host_buffer_forget(buf)
        buf[:] = x
        lib.ntt("coset_fft", buf)
        assert (buf == expect).all()
        # MSM scalars through the cache as well
        m = 1 << 16
        table, a0, d = H.generator_multiples_table(91, m)
        sc = H.random_scalars_mont(92, m)
        for _ in range(8):
            assert (lib.msm(sc, table) == H.closed_form_msm(sc, a0, d)).all()
        assert lib.host_register_stats()["registrations"] >= 2
        lib.host_buffer_forget(sc)
        lib.host_buffer_forget(table)
    finally:
        lib.set_host_register_cache(False)


def test_msm_plain_points_entry(lib):
    """bbg_msm_g1_points (behind pippenger_low_memory / pippenger_precomputed): n plain points, table built on the device"""
    import ctypes as C

    lib.lib.bbg_msm_g1_points.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]
    for n in (1, 2, 999, 5000):
        table, a0, d = H.generator_multiples_table(33 + n, n)
        plain = np.ascontiguousarray(table[0::2])
        sc = H.random_scalars_mont(5 + n, n)
        out = np.zeros(12, dtype=np.uint64)
        lib.check(lib.lib.bbg_msm_g1_points(sc.ctypes.data_as(C.c_void_p), plain.ctypes.data_as(C.c_void_p), n, out.ctypes.data_as(C.c_void_p)))
        assert (out == H.oracle_msm(sc, table)).all(), n


def test_multi_device_instance():
    """bbg_init_multi over every GPU of the box in a process of its own (tests/multi_device_check.py): host, device,
    batched and launched MSMs over registered tables fanned out by point range must equal the single-device results and the
    closed form.  With one GPU the script still runs (a one-device list) and checks the same identities."""
    import subprocess
    import sys

    out = subprocess.run([sys.executable, os.path.join(H.ROOT, "tests", "multi_device_check.py")], cwd=H.ROOT, capture_output=True, text=True, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "multi-device check ok" in out.stdout
