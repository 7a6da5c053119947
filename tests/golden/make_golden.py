"""Generates tests/golden/ref_vectors.npz from the UNMODIFIED reference compiled into
oracle/_ref/libbb_ref.so (run in the build container: `python tests/golden/make_golden.py`).
Inputs are seeded (helpers.splitmix64), so the file is reproducible."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import helpers as H  # noqa: E402
from helpers import ptr  # noqa: E402


def main():
    r = H.ref()
    assert r is not None, "build oracle/_ref first: make -C oracle ref"
    out = {}
    # MSM: n = 600 points (a0 + i d)G, seeded scalars incl. zero / repeated / lazily reduced
    n = 600
    pts = np.zeros((n, 8), dtype=np.uint64)
    r.ref_g1_arith_progression(ptr(H.to_limbs(H.mont(0xABCDEF12345))), ptr(H.to_limbs(H.mont(0x13579BDF))), ptr(pts), n)
    table = np.zeros((2 * n, 8), dtype=np.uint64)
    r.ref_generate_pippenger_point_table(ptr(pts), ptr(table), n)
    sc = H.random_scalars_mont(2024, n)
    sc[10] = 0
    sc[11] = sc[12]
    sc[13] = H.to_limbs(H.from_limbs(sc[13]) + H.FR_MODULUS)
    jac = np.zeros(12, dtype=np.uint64)
    r.ref_pippenger(ptr(sc), ptr(table), n, 0, ptr(jac))
    r.ref_g1_batch_normalize(ptr(jac), 1)
    out.update(msm_n=np.int64(n), msm_scalars=sc, msm_table=table, msm_out_normalized=jac)
    # NTT: n = 2^10, all seven ops
    m = 1 << 10
    x = H.random_scalars_mont(4048, m)
    x[5] = H.to_limbs(H.from_limbs(x[5]) + H.FR_MODULUS)
    k = H.random_scalars_mont(9, 1)[0]
    rd = r.ref_domain_new(m)
    buf = r.ref_aligned_alloc(32 * m)
    view = np.ctypeslib.as_array((H.C.c_uint64 * (4 * m)).from_address(buf)).reshape(m, 4)
    out.update(ntt_in=x, ntt_constant=k)
    for name, op in H.NTT_OPS.items():
        view[:] = x
        r.ref_ntt(rd, op, buf, ptr(k))
        out["ntt_" + name] = view.copy()
    r.ref_aligned_free(buf)
    r.ref_domain_free(rd)
    # field muls
    for fname, field, fn in (("fr", H.FR, r.ref_fr_mul_n), ("fq", H.FQ, r.ref_fq_mul_n)):
        a, b = H.random_field_raw(11, 256, field), H.random_field_raw(12, 256, field)
        a[0] = H.to_limbs(H.from_limbs(a[0]) + H.MODULUS[field])
        res = np.zeros_like(a)
        fn(ptr(a), ptr(b), ptr(res), 256)
        out.update({fname + "_mul_a": a, fname + "_mul_b": b, fname + "_mul_r": res})
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
