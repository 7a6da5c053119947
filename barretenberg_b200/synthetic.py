"""Seeded synthetic inputs (numpy, host side): the workload generator of bench.py and the tests.

Scalars: splitmix64(seed) -> 4 limbs, top two bits cleared, one conditional subtraction of the modulus.  The
limbs are used directly as canonical Montgomery residues (uniform residues are uniform in either form)."""
import numpy as np

FR_MODULUS = 0x30644E72E131A029B85045B68181585D2833E84879B9709143E1F593F0000001
FQ_MODULUS = 0x30644E72E131A029B85045B68181585D97816A916871CA8D3C208C16D87CFD47
R_MONT = 1 << 256


def to_limbs(x, n=4):
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(n)], dtype=np.uint64)


def from_limbs(a):
    a = np.asarray(a, dtype=np.uint64).reshape(-1)
    return sum(int(v) << (64 * i) for i, v in enumerate(a))


def splitmix64(seed, count):
    """Vectorised splitmix64 stream: `count` uint64 values from `seed`."""
    with np.errstate(over="ignore"):
        idx = np.arange(1, count + 1, dtype=np.uint64)
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def random_field(seed, n, modulus=FR_MODULUS, out=None):
    """(n, 4) uint64: n seeded canonical residues < modulus."""
    raw = splitmix64(seed, 4 * n).reshape(n, 4)
    if out is not None:
        out[:] = raw
        raw = out
    else:
        raw = raw.copy()
    raw[:, 3] &= np.uint64(0x3FFFFFFFFFFFFFFF)  # < 2^254 < 2p
    p = to_limbs(modulus)
    ge = np.zeros(n, dtype=bool)
    eq = np.ones(n, dtype=bool)
    for i in (3, 2, 1, 0):
        ge |= eq & (raw[:, i] > p[i])
        eq &= raw[:, i] == p[i]
    ge |= eq
    if ge.any():
        borrow = np.zeros(n, dtype=np.uint64)
        with np.errstate(over="ignore"):
            for i in range(4):
                a = raw[:, i].copy()
                d = a - p[i] - borrow
                nb = ((a < p[i]) | ((a == p[i]) & (borrow == 1))).astype(np.uint64)
                raw[:, i] = np.where(ge, d, a)
                borrow = nb
    return raw


def mont(x, modulus=FR_MODULUS):
    return (x * R_MONT) % modulus


def dot_mod_r(scalars_mont, a0, d):
    """(sum_i k_i * (a0 + i d)) mod r for Montgomery-form scalar limbs (value k_i = limbs * R^-1).

    Exact and vectorised: sum_i k_i (a0 + i d) = a0 * sum k_i + d * sum i k_i, with k_i cut into 16-bit columns and i
    into bytes so every partial dot product fits a uint64 (n <= 2^26: 2^26 * 2^16 * 2^8 = 2^50)."""
    n = scalars_mont.shape[0]
    if n == 0:
        return 0
    if n > (1 << 32):
        raise ValueError("dot_mod_r supports n <= 2^32")
    cols = np.ascontiguousarray(scalars_mont).view(np.uint16).reshape(n, 16)
    idx = np.arange(n, dtype=np.uint64)
    idx_bytes = [((idx >> np.uint64(8 * t)) & np.uint64(0xFF)) for t in range(4)]
    s0 = 0
    s1 = 0
    for c in range(16):
        col = cols[:, c].astype(np.uint64)
        s0 += int(col.sum(dtype=np.uint64)) << (16 * c)
        for t in range(4):
            s1 += int(np.dot(idx_bytes[t], col)) << (8 * t + 16 * c)
    return ((a0 * s0 + d * s1) * pow(R_MONT, -1, FR_MODULUS)) % FR_MODULUS
