"""Algebra behind DESIGN.md 7.2 (frequency split of one polynomial product over two CTAs), in numpy on the CPU.

The transform of DESIGN.md 3.1 evaluates z(Y) = sum_j (a_j + i a_{j+M}) Y^j at the M roots of Y^M = i by a shifted
decimation: Y^m - c = (Y^(m/2) - w)(Y^(m/2) + w), w = sqrt(c).  Stage 0 therefore separates the product a * b mod
X^N + 1 into two INDEPENDENT half-size products, mod Y^(M/2) - w0 and mod Y^(M/2) + w0: a CTA that computes
x_lo + w0 x_hi (or x_lo - w0 x_hi) of every operand can transform, contract and inverse-transform its half without
talking to the other CTA, and the only exchange is the last inverse stage,
    x_lo = (A + B) / 2,   x_hi = (A - B) / (2 w0).
This script checks exactly that against a schoolbook negacyclic product, on 17-bit digits x a ternary secret (the
encryption kernel's product) and on 17-bit x 17-bit digits accumulated over 6 rows (the external product's)."""
import numpy as np


def fwd(z, c):
    if len(z) == 1:
        return z
    h = len(z) // 2
    w = np.sqrt(complex(c))
    return np.concatenate([fwd(z[:h] + w * z[h:], w), fwd(z[:h] - w * z[h:], -w)])


def inv(s, c):
    if len(s) == 1:
        return s
    h = len(s) // 2
    w = np.sqrt(complex(c))
    a, b = inv(s[:h], w), inv(s[h:], -w)
    return np.concatenate([(a + b) / 2, (a - b) / (2 * w)])


def fold(a):
    m = len(a) // 2
    return a[:m].astype(np.float64) + 1j * a[m:].astype(np.float64)


def unfold(z):
    return np.concatenate([np.rint(z.real), np.rint(z.imag)]).astype(np.int64)


def schoolbook(a, b):
    n = len(a)
    full = np.convolve(a.astype(object), b.astype(object))
    out = full[:n].copy()
    out[: n - 1] -= full[n:]
    return out.astype(np.int64)


def split_product(rows_a, rows_b):
    """sum_r a_r * b_r mod X^N + 1 with the two halves of the spectrum handled by two independent parties"""
    m = len(rows_a[0]) // 2
    h = m // 2
    w0 = np.sqrt(1j)
    halves = []
    for sign, c in ((+1, w0), (-1, -w0)):  # party 0: mod Y^h - w0, party 1: mod Y^h + w0
        acc = np.zeros(h, dtype=np.complex128)
        for a, b in zip(rows_a, rows_b):
            za, zb = fold(a), fold(b)
            xa = za[:h] + sign * w0 * za[h:]          # this party's side of stage 0: reads every coefficient
            xb = zb[:h] + sign * w0 * zb[h:]
            acc += fwd(xa, c) * fwd(xb, c)            # half-size transform and contraction, no communication
        halves.append(inv(acc, c))                    # half-size inverse, no communication
    A, B = halves
    return unfold(np.concatenate([(A + B) / 2, (A - B) / (2 * w0)]))  # the one exchange


def main(n=4096, seed=0):
    rng = np.random.default_rng(seed)
    a = rng.integers(-(1 << 16), 1 << 16, size=n)
    s = rng.integers(-1, 2, size=n)
    assert np.array_equal(unfold(inv(fwd(fold(a), 1j) * fwd(fold(s), 1j), 1j)), schoolbook(a, s))
    assert np.array_equal(split_product([a], [s]), schoolbook(a, s))
    n2 = min(n, 512)  # 17 x 17-bit digits over 6 rows: schoolbook in Python integers, kept small
    ra = [rng.integers(-(1 << 16), 1 << 16, size=n2) for _ in range(6)]
    rb = [rng.integers(-(1 << 16), 1 << 16, size=n2) for _ in range(6)]
    want = sum(schoolbook(x, y).astype(object) for x, y in zip(ra, rb)).astype(np.int64)
    assert np.array_equal(split_product(ra, rb), want)
    return True


if __name__ == "__main__":
    main()
    print("frequency split: both halves independent, one exchange at the end, integers exact")
