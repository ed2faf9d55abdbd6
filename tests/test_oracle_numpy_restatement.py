"""CPU: a third, independent restatement of the two vmp-class operations in plain numpy (schoolbook negacyclic
convolution with np.convolve on int64, no NTT, no FFT), written from the operation semantics of SURVEY.md Appendix A.2,
checked against the C oracle.  Guards the oracle's own transforms and index conventions."""
import numpy as np
import pytest

N, K, ROWS, LOUT, LRES = 4096, 17, 3, 4, 3


def negacyclic(a, b):
    full = np.convolve(a.astype(np.int64), b.astype(np.int64))      # |coeff| <= 4096 * 2^32 < 2^45
    out = full[:N].copy()
    out[: N - 1] -= full[N:]
    return out


def normalize(big):
    """vec_znx_big_normalize of one column: big[l], l = 0 (most significant) .. LOUT-1 -> LRES balanced digits;
    the carry runs from the last limb up and the top carry is dropped"""
    carry = np.zeros(N, dtype=np.int64)
    out = np.zeros((LRES, N), dtype=np.int64)
    for l in range(LOUT - 1, -1, -1):
        t = big[l] + carry
        carry = (t + (1 << (K - 1))) >> K
        if l < LRES:
            out[l] = t - (carry << K)
    return out


def glwe_view(ct):      # [limb][col][N]
    return ct.reshape(3, 2, N)


def external_product_np(ct, ggsw):
    """res = ct (x) GGSW: rows = (digit r, input column ci); GGSW raw layout [row r][ci][limb l][co][N]"""
    a = glwe_view(ct)
    g = ggsw.reshape(ROWS, 2, LOUT, 2, N)
    res = np.zeros((3, 2, N), dtype=np.int64)
    for co in range(2):
        big = np.zeros((LOUT, N), dtype=np.int64)
        for l in range(LOUT):
            for r in range(ROWS):
                for ci in range(2):
                    big[l] += negacyclic(a[r, ci], g[r, ci, l, co])
        res[:, co, :] = normalize(big)
    return res.reshape(-1)


def automorphism_np(ct, key, gal):
    """glwe_automorphism: phi_g(normalize(KS(ct))), KS = sum_r mask digit r * key row r + body on column 0;
    key raw layout [row r][limb l][co][N]; phi_g: coefficient i -> position i g mod 2N with a sign flip past N"""
    a = glwe_view(ct)
    kk = key.reshape(ROWS, LOUT, 2, N)
    res = np.zeros((3, 2, N), dtype=np.int64)
    idx = (np.arange(N, dtype=np.int64) * (gal % (2 * N))) % (2 * N)
    pos, neg = idx % N, idx >= N
    for co in range(2):
        big = np.zeros((LOUT, N), dtype=np.int64)
        for l in range(LOUT):
            for r in range(ROWS):
                big[l] += negacyclic(a[r, 1], kk[r, l, co])
            if co == 0 and l < 3:
                big[l] += a[l, 0]
        dig = normalize(big)
        out = np.zeros((3, N), dtype=np.int64)
        out[:, pos] = np.where(neg, -dig, dig)
        res[:, co, :] = out
    return res.reshape(-1)


def test_external_product_against_schoolbook_numpy(scenario):
    s = scenario()
    rng = np.random.default_rng(31)
    ct = rng.integers(-(1 << 16), 1 << 16, size=s.params.glwe_len(), dtype=np.int64)
    ggsw = s.address(2024).data[: s.params.ggsw_len()]
    assert np.array_equal(s.orc.external_product(ct, ggsw), external_product_np(ct, ggsw))


@pytest.mark.parametrize("gi", [0, 3])
def test_automorphism_against_schoolbook_numpy(scenario, gi):
    s = scenario()
    rng = np.random.default_rng(32 + gi)
    ct = rng.integers(-(1 << 16), 1 << 16, size=s.params.glwe_len(), dtype=np.int64)
    L = s.params.atk_len()
    key = s.evk.atk_glwe[gi * L:(gi + 1) * L]
    gal = s.params.trace_galois_elements()[gi]
    assert np.array_equal(s.orc.automorphism(s.okeys, gi, 0, ct), automorphism_np(ct, key, gal))
