"""CPU: pins the oracle against every known-answer test the reference holds for this path
(src/base.rs:114-438, src/parameters.rs:296-323, src/lib.rs:23-26) and against the acceptance
scenario of examples/fhe-ram.rs (decrypt == plaintext, noise bound).  Limb-level parity with real
Poulpy output is UNPINNED (no golden ciphertexts exist; see oracle/SPEC.md)."""
import ctypes as C

import numpy as np
import pytest

from oracle.oracle import Oracle, Params


@pytest.fixture(scope="module")
def orc(built):
    return Oracle()


def _arr(v):
    return (C.c_int32 * len(v))(*v)


def _decomp(o, base, value):
    out = (C.c_uint8 * len(base))()
    o.lib.orc_base1d_decomp(_arr(base), len(base), value, out)
    return list(out)


def _recomp(o, base, digits):
    return o.lib.orc_base1d_recomp(_arr(base), len(base), (C.c_uint8 * len(digits))(*digits))


def test_base1d_max(orc):  # src/base.rs:114-129
    for base, want in (([4, 4, 4], 1 << 12), ([8, 8], 1 << 16), ([12], 1 << 12), ([1, 1, 1, 1], 1 << 4)):
        assert orc.lib.orc_base1d_max(_arr(base), len(base)) == want


def test_base1d_roundtrip_and_known_values(orc):  # src/base.rs:131-181
    base = [4, 4, 4]
    for v in (0, 1, 15, 255, 1000, 4095):
        d = _decomp(orc, base, v)
        assert len(d) == 3 and all(x < 16 for x in d)
        assert _recomp(orc, base, d) == v
    assert _decomp(orc, base, 0b0000_0000_1111) == [15, 0, 0]
    assert _decomp(orc, base, 0b1010_1100_1111) == [15, 12, 10]
    assert _recomp(orc, base, [15, 12, 10]) == 0b1010_1100_1111


def test_base1d_gap(orc):  # src/base.rs:183-199
    for base in ([4, 4, 4], [6, 6], [3, 3, 3, 3]):
        assert orc.lib.orc_base1d_gap(_arr(base), len(base), 12) == 1


def test_get_base_2d_layouts(orc):  # src/base.rs:84-108, SURVEY.md 8a
    assert orc.base2d(1 << 14, [3, 3, 3, 3]) == [[3, 3, 3, 3], [2]]
    assert orc.base2d(1 << 18, [3, 3, 3, 3]) == [[3, 3, 3, 3], [3, 3]]
    assert orc.base2d(1 << 22, [3, 3, 3, 3]) == [[3, 3, 3, 3], [3, 3, 3, 1]]
    assert orc.base2d(1 << 12, [3, 3, 3, 3]) == [[3, 3, 3, 3]]
    assert orc.base2d(1 << 5, [3, 3, 3, 3]) == [[3, 2]]
    assert orc.base2d(1 << 12, [4, 4, 4]) == [[4, 4, 4]]
    assert orc.base2d(1, [3, 3, 3, 3]) == []
    assert orc.base2d(2, [3, 3, 3, 3]) == [[1]]
    assert orc.base2d(5000, [3, 3, 3, 3]) == [[3, 3, 3, 3], [1]]


def test_reverse_bits_msb(orc):  # src/lib.rs:23-26
    assert orc.lib.orc_reverse_bits_msb(1, 12) == 2048
    assert orc.lib.orc_reverse_bits_msb(0b110, 3) == 0b011
    assert [orc.lib.orc_reverse_bits_msb(i, 2) for i in range(4)] == [0, 2, 1, 3]


def test_parameters_defaults(orc):  # src/parameters.rs:296-323
    p = orc.params
    assert (p.base2k, p.k_ct, p.k_pt, p.word_size, p.max_addr) == (17, 51, 3, 4, 1 << 14)
    assert (p.k_addr, p.k_evk_ggsw_inv, p.k_evk_trace) == (68, 85, 68)
    assert list(p.decomp_n[:p.n_decomp]) == [3, 3, 3, 3]
    assert orc.n == 4096 and orc.size_ct == 3
    r = Oracle.readme_params()
    assert (r.max_addr, r.k_pt) == (1 << 18, 9)


def test_cast_u8_to_signed(orc):  # examples/fhe-ram.rs:25-32
    assert orc.cast_u8_to_signed(0b101, 3) == -3
    assert orc.cast_u8_to_signed(0xFF, 8) == -1
    assert orc.cast_u8_to_signed(0x7F, 8) == 127
    assert orc.cast_u8_to_signed(0b0111_0011, 3) == 3


def test_trace_galois_elements(orc):
    g = orc.gal_els()
    assert len(g) == 12 and g[0] == -1 and g[1] == 5 and g[2] == 25
    for i in range(2, 12):
        assert g[i] == (g[i - 1] * g[i - 1]) % 8192


def test_normalize_and_rsh_semantics(orc):
    """rsh(1) = balanced digits of ceil(X/2); normalize is idempotent and value-preserving."""
    rng = np.random.default_rng(0)
    n, K = orc.n, 17
    g = rng.integers(-(1 << 18), 1 << 18, size=orc.glwe_len, dtype=np.int64)
    nz = orc.glwe_normalize(g)
    assert nz.min() >= -(1 << 16) and nz.max() < (1 << 16)
    assert np.array_equal(orc.glwe_normalize(nz), nz)
    val = lambda a: sum(a.reshape(3, 2, n)[l].astype(object) * (1 << (K * (2 - l))) for l in range(3))
    mod = 1 << (3 * K)
    assert np.all((val(g) - val(nz)) % mod == 0)
    r = orc.glwe_rsh(1, g)
    assert r.min() >= -(1 << 16) and r.max() < (1 << 16)
    assert np.all((val(r) - (-((-val(g)) // 2))) % mod == 0)


def test_rotate_and_automorphism_algebra(orc):
    rng = np.random.default_rng(1)
    g = rng.integers(-(1 << 16), 1 << 16, size=orc.glwe_len, dtype=np.int64)
    assert np.array_equal(orc.glwe_rotate(-5, orc.glwe_rotate(5, g)), g)
    assert np.array_equal(orc.glwe_rotate(orc.n, g), -g)
    a = orc.glwe_small_automorphism(5, g)
    inv5 = pow(5, -1, 2 * orc.n)
    assert np.array_equal(orc.glwe_small_automorphism(inv5, a), g)
    assert np.array_equal(orc.glwe_small_automorphism(-1, orc.glwe_small_automorphism(-1, g)), g)


@pytest.mark.parametrize("backend", ["exact", "fft64"])
def test_example_scenario(built, backend):
    """examples/fhe-ram.rs:34-177 at max_addr = 2^13, word_size = 2 (seeds as :37-39,66)."""
    o = Oracle(backend=backend, max_addr=1 << 13, word_size=2, k_pt=8)
    xs, xa, xe = o.source(0), o.source(0), o.source(0)
    sk = o.secret_gen(xs)
    keys = o.keys_prepare(*o.keygen(sk, xa, xe))
    src = o.source(5)
    data = o.source_bytes(src, o.params.max_addr * o.word_size)
    ram = o.ram_new(o.ram_encrypt(data, sk, xa, xe))
    idx = o.source_u32(src) % o.params.max_addr
    addr = o.address_encrypt(idx, sk, xa, xe)

    def check(cts, d, i):
        for b in range(o.word_size):
            want = o.cast_u8_to_signed(d[b + o.word_size * i], 8)
            v, noise = o.decrypt_glwe(cts[b], sk, want)
            assert v == want and noise < -(8 + 1)

    rc, out = o.ram_read(ram, addr, keys)
    assert rc == 0
    check(out, data, idx)
    assert o.op_counters() == (2 * (2 * 4 + 1), 2 * (2 * 11 + 1 + 12))  # SURVEY.md 3.2 op model
    rc, out = o.ram_read_prepare_write(ram, addr, keys)
    assert rc == 0 and o.lib.orc_ram_state(ram) == 1
    check(out, data, idx)
    assert o.ram_read(ram, addr, keys)[0] == -2          # src/ram.rs:393-396
    val = o.source_bytes(src, o.word_size)
    w = np.concatenate([o.encrypt_byte(v, sk, o.source(1), o.source(1)) for v in val])
    assert o.ram_write(ram, w, addr, keys) == 0 and o.lib.orc_ram_state(ram) == 0
    assert o.ram_write(ram, w, addr, keys) == -3         # src/ram.rs:555-558
    data2 = data.copy()
    data2[idx * 2: idx * 2 + 2] = val
    rc, out = o.ram_read(ram, addr, keys)
    check(out, data2, idx)
    other = (idx + 4097) % o.params.max_addr
    rc, out = o.ram_read(ram, o.address_encrypt(other, sk, xa, xe), keys)
    check(out, data2, other)


def test_exact_and_fft64_backends_agree(built):
    """The FFT64 restatement (the reference's arithmetic) equals the exact-integer one limb for limb."""
    res = []
    for backend in ("exact", "fft64"):
        o = Oracle(backend=backend, max_addr=1 << 13, word_size=1, k_pt=8)
        sk = o.secret_gen(o.source(0))
        xa, xe = o.source(1), o.source(2)
        raw = o.keygen(sk, xa, xe)
        keys = o.keys_prepare(*raw)
        data = o.source_bytes(o.source(5), o.params.max_addr)
        cts = o.ram_encrypt(data, sk, xa, xe)
        ram = o.ram_new(cts)
        addr = o.address_encrypt(5000, sk, xa, xe)
        _, a = o.ram_read_prepare_write(ram, addr, keys)
        w = o.encrypt_byte(77, sk, o.source(1), o.source(1))
        o.ram_write(ram, w, addr, keys)
        _, b = o.ram_read(ram, addr, keys)
        res.append((raw[0], cts, addr, a, o.ram_store(ram), b))
    for x, y in zip(*res):
        assert np.array_equal(x, y)


def test_packer_places_coefficient_zero(orc):
    """GLWEPacker semantics: coefficient h of the output = coefficient 0 of input h."""
    o = Oracle(max_addr=1 << 13, word_size=1, k_pt=8)
    sk = o.secret_gen(o.source(3))
    xa, xe = o.source(1), o.source(2)
    keys = o.keys_prepare(*o.keygen(sk, xa, xe))
    vals = [17, -5, 100, -128]
    cts = [o.encrypt_byte(v & 0xFF, sk, xa, xe) for v in vals]
    feed = []
    for j in range(o.n):
        jr = o.lib.orc_reverse_bits_msb(j, 12)
        feed.append(cts[jr] if jr < 4 else None)
    packed = o.pack(keys, feed)
    pt = o.glwe_decrypt(packed, sk)
    got = [int(round(int(pt[0, h]) / (1 << 9))) for h in range(4)]
    assert got == vals


def test_ggsw_inverse_decrypts_to_inverse_monomial(orc):
    """prepare_inv (coordinate_prepared.rs:121-142): GGSW(X^-e) -> GGSW(X^+e)."""
    o = Oracle(max_addr=1 << 13, word_size=1, k_pt=8)
    sk = o.secret_gen(o.source(3))
    xa, xe = o.source(1), o.source(2)
    keys = o.keys_prepare(*o.keygen(sk, xa, xe))
    addr = o.address_encrypt(5, sk, xa, xe)            # digit 0 of coordinate 0: X^-5
    g = addr[: o.ggsw_len]
    pt = o.ggsw_decrypt_row(g, 0, 0, sk)
    assert pt[0, o.n - 5] == -1 and np.count_nonzero(pt[0]) == 1
    inv = o.ggsw_automorphism_inv(keys, g)
    pt = o.ggsw_decrypt_row(inv, 0, 0, sk)
    assert pt[0, 5] == 1 and np.count_nonzero(pt[0]) == 1
    pt = o.ggsw_decrypt_row(inv, 2, 0, sk)
    assert pt[2, 5] == 1 and np.count_nonzero(pt[2]) == 1
