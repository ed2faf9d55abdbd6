"""CPU: the C-ABI library loads and exports every symbol include/fheram.h declares (no compute
calls without a GPU), the host-side logic (parameters, digit layout, sizes) matches the reference's
KATs, the client side matches the oracle's restatement bit for bit, and the product fails loudly
without a device (no CPU fallback)."""
import ctypes as C
import os
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "fheram.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fheram_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(built):
    from fhe_ram_b200 import api
    lib = C.CDLL(str(api._LIB_PATH))
    names = _declared()
    assert len(names) > 50
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/fheram.h but not exported"
    # and the ctypes mirror binds all of them
    missing = [n for n in names if n not in api.SIGNATURES]
    assert not missing, missing


def test_library_has_sm100a_code_only(built):
    """the shipped .so carries sm_100a SASS (no multi-arch fallback)"""
    import subprocess
    from fhe_ram_b200 import api
    out = subprocess.run(["cuobjdump", "-lelf", str(api._LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert not re.search(r"sm_(?!100a)\d+", out), out


def test_parameters_kat(built):  # src/parameters.rs:296-323
    import fhe_ram_b200 as fr
    p = fr.Parameters.new()
    assert p.basek() == 17 and p.k_glwe_ct() == 51 and p.k_glwe_pt() == 3 and p.rank() == 1
    assert p.word_size() == 4 and p.max_addr() == 1 << 14 and p.n() == 4096
    assert p.k_ggsw_addr() == 68 and p.k_evk_ggsw_inv() == 85 and p.k_evk_trace() == 68
    assert p.decomp_n() == [3, 3, 3, 3] and p.dnum_ct() == 3 and p.dnum_ggsw() == 4
    r = fr.Parameters.readme()
    assert r.max_addr() == 1 << 18 and r.k_glwe_pt() == 9


def test_base2d_and_sizes(built):  # src/base.rs:84-108, SURVEY.md 8
    import fhe_ram_b200 as fr
    assert fr.Parameters.new().base2d() == [[3, 3, 3, 3], [2]]
    assert fr.Parameters.readme().base2d() == [[3, 3, 3, 3], [3, 3]]
    assert fr.Parameters.new(max_addr=1 << 22).base2d() == [[3, 3, 3, 3], [3, 3, 3, 1]]
    assert fr.Parameters.new(max_addr=1 << 12).base2d() == [[3, 3, 3, 3]]
    p = fr.Parameters.readme()
    assert p.glwe_len() * 8 == 192 * 1024 and p.ggsw_len() * 8 == 1536 * 1024
    assert p.atk_len() * 8 == 768 * 1024 and p.evk_inv_len() * 8 == 1280 * 1024
    assert p.n_ggsw() == 6 and p.n_glwe() == 64 and p.n_trace_keys() == 12
    g = p.trace_galois_elements()
    assert g[:3] == [-1, 5, 25] and all(g[i] == g[i - 1] ** 2 % 8192 for i in range(2, 12))


def test_cast_u8_to_signed(built):  # examples/fhe-ram.rs:25-32
    import fhe_ram_b200 as fr
    assert fr.cast_u8_to_signed(0b101, 3) == -3 and fr.cast_u8_to_signed(0xFF, 8) == -1
    assert fr.cast_u8_to_signed(0x7F, 8) == 127 and fr.cast_u8_to_signed(0b0111_0011, 3) == 3
    with pytest.raises(AssertionError):
        fr.cast_u8_to_signed(1, 9)


def test_client_side_matches_oracle(built):
    """same seeds -> identical secret, evaluation keys, RAM, address and word ciphertexts."""
    import fhe_ram_b200 as fr
    from fhe_ram_b200 import api
    from oracle.oracle import Oracle
    params = fr.Parameters.new(max_addr=1 << 13, word_size=2, k_pt=8)
    o = Oracle(max_addr=1 << 13, word_size=2, k_pt=8)
    sk = fr.GLWESecret.fill_ternary_prob(params, 0.5, fr.Source(0))
    osk = o.secret_gen(o.source(0))
    assert np.array_equal(sk.data, osk) and 1500 < np.count_nonzero(osk) < 2600
    evk = fr.EvaluationKeys.encrypt_sk(params, sk, fr.Source(1), fr.Source(2))
    atk, tsk, inv = o.keygen(osk, o.source(1), o.source(2))
    assert np.array_equal(evk.atk_glwe, atk) and np.array_equal(evk.gglwe_to_ggsw_key, tsk)
    assert np.array_equal(evk.atk_ggsw_inv, inv)
    data = fr.Source(5).fill_bytes(params.max_addr() * 2)
    assert np.array_equal(data, o.source_bytes(o.source(5), params.max_addr() * 2))
    xa, xe = fr.Source(7), fr.Source(8)
    cts = np.zeros(2 * params.n_glwe() * params.glwe_len(), dtype=np.int64)
    api._check(api.lib().fheram_encrypt_ram(C.byref(params.c), data.ctypes.data_as(api._PU8), api._p(sk.data),
                                            xa.h, xe.h, api._p(cts)))
    assert np.array_equal(cts, o.ram_encrypt(data, osk, o.source(7), o.source(8)))
    a = fr.Address.alloc(params).encrypt_sk(params, 5000, sk, fr.Source(9), fr.Source(10))
    assert np.array_equal(a.data, o.address_encrypt(5000, osk, o.source(9), o.source(10)))
    w = fr.encrypt_glwe(params, 200, sk)
    assert np.array_equal(w, o.encrypt_byte(200, osk, o.source(1), o.source(1)))
    assert fr.decrypt_glwe(params, w, -56, sk) == o.decrypt_glwe(w, osk, -56)
    with pytest.raises(fr.FheRamError):
        fr.Address.alloc(params).encrypt_sk(params, 1 << 13, sk, fr.Source(9), fr.Source(10))


def test_address_encodes_negated_digits(built):
    """src/address.rs:102-108 + src/coordinate.rs:151-160: digit d of coordinate c is GGSW(X^-(digit<<3d))."""
    import fhe_ram_b200 as fr
    from oracle.oracle import Oracle
    params = fr.Parameters.new(max_addr=1 << 14, word_size=1, k_pt=8)
    o = Oracle(max_addr=1 << 14, word_size=1, k_pt=8)
    sk = fr.GLWESecret.fill_ternary_prob(params, 0.5, fr.Source(0))
    idx = (2 << 12) | (5 << 9) | (0 << 6) | (7 << 3) | 1
    a = fr.Address.alloc(params).encrypt_sk(params, idx, sk, fr.Source(9), fr.Source(10))
    L = params.ggsw_len()
    expect = [1, 7 << 3, 0, 5 << 9, 2]
    for g, e in enumerate(expect):
        pt = o.ggsw_decrypt_row(a.data[g * L:(g + 1) * L], 0, 0, sk.data)
        nz = np.nonzero(pt[0])[0]
        assert len(nz) == 1
        if e == 0:
            assert nz[0] == 0 and pt[0, 0] == 1
        else:
            assert nz[0] == 4096 - e and pt[0, nz[0]] == -1


def test_source_position_and_skip(built):
    """The Source is a counter-based stream: skipping n words equals drawing them, from any alignment to the 16-word
    ChaCha20 block.  The device encryption path (fheram_ram_encrypt_sk / fheram_address_encrypt_sk) regenerates the
    draws from (key, position) and then skips the Source over them."""
    import fhe_ram_b200 as fr
    rng = np.random.default_rng(1)
    for pre in (0, 1, 2, 15, 16, 17, 31, 32, 33):
        for n in (0, 1, 2, 14, 15, 16, 17, 30, 31, 32, 33, 1000, int(rng.integers(1, 50000))):
            a, b = fr.Source(7), fr.Source(7)
            for _ in range(pre):
                a.next_u32(); b.next_u32()
            for _ in range(n):
                a.next_u32()
            b.skip(n)
            assert a.position() == b.position() == pre + n
            assert [a.next_u32() for _ in range(35)] == [b.next_u32() for _ in range(35)], (pre, n)
    # position counts the words of the byte and 64-bit draws as well (fill_bytes: one word per 4 bytes, rounded up)
    s = fr.Source(3)
    s.fill_bytes(10)
    assert s.position() == 3


@pytest.mark.skipif(any(os.path.exists(f"/dev/nvidia{i}") for i in range(8)), reason="GPU present")
def test_no_cpu_fallback(built):
    """without a CUDA device the product path raises instead of computing elsewhere."""
    import fhe_ram_b200 as fr
    p = fr.Parameters.new()
    with pytest.raises(fr.FheRamError) as e:
        fr.Ram.new(p)
    assert e.value.code == -5 and "no CPU fallback" in str(e.value)


def test_unsupported_parameters_rejected_before_device(built):
    import fhe_ram_b200 as fr
    for over in ({"base2k": 16}, {"log_n": 11}, {"k_ct": 34}):
        with pytest.raises(fr.FheRamError) as e:
            fr.Parameters.new(**over).module()
        assert e.value.code in (-1, -5)


def test_pack17_roundtrip_and_range(built):
    """packed host format: 17-bit two's-complement fields, limb i at bits [17 i, 17 i + 17)"""
    import ctypes as C
    from fhe_ram_b200 import api
    rng = np.random.default_rng(17)
    a = rng.integers(-(1 << 16), 1 << 16, size=64 * 4096, dtype=np.int64)
    a[:4] = [-(1 << 16), (1 << 16) - 1, 0, -1]
    p = api.pack17(a)
    assert p.dtype == np.uint32 and p.size == a.size * 17 // 32
    assert np.array_equal(api.unpack17(p, a.size), a)
    # bit layout pinned independently of the C code
    bits = np.zeros(a[:64].size * 17, dtype=np.uint8)
    for i, v in enumerate(a[:64]):
        for b in range(17):
            bits[17 * i + b] = (int(v) >> b) & 1
    words = np.packbits(bits.reshape(-1, 32)[:, ::-1], axis=1).view(">u4").reshape(-1)
    assert np.array_equal(words.astype(np.uint32), p[:words.size])
    bad = a.copy()
    bad[5] = 1 << 16
    with pytest.raises(api.FheRamError) as e:
        api.pack17(bad)
    assert e.value.code == -6
