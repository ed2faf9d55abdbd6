"""-m gpu: LIMB-level parity at BASELINE.json's own configurations.

  config 1/3  2^18 x 4 B: read, read_prepare_write, write, read-back and a batch of reads against the oracle's
              FFT64 backend, every limb, default kernel selection (the wide kernels that produce the bench number).
              Mirrors examples/fhe-ram.rs:97-176.
  config 2    4096 GLWE(k=51) x one prepared GGSW(k=68): every ciphertext against the oracle.
  config 5    2^22 x 4 B: one read, one read_prepare_write + write (RAM and tree limbs), word_size 4.
The oracle runs the four sub-RAMs of a call on four threads (orc_set_ram_threads: same limbs, shorter test)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _acceptance(s, keys, idx, batch_idxs):
    fr, p = s.fr, s.params
    ws = p.word_size()
    s.orc.set_ram_threads(4)
    ram = fr.Ram.new(p)
    ram.load(s.cts)
    oram = s.orc.ram_new(s.cts.copy())
    try:
        addr = s.address(idx)
        got = ram.read(addr, keys)                                           # examples/fhe-ram.rs:99
        rc, want = s.orc.ram_read(oram, addr.data, s.okeys)
        assert rc == 0 and np.array_equal(got, want), f"read: {np.count_nonzero(got != want)} limbs differ"
        s.check_decrypt(got, idx)

        if batch_idxs:
            addrs = [s.address(i) for i in batch_idxs]
            gotb = ram.read_batch(fr.Address.batch(p, addrs), keys)
            rc, wantb = s.orc.ram_read_many(oram, np.concatenate([a.data for a in addrs]), len(addrs), s.okeys, 8)
            assert rc == 0
            for b, i in enumerate(batch_idxs):
                assert np.array_equal(gotb[b], wantb[b]), f"batched read {b} (address {i}): {np.count_nonzero(gotb[b] != wantb[b])} limbs differ"
                s.check_decrypt(gotb[b], i)

        got = ram.read_prepare_write(addr, keys)                             # :119
        rc, want = s.orc.ram_read_prepare_write(oram, addr.data, s.okeys)
        assert rc == 0 and np.array_equal(got, want), "read_prepare_write result"
        assert ram.state()
        assert np.array_equal(ram.store(), s.orc.ram_store(oram)), "RAM limbs after read_prepare_write"
        assert np.array_equal(ram.tree_store(), s.orc.ram_tree_store(oram)), "tree limbs after read_prepare_write"

        # bytes below 128: the example encodes the written byte unsigned (examples/fhe-ram.rs:196) but checks it as a
        # signed byte (:165); at k_pt = 9 the two agree only on 0..127
        value = np.array([(7 * idx + 3 * i + 1) % 128 for i in range(ws)], dtype=np.uint8)
        w = np.stack([fr.encrypt_glwe(p, int(v), s.sk) for v in value])      # :141-149
        ram.write(w, addr, keys)                                             # :152
        assert s.orc.ram_write(oram, w.reshape(-1), addr.data, s.okeys) == 0
        assert not ram.state()
        assert np.array_equal(ram.store(), s.orc.ram_store(oram)), "RAM limbs after write"
        assert np.array_equal(ram.tree_store(), s.orc.ram_tree_store(oram)), "tree limbs after write"

        data2 = s.data.copy()
        data2[idx * ws:(idx + 1) * ws] = value
        got = ram.read(addr, keys)                                           # :162
        rc, want = s.orc.ram_read(oram, addr.data, s.okeys)
        assert rc == 0 and np.array_equal(got, want), "read-back"
        s.check_decrypt(got, idx, data2)
    finally:
        s.orc.set_ram_threads(1)
        ram.close()


def test_config1_2pow18_x4_limb_parity(scenario, gpu_keys):
    s = scenario(1 << 18, 4, 9)
    _acceptance(s, gpu_keys(s), 200001, [0, 4095, 4096, (1 << 18) - 1, 123457])


def test_config5_2pow22_x4_limb_parity(scenario, gpu_keys):
    s = scenario(1 << 22, 4, 9)
    _acceptance(s, gpu_keys(s), 3333333, [])


def test_config2_ext_product_batch_4096_every_ciphertext(scenario):
    """BASELINE config 2: 4096 GLWE x 1 prepared GGSW at N = 2^12, every output ciphertext against the oracle."""
    from fhe_ram_b200 import api
    s = scenario(1 << 18, 4, 9)
    p = s.params
    rng = np.random.default_rng(42)
    n = 4096
    cts = rng.integers(-(1 << 16), 1 << 16, size=(n, p.glwe_len()), dtype=np.int64)
    ggsw = rng.integers(-(1 << 16), 1 << 16, size=p.ggsw_len(), dtype=np.int64)
    got = api.external_product_batch(p, cts, ggsw)
    want = s.orc.external_product_many(cts, ggsw, 8)
    bad = [i for i in range(n) if not np.array_equal(got[i], want[i])]
    assert not bad, f"{len(bad)} of {n} ciphertexts differ, first {bad[:5]}"


def test_max_addr_n_squared_2pow24(scenario, gpu_keys):
    """max_addr = N^2 (n_glwe = N: every packer level is two-sided, no trace chain before the tree):
    read against the oracle limb for limb, read_prepare_write / write / read-back at the decrypt level."""
    s = scenario(1 << 24, 1, 8)
    fr, p = s.fr, s.params
    keys = gpu_keys(s)
    ram = fr.Ram.new(p)
    ram.load(s.cts)
    oram = s.orc.ram_new(s.cts)
    idx = 13371337 % (1 << 24)
    addr = s.address(idx)
    got = ram.read(addr, keys)
    rc, want = s.orc.ram_read(oram, addr.data, s.okeys)
    assert rc == 0 and np.array_equal(got, want), np.count_nonzero(got != want)
    s.check_decrypt(got, idx)
    s.check_decrypt(ram.read_prepare_write(addr, keys), idx)
    w = np.stack([fr.encrypt_glwe(p, 77, s.sk)])
    ram.write(w, addr, keys)
    data2 = s.data.copy()
    data2[idx] = 77
    s.check_decrypt(ram.read(addr, keys), idx, data2)
    other = (idx + 4097) % (1 << 24)
    s.check_decrypt(ram.read(s.address(other), keys), other, data2)
    ram.close()
