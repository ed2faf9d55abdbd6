"""Bulk encryption on the device: k_glwe_encrypt behind fheram_ram_encrypt_sk /
fheram_address_encrypt_sk must give, limb for limb, what the CPU client side (client.cpp:
Ram::encrypt_sk src/ram.rs:129-167, Address::encrypt_sk src/address.rs:86-109) gives from the same
Sources, and leave both Sources at the same stream position.  The CPU client side is itself checked
against the oracle in tests/test_abi.py::test_client_side_matches_oracle."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def _sources_agree(a, b):
    return [a.next_u32() for _ in range(5)] == [b.next_u32() for _ in range(5)]


@pytest.mark.gpu
@pytest.mark.parametrize("max_addr,word_size,k_pt,skew", [(1 << 13, 2, 8, 0), (1 << 12, 1, 3, 1), (1 << 14, 4, 9, 3),
                                                          (1 << 10, 1, 8, 16)])
def test_ram_encrypt_sk_on_device_equals_client_side(built, max_addr, word_size, k_pt, skew):
    import fhe_ram_b200 as fr
    params = fr.Parameters.new(max_addr=max_addr, word_size=word_size, k_pt=k_pt)
    sk, _ = fr.gen_keys(params)
    data = fr.Source(5).fill_bytes(max_addr * word_size)
    xa0, xe0, xa1, xe1 = fr.Source(21), fr.Source(22), fr.Source(21), fr.Source(22)
    for s in (xa0, xa1):  # mask stream not aligned to a ChaCha block / to a 64-bit draw
        for _ in range(skew):
            s.next_u32()
    ram0, ram1 = fr.Ram.new(params), fr.Ram.new(params)
    want = ram0.encrypt_sk(data, sk, xa0, xe0)
    ram1.encrypt_sk_gpu(data, sk, xa1, xe1)
    got = ram1.store()
    assert np.array_equal(got, want)
    assert _sources_agree(xa0, xa1) and _sources_agree(xe0, xe1)
    # and it decrypts: word 0 and the last word
    cts = got.reshape(word_size, params.n_glwe(), -1)
    for idx in (0, max_addr - 1):
        h, j = divmod(idx, params.n())
        if j != 0:
            continue
        for i in range(word_size):
            w = fr.cast_u8_to_signed(int(data[i + word_size * idx]), min(8, k_pt))
            v, noise = fr.decrypt_glwe(params, cts[i, h], w, sk)
            assert v == w and noise < -(k_pt + 1)


@pytest.mark.gpu
def test_sharded_ram_encrypts_its_own_polynomials(built):
    import fhe_ram_b200 as fr
    params = fr.Parameters.new(max_addr=1 << 14, word_size=2, k_pt=8)
    sk, _ = fr.gen_keys(params)
    data = fr.Source(5).fill_bytes(params.max_addr() * 2)
    want = fr.Ram.new(params).encrypt_sk(data, sk, fr.Source(21), fr.Source(22)).reshape(2, params.n_glwe(), -1)
    for shard in range(2):
        xa, xe = fr.Source(21), fr.Source(22)
        r = fr.Ram(params, shard, 2)
        r.encrypt_sk_gpu(data, sk, xa, xe)
        # fheram_ram_store of a shard scatters its polynomials into the full layout (others stay zero)
        got = r.store().reshape(2, params.n_glwe(), -1)
        assert np.array_equal(got[:, shard::2], want[:, shard::2])


@pytest.mark.gpu
@pytest.mark.parametrize("max_addr", [1 << 13, 1 << 18])
def test_address_encrypt_sk_on_device_equals_client_side(built, max_addr):
    import fhe_ram_b200 as fr
    params = fr.Parameters.new(max_addr=max_addr, word_size=1, k_pt=8)
    sk, _ = fr.gen_keys(params)
    values = [0, 1, max_addr - 1, 4096 % max_addr, 0x2a5f % max_addr, 7 * 4096 % max_addr + 5]
    # one Source pair for all addresses in turn
    xa0, xe0, xa1, xe1 = fr.Source(31), fr.Source(32), fr.Source(31), fr.Source(32)
    want = np.concatenate([fr.Address.alloc(params).encrypt_sk(params, v, sk, xa0, xe0).data for v in values])
    dev = fr.Address.encrypt_sk_gpu(params, values, sk, xa1, xe1)
    assert np.array_equal(dev.download_raw(), want)
    assert _sources_agree(xa0, xa1) and _sources_agree(xe0, xe1)
    # one pair per address
    want2 = np.concatenate([fr.Address.alloc(params).encrypt_sk(params, v, sk, fr.Source(100 + i), fr.Source(200 + i)).data
                            for i, v in enumerate(values)])
    xas = [fr.Source(100 + i) for i in range(len(values))]
    xes = [fr.Source(200 + i) for i in range(len(values))]
    dev2 = fr.Address.encrypt_sk_gpu(params, values, sk, xas, xes)
    assert np.array_equal(dev2.download_raw(), want2)
    dev.close(); dev2.close()


@pytest.mark.gpu
def test_reads_with_device_encrypted_inputs_match_oracle(scenario, gpu_keys):
    """read on a RAM and an address set that never left the device equals the oracle's read on the
    client-side limbs of the same seeds."""
    sc = scenario()
    fr, params = sc.fr, sc.params
    keys = gpu_keys(sc)
    ram = fr.Ram.new(params)
    ram.encrypt_sk_gpu(sc.data, sc.sk, fr.Source(11), fr.Source(12))  # the scenario's seeds (conftest.py)
    assert np.array_equal(ram.store(), sc.cts)
    idxs = [3, params.max_addr() - 2]
    xa, xe = fr.Source(41), fr.Source(42)
    dev = fr.Address.encrypt_sk_gpu(params, idxs, sc.sk, xa, xe)
    got = ram.read_batch(dev, keys)
    xa, xe = fr.Source(41), fr.Source(42)
    oram = sc.orc.ram_new(sc.cts.copy())
    for b, idx in enumerate(idxs):
        a = fr.Address.alloc(params).encrypt_sk(params, idx, sc.sk, xa, xe)
        rc, want = sc.orc.ram_read(oram, a.data, sc.okeys)
        assert rc == 0 and np.array_equal(np.asarray(got[b]).reshape(-1), np.asarray(want).reshape(-1))
        sc.check_decrypt(got[b], idx)


@pytest.mark.gpu
def test_address_out_of_range_is_rejected_before_sources_move(built):
    import fhe_ram_b200 as fr
    params = fr.Parameters.new(max_addr=1 << 13, word_size=1, k_pt=8)
    sk, _ = fr.gen_keys(params)
    xa, xe = fr.Source(1), fr.Source(2)
    with pytest.raises(fr.FheRamError):
        fr.Address.encrypt_sk_gpu(params, [1, 1 << 13], sk, xa, xe)
    assert _sources_agree(xa, fr.Source(1)) and _sources_agree(xe, fr.Source(2))


_NOISE_SCRIPT = r"""
import json, sys
import numpy as np
sys.path.insert(0, %r)
import fhe_ram_b200 as fr
params = fr.Parameters.new(max_addr=1 << 13, word_size=2, k_pt=8)
sk, _ = fr.gen_keys(params)
data = fr.Source(5).fill_bytes(params.max_addr() * 2)
ram0, ram1 = fr.Ram.new(params), fr.Ram.new(params)
xa0, xe0, xa1, xe1 = fr.Source(21), fr.Source(22), fr.Source(21), fr.Source(22)
want = ram0.encrypt_sk(data, sk, xa0, xe0)
ram1.encrypt_sk_gpu(data, sk, xa1, xe1)
ok = bool(np.array_equal(ram1.store(), want))
ok &= [xe0.next_u32() for _ in range(4)] == [xe1.next_u32() for _ in range(4)]
values = list(range(0, 8192, 683))
want = np.concatenate([fr.Address.alloc(params).encrypt_sk(params, v, sk, fr.Source(100 + i), fr.Source(200 + i)).data
                       for i, v in enumerate(values)])
xes = [fr.Source(200 + i) for i in range(len(values))]
dev = fr.Address.encrypt_sk_gpu(params, values, sk, [fr.Source(100 + i) for i in range(len(values))], xes)
ok &= bool(np.array_equal(dev.download_raw(), want))
ref = [fr.Source(200 + i) for i in range(len(values))]
for i, v in enumerate(values):  # noise Sources end where the client side leaves them
    fr.Address.alloc(params).encrypt_sk(params, v, sk, fr.Source(100 + i), ref[i])
ok &= all(a.next_u32() == b.next_u32() for a, b in zip(xes, ref))
print(json.dumps({"ok": ok, "streams": 1 + len(values), **params.encrypt_stats()}))
"""


def _run_noise_script(env_over):
    env = dict(os.environ)
    env.update(env_over)
    r = subprocess.run([sys.executable, "-c", _NOISE_SCRIPT % str(ROOT)], cwd=ROOT, env=env, capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return json.loads(r.stdout.strip().splitlines()[-1])


@pytest.mark.gpu
def test_noise_is_sampled_on_the_device_by_default(built):
    st = _run_noise_script({})
    assert st["ok"]
    assert st["device_draws"] > 0 and st["host_streams"] == 0 and st["host_redraws"] <= 2, st


@pytest.mark.gpu
def test_host_redraws_of_reported_samples_keep_the_limbs(built):
    # 0.1 % of the draws fall in a report band this wide: the host re-draws them one by one
    st = _run_noise_script({"FHERAM_ENC_GUARD": "0.0005"})
    assert st["ok"] and st["host_redraws"] > 100 and st["host_streams"] == 0, st


@pytest.mark.gpu
def test_streams_with_a_possible_rejection_are_sampled_by_the_host(built):
    # report band of the rejection bound widened to |z| > 4.6 sigma: about 4 in 10 address streams (122 880 draws
    # each) hold such a draw and go to the host sampler, the others stay on the device
    st = _run_noise_script({"FHERAM_ENC_BOUND_GUARD": "4.48"})
    assert st["ok"] and 0 < st["host_streams"] < st["streams"] and st["device_draws"] > 0, st


@pytest.mark.gpu
def test_host_only_noise_mode(built):
    st = _run_noise_script({"FHERAM_ENC_NOISE": "host"})
    assert st["ok"] and st["device_draws"] == 0 and st["host_streams"] == st["streams"], st


@pytest.mark.gpu
def test_evaluation_keys_encrypt_sk_on_device_matches_client_side(built):
    """EvaluationKeys::encrypt_sk (src/keys.rs:135-180) on the device: 12 trace keys, the GGLWE -> GGSW key and the
    automorphism key p = -1, limb for limb what the CPU client side makes from the same Sources; Sources left in the
    same state; a read with the device-made keys equals a read with the uploaded ones."""
    import fhe_ram_b200 as fr
    p = fr.Parameters.new(max_addr=1 << 13, word_size=1, k_pt=8)
    sk = fr.GLWESecret.fill_ternary_prob(p, 0.5, fr.Source(0))
    xa, xe = fr.Source(3), fr.Source(4)
    cpu = fr.EvaluationKeys.encrypt_sk(p, sk, xa, xe)
    ya, ye = fr.Source(3), fr.Source(4)
    dev = fr.EvaluationKeysPrepared.encrypt_sk_gpu(p, sk, ya, ye)
    raw = dev.download_raw()
    assert np.array_equal(raw.atk_glwe, cpu.atk_glwe), np.count_nonzero(raw.atk_glwe != cpu.atk_glwe)
    assert np.array_equal(raw.gglwe_to_ggsw_key, cpu.gglwe_to_ggsw_key)
    assert np.array_equal(raw.atk_ggsw_inv, cpu.atk_ggsw_inv)
    assert ya.position() == xa.position() and ye.position() == xe.position()
    keys = fr.EvaluationKeysPrepared.alloc(p).prepare(cpu)
    data = fr.Source(5).fill_bytes(1 << 13)
    ram = fr.Ram.new(p)
    ram.encrypt_sk(data, sk, fr.Source(11), fr.Source(12))
    addr = fr.Address.alloc(p).encrypt_sk(p, 4242, sk, fr.Source(21), fr.Source(22))
    a, b = ram.read(addr, keys), ram.read(addr, dev)
    assert np.array_equal(a, b)
    # write path (GGSW inversion keys) too
    ram.read_prepare_write(addr, dev)
    ram.write(np.stack([fr.encrypt_glwe(p, 99, sk)]), addr, dev)
    v, noise = fr.decrypt_glwe(p, ram.read(addr, dev)[0], 99, sk)
    assert v == 99 and noise < -9
    ram.close()
