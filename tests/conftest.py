import os
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def _has_gpu() -> bool:
    return any(os.path.exists(f"/dev/nvidia{i}") for i in range(8))


@pytest.fixture(scope="session")
def built():
    import __graft_entry__ as g
    g.build()
    return True


class Scenario:
    """Keys + RAM + addresses for one parameter set, generated once with the product's client
    side (fixed seeds, examples/fhe-ram.rs:37-39,66) and shared by the oracle and the GPU path."""

    def __init__(self, max_addr, word_size, k_pt=8, backend="fft64"):
        import fhe_ram_b200 as fr
        from oracle.oracle import Oracle
        self.fr = fr
        self.params = fr.Parameters.new(max_addr=max_addr, word_size=word_size, k_pt=k_pt)
        self.sk, self.evk = fr.gen_keys(self.params)
        self.src = fr.Source(5)
        self.data = self.src.fill_bytes(max_addr * word_size)
        self.xa, self.xe = fr.Source(11), fr.Source(12)
        import ctypes as C
        from fhe_ram_b200 import api
        self.cts = np.zeros(word_size * self.params.n_glwe() * self.params.glwe_len(), dtype=np.int64)
        api._check(api.lib().fheram_encrypt_ram(C.byref(self.params.c), self.data.ctypes.data_as(api._PU8),
                                                api._p(self.sk.data), self.xa.h, self.xe.h, api._p(self.cts)))
        self.orc = Oracle(backend=backend, max_addr=max_addr, word_size=word_size, k_pt=k_pt)
        self.okeys = self.orc.keys_prepare(self.evk.atk_glwe, self.evk.gglwe_to_ggsw_key, self.evk.atk_ggsw_inv)

    def address(self, idx):
        return self.fr.Address.alloc(self.params).encrypt_sk(self.params, idx, self.sk, self.xa, self.xe)

    def want(self, idx, i):
        return self.fr.cast_u8_to_signed(int(self.data[i + self.params.word_size() * idx]),
                                         min(8, self.params.k_glwe_pt()))

    def check_decrypt(self, cts, idx, data=None):
        data = self.data if data is None else data
        p = self.params
        for i in range(p.word_size()):
            w = self.fr.cast_u8_to_signed(int(data[i + p.word_size() * idx]), min(8, p.k_glwe_pt()))
            v, noise = self.fr.decrypt_glwe(p, cts[i], w, self.sk)
            # examples/fhe-ram.rs:107-114
            assert v == w, (idx, i, v, w)
            assert noise < -(p.k_glwe_pt() + 1), (idx, i, noise)


_scen = {}


@pytest.fixture(scope="session")
def scenario(built):
    def get(max_addr=1 << 13, word_size=2, k_pt=8, backend="fft64"):
        key = (max_addr, word_size, k_pt, backend)
        if key not in _scen:
            _scen[key] = Scenario(*key)
        return _scen[key]
    return get


@pytest.fixture(scope="session")
def gpu_keys(scenario):
    def get(s):
        if not hasattr(s, "_gkeys"):
            s._gkeys = s.fr.EvaluationKeysPrepared.alloc(s.params).prepare(s.evk)
        return s._gkeys
    return get
