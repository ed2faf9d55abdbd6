"""Committed golden vectors (tests/golden/fheram_golden.json, made by tests/golden/make_golden.py with the oracle's
exact-integer backend): the oracle's FFT64 backend (CPU) and the CUDA path (-m gpu, through the C ABI) must reproduce
every digest on the same fixed-seed inputs."""
import json
from pathlib import Path

import pytest

GOLDEN = json.loads((Path(__file__).parent / "golden" / "fheram_golden.json").read_text())


def _scenario(scenario, backend):
    p = GOLDEN["params"]
    return scenario(p["max_addr"], p["word_size"], p["k_pt"], backend)


def _compare(got):
    want = GOLDEN["vectors"]
    assert set(got) == set(want)
    bad = [k for k in want if got[k] != want[k]]
    assert not bad, {k: (got[k]["head"], want[k]["head"]) for k in bad}


def test_oracle_fft64_backend_reproduces_golden(scenario):
    from golden.engines import OracleEngine
    from golden.make_golden import compute
    s = _scenario(scenario, "fft64")
    _compare(compute(s, OracleEngine(s)))


@pytest.mark.gpu
def test_cuda_path_reproduces_golden(scenario, gpu_keys):
    from golden.engines import GpuEngine
    from golden.make_golden import compute
    s = _scenario(scenario, "fft64")
    _compare(compute(s, GpuEngine(s, gpu_keys(s))))
