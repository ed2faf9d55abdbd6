"""-m gpu: the kernel that serves a launch depends on its width: clusters of eight SMs per chain for the
narrowest launches (k_ks8<8> / k_ext9<8>), clusters of four up to a quarter of the SM count (k_ks8<4> / k_ext9<4>), the
two- and one-SM kernels above that, the throughput kernels beyond the SM count.  The small parity cases of
test_gpu_parity.py sit in the first class only, so the same three operations are run here, with the DEFAULT selection, at
widths on both sides of every boundary, and checked limb for limb against the CPU oracle on sampled ciphertexts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

WIDTHS = [4, 12, 16, 24, 37, 40, 70, 100, 160]


def _rand_glwe(rng, params, n):
    return rng.integers(-(1 << 16), 1 << 16, size=(n, params.glwe_len()), dtype=np.int64)


def _sample(n):
    return sorted({0, 1, n // 2, n - 2, n - 1} & set(range(n)))


@pytest.mark.parametrize("n", WIDTHS)
def test_trace_chain_at_every_launch_width(scenario, gpu_keys, n):
    from fhe_ram_b200 import api
    s = scenario()
    keys = gpu_keys(s)
    cts = _rand_glwe(np.random.default_rng(100 + n), s.params, n)
    got = api.glwe_trace(s.params, keys, cts, 8, 12)  # 4-step chains, the last galois elements
    for i in _sample(n):
        want = s.orc.trace(s.okeys, cts[i], 8, 12)
        assert np.array_equal(got[i], want), f"n={n} ct {i}: {np.count_nonzero(got[i] != want)} limbs differ"


@pytest.mark.parametrize("n", WIDTHS)
def test_external_product_chain_at_every_launch_width(scenario, n):
    from fhe_ram_b200 import api
    s = scenario()
    cts = _rand_glwe(np.random.default_rng(200 + n), s.params, n)
    addr = s.address(4321)
    ggsws = addr.data[: 2 * s.params.ggsw_len()]
    got = api.coordinate_product(s.params, cts, ggsws, 2)
    for i in _sample(n):
        want = s.orc.coordinate_product(cts[i], ggsws, 2)
        assert np.array_equal(got[i], want), f"n={n} ct {i}: {np.count_nonzero(got[i] != want)} limbs differ"


@pytest.mark.parametrize("n", [8, 32, 64, 128, 256])
def test_packer_levels_at_every_launch_width(scenario, gpu_keys, n):
    """n inputs: the one-sided levels as one trace-chain launch of width n (with a source map), then two-sided
    combines on n/2, n/4, ... 1 pairs: every kernel class on the way down"""
    from fhe_ram_b200 import api
    s = scenario()
    keys = gpu_keys(s)
    cts = _rand_glwe(np.random.default_rng(300 + n), s.params, n)
    got = api.glwe_pack(s.params, keys, cts)
    N, log_n = s.params.n(), s.params.log_n()
    feed = []
    for j in range(N):
        jr = int(format(j, f"0{log_n}b")[::-1], 2)
        feed.append(cts[jr] if jr < n else None)
    want = s.orc.pack(s.okeys, feed)
    assert np.array_equal(got, want), np.count_nonzero(got != want)
