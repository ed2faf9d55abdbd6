"""Per-phase SM-cycle breakdown of the fused kernels (fheram_debug_phase_cycles); run on the GPU box."""
import sys, ctypes as C
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import __graft_entry__ as g
g.build()
import fhe_ram_b200 as fr
from fhe_ram_b200 import api
p = fr.Parameters.readme()
sk, evk = fr.gen_keys(p)
keys = fr.EvaluationKeysPrepared.alloc(p).prepare(evk)
rng = np.random.default_rng(0)
n = 148 * 8
cts = rng.integers(-(1 << 16), 1 << 16, size=(n, p.glwe_len()), dtype=np.int64)
addr = fr.Address.alloc(p).encrypt_sk(p, 12345, sk, fr.Source(1), fr.Source(2))
names = ["prologue", "fwd_p1", "fwd_warp", "contract", "inverse", "epilogue", "rest"]
def run(label, fn, ops):
    fn()
    out = (C.c_longlong * 8)()
    api._check(api.lib().fheram_debug_phase_cycles(p.module(), 1, out))
    fn()
    api._check(api.lib().fheram_debug_phase_cycles(p.module(), 0, out))
    tot = sum(out)
    print(label, "cycles/op:", round(tot / ops), {nm: round(out[i] / ops) for i, nm in enumerate(names)})
run("EXT x4 chain", lambda: api.coordinate_product(p, cts, addr.data[:4 * p.ggsw_len()], 4), n * 4)
run("TRACE x12", lambda: api.glwe_trace(p, keys, cts), n * 12)
run("PACK 8->1 (combine2 levels)", lambda: api.glwe_pack(p, keys, cts[:1024]), 1)
