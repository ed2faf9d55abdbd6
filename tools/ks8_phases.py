"""Per-phase cycles of one CTA of k_ks8 (12-step trace on 4 ciphertexts); FHERAM_STAGGER picks the CTA rank (0..7)."""
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
cts = rng.integers(-(1 << 16), 1 << 16, size=(4, p.glwe_len()), dtype=np.int64)
names = ["g0 phaseA", "g0 forward", "g0 base/clear", "g0 barrier2", "g3 wait spectra", "g3 contract", "g3 inverse", "g3 words"]
api.glwe_trace(p, keys, cts)
out = (C.c_longlong * 8)()
api._check(api.lib().fheram_debug_phase_cycles(p.module(), 1, out))
api.glwe_trace(p, keys, cts)
api._check(api.lib().fheram_debug_phase_cycles(p.module(), 0, out))
print({nm: round(out[i] / 12) for i, nm in enumerate(names)})
p.profile(True); api.glwe_trace(p, keys, cts); print(p.profile_get()["trace"]); p.profile(False)
