"""Single read / read_prepare_write / write latency at 2^18 x 4 B through the C ABI (host clock around call + synchronize,
median of 7), as bench.py measures them; run on the GPU box."""
import sys, time, json
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
data = fr.Source(5).fill_bytes(p.max_addr() * 4)
ram = fr.Ram.new(p)
ram.encrypt_sk_gpu(data, sk, fr.Source(1), fr.Source(2))
a = fr.Address.alloc(p).encrypt_sk(p, 123456, sk, fr.Source(3), fr.Source(4))
a.device()
w = np.stack([fr.encrypt_glwe(p, v, sk) for v in (1, 2, 3, 4)])
rd, rpw, wr = [], [], []
for _ in range(8):
    p.synchronize(); t0 = time.perf_counter(); ram.read(a, keys); p.synchronize(); rd.append((time.perf_counter() - t0) * 1e3)
    p.synchronize(); t0 = time.perf_counter(); ram.read_prepare_write(a, keys); p.synchronize(); rpw.append((time.perf_counter() - t0) * 1e3)
    t0 = time.perf_counter(); ram.write(w, a, keys); p.synchronize(); wr.append((time.perf_counter() - t0) * 1e3)
got = ram.read(a, keys)
ok = all(fr.decrypt_glwe(p, got[i], v, sk)[0] == v for i, v in enumerate((1, 2, 3, 4)))
print(json.dumps({"read_ms": float(np.median(rd[1:])), "read_prepare_write_ms": float(np.median(rpw[1:])),
                  "write_ms": float(np.median(wr[1:])), "pair_ms": float(np.median(np.array(rpw[1:]) + np.array(wr[1:]))),
                  "read_back_ok": bool(ok)}))
