"""Per-launch device time of one read / read_prepare_write / write at 2^18 x 4 B (CUDA events)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import __graft_entry__ as g
g.build()
import fhe_ram_b200 as fr
p = fr.Parameters.readme()
sk, evk = fr.gen_keys(p)
keys = fr.EvaluationKeysPrepared.alloc(p).prepare(evk)
data = fr.Source(5).fill_bytes(p.max_addr() * 4)
ram = fr.Ram.new(p)
ram.encrypt_sk(data, sk, fr.Source(1), fr.Source(2))
a = fr.Address.alloc(p).encrypt_sk(p, 123456, sk, fr.Source(3), fr.Source(4))
a.device()
w = np.stack([fr.encrypt_glwe(p, v, sk) for v in (1, 2, 3, 4)])
for name, fn in (("read", lambda: ram.read(a, keys)), ("rpw", lambda: ram.read_prepare_write(a, keys)),
                 ("write", lambda: ram.write(w, a, keys))):
    if name == "write":
        pass
    else:
        fn() if name == "read" else None
    p.profile(True)
    fn()
    recs = p.profile_records()
    p.profile(False)
    print(name, "total kernel ms %.3f" % sum(r[1] for r in recs))
    for r in recs:
        print("   %-9s %8.1f us  items %5d steps %2d" % (r[0], r[1] * 1e3, r[2], r[3]))
