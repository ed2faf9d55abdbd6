"""One read, one read_prepare_write and one write at 2^18 x 4 B plus one small batch: launches every kernel of the
library at least once (used for the per-kernel ncu capture in profiles/)."""
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
addrs = [fr.Address.alloc(p).encrypt_sk(p, 1000 * i + 7, sk, fr.Source(3 + i), fr.Source(40 + i)) for i in range(2)]
ram.read(addrs[0], keys)
ram.read_batch(fr.Address.batch(p, addrs), keys)
ram.read_prepare_write(addrs[0], keys)
w = np.stack([fr.encrypt_glwe(p, v, sk) for v in (1, 2, 3, 4)])
ram.write(w, addrs[0], keys)
print("done", p.launch_count(), "launches")
