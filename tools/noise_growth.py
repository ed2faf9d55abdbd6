"""Noise growth under repeated read_prepare_write / write cycles (README.md:36 claims "~40 mio
read/write without refresh"; examples/fhe-ram.rs:108,131,169 print the per-access noise).
Runs n cycles at random addresses on the GPU and prints log2 of the decryption noise of a read
of a never-written word and of the last written word after each cycle."""
import argparse
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import __graft_entry__ as g


def run(max_addr_log2=14, word_size=1, cycles=20, k_pt=8, seed=7, verbose=True):
    g.build()
    import fhe_ram_b200 as fr
    p = fr.Parameters.new(max_addr=1 << max_addr_log2, word_size=word_size, k_pt=k_pt)
    sk, evk = fr.gen_keys(p)
    keys = fr.EvaluationKeysPrepared.alloc(p).prepare(evk)
    data = fr.Source(5).fill_bytes(p.max_addr() * word_size)
    ram = fr.Ram.new(p)
    xa, xe = fr.Source(11), fr.Source(12)
    ram.encrypt_sk(data, sk, xa, xe)
    rng = np.random.default_rng(seed)
    probe = int(rng.integers(0, p.max_addr()))
    a_probe = fr.Address.alloc(p).encrypt_sk(p, probe, sk, xa, xe)
    rows = []
    for c in range(cycles + 1):
        if c > 0:
            idx = int(rng.integers(0, p.max_addr()))
            while idx == probe:
                idx = int(rng.integers(0, p.max_addr()))
            a = fr.Address.alloc(p).encrypt_sk(p, idx, sk, xa, xe)
            ram.read_prepare_write(a, keys)
            val = rng.integers(0, 256, size=word_size)
            ram.write(np.stack([fr.encrypt_glwe(p, int(v), sk) for v in val]), a, keys)
            data[idx * word_size:(idx + 1) * word_size] = val
            got = ram.read(a, keys)
            nw = max(fr.decrypt_glwe(p, got[i], fr.cast_u8_to_signed(int(val[i]), 8), sk)[1] for i in range(word_size))
            a.close()
        else:
            nw = float("nan")
        got = ram.read(a_probe, keys)
        res = [fr.decrypt_glwe(p, got[i], fr.cast_u8_to_signed(int(data[probe * word_size + i]), 8), sk) for i in range(word_size)]
        assert all(r[0] == fr.cast_u8_to_signed(int(data[probe * word_size + i]), 8) for i, r in enumerate(res)), "probe word corrupted"
        npr = max(r[1] for r in res)
        rows.append((c, npr, nw))
        if verbose:
            print(f"cycle {c:4d}: untouched word noise 2^{npr:6.2f}   last written word 2^{nw:6.2f}   (budget 2^-{k_pt + 1})")
    return rows


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-addr-log2", type=int, default=14)
    ap.add_argument("--cycles", type=int, default=20)
    ap.add_argument("--word-size", type=int, default=1)
    a = ap.parse_args()
    run(a.max_addr_log2, a.word_size, a.cycles)
