"""Generates tests/golden/fheram_golden.json: SHA-256 digests (+ the first limbs) of the outputs of the hot path on
fixed-seed inputs, computed with the oracle's EXACT-integer backend (negacyclic products by NTT, no floating point).

There are no golden ciphertexts in the reference and its binary cannot be built offline, so these
vectors do not pin the oracle to Poulpy; they pin (a) the oracle's two backends to each other, (b) the CUDA path to
the oracle without needing the oracle at test time, and (c) every later change of either against today's limbs.

    python tests/golden/make_golden.py          # rewrites the JSON; commit the result
"""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))

OUT = Path(__file__).with_name("fheram_golden.json")
MAX_ADDR, WORD_SIZE, K_PT = 1 << 13, 2, 8
IDX_READ, IDX_RPW = 1234, 4321
WRITE_BYTES = (0x5A, 0xC3)


def digest(a) -> dict:
    a = np.ascontiguousarray(a, dtype="<i8").reshape(-1)
    return {"sha256": hashlib.sha256(a.tobytes()).hexdigest(), "n": int(a.size), "head": [int(x) for x in a[:8]]}


def inputs(s):
    """the fixed-seed inputs every producer (oracle, CUDA path) is run on"""
    rng = np.random.default_rng(2026)
    p = s.params
    cts = rng.integers(-(1 << 16), 1 << 16, size=(4, p.glwe_len()), dtype=np.int64)
    # fresh fixed-seed sources: the session-wide Scenario's own sources are advanced by other tests
    fr = s.fr
    a_read = fr.Address.alloc(p).encrypt_sk(p, IDX_READ, s.sk, fr.Source(101), fr.Source(102))
    a_rpw = fr.Address.alloc(p).encrypt_sk(p, IDX_RPW, s.sk, fr.Source(103), fr.Source(104))
    w = np.stack([s.fr.encrypt_glwe(p, int(v), s.sk) for v in WRITE_BYTES[:p.word_size()]])
    return cts, a_read, a_rpw, w


def compute(s, backend_engine):
    """backend_engine: tests/golden/engines.py object exposing the ops on int64 limb arrays"""
    cts, a_read, a_rpw, w = inputs(s)
    p = s.params
    e = backend_engine
    out = {}
    out["inputs.ram"] = digest(s.cts)
    out["inputs.atk_glwe"] = digest(s.evk.atk_glwe)
    out["inputs.address_read"] = digest(a_read.data)
    out["inputs.write_words"] = digest(w)
    ggsw = a_read.data[: p.ggsw_len()]
    out["external_product"] = digest(e.external_product(cts[0], ggsw))
    nd = len(p.base2d()[0])
    out["coordinate_product"] = digest(e.coordinate_product(cts[1], a_read.data[: nd * p.ggsw_len()], nd))
    out["trace"] = digest(e.trace(cts[2]))
    out["pack2"] = digest(e.pack(cts[:2]))
    ram = e.ram_new(s.cts.copy())
    out["read"] = digest(e.ram_read(ram, a_read))
    out["read_prepare_write"] = digest(e.ram_rpw(ram, a_rpw))
    e.ram_write(ram, w, a_rpw)
    out["ram_after_write"] = digest(e.ram_store(ram))
    out["read_back"] = digest(e.ram_read(ram, a_rpw))
    return out


def main():
    from conftest import Scenario
    from golden.engines import OracleEngine
    import __graft_entry__ as g
    g.build()
    s = Scenario(MAX_ADDR, WORD_SIZE, K_PT, backend="exact")
    res = compute(s, OracleEngine(s))
    doc = {"generator": "tests/golden/make_golden.py (oracle, exact-integer backend)",
           "params": {"max_addr": MAX_ADDR, "word_size": WORD_SIZE, "k_pt": K_PT, "idx_read": IDX_READ, "idx_rpw": IDX_RPW,
                      "write_bytes": list(WRITE_BYTES)},
           "vectors": res}
    OUT.write_text(json.dumps(doc, indent=1) + "\n")
    print(f"wrote {OUT} ({len(res)} vectors)")


if __name__ == "__main__":
    main()
