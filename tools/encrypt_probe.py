"""Where the time of the device-side Address::encrypt_sk goes: allocation / encryption / prepare (B200)."""
import ctypes as C
import sys
import time
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import fhe_ram_b200 as fr  # noqa: E402
from fhe_ram_b200 import api  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
params = fr.Parameters.readme(max_addr=1 << 18, word_size=4, k_pt=9)
sk, _ = fr.gen_keys(params)
values = np.random.default_rng(7).integers(0, 1 << 18, size=n).astype(np.uint32)
for rep in range(3):
    xas, xes = [fr.Source(1000 + j) for j in range(n)], [fr.Source(5000 + j) for j in range(n)]
    ha = (C.c_void_p * n)(*[x.h for x in xas])
    he = (C.c_void_p * n)(*[x.h for x in xes])
    params.synchronize(); t0 = time.perf_counter()
    a = fr.Address.device_alloc(params, n)
    params.synchronize(); t1 = time.perf_counter()
    api._check(api.lib().fheram_address_encrypt_sk(a.h, 0, n, values.ctypes.data_as(C.POINTER(C.c_uint32)),
                                                   api._p(sk.data), ha, he, n))
    params.synchronize(); t2 = time.perf_counter()
    a.prepare()
    params.synchronize(); t3 = time.perf_counter()
    a.close()
    t4 = time.perf_counter()
    print(f"n={n} rep={rep}: alloc {1e3*(t1-t0):.1f} ms, encrypt {1e3*(t2-t1):.1f} ms, prepare {1e3*(t3-t2):.1f} ms, "
          f"free {1e3*(t4-t3):.1f} ms -> {n/(t3-t0):.0f} addresses/s ({n/(t2-t1):.0f}/s encryption alone)")
