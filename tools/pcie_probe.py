"""Host->device and device->host copy rates seen by the C ABI's pinned-buffer path."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
a = torch.empty(1 << 28, dtype=torch.int64).pin_memory()      # 2 GiB pinned
d = torch.empty_like(a, device="cuda")
for name, fn in (("H2D pinned", lambda: d.copy_(a, non_blocking=True)), ("D2H pinned", lambda: a.copy_(d, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f"{name}: {a.numel() * 8 / dt / 1e9:.1f} GB/s")
b = np.zeros(1 << 27, dtype=np.int64)
from fhe_ram_b200 import api
import __graft_entry__ as g
g.build()
api.host_register(b)
tb = torch.from_numpy(b)
d2 = torch.empty(1 << 27, dtype=torch.int64, device="cuda")
d2.copy_(tb, non_blocking=True); torch.cuda.synchronize()
t0 = time.perf_counter(); d2.copy_(tb, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D cudaHostRegister'ed numpy: {b.nbytes / dt / 1e9:.1f} GB/s")
