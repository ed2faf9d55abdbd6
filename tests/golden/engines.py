"""The two producers of the golden vectors behind one interface: the CPU oracle (test infrastructure) and the CUDA
path through the C ABI (the product)."""
import numpy as np


def _feed(params, cts):
    """GLWEPacker feed order of src/ram.rs:424-449: bit-reversed positions, None elsewhere"""
    N, log_n = params.n(), params.log_n()
    n = len(cts)
    feed = []
    for j in range(N):
        jr = int(format(j, f"0{log_n}b")[::-1], 2)
        feed.append(cts[jr] if jr < n else None)
    return feed


class OracleEngine:
    def __init__(self, s):
        self.s, self.o, self.k = s, s.orc, s.okeys

    def external_product(self, ct, ggsw): return self.o.external_product(ct, ggsw)
    def coordinate_product(self, ct, ggsws, nd): return self.o.coordinate_product(ct, ggsws, nd)
    def trace(self, ct): return self.o.trace(self.k, ct)
    def pack(self, cts): return self.o.pack(self.k, _feed(self.s.params, cts))
    def ram_new(self, cts): return self.o.ram_new(cts)

    def ram_read(self, ram, addr):
        rc, out = self.o.ram_read(ram, addr.data, self.k)
        assert rc == 0
        return out

    def ram_rpw(self, ram, addr):
        rc, out = self.o.ram_read_prepare_write(ram, addr.data, self.k)
        assert rc == 0
        return out

    def ram_write(self, ram, w, addr):
        assert self.o.ram_write(ram, np.ascontiguousarray(w).reshape(-1), addr.data, self.k) == 0

    def ram_store(self, ram): return self.o.ram_store(ram)


class GpuEngine:
    def __init__(self, s, keys):
        from fhe_ram_b200 import api
        self.s, self.api, self.keys = s, api, keys

    def external_product(self, ct, ggsw): return self.api.external_product_batch(self.s.params, ct[None, :], ggsw)[0]
    def coordinate_product(self, ct, ggsws, nd): return self.api.coordinate_product(self.s.params, ct[None, :], ggsws, nd)[0]
    def trace(self, ct): return self.api.glwe_trace(self.s.params, self.keys, ct[None, :])[0]
    def pack(self, cts): return self.api.glwe_pack(self.s.params, self.keys, cts)

    def ram_new(self, cts):
        ram = self.s.fr.Ram.new(self.s.params)
        ram.load(cts)
        return ram

    def ram_read(self, ram, addr): return ram.read(addr, self.keys)
    def ram_rpw(self, ram, addr): return ram.read_prepare_write(addr, self.keys)
    def ram_write(self, ram, w, addr): ram.write(w, addr, self.keys)
    def ram_store(self, ram): return ram.store()
