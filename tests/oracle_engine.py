"""Oracle-backed engine for fhe_ram_b200.sharded.ShardedRam (test infrastructure).

Restates the level-parallel / sharded schedule of the CUDA host code (ram_local_stage,
ram_finish_stage, fheram_ram_write in fhe_ram_b200/csrc/fheram_cuda.cu) with the CPU oracle's
arithmetic, on CPU tensors, so the multi-rank plumbing can run under gloo without a GPU."""
import numpy as np
import torch


def rev(x, n):
    r = 0
    for i in range(n):
        r |= ((x >> i) & 1) << (n - 1 - i)
    return r


class HostAddress:
    def __init__(self, limbs, count):
        self.data = np.ascontiguousarray(limbs, dtype=np.int64).reshape(count, -1)
        self.count = count


class OracleEngine:
    def __init__(self, orc, okeys, rank, world, cts_full):
        self.o, self.k, self.rank, self.world = orc, okeys, rank, world
        o = orc
        self.ws, self.G, self.L = o.word_size, o.n_glwe, o.glwe_len
        self.nl = self.G // world
        full = cts_full.reshape(self.ws, self.G, self.L)
        self.data = np.stack([[full[s, rank + world * hp].copy() for hp in range(self.nl)] for s in range(self.ws)])
        self.tree = np.zeros((self.ws, self.L), dtype=np.int64)
        self.state = False
        self.base2d = o.base2d()
        self.log_n = int(o.params.log_n)
        self.lg_total = (self.G - 1).bit_length() if self.G > 1 else 0
        self.one_sided = self.log_n - self.lg_total

    def _coord(self, addr_row, c):
        first = sum(len(b) for b in self.base2d[:c])
        nd = len(self.base2d[c])
        return addr_row[first * self.o.ggsw_len:(first + nd) * self.o.ggsw_len], nd

    def _pack_levels(self, buf, first_level):
        level = first_level
        while len(buf) > 1:
            buf = [self.o.packer_combine(self.k, level, buf[2 * i], buf[2 * i + 1]) for i in range(len(buf) // 2)]
            level += 1
        return buf[0]

    def _local(self, addr_row, inplace):
        g0, nd0 = self._coord(addr_row, 0)
        lgl = (self.nl - 1).bit_length() if self.nl > 1 else 0
        out = []
        for s in range(self.ws):
            rot = [self.o.coordinate_product(self.data[s, hp], g0, nd0) for hp in range(self.nl)]
            if inplace:
                for hp in range(self.nl):
                    self.data[s, hp] = rot[hp]
            if len(self.base2d) == 1:
                out.append(rot[0])
                continue
            buf = [self.o.trace(self.k, rot[rev(m, lgl)], 0, self.one_sided) for m in range(self.nl)]
            out.append(self._pack_levels(buf, self.one_sided))
        return np.stack(out)

    def read_local(self, addrs, keys):
        return torch.from_numpy(np.stack([self._local(addrs.data[b], False) for b in range(addrs.count)]).reshape(-1))

    def _finish(self, parts, addr_row, store_tree):
        """parts: [S][ws][L] partials of one read, rank-major"""
        S = self.world
        lgS = (S - 1).bit_length() if S > 1 else 0
        out = []
        for s in range(self.ws):
            if len(self.base2d) == 1:
                packed = parts[0][s]
                res = packed
            else:
                packed = self._pack_levels([parts[rev(blk, lgS)][s] for blk in range(S)], self.log_n - lgS)
                g1, nd1 = self._coord(addr_row, 1)
                res = self.o.coordinate_product(packed, g1, nd1)
                if store_tree:
                    self.tree[s] = res
            out.append(self.o.trace(self.k, res))
        return np.stack(out)

    def read_finish(self, gathered, n_entries, addrs, addr_first, keys):
        g = gathered.numpy().reshape(self.world, n_entries, self.ws, self.L)
        res = [self._finish(g[:, i], addrs.data[addr_first + i], False) for i in range(n_entries)]
        return torch.from_numpy(np.stack(res).reshape(-1))

    def rpw_local(self, addr, keys):
        assert not self.state
        return torch.from_numpy(self._local(addr.data[0], True).reshape(-1))

    def rpw_finish(self, gathered, addr, keys):
        g = gathered.numpy().reshape(self.world, self.ws, self.L)
        self.state = True
        return torch.from_numpy(self._finish(g, addr.data[0], True).reshape(-1))

    def write(self, w, addr, keys):
        assert self.state
        o, k = self.o, self.k
        w = np.asarray(w, dtype=np.int64).reshape(self.ws, self.L)
        row = addr.data[0]
        inv = np.concatenate([o.ggsw_automorphism_inv(k, row[g * o.ggsw_len:(g + 1) * o.ggsw_len])
                              for g in range(o.n_ggsw)])
        two = len(self.base2d) > 1
        for s in range(self.ws):
            to = self.tree[s] if two else self.data[s, 0]
            to = o.glwe_normalize(to - o.trace(k, to) + w[s])
            if two:
                g1, nd1 = self._coord(inv, 1)
                lo = o.coordinate_product(to, g1, nd1)
                for hp in range(self.nl):
                    h = self.rank + self.world * hp
                    t1 = o.trace(k, self.data[s, hp])
                    t2 = o.trace(k, o.glwe_rotate(-h, lo))
                    self.data[s, hp] = o.glwe_normalize(self.data[s, hp] - t1 + t2)
                self.tree[s] = o.glwe_rotate(-self.G, lo)
            else:
                self.data[s, 0] = to
            g0, nd0 = self._coord(inv, 0)
            for hp in range(self.nl):
                self.data[s, hp] = o.coordinate_product(self.data[s, hp], g0, nd0)
        self.state = False

    def empty(self, n):
        return torch.empty(n, dtype=torch.int64)

    def to_host(self, t):
        return t.numpy().astype(np.int64)

    def store_full(self, dist=None):
        return self.data
