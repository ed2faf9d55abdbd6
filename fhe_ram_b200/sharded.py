"""Sharded multi-GPU FHE-RAM: one process per GPU, torch.distributed for the plumbing.

Partition: rank g keeps the polynomials h == g (mod G) of every sub-RAM.  The packer feeds its
inputs in bit-reversed order (src/ram.rs:426,512), so the polynomials of one rank form one
contiguous block of the packing tree and everything below the top log2(G) levels is local
(rotation by the first coordinate, one-sided levels, local two-sided levels).  One exchange step
per batch of reads: every rank ends with one partial ciphertext per (read, sub-RAM); an
all-to-all hands read q's G partials to rank q*G/B, which finishes it (top log2 G levels, second
coordinate, trace).  Never a floating-point reduction: each combine runs exactly once, on integer
limbs, so the result is bit-identical to the single-GPU / reference order.

read_prepare_write all-gathers instead (one read) and finishes on every rank, so every rank
holds tree[0][0]; Ram::write then needs no communication at all (each rank rebuilds the new
packed polynomial redundantly and updates its own slice).

The arithmetic lives behind an `engine` object (GpuEngine = the C ABI; tests substitute an
oracle-backed engine to exercise this host logic with gloo on CPU).
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_of_read(q: int, n_reads: int, world: int) -> int:
    """rank that finishes read q (reads are split into `world` equal contiguous slices)"""
    return q // (n_reads // world)


class _DevView:
    """exposes a raw device pointer to torch through __cuda_array_interface__ (no copy)"""

    def __init__(self, ptr: int, n: int):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<i4", "data": (ptr, False), "version": 2}


class GpuEngine:
    """Local / finishing stages through libfheram_cuda.so (include/fheram.h)."""

    def __init__(self, params, rank: int, world: int, cts_full: np.ndarray):
        from . import api
        self.api, self.params, self.rank, self.world = api, params, rank, world
        self.ram = api.Ram(params, shard=rank, n_shards=world)
        self.ram.load(cts_full)
        self.L = params.word_size() * params.glwe_len()     # int32 limbs per read result / partial

    def _tensor(self, ptr, n):
        import torch
        return torch.as_tensor(_DevView(int(ptr), n), device=f"cuda:{self.params.device}")

    def read_local(self, addrs, keys):
        d = C.c_void_p()
        self.api._check(self.api.lib().fheram_ram_read_local_device(self.ram.h, addrs.device(), keys.h, C.byref(d)))
        return self._tensor(d.value, addrs.count * self.L)

    def read_finish(self, gathered, n_entries, addrs, addr_first, keys):
        d = C.c_void_p()
        self.api._check(self.api.lib().fheram_ram_read_finish_device(
            self.ram.h, C.c_void_p(gathered.data_ptr()), n_entries, 0, n_entries, addrs.device(), addr_first,
            keys.h, C.byref(d)))
        return self._tensor(d.value, n_entries * self.L)

    def rpw_local(self, addr, keys):
        d = C.c_void_p()
        self.api._check(self.api.lib().fheram_ram_rpw_local_device(self.ram.h, addr.device(), keys.h, C.byref(d)))
        return self._tensor(d.value, self.L)

    def rpw_finish(self, gathered, addr, keys):
        d = C.c_void_p()
        self.api._check(self.api.lib().fheram_ram_rpw_finish_device(
            self.ram.h, C.c_void_p(gathered.data_ptr()), addr.device(), keys.h, C.byref(d)))
        return self._tensor(d.value, self.L)

    def write(self, w, addr, keys):
        self.ram.write(w, addr, keys)

    def empty(self, n):
        import torch
        return torch.empty(n, dtype=torch.int32, device=f"cuda:{self.params.device}")

    def to_host(self, t):
        return t.cpu().numpy().astype(np.int64)

    def store(self):
        return self.ram.store()


class ShardedRam:
    """Ram (src/ram.rs:25-29) sharded over the ranks of the default process group."""

    def __init__(self, engine, rank: int, world: int):
        assert world & (world - 1) == 0, "world size must be a power of two"
        self.e, self.rank, self.world = engine, rank, world

    # ---- batched reads ------------------------------------------------------------------
    def read_batch_local_slice(self, addrs, keys):
        """B = addrs.count independent reads.  Returns this rank's slice of the results:
        reads [rank*B/G, (rank+1)*B/G), as a flat int32 tensor [B/G][word_size][glwe limbs]."""
        import torch.distributed as dist
        B, G = addrs.count, self.world
        assert B % G == 0, "batch must be a multiple of the world size"
        part = self.e.read_local(addrs, keys)                  # [B][ws] partials of the local slice
        if G == 1:
            return self.e.read_finish(part, B, addrs, 0, keys)
        recv = self.e.empty(part.numel())                      # [G shards][B/G][ws]
        dist.all_to_all_single(recv, part)                     # chunk r of `part` = reads of rank r
        return self.e.read_finish(recv, B // G, addrs, self.rank * (B // G), keys)

    def read_batch(self, addrs, keys):
        """all B results on every rank ([B][word_size][limbs] int64 numpy), for tests"""
        import torch.distributed as dist
        mine = self.read_batch_local_slice(addrs, keys)
        if self.world == 1:
            return self.e.to_host(mine)
        full = self.e.empty(mine.numel() * self.world)
        dist.all_gather_into_tensor(full, mine.contiguous())
        return self.e.to_host(full)

    # ---- read_prepare_write / write -----------------------------------------------------
    def read_prepare_write(self, addr, keys):
        import torch.distributed as dist
        part = self.e.rpw_local(addr, keys)
        if self.world > 1:
            gathered = self.e.empty(part.numel() * self.world)  # [G][1][ws]
            dist.all_gather_into_tensor(gathered, part.contiguous())
        else:
            gathered = part
        return self.e.to_host(self.e.rpw_finish(gathered, addr, keys))

    def write(self, w, addr, keys):
        """Ram::write (src/ram.rs:226-294): no communication (see module docstring)."""
        self.e.write(w, addr, keys)

    # ---- bench helpers ------------------------------------------------------------------
    def bench_closures(self, api, addr_limbs, keys, B, out_host):
        """(run_resident, run_e2e) closures for bench.py"""
        params = self.e.params
        resident = api.Address.from_limbs(params, addr_limbs, B)
        resident.device()

        def run_resident():
            return self.read_batch_local_slice(resident, keys)

        import torch
        import torch.distributed as dist
        G, rank = self.world, self.rank
        per = params.n_ggsw() * params.ggsw_len()          # int32 limbs per address on the device

        # host-buffer path, pipelined in chunks of reads per rank: every rank uploads only its own addresses over
        # PCIe on a copy stream (the others arrive over NVLink with one all-gather per chunk), so the upload of
        # chunk k+1 overlaps prepare / read / exchange / finish of chunk k.  The first chunks are small (their
        # upload overlaps nothing), then the size doubles up to `chunk`.
        cnt = B // G
        chunk = max(1, min(64, max(16, cnt // 4), cnt))
        sched = []                                  # (first read of the rank's slice, reads) per chunk
        b, sz = 0, min(8, chunk)
        while b < cnt:
            nb = min(sz, cnt - b)
            sched.append((b, nb))
            b += nb
            sz = min(2 * sz, chunk)
        # two address sets per chunk size (double buffering), each sized exactly: G * nb addresses
        sets = {}
        for _, nb in sched:
            if nb not in sets:
                sets[nb] = [api.Address.device_alloc(params, nb * G) for _ in range(2)]
        use = []                                    # address set of every chunk (alternating within a size)
        seen = {}
        for _, nb in sched:
            k = seen.get(nb, 0)
            use.append(sets[nb][k & 1])
            seen[nb] = k + 1

        def run_e2e():
            def issue_upload(ci):
                b0, nb = sched[ci]
                # chunk layout on the device: [G ranks][nb] so that one all-gather completes it
                use[ci].upload_slice_async(addr_limbs[rank * cnt + b0:rank * cnt + b0 + nb], rank * nb, nb)

            issue_upload(0)
            for ci, (b0, nb) in enumerate(sched):
                if ci + 1 < len(sched):
                    issue_upload(ci + 1)
                a = use[ci]
                a.wait_upload()
                if G > 1:
                    full = torch.as_tensor(_DevView(a.raw_ptr(), G * nb * per), device=f"cuda:{params.device}")
                    dist.all_gather_into_tensor(full, full[rank * nb * per:(rank + 1) * nb * per])
                a.prepare()
                # the chunk holds G*nb addresses, [r][j] = read (r*cnt + b0 + j): this rank finishes the nb reads
                # of its own row
                part = self.e.read_local(a, keys)                      # [G*nb][ws] local partials
                if G > 1:
                    recv = self.e.empty(part.numel())
                    dist.all_to_all_single(recv, part)
                else:
                    recv = part
                mine = self.e.read_finish(recv, nb, a, rank * nb, keys)
                a.release()
                api._check(api.lib().fheram_download_glwe(params.module(), C.c_void_p(mine.data_ptr()),
                                                          nb * params.word_size(), api._p(out_host[b0:b0 + nb])))
            return out_host

        return run_resident, run_e2e
