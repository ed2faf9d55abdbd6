"""CPU: the sharded multi-rank path (fhe_ram_b200/sharded.py) under gloo, world_size 2, with the
oracle's arithmetic standing in for the GPU kernels.  Checks that the partition (h mod G), the
all-to-all / all-gather order and the replicated write give limbs identical to the sequential
reference order (oracle Ram::read / read_prepare_write / write)."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def _scenario(max_addr, ws):
    from oracle.oracle import Oracle
    o = Oracle(backend="fft64", max_addr=max_addr, word_size=ws, k_pt=8)
    sk = o.secret_gen(o.source(0))
    xa, xe = o.source(1), o.source(2)
    keys = o.keys_prepare(*o.keygen(sk, xa, xe))
    data = o.source_bytes(o.source(5), max_addr * ws)
    cts = o.ram_encrypt(data, sk, xa, xe)
    idxs = [3, max_addr - 1, 4097 % max_addr, 2 * 4096 + 17]
    addrs = np.stack([o.address_encrypt(i, sk, xa, xe) for i in idxs])
    w = np.stack([o.encrypt_byte(40 + i, sk, o.source(1), o.source(1)) for i in range(ws)])
    return o, sk, keys, data, cts, idxs, addrs, w


def _worker(rank, world, port, max_addr, ws, out_dir):
    import torch.distributed as dist
    from fhe_ram_b200.sharded import ShardedRam
    from oracle_engine import HostAddress, OracleEngine
    if world > 1:
        dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    o, sk, keys, data, cts, idxs, addrs, w = _scenario(max_addr, ws)
    eng = OracleEngine(o, keys, rank, world, cts)
    sram = ShardedRam(eng, rank, world)
    batch = HostAddress(addrs, len(idxs))
    got = sram.read_batch(batch, None).reshape(len(idxs), ws, -1)
    one = HostAddress(addrs[2], 1)
    rpw = sram.read_prepare_write(one, None).reshape(ws, -1)
    sram.write(w, one, None)
    after = sram.read_batch(batch, None).reshape(len(idxs), ws, -1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), got=got, rpw=rpw, after=after, data=eng.data, tree=eng.tree)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,max_addr,ws", [(1, 1 << 14, 1), (2, 1 << 14, 1), (2, 1 << 13, 2)])
def test_sharded_equals_sequential_reference_order(built, tmp_path, world, max_addr, ws):
    if world == 1:
        _worker(0, 1, 0, max_addr, ws, str(tmp_path))
    else:
        mp.spawn(_worker, args=(world, _free_port(), max_addr, ws, str(tmp_path)), nprocs=world, join=True)
    o, sk, keys, data, cts, idxs, addrs, w = _scenario(max_addr, ws)
    ram = o.ram_new(cts.copy())
    want = np.stack([o.ram_read(ram, a, keys)[1] for a in addrs])
    rc, want_rpw = o.ram_read_prepare_write(ram, addrs[2], keys)
    assert rc == 0
    assert o.ram_write(ram, w.reshape(-1), addrs[2], keys) == 0
    want_after = np.stack([o.ram_read(ram, a, keys)[1] for a in addrs])
    full = o.ram_store(ram).reshape(ws, o.n_glwe, -1)
    tree = o.ram_tree_store(ram).reshape(ws, -1)
    for r in range(world):
        z = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(z["got"], want), f"rank {r}: batched sharded read differs"
        assert np.array_equal(z["rpw"], want_rpw), f"rank {r}: read_prepare_write differs"
        assert np.array_equal(z["after"], want_after), f"rank {r}: read after write differs"
        for s in range(ws):
            for hp in range(o.n_glwe // world):
                assert np.array_equal(z["data"][s, hp], full[s, r + world * hp]), (r, s, hp)
        if len(o.base2d()) > 1:
            assert np.array_equal(z["tree"], tree), f"rank {r}: tree[0][0] differs"
    # and the written word decrypts
    q = idxs[2]
    for b in range(ws):
        v, noise = o.decrypt_glwe(want_after[2, b], sk, o.cast_u8_to_signed(40 + b, 8))
        assert v == 40 + b and noise < -9
