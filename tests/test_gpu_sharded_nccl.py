"""-m gpu, needs >= 2 GPUs: the sharded path end to end over NCCL (one process per GPU), limb for limb against the
unsharded RAM on the same GPU.  Product path = ShardedRamLib (the exchange steps run inside libfheram_cuda.so on its
own communicator); the torch.distributed variant (ShardedRam + GpuEngine halves) is checked beside it.
Run by hand on a multi-GPU box: `gpurun --gpus 8 -- python -m pytest tests/test_gpu_sharded_nccl.py -m gpu -q`
(the log of such a run is committed under profiles/)."""
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, str(ROOT))
    import torch
    import torch.distributed as dist
    import fhe_ram_b200 as fr
    from fhe_ram_b200.sharded import GpuEngine, ShardedRam, ShardedRamLib
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ws = 2
    p = fr.Parameters.new(device=rank, max_addr=1 << 15, word_size=ws, k_pt=8)
    sk, evk = fr.gen_keys(p)
    keys = fr.EvaluationKeysPrepared.alloc(p).prepare(evk)
    data = fr.Source(5).fill_bytes(p.max_addr() * ws)
    ref = fr.Ram.new(p)                                    # the unsharded RAM: what one GPU computes
    cts = ref.encrypt_sk(data, sk, fr.Source(11), fr.Source(12))
    B = 2 * world                                          # two reads per rank
    idxs = [(9973 * i + 1) % (1 << 15) for i in range(B)]
    idxs[1], idxs[-1] = 4097, (1 << 15) - 1
    xa, xe = fr.Source(21), fr.Source(22)
    addrs = [fr.Address.alloc(p).encrypt_sk(p, i, sk, xa, xe) for i in idxs]
    batch = fr.Address.batch(p, addrs)
    want = ref.read_batch(batch, keys)
    lo, hi = rank * (B // world), (rank + 1) * (B // world)
    fails = []

    # ---- product path: the library's own communicator -------------------------------------------------
    lib = ShardedRamLib(p, rank, world, cts)
    got = lib.read_batch(batch, keys)                      # this rank's slice
    if not np.array_equal(got, want[lo:hi]):
        fails.append("lib read_batch")
    limbs = np.stack([a.data for a in addrs[lo:hi]])
    got = lib.read_batch_host(limbs, hi - lo, keys)
    if not np.array_equal(got, want[lo:hi]):
        fails.append("lib read_batch_host (int64)")
    got = lib.read_batch_host(limbs.astype(np.int32), hi - lo, keys, fmt="i32")
    if not np.array_equal(got.astype(np.int64), want[lo:hi]):
        fails.append("lib read_batch_host (int32)")
    from fhe_ram_b200 import api
    got = lib.read_batch_host(api.pack17(limbs), hi - lo, keys, fmt="p17")
    if not np.array_equal(got.astype(np.int64), want[lo:hi]):
        fails.append("lib read_batch_host (packed 17-bit)")
    a_w = addrs[2 % B]
    rpw = lib.read_prepare_write(a_w, keys)
    want_rpw = ref.read_prepare_write(a_w, keys)
    if not np.array_equal(rpw, want_rpw):
        fails.append("lib read_prepare_write")
    w = np.stack([fr.encrypt_glwe(p, 60 + i, sk) for i in range(ws)])
    lib.write(w if rank == 0 else None, a_w, keys)         # rank 0's word is broadcast
    ref.write(w, a_w, keys)
    want_after = ref.read_batch(batch, keys)
    after = lib.read_batch(batch, keys)
    if not np.array_equal(after, want_after[lo:hi]):
        fails.append("lib read after write")
    # the shards, gathered, are the RAM of one GPU after the same write
    mine = lib.store().reshape(ws, p.n_glwe(), -1)
    full = ref.store().reshape(ws, p.n_glwe(), -1)
    if not np.array_equal(mine[:, rank::world], full[:, rank::world]):
        fails.append("lib RAM limbs after write")
    if 2 % B >= lo and 2 % B < hi:
        for i in range(ws):
            v, noise = fr.decrypt_glwe(p, after[2 % B - lo, i], fr.cast_u8_to_signed(60 + i, 8), sk)
            if v != 60 + i:
                fails.append("decrypt after write")
    lib.close()

    # ---- the halves of the C ABI with torch.distributed doing the exchange ----------------------------
    ref2 = fr.Ram.new(p)
    ref2.load(cts)
    sram = ShardedRam(GpuEngine(p, rank, world, cts), rank, world)
    got = sram.read_batch(batch, keys).reshape(want.shape)
    if not np.array_equal(got, want):
        fails.append("torch read_batch")
    rpw = sram.read_prepare_write(a_w, keys)
    if not np.array_equal(rpw.reshape(want_rpw.shape), ref2.read_prepare_write(a_w, keys)):
        fails.append("torch read_prepare_write")
    sram.write(w, a_w, keys)
    if not np.array_equal(sram.read_batch(batch, keys).reshape(want.shape), want_after):
        fails.append("torch read after write")

    Path(f"{out_dir}/r{rank}.txt").write_text("\n".join(fails))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_reads_and_write_over_nccl(built, tmp_path, world):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < world:
        pytest.skip(f"needs >= {world} GPUs")
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        fails = (tmp_path / f"r{r}.txt").read_text()
        assert fails == "", f"rank {r}: {fails}"
