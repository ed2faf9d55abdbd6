"""-m gpu, needs >= 2 GPUs: the sharded path end to end over NCCL (one process per GPU)."""
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
    from fhe_ram_b200.sharded import GpuEngine, ShardedRam
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    p = fr.Parameters.new(device=rank, max_addr=1 << 15, word_size=2, k_pt=8)
    sk, evk = fr.gen_keys(p)
    keys = fr.EvaluationKeysPrepared.alloc(p).prepare(evk)
    data = fr.Source(5).fill_bytes(p.max_addr() * 2)
    tmp = fr.Ram.new(p)
    cts = tmp.encrypt_sk(data, sk, fr.Source(11), fr.Source(12))
    idxs = [1, 4097, 20000, (1 << 15) - 1]
    xa, xe = fr.Source(21), fr.Source(22)
    addrs = [fr.Address.alloc(p).encrypt_sk(p, i, sk, xa, xe) for i in idxs]
    batch = fr.Address.batch(p, addrs)
    want = tmp.read_batch(batch, keys)
    stream = torch.cuda.ExternalStream(p.stream(), device=rank)
    with torch.cuda.stream(stream):
        sram = ShardedRam(GpuEngine(p, rank, world, cts), rank, world)
        got = sram.read_batch(batch, keys).reshape(want.shape)
        rpw = sram.read_prepare_write(addrs[2], keys)
        w = np.stack([fr.encrypt_glwe(p, 60 + i, sk) for i in range(2)])
        sram.write(w, addrs[2], keys)
        after = sram.read_batch(batch, keys).reshape(want.shape)
    want_rpw = tmp.read_prepare_write(addrs[2], keys)
    tmp.write(w, addrs[2], keys)
    want_after = tmp.read_batch(batch, keys)
    ok = (np.array_equal(got, want) and np.array_equal(rpw.reshape(want_rpw.shape), want_rpw)
          and np.array_equal(after, want_after))
    for i in range(2):
        v, noise = fr.decrypt_glwe(p, after[2, i], fr.cast_u8_to_signed(60 + i, 8), sk)
        ok = ok and v == 60 + i
    np.save(f"{out_dir}/ok{rank}.npy", np.array([int(ok)]))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_reads_and_write_over_nccl(built, tmp_path):
    import torch
    import torch.multiprocessing as mp
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        assert np.load(tmp_path / f"ok{r}.npy")[0] == 1, f"rank {r} mismatch"
