"""The N>1 host logic on CPU: two gloo ranks each own a row band, exchange their per-row records
with the same all_gather the NCCL path uses, and must reassemble exactly the single-rank answer."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _worker(rank, world, port, n, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from taxondna_b200 import shard
    rec = 120
    full = np.load(os.path.join(tmp, "full.npy"))
    b0, b1 = shard.row_band(n, rank, world)
    local = np.zeros_like(full)
    local[b0:b1] = full[b0:b1]  # this rank only knows its own band
    got = shard.gather_rows(torch.from_numpy(local), n, rank, world)
    assert got.shape == (n, rec) and np.array_equal(got.numpy(), full)
    # variable-length pieces (class distances / edges): concatenation in rank order
    piece = torch.arange(rank * 1000, rank * 1000 + 7 * (rank + 1), dtype=torch.float32)
    allp = shard.gather_varlen(piece)
    exp = np.concatenate([np.arange(r * 1000, r * 1000 + 7 * (r + 1), dtype=np.float32) for r in range(world)])
    assert np.array_equal(allp.numpy(), exp)
    # streaming Best Match exchange: records concatenated in rank order, counts summed, flags ORed
    rng = np.random.default_rng(100 + rank)
    k = 0 if rank == 1 and world > 2 else 5 + 3 * rank  # one rank may have nothing to contribute
    rows = torch.from_numpy(rng.integers(0, n, size=k).astype(np.int32))
    cols = torch.from_numpy(rng.integers(0, n, size=k).astype(np.int32))
    d = torch.from_numpy(rng.random(k))
    inval = torch.from_numpy(rng.integers(0, 4, size=n).astype(np.int32))
    flags = torch.from_numpy(rng.integers(0, 4, size=n).astype(np.int32))
    ur, uc, ud, ui, uf = shard.exchange_candidates(rows, cols, d, inval, flags)
    er, ec, ed, ei, ef = [], [], [], np.zeros(n, np.int64), np.zeros(n, np.int64)
    for r in range(world):
        g = np.random.default_rng(100 + r)
        kk = 0 if r == 1 and world > 2 else 5 + 3 * r
        er.append(g.integers(0, n, size=kk).astype(np.int32))
        ec.append(g.integers(0, n, size=kk).astype(np.int32))
        ed.append(g.random(kk))
        ei += g.integers(0, 4, size=n)
        ef |= g.integers(0, 4, size=n)
    assert np.array_equal(ur.numpy(), np.concatenate(er)) and np.array_equal(uc.numpy(), np.concatenate(ec))
    assert np.array_equal(ud.numpy().view(np.uint64), np.concatenate(ed).view(np.uint64))
    assert np.array_equal(ui.numpy(), ei) and np.array_equal(uf.numpy(), ef)
    dist.destroy_process_group()


def test_row_bands_cover_the_rows_exactly():
    from taxondna_b200 import shard
    for n in (1, 127, 128, 129, 2185, 8000, 50000, 1000000):
        for world in (1, 2, 3, 4, 8):
            bands = [b for b in (shard.row_band(n, r, world) for r in range(world)) if b[1] > b[0]]
            assert bands[0][0] == 0 and bands[-1][1] == n
            for (a0, a1), (b0, b1) in zip(bands, bands[1:]):
                assert a1 == b0 and a0 < a1
            assert all(b[0] % 128 == 0 for b in bands)


def test_two_rank_gather_reassembles_single_rank_answer(tmp_path):
    n, world = 300, 2
    rng = np.random.default_rng(0)
    full = rng.integers(0, 256, size=(n, 120), dtype=np.uint8)
    np.save(os.path.join(tmp_path, "full.npy"), full)
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)


def test_three_rank_exchange_with_an_empty_rank(tmp_path):
    n, world = 200, 3
    rng = np.random.default_rng(1)
    np.save(os.path.join(tmp_path, "full.npy"), rng.integers(0, 256, size=(n, 120), dtype=np.uint8))
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, n, str(tmp_path)), nprocs=world, join=True)
