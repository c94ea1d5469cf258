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


def _routed_worker(rank, world, port, n, seg):
    """The torch.distributed calls of bench.py's routed exchange, on gloo, with numpy standing in for the library: records are
    bucketed by the rank that owns their row (owner = row // rows_per_part), one equal segment per destination travels through
    all_to_all_single, the per-row state is SUM-reduced, every rank finalises its own rows and the slices are all-gathered."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rpp = -(-(-(-n // world)) // 128) * 128  # tdna_rows_per_part: ceil(n / world) rounded up to 128
    assert rpp * world >= n
    # every rank's records (row, col, d): what its share of the tiles would have emitted
    per_rank = []
    for r in range(world):
        g = np.random.default_rng(500 + r)
        k = 40 + 17 * r
        per_rank.append((g.integers(0, n, size=k), g.integers(0, n, size=k), g.random(k), g.integers(0, 3, size=n)))
    rows, cols, d, st = per_rank[rank]
    rec = torch.zeros((world, seg, 2), dtype=torch.int64)
    counts = torch.zeros(world, dtype=torch.int64)
    for q, j, dd in zip(rows, cols, d):  # tdna_best_match_part_routed: segment = owner of the row
        dest = int(q) // rpp
        pos = int(counts[dest])
        rec[dest, pos, 0] = (int(q) << 32) | int(j)
        rec[dest, pos, 1] = int(np.float64(dd).view(np.int64))
        counts[dest] += 1
    state = torch.from_numpy(st.astype(np.int64))
    recv = torch.zeros_like(rec)
    recv_counts = torch.zeros_like(counts)
    dist.all_to_all_single(recv_counts, counts)
    dist.all_to_all_single(recv.view(-1), rec.view(-1))
    dist.all_reduce(state, op=dist.ReduceOp.SUM)
    # what arrived: segment s = the records rank s had for MY rows, in rank s's order
    q0, q1 = min(n, rank * rpp), min(n, rank * rpp + rpp)
    got = []
    for s in range(world):
        for p in range(int(recv_counts[s])):
            key, bits = int(recv[s, p, 0]), int(recv[s, p, 1])
            got.append((key >> 32, key & 0xffffffff, bits))
    exp = []
    for s in range(world):
        r_, c_, d_, _ = per_rank[s]
        exp += [(int(q), int(j), int(np.float64(dd).view(np.int64))) for q, j, dd in zip(r_, c_, d_) if q0 <= q < q1]
    assert got == exp
    assert np.array_equal(state.numpy(), sum(p[3] for p in per_rank))
    # tdna_best_match_merge_routed stand-in: one 8-byte "record" per own row = how many candidates it received
    own = torch.zeros(rpp, dtype=torch.int64)
    for q, _, _ in got:
        own[q - q0] += 1
    allrows = torch.zeros(world * rpp, dtype=torch.int64)
    dist.all_gather_into_tensor(allrows, own)
    want = np.zeros(n, dtype=np.int64)
    for r_, _, _, _ in per_rank:
        np.add.at(want, r_, 1)
    assert np.array_equal(allrows.numpy()[:n], want) and not allrows.numpy()[n:].any()
    dist.destroy_process_group()


def test_routed_exchange_choreography_on_three_ranks():
    port = 33500 + (os.getpid() % 2000)
    mp.spawn(_routed_worker, args=(3, port, 700, 256), nprocs=3, join=True)
