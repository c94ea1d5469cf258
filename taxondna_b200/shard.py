"""Host-side helpers for multi-GPU layouts that run OUTSIDE the library (one process per GPU, torch.distributed; NCCL over NVLink on
GPUs, gloo in the CPU tests).  The streaming passes need none of this any more: with a communicator (tdna_comm_init /
tdna_init_multi) the library shares a whole-range call among the ranks itself -- units of 64 tile pairs dealt cyclically, the
candidate records routed to the rank that owns their row.  What is left here:

* matrix-shaped outputs: rank r owns the contiguous row band row_band(n, r, world) against every column; per-row records are
  all_gathered (gather_rows), variable-length outputs concatenated (gather_varlen);
* exchange_candidates: the gathered exchange of the low-level part / merge entry points (tdna_best_match_part ->
  tdna_best_match_merge), kept for hosts that run the exchange themselves and exercised over gloo by tests/test_multigpu_cpu.py."""
import numpy as np

TILE = 128  # row bands start on tile boundaries (include/taxondna_cuda.h: tdna_set_row_range)


def rows_per_rank(n, world):
    per = -(-n // world)
    return -(-per // TILE) * TILE


def row_band(n, rank, world):
    per = rows_per_rank(n, world)
    if rank * per >= n:
        return 0, 0  # nothing left for this rank (row_begin must stay a multiple of 128)
    return rank * per, min(n, (rank + 1) * per)


def gather_rows(local, n, rank, world, group=None):
    """local: torch uint8 tensor [n, record_bytes] whose rows row_band(n, rank, world) are valid.
    Returns the full [n, record_bytes] tensor on every rank."""
    import torch
    import torch.distributed as dist
    per = rows_per_rank(n, world)
    b0, b1 = row_band(n, rank, world)
    rec = local.shape[1]
    mine = torch.zeros((per, rec), dtype=local.dtype, device=local.device)
    mine[: b1 - b0].copy_(local[b0:b1])
    out = torch.empty((world * per, rec), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), mine.view(-1), group=group)
    return out[:n]


def gather_varlen(local, group=None):
    """Concatenates 1-D tensors of different lengths from all ranks (class distances, edge lists)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    cnt = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    cnts = [torch.zeros_like(cnt) for _ in range(world)]
    dist.all_gather(cnts, cnt, group=group)
    m = int(max(int(c.item()) for c in cnts))
    pad = torch.zeros(max(m, 1), dtype=local.dtype, device=local.device)
    pad[: local.numel()] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[: int(c.item())] for o, c in zip(outs, cnts)])


def exchange_candidates(rows, cols, d, inval, flags, group=None):
    """rows/cols: int32 [k], d: float64 [k] -- this rank's candidate records (k may be 0);
    inval: int32 [n] invalid-pair counts, flags: int32 [n] TDNA_ROW_* bits seen by this rank.
    Returns (rows, cols, d, inval, flags) of the union: records concatenated in rank order
    (all_gather), counts summed, flags ORed (all_reduce)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    dev = inval.device
    k = int(rows.numel())
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, torch.tensor([k], dtype=torch.int64, device=dev), group=group)
    inval = inval.clone()
    dist.all_reduce(inval, op=dist.ReduceOp.SUM, group=group)
    bits = torch.stack([flags & 1, (flags >> 1) & 1])  # NCCL has no bitwise-or reduction: OR == max per bit
    dist.all_reduce(bits, op=dist.ReduceOp.MAX, group=group)
    flags = (bits[0] | (bits[1] << 1)).to(torch.int32).contiguous()
    cl = [int(x) for x in counts.tolist()]
    m = max(max(cl), 1)
    rec = torch.zeros((m, 2), dtype=torch.int64, device=dev)  # one padded 16-byte record per candidate
    if k:
        rec[:k, 0] = (rows.to(torch.int64) << 32) | cols.to(torch.int64)
        rec[:k, 1] = d.contiguous().view(torch.int64)
    allrec = torch.empty((world, m, 2), dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(allrec.view(-1), rec.view(-1), group=group)
    pieces = [allrec[w, : cl[w]] for w in range(world) if cl[w]]
    cat = torch.cat(pieces) if pieces else torch.zeros((0, 2), dtype=torch.int64, device=dev)
    urows = (cat[:, 0] >> 32).to(torch.int32).contiguous()
    ucols = (cat[:, 0] & 0xFFFFFFFF).to(torch.int32).contiguous()
    ud = cat[:, 1].contiguous().view(torch.float64)
    return urows, ucols, ud, inval.contiguous(), flags
