"""Multi-GPU plumbing for the row-band decomposition (one process per GPU, torch.distributed).

The pair space is sharded by QUERY ROW: rank r owns the contiguous band row_band(n, r, world) and
computes those rows against every column, so each per-row result is complete on exactly one rank and
the only exchange is one all_gather of the fixed-size per-row records (NCCL over NVLink on GPUs;
gloo in the CPU tests).  Histogram-like outputs (class distances, edges) are concatenated."""
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
