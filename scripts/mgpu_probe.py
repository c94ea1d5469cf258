"""Development probe (torchrun): where a multi-GPU streaming Best-Match step spends its time."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import shard, synth

world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
base = synth.CONFIGS["H"]
n1 = base["genera"] * base["species"] * base["reps"]
data = synth.generate("H", names=False, scale_n=n1 * world ** 0.5)
sym = data["symbols"]
n, L = sym.shape
ctx = t.Context(local)
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(0, 300, True)
stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)


class Dev:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(6):
    dist.barrier()
    ts = [sync()]
    mins = aln.best_match_seed(rank, world); ts.append(sync())
    with torch.cuda.stream(stream):
        mt = torch.as_tensor(Dev(mins, 3 * n, "<i8"), device=dev)
        dist.all_reduce(mt, op=dist.ReduceOp.MIN)
    ts.append(sync())
    part = aln.best_match_part(rank, world, seeded=True); ts.append(sync())
    k = int(part.n_cand)
    with torch.cuda.stream(stream):
        rows = torch.as_tensor(Dev(part.cand_row, k, "<i4"), device=dev)
        cols = torch.as_tensor(Dev(part.cand_col, k, "<i4"), device=dev)
        d = torch.as_tensor(Dev(part.cand_d, k, "<f8"), device=dev)
        inval = torch.as_tensor(Dev(part.invalid_count, n, "<i4"), device=dev)
        flags = torch.as_tensor(Dev(part.flags, n, "<i4"), device=dev)
        ur, uc, ud, ui, uf = shard.exchange_candidates(rows, cols, d, inval, flags)
    ts.append(sync())
    res = aln.best_match_merge(ur.data_ptr(), uc.data_ptr(), ud.data_ptr(), ur.numel(), ui.data_ptr(), uf.data_ptr()); ts.append(sync())
    if rank == 0:
        names = ["seed", "allreduce_min", "bulk", "exchange", "merge+d2h"]
        print("rep", rep, " ".join("%s=%.3f" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, ts, ts[1:])), "total=%.3f ms, cand/rank %d" % ((ts[-1] - ts[0]) * 1e3, k), flush=True)
dist.destroy_process_group()
