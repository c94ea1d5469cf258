"""Development probe (torchrun): where a multi-GPU streaming Best-Match step (the packed exchange bench.py uses) spends its time."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

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
seg = max(1 << 14, (48 * n) // (world * world))
rpp = aln.rows_per_part(world)
RS = t.ROW_RESULT_DTYPE.itemsize
xb = {"rec": torch.empty((world, seg, 2), dtype=torch.int64, device=dev), "counts": torch.zeros(world, dtype=torch.int64, device=dev),
      "recv": torch.empty((world, seg, 2), dtype=torch.int64, device=dev), "recv_counts": torch.zeros(world, dtype=torch.int64, device=dev),
      "state": torch.empty(n, dtype=torch.int64, device=dev),
      "slice": torch.empty(rpp * RS, dtype=torch.uint8, device=dev), "rows": torch.empty(world * rpp * RS, dtype=torch.uint8, device=dev)}


class Dev:
    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def sync():
    torch.cuda.synchronize()
    return time.perf_counter()


for rep in range(6):
    aln.set_config(0, 300, True)
    dist.barrier()
    ts = [sync()]
    mins = aln.best_match_seed(rank, world); ts.append(sync())
    with torch.cuda.stream(stream):
        mt = torch.as_tensor(Dev(mins, 3 * n, "<i8"), device=dev)
        dist.all_reduce(mt, op=dist.ReduceOp.MIN)
    ts.append(sync())
    aln.best_match_part_routed(rank, world, True, xb["rec"].data_ptr(), seg, xb["counts"].data_ptr(), xb["state"].data_ptr()); ts.append(sync())
    tile_ms = ctx.last_tile_pass_ms()
    with torch.cuda.stream(stream):
        dist.all_to_all_single(xb["recv_counts"], xb["counts"])
        dist.all_to_all_single(xb["recv"].view(-1), xb["rec"].view(-1))
        dist.all_reduce(xb["state"], op=dist.ReduceOp.SUM)
    ts.append(sync())
    aln.best_match_merge_routed(rank, world, xb["recv"].data_ptr(), seg, xb["recv_counts"].data_ptr(), xb["state"].data_ptr(), xb["slice"].data_ptr()); ts.append(sync())
    with torch.cuda.stream(stream):
        dist.all_gather_into_tensor(xb["rows"], xb["slice"])
    ts.append(sync())
    if rank == 0:
        names = ["seed", "allreduce_min", "part_routed", "all_to_all", "merge_own_rows", "all_gather_rows"]
        print("rep", rep, "n", n, " ".join("%s=%.3f" % (nm, (b - a) * 1e3) for nm, a, b in zip(names, ts, ts[1:])),
              "total=%.3f ms (tile pass %.3f), records/rank %d, seg %d" % ((ts[-1] - ts[0]) * 1e3, tile_ms, int(xb["counts"].sum().item()), seg), flush=True)
dist.destroy_process_group()
