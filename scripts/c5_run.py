"""C5 (BASELINE.json configs[4]): 1,000,000 x 658 synthetic COI, 5e11 pairs -- streaming per-row Best Match + the intraspecific and
congeneric-interspecific histograms from ONE pass, on 1 GPU (python scripts/c5_run.py) or N (torchrun ... scripts/c5_run.py: the
ranks share the pass inside the library, tdna_comm_init).  Prints one JSON line."""
import ctypes
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=dev)
t0 = time.time()
data = synth.generate("C5", names=False)
sym = data["symbols"]
n, L = sym.shape
t_gen = time.time() - t0
ctx = t.Context(local)
if world > 1:
    ctx.comm_init_torch(dist, dev)
t0 = time.time()
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(0, 300, True)
t_create = time.time() - t0
Lb = t.load()
rows = np.zeros(n, dtype=t.ROW_RESULT_DTYPE)
hi = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
he = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
times, rows_only = [], []
for rep in range(3):
    io = t.Analysis()
    io.want = t.WANT_BEST_MATCH | t.WANT_HIST_INTRA | t.WANT_HIST_INTER
    io.rows = rows.ctypes.data
    io.hist_intra, io.hist_intra_cap = hi.ctypes.data, hi.size
    io.hist_inter, io.hist_inter_cap = he.ctypes.data, he.size
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    t.check(Lb.tdna_analyze(aln.h, ctypes.byref(io)))
    times.append(time.time() - t0)
for rep in range(2):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    aln.set_config(0, 300, True)
    aln.best_match_compute()
    rows_only.append(time.time() - t0)
tt = torch.tensor([min(times), min(rows_only)], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
pairs = n * (n - 1) // 2
digest = hashlib.sha256(rows.tobytes()).hexdigest()
if rank == 0:
    line = {"workload": "C5: 1,000,000 x 658 bp synthetic COI, uncorrected p, streaming per-row Best Match + intra / congeneric-inter histograms in one pass",
            "n_gpus": world, "n": n, "L": L, "pairs": pairs, "count_kernel": "tensor" if aln.last_count_kernel() == t.KERNEL_TENSOR else "popc",
            "seconds_rows_and_histograms": float(tt[0]), "seconds_rows_only_device": float(tt[1]), "tile_pass_ms_rank0": ctx.last_tile_pass_ms(),
            "gpairs_per_s_rows_and_histograms": pairs / float(tt[0]) / 1e9, "gpairs_per_s_rows_only": pairs / float(tt[1]) / 1e9,
            "hist_intra": {"distinct": int(io.n_hist_intra), "pushes": int(io.hist_intra_pushes)},
            "hist_inter": {"distinct": int(io.n_hist_inter), "pushes": int(io.hist_inter_pushes)},
            "rows_with_match": int((rows["best"] >= 0).sum()),
            "rows_best_is_conspecific": int((data["species_id"][np.maximum(rows["best"], 0)] == data["species_id"]).sum()),
            "near_tie_rows": int((rows["near_best"] > 0).sum()), "rows_sha256": digest, "generate_s": t_gen, "create_s": t_create,
            "phases_ms_rank0": ctx.multi_phase_ms() if world > 1 else None}
    print(json.dumps(line), flush=True)
aln.close()
if world > 1:
    ctx.comm_destroy()
ctx.close()
if world > 1:
    dist.destroy_process_group()
