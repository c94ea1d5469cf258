"""C5 (BASELINE.json configs[4]) on ONE B200: 1,000,000 x 658 synthetic COI, 5e11 pairs, streaming per-row Best Match +
intraspecific distances from one pass (tdna_analyze); kernel (b).  Prints one JSON line."""
import ctypes
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

t0 = time.time()
data = synth.generate("C5", names=False)
sym = data["symbols"]
n, L = sym.shape
t_gen = time.time() - t0
ctx = t.Context(0)
t0 = time.time()
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(0, 300, True)
t_create = time.time() - t0
Lb = t.load()
c64 = ctypes.c_int64()
rows = np.zeros(n, dtype=t.ROW_RESULT_DTYPE)
# intraspecific pairs: 100,000 species x C(10, 2)
cap = 100000 * 45 * 2
intra = torch.empty(cap, dtype=torch.float32, device="cuda")
times = []
for rep in range(2):
    io = t.Analysis()
    io.want = t.WANT_BEST_MATCH | t.WANT_INTRA
    io.rows = rows.ctypes.data
    io.intra, io.intra_cap = intra.data_ptr(), cap
    t0 = time.time()
    t.check(Lb.tdna_analyze(aln.h, ctypes.byref(io)))
    times.append(time.time() - t0)
pairs = n * (n - 1) // 2
line = {"workload": "C5: 1,000,000 x 658 bp synthetic COI, uncorrected p, streaming per-row Best Match + intraspecific distances, 1 x B200",
        "n": n, "L": L, "pairs": pairs, "count_kernel": "tensor" if aln.last_count_kernel() == t.KERNEL_TENSOR else "popc",
        "seconds": times, "tile_pass_ms": ctx.last_tile_pass_ms(), "gpairs_per_s": pairs / min(times) / 1e9,
        "n_intra": int(io.n_intra), "rows_with_match": int((rows["best"] >= 0).sum()),
        "rows_best_is_conspecific": int((data["species_id"][np.maximum(rows["best"], 0)] == data["species_id"]).sum()),
        "near_tie_rows": int((rows["near_best"] > 0).sum()), "generate_s": t_gen, "create_s": t_create}
print(json.dumps(line), flush=True)
