"""Quick on-GPU measurements used while developing (not the bench contract)."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

ctx = t.Context(0)
out = {}
popc, lop3 = ctx.measure_int_peaks()
out["popc_per_s"] = popc
out["lop3_per_s"] = lop3
print("int peaks: popc %.3e /s  lop3 %.3e /s" % (popc, lop3), flush=True)
for cfg, mode in (("C2", 0), ("C3", 1), ("C3", 0), ("H", 0), ("H", 1)):
    if len(sys.argv) > 1 and cfg not in sys.argv[1:]:
        continue
    data = synth.generate(cfg, names=False)
    sym = data["symbols"]
    n, L = sym.shape
    t0 = time.time()
    aln = t.Alignment(ctx, sym)
    aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
    aln.set_config(mode, 300, True)
    t_create = time.time() - t0
    ms = aln.time_matrix_kernel(5)
    pairs = n * (n - 1) / 2
    t0 = time.time()
    aln.best_match_compute()
    res = aln.best_match()
    t_bm = time.time() - t0
    print("%s mode %d: n=%d L=%d  matrix kernel %.3f ms -> %.2f Gpairs/s;  create %.2fs; best_match (matrix cached) %.3fs; near_best rows %d"
          % (cfg, mode, n, L, ms, pairs / ms / 1e6, t_create, t_bm, int((res["near_best"] > 0).sum())), flush=True)
    out["%s_mode%d" % (cfg, mode)] = dict(n=n, L=L, ms=ms, gpairs=pairs / ms / 1e6)
    aln.close()
json.dump(out, open("gpurun_out/quick_gpu.json", "w"), indent=1)
