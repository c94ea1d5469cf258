"""tdna_row_extremes / tdna_pairs_between at bench sizes: wall time of the call (records back on the host)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

for cfg, mode in (("H", 0), ("Hg", 0), ("C3", 1)):
    d = synth.generate(cfg, names=False)
    ctx = t.Context(0)
    aln = t.Alignment(ctx, d["symbols"])
    aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    aln.set_config(mode, 300, True)
    best = 1e9
    for rep in range(3):
        t0 = time.time()
        r = aln.row_extremes()
        best = min(best, time.time() - t0)
    G, S, R, L = d["shape"]
    evals = aln.n * S * R
    undecided = int(((r["near_intra"] > 0) | (r["near_inter"] > 0) | ((r["flags"] & (t.ROW_NONFINITE | t.EXT_TOO_MANY)) != 0)).sum())
    sum_undecided = int(((r["near_inter"] > 0) | (r["near_order"] > 0) | ((r["flags"] & (t.ROW_NONFINITE | t.EXT_TOO_MANY)) != 0)).sum())
    print("%s mode %d: n = %d, %d pair evaluations in %.2f ms (%.1f Gpairs/s); rows left to the exact sort: ExtremePairwise %d, DistanceAnalysis %d"
          % (cfg, mode, aln.n, evals, best * 1e3, evals / best / 1e9, undecided, sum_undecided), flush=True)
    aln.close()
    ctx.close()
