"""Clusters of four (TDNA_OPT_CLUSTER = 4) against the pairs: same bytes? how fast?"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "tiny"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
kw = {}
if len(sys.argv) > 3:
    kw["scale_n"] = int(sys.argv[3])
d = synth.generate(cfg, names=False, **kw)
ctx = t.Context(0)
ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_TENSOR)
res = {}
for cl in (2, 3, 4, 2, 3):  # 3 = cta_group::2 (the b-side operands are encoded in 128-row panels: a fresh alignment per mode)
    ctx.set_option(t.OPT_CLUSTER, cl)
    aln = t.Alignment(ctx, d["symbols"])
    aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    aln.set_config(mode, 300, True)
    best, steps = 1e9, []
    for rep in range(5):
        t0 = time.perf_counter()
        r = aln.best_match_compute() if hasattr(aln, "best_match_compute") else aln.best_match()
        steps.append(time.perf_counter() - t0)
        best = min(best, ctx.last_tile_pass_ms())
    r = aln.best_match()
    res.setdefault(cl, r.tobytes())
    print("n = %d, cluster mode %d: tile pass %.3f ms, step %.3f ms, identical to pairs: %s" % (aln.n, cl, best, min(steps) * 1e3, r.tobytes() == res[2]), flush=True)
    aln.close()
