"""Development probe: candidates and timings of the streaming pass per count kernel."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "H"
data = synth.generate(cfg, names=False)
sym = data["symbols"]
n, L = sym.shape
ctx = t.Context(0)
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(0, 300, True)
for kern in (t.KERNEL_POPC, t.KERNEL_TENSOR):
    ctx.set_option(t.OPT_COUNT_KERNEL, kern)
    for rep in range(3):
        t0 = time.time()
        p = aln.best_match_part(0, 1)
        t1 = time.time()
        ms = ctx.last_tile_pass_ms()
        r = aln.best_match()
        t2 = time.time()
        print("kernel %d rep %d: part %.2f ms (tile pass %.2f ms), n_cand %d, full best_match %.2f ms" % (kern, rep, (t1 - t0) * 1e3, ms, p.n_cand, (t2 - t1) * 1e3), flush=True)
