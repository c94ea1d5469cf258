"""Development probe: lists whose species blocks are wider than the seed window (hundreds of conspecifics in a row)."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 400
data = synth.generate("H", names=False, genera=25, species=max(1, 2000 // reps), reps=reps)
sym = data["symbols"]
n, L = sym.shape
ctx = t.Context(0)
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(0, 300, True)
out = {}
for name, k in (("auto", t.KERNEL_AUTO), ("popc", t.KERNEL_POPC)):
    ctx.set_option(t.OPT_COUNT_KERNEL, k)
    for rep in range(3):
        aln.set_config(0, 300, True)
        t0 = time.time()
        out[name] = aln.best_match()
        dt = time.time() - t0
    print("%s: n=%d reps=%d  %.2f ms  kernel used %d  tile pass %.2f ms" % (name, n, reps, dt * 1e3, aln.last_count_kernel(), ctx.last_tile_pass_ms()), flush=True)
print("identical:", out["auto"].tobytes() == out["popc"].tobytes())
