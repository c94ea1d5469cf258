"""Development check: streaming Best Match at n = 200,000 (2e10 pairs) -- kernel (b) against kernel (a), byte for byte.
usage: big_n_check.py [n] [config] [mode]   (config C4: ambiguity letters, gaps, '?'; mode 1: K2P)"""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
cfg = sys.argv[2] if len(sys.argv) > 2 else "H"
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
data = synth.generate(cfg, names=False, scale_n=n_target)
sym = data["symbols"]
n, L = sym.shape
ctx = t.Context(0)
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(mode, 300, True)
ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
out = {}
for name, k in (("tensor", t.KERNEL_TENSOR), ("popc", t.KERNEL_POPC)):
    ctx.set_option(t.OPT_COUNT_KERNEL, k)
    aln.best_match()
    t0 = time.time()
    out[name] = aln.best_match()
    dt = time.time() - t0
    print("%s: n=%d pairs=%.3e  %.1f ms  %.1f Gpairs/s (incl. D2H)  tile pass %.2f ms" % (name, n, n * (n - 1) / 2, dt * 1e3, n * (n - 1) / 2 / dt / 1e9, ctx.last_tile_pass_ms()), flush=True)
same = out["tensor"].tobytes() == out["popc"].tobytes()
print("records identical:", same)
assert same
