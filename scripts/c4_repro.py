"""Development probe: class-distance size query on C4 (ambiguity codes -> popc path)."""
import ctypes
import sys

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

kern = int(sys.argv[1]) if len(sys.argv) > 1 else 0
cfg = sys.argv[2] if len(sys.argv) > 2 else "C4"
data = synth.generate(cfg, names=False)
sym = data["symbols"]
ctx = t.Context(0)
ctx.set_option(t.OPT_COUNT_KERNEL, kern)
aln = t.Alignment(ctx, sym)
aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
aln.set_config(0, 300, True)
L = t.load()
c64 = ctypes.c_int64()
print("size query", flush=True)
rc = L.tdna_class_distances(aln.h, t.PD_INTRA, None, 0, ctypes.byref(c64))
print("rc", rc, "n_intra", c64.value, "kernel", aln.last_count_kernel(), flush=True)
x = aln.class_distances(t.PD_INTRA)
print("intra", len(x), float(np.sort(x)[-1]), flush=True)
r = aln.best_match()
print("best match ok", r.shape, flush=True)
