"""Rows only vs rows + both class histograms in one pass, C5-shaped at n = 200,000 (for an ncu launch list)."""
import ctypes
import sys
import time

import numpy as np

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

n_target = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
d = synth.generate("C5", names=False, scale_n=n_target)
ctx = t.Context(0)
aln = t.Alignment(ctx, d["symbols"])
aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
aln.set_config(0, 300, True)
n = aln.n
hi = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
he = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
wants = (t.WANT_BEST_MATCH, t.WANT_BEST_MATCH | t.WANT_HIST_INTRA | t.WANT_HIST_INTER, t.WANT_HIST_INTRA | t.WANT_HIST_INTER)
if len(sys.argv) > 2:
    wants = (int(sys.argv[2]),)
for want in wants:
    for rep in range(2):
        io = t.Analysis()
        io.want = want
        io.hist_intra, io.hist_intra_cap = hi.ctypes.data, hi.size
        io.hist_inter, io.hist_inter_cap = he.ctypes.data, he.size
        aln.set_config(0, 300, True)
        t0 = time.time()
        t.check(ctx.L.tdna_analyze(aln.h, ctypes.byref(io)))
        dt = time.time() - t0
    print("want=%d  %.2f ms  tile pass %.2f ms  inter pushes %d" % (want, dt * 1e3, ctx.last_tile_pass_ms(), io.hist_inter_pushes), flush=True)
