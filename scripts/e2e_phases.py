"""Phase breakdown of the host-buffer (e2e) path."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
mode = {"C2": 0, "C3": 1, "H": 0}[cfg]
data = synth.generate(cfg, names=False)
sym = torch.from_numpy(data["symbols"]).pin_memory()
n, L = data["symbols"].shape
labels = [data[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
ctx = t.Context(0)
out = torch.empty((n, n), dtype=torch.float64).pin_memory() if cfg == "C2" else None
res = torch.empty((n, 120), dtype=torch.uint8).pin_memory()
Lb = t.load()
for rep in range(4):
    ts = [time.perf_counter()]
    a = t.Alignment(ctx, sym.numpy()); ts.append(time.perf_counter())
    a.set_labels(*labels); ts.append(time.perf_counter())
    a.set_config(mode, 300, True); ts.append(time.perf_counter())
    if cfg == "C2":
        t.check(Lb.tdna_matrix(a.h, out.data_ptr(), None)); ts.append(time.perf_counter())
    else:
        t.check(Lb.tdna_best_match(a.h, res.data_ptr())); ts.append(time.perf_counter())
    x = a.class_distances(t.PD_INTRA); ts.append(time.perf_counter())
    a.close(); ts.append(time.perf_counter())
    names = ["create", "labels", "config", "matrix/bestmatch", "class_intra", "close"]
    print(cfg, "rep", rep, " ".join("%s=%.3fms" % (nm, (b - a_) * 1e3) for nm, a_, b in zip(names, ts, ts[1:])), "total=%.3fms" % ((ts[-1] - ts[0]) * 1e3), flush=True)
