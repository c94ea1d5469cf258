"""Host-side phases of tdna_analyze_host (the pass pipelined with the upload) on a bench workload."""
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "H"
mode = int(sys.argv[2]) if len(sys.argv) > 2 else 0
d = synth.generate(cfg, names=False)
sym = torch.from_numpy(d["symbols"]).pin_memory()
labels = [d[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
ctx = t.Context(0)
if len(sys.argv) > 3:
    ctx.set_option(t.OPT_PIPELINE_CHUNKS, int(sys.argv[3]))
out = torch.empty((sym.shape[0], 120), dtype=torch.uint8).pin_memory()
res = out.numpy().view(t.ROW_RESULT_DTYPE).reshape(-1)
for rep in range(int(sys.argv[4]) if len(sys.argv) > 4 else 6):
    t0 = time.perf_counter()
    t.best_match_host(ctx, sym.numpy(), labels, mode, 300, out=res)
    dt = (time.perf_counter() - t0) * 1e3
    print("%s mode %d: %.3f ms  pipelined %d  %s" % (cfg, mode, dt, ctx.pipelined_passes(), ctx.host_call_phase_ms()), flush=True)
ctx.close()
