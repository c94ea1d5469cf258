"""Host-side phases of tdna_analyze_host when a communicator shares the pass (torchrun, one rank per GPU)."""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
d = synth.generate("H", names=False)
sym = torch.from_numpy(d["symbols"]).pin_memory()
labels = [d[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
ctx = t.Context(local)
ctx.comm_init_torch(dist, dev)
out = torch.empty((sym.shape[0], 120), dtype=torch.uint8).pin_memory()
res = out.numpy().view(t.ROW_RESULT_DTYPE).reshape(-1)
for rep in range(8):
    dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t.best_match_host(ctx, sym.numpy(), labels, 0, 300, out=res)
    dt = (time.perf_counter() - t0) * 1e3
    ph = ctx.host_call_phase_ms()
    allv = [None] * dist.get_world_size()
    dist.all_gather_object(allv, (dt, ph))
    if rank == 0 and rep >= 2:
        print("rep %d: max %.3f ms min %.3f ms; rank 0 phases %s" % (rep, max(v[0] for v in allv), min(v[0] for v in allv), {k: v for k, v in ph.items() if v >= 0}), flush=True)
        print("        multi phases (ms): %s" % {k: round(v, 3) for k, v in ctx.multi_phase_ms().items()}, flush=True)
ctx.close()
dist.destroy_process_group()
