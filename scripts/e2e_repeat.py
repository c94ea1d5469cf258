"""Looks for slow tdna_analyze_host calls under the conditions bench.py creates (a resident alignment, NVML initialised, ...)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth
what = sys.argv[1] if len(sys.argv) > 1 else "plain"
d = synth.generate("H", names=False)
sym = torch.from_numpy(d["symbols"]).pin_memory()
labels = [d[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
ctx = t.Context(0)
out = torch.empty((sym.shape[0], 120), dtype=torch.uint8).pin_memory()
res = out.numpy().view(t.ROW_RESULT_DTYPE).reshape(-1)
if "resident" in what:
    keep = t.Alignment(ctx, sym.numpy()); keep.set_labels(*labels); keep.set_config(0, 300, True); keep.best_match()
if "nvml" in what:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    for _ in range(20): pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
if "flush" in what:
    buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(5): buf.fill_(1)
    torch.cuda.synchronize()
if "mm" in what:
    a8 = torch.randint(-5, 5, (8192, 8192), dtype=torch.int8, device="cuda")
    for _ in range(10): torch._int_mm(a8, a8)
    torch.cuda.synchronize()
for _ in range(3):
    a = t.Alignment(ctx, sym.numpy()); a.set_labels(*labels); a.set_config(0, 300, True); a.best_match_into(out.data_ptr()); a.close()
ts = []
for rep in range(60):
    t0 = time.perf_counter()
    t.best_match_host(ctx, sym.numpy(), labels, 0, 300, out=res)
    ts.append((time.perf_counter() - t0) * 1e3)
print("%s: one-call ms: min %.3f median %.3f max %.3f mean %.3f; slow calls (> 4.5 ms): %d" % (what, min(ts), sorted(ts)[len(ts)//2], max(ts), sum(ts)/len(ts), sum(x > 4.5 for x in ts)), flush=True)
