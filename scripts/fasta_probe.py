"""Development probe: the device FASTA parser on an H-sized file (50,000 x 658 as 60-column FASTA), text resident on the device,
next to the host layer's parser (taxondna_b200/host: SequenceList.fromFastaText)."""
import ctypes
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import taxondna_b200 as t
from taxondna_b200 import synth

d = synth.generate("H")
sym = d["symbols"].copy()
sym[sym == ord("_")] = ord("-")
parts = []
for nm, row in zip(d["names"], sym):
    s = row.tobytes().decode()
    parts.append(">" + nm + "\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
text = "".join(parts).encode("latin-1")
ctx = t.Context(0)
L = t.load()
dev = torch.frombuffer(bytearray(text), dtype=torch.uint8).cuda()
for rep in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    h = ctypes.c_void_p()
    t.check(L.tdna_fasta_parse(ctx.h, dev.data_ptr(), len(text), ctypes.byref(h)))
    t1 = time.perf_counter()
    n, w = L.tdna_fasta_count(h), L.tdna_fasta_width(h)
    L.tdna_fasta_destroy(h)
    print("device parse: %d records x %d, %.1f MB in %.3f ms = %.1f GB/s" % (n, w, len(text) / 1e6, (t1 - t0) * 1e3, len(text) / (t1 - t0) / 1e9), flush=True)
t0 = time.perf_counter()
f = t.Fasta(ctx, text)
t1 = time.perf_counter()
print("from host bytes (H2D of the text included): %.3f ms" % ((t1 - t0) * 1e3))
assert np.array_equal(f.symbols(), d["symbols"])
try:
    from taxondna_b200 import _host as H
    t0 = time.perf_counter()
    lst = H.SequenceList.fromFastaText(text.decode("latin-1"))
    t1 = time.perf_counter()
    print("host layer parser (C++, one thread): %d sequences in %.1f ms" % (lst.count(), (t1 - t0) * 1e3))
except Exception as e:  # noqa: BLE001
    print("host parser unavailable:", e)
