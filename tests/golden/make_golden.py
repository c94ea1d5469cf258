"""Regenerates tests/golden/ from the reference's data fixtures (run HERE, where
/root/reference exists; the GPU box only ever sees the committed outputs).

Inputs are the reference's *data* files under /root/reference/Tests (no source
code is copied).  Outputs:
  c1_best_match_fixture.txt      verbatim copy of "Tests/best match/doesnt_add_up_to_100_bug.txt"
  diptera_small_nogaps.fasta.gz  "Tests/files/Diptera COI - small - no gaps.fasta"
  diptera_coi.fasta.gz           "Tests/files/Diptera COI.fasta"
  checksums.json                 oracle checksums over all unordered pairs (the same
                                 figures as SURVEY.md Appendix B, which an independent
                                 numpy restatement produced) + per-config digests
"""
import gzip
import hashlib
import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from oracle import consumers as C  # noqa: E402

REF = "/root/reference/Tests"


def gz_copy(src, dst):
    with open(src, "rb") as f, gzip.GzipFile(dst, "wb", mtime=0) as g:
        shutil.copyfileobj(f, g)


def main():
    shutil.copyfile(os.path.join(REF, "best match", "doesnt_add_up_to_100_bug.txt"),
                    os.path.join(HERE, "c1_best_match_fixture.txt"))
    files = {"diptera_small_nogaps": "Diptera COI - small - no gaps.fasta", "diptera_coi": "Diptera COI.fasta"}
    sums = {}
    for key, fn in files.items():
        src = os.path.join(REF, "files", fn)
        gz_copy(src, os.path.join(HERE, key + ".fasta.gz"))
        seqs = C.read_fasta(src)
        rows, lens = oracle.pack_rows([s.seq for s in seqs])
        iu = np.triu_indices(len(seqs), 1)
        ent = {"n": len(seqs), "L": int(rows.shape[1])}
        for mode, nm in ((0, "p"), (1, "k2p"), (2, "trans")):
            r = oracle.all_pairs(rows, lens, mode, 300, True, threads=8)
            d = r["dist"][iu]
            ent[nm] = {
                "sum_shared": int(r["shared"][iu].sum()), "sum_c1": int(r["c1"][iu].sum()),
                "sum_c2": int(r["c2"][iu].sum()), "sum_c3": int(r["c3"][iu].sum()),
                "valid_pairs": int((~(d < 0)).sum()), "exact_zero": int((d == 0).sum()),
                "max_d": float(np.nanmax(d)), "nan": int(np.isnan(d).sum()),
                "sha256_dist_upper": hashlib.sha256(np.ascontiguousarray(d).tobytes()).hexdigest(),
            }
        sums[key] = ent
    with open(os.path.join(HERE, "checksums.json"), "w") as f:
        json.dump(sums, f, indent=1, sort_keys=True)
    print(json.dumps(sums, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
