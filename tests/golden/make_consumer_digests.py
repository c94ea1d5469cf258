"""Regenerates tests/golden/consumer_digests.json: sha256 of the report text every SpeciesIdentifier module restated in
oracle/consumers.py produces on the reference's small Diptera fixture (uncorrected p and K2P, minOverlap 300, threshold 3 %).
The reference stores no expected output for these modules, so the digests pin the ORACLE against drift -- the
GPU host layer is then checked against the same digests (tests/test_host_gpu.py) without the oracle in the loop."""
import gzip
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import consumers as C  # noqa: E402


def digests(text, mode, overlap=300, threshold=3.0):
    seqs = C.sort_by_name(C.read_fasta_text(text))
    D = C.Dist(seqs, C.Config(mode, overlap, True), threads=8)
    sha = lambda s: hashlib.sha256(s.encode("utf-8")).hexdigest()  # noqa: E731
    out = {"n": len(seqs)}
    out["best_match"] = sha(C.best_match(D, threshold)[0])
    out["block_analysis"] = sha(C.block_analysis(D, threshold)[0])
    out["pairwise_summary"] = sha(C.pairwise_summary(D)[0])
    out["extreme_pairwise"] = sha(C.extreme_pairwise(D)[0])
    try:
        out["distance_analysis"] = sha(C.distance_analysis(D)[0])
    except C.JavaNPE:
        out["distance_analysis"] = "NullPointerException"
    clusters, _ = C.cluster(D, threshold / 100)
    out["cluster"] = sha(json.dumps(clusters))
    return out


def main():
    text = gzip.open(os.path.join(HERE, "diptera_small_nogaps.fasta.gz"), "rt", encoding="latin-1").read()
    res = {"diptera_small_nogaps": {"p": digests(text, 0), "k2p": digests(text, 1)}}
    with open(os.path.join(HERE, "consumer_digests.json"), "w") as f:
        json.dump(res, f, indent=1, sort_keys=True)
    print(json.dumps(res, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
