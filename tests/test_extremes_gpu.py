"""tdna_row_extremes: the per-row records ExtremePairwise.run (SpeciesIdentifier/ExtremePairwise.java:
99-170) and DistanceAnalysis.run (SpeciesIdentifier/DistanceAnalysis.java:183-230) read, through the C ABI against the oracle's
exact restatement (whole-list sortAgainst under the reference's comparator, TimSort), and the host layer's use of them."""
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import consumers as C

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def synth_list(cfg, **kw):
    from taxondna_b200 import synth
    d = synth.generate(cfg, **kw)
    sym = d["symbols"].copy()
    sym[sym == ord("_")] = ord("-")  # what a file holds; normalisation turns the external runs back into '_'
    oseqs = [C.OSeq(nm, row.tobytes().decode()) for nm, row in zip(d["names"], sym)]
    return d, oseqs


def oracle_rows(D):
    """Per query, from the exact sorted list: (largest-intra position, smallest-inter position, [inter distances in sorted order])."""
    seqs = D.seqs
    out = []
    for q, seq in enumerate(seqs):
        if seq.species_name() is None:
            out.append(None)
            continue
        ss = C.sort_against(D, q)
        largest = smallest = None
        inters = []
        for x in ss:
            s2 = seqs[x]
            if s2.species_name() is None:
                continue
            if seq.genus and s2.genus == seq.genus and s2.species_name() != seq.species_name():
                if smallest is None:
                    smallest = x
                inters.append(D.pairwise(x, q))
        for k in range(len(ss) - 1, 0, -1):
            x = ss[k]
            if seqs[x].species_name() is not None and seqs[x].species_name() == seq.species_name():
                largest = x
                break
        out.append((largest, smallest, inters))
    return out


@pytest.mark.parametrize("cfg,kw,mode,overlap", [
    ("C3", dict(genera=5, species=4, reps=6), 0, 300),
    ("C3", dict(genera=5, species=4, reps=6), 1, 300),
    ("C3", dict(genera=5, species=4, reps=6), 2, 300),
    ("C4", dict(genera=4, species=4, reps=5, L=700), 0, 300),
    ("C4", dict(genera=4, species=4, reps=5, L=700), 1, 420),  # many pairs under minOverlap
    ("C3", dict(genera=2, species=1, reps=9), 0, 300),         # one species per genus: no congeneric entries
])
def test_row_extremes_against_the_exact_sort(cfg, kw, mode, overlap):
    import taxondna_b200 as t
    d, oseqs = synth_list(cfg, seed=23, **kw)
    D = C.Dist(oseqs, C.Config(mode, overlap, True), threads=8)
    exp = oracle_rows(D)
    ctx = t.Context(0)
    aln = t.Alignment(ctx, d["symbols"])
    try:
        aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
        aln.set_config(mode, overlap, True)
        rec = aln.row_extremes()
    finally:
        aln.close()
        ctx.close()
    decided = 0
    for q, e in enumerate(exp):
        r = rec[q]
        if e is None:
            assert r["flags"] & t.EXT_UNNAMED
            continue
        largest, smallest, inters = e
        assert not (r["flags"] & (t.EXT_UNNAMED | t.EXT_TOO_MANY))
        self_valid = not (D.pairwise(q, q) < 0)
        assert bool(r["flags"] & t.ROW_SELF_VALID) == self_valid
        finite = [x for x in inters if x == x and x != float("inf")]
        if r["flags"] & t.ROW_NONFINITE:
            continue
        assert r["n_inter"] == len(inters)
        if r["near_intra"] == 0 and (self_valid or r["n_intra"] + r["n_inter"] == 0):
            assert r["max_intra"] == (-1 if largest is None else largest), q
            if largest is not None:
                assert np.float64(r["d_max_intra"]).view(np.uint64) == np.float64(D.pairwise(largest, q)).view(np.uint64)
                assert r["shared_intra"] == D.shared_length(largest, q)
        if r["near_inter"] == 0:
            assert r["min_inter"] == (-1 if smallest is None else smallest), q
            if smallest is not None:
                assert np.float64(r["d_min_inter"]).view(np.uint64) == np.float64(D.pairwise(smallest, q)).view(np.uint64)
                assert r["shared_inter"] == D.shared_length(smallest, q)
        if r["near_inter"] == 0 and r["near_order"] == 0 and inters:
            tot = 0.0
            for x in finite:
                tot += x
            assert np.float64(r["sum_inter"]).view(np.uint64) == np.float64(tot).view(np.uint64), q
            decided += 1
    if cfg == "C3" and kw.get("species", 0) > 1 and mode != 2:
        assert decided > 0  # the device records carry real rows on clean data (transversion-only distances tie all the time)


def test_host_modules_use_the_records_and_agree_with_the_exact_sort():
    from taxondna_b200 import _host as H
    from taxondna_b200 import synth
    d = synth.generate("C3", seed=29, genera=6, species=5, reps=6)
    H.Sequence.setPairwiseDistanceMethod(0)
    H.Sequence.setMinOverlap(300)
    H.Sequence.ambiguousBasesAllowed(True)
    lst = H.SequenceList([H.Sequence(nm, row.tobytes().decode()) for nm, row in zip(d["names"], d["symbols"])])
    n = lst.count()
    ep = H.ExtremePairwise(lst)
    fast = ep.run()
    assert ep.rows_resolved_on_host < n  # most queries never see a sort
    launches = H.launchCount()
    ep.force_exact = True
    assert ep.run() == fast
    assert H.launchCount() - launches >= n  # ... which costs a row launch per query
    da = H.DistanceAnalysis(lst)
    fast = da.run()
    assert da.rows_resolved_on_host < n
    da.force_exact = True
    assert da.run() == fast


def test_a_genus_beyond_the_kernels_capacity_is_left_to_the_host():
    """More than 4,096 rows in one genus: the record says so (TDNA_EXT_TOO_MANY) and the host module sorts such queries exactly."""
    import taxondna_b200 as t
    rng = np.random.default_rng(41)
    n, L = 4200, 64
    root = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    sym = np.tile(root, (n, 1))
    m = rng.random(sym.shape) < 0.1
    sym[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
    species = (np.arange(n) // 7).astype(np.int32)
    genus = np.zeros(n, dtype=np.int32)
    genus[-50:] = 1  # a second, small genus
    ctx = t.Context(0)
    aln = t.Alignment(ctx, sym)
    try:
        aln.set_labels(species, genus, species, np.arange(n, dtype=np.int32))
        aln.set_config(0, 30, True)
        rec = aln.row_extremes()
        d = aln.matrix()
    finally:
        aln.close()
        ctx.close()
    assert ((rec["flags"][:-50] & t.EXT_TOO_MANY) != 0).all()
    small = rec[-50:]
    assert ((small["flags"] & t.EXT_TOO_MANY) == 0).all()
    for q in range(n - 50, n):  # the small genus is decided on the device: check it against the matrix
        r = rec[q]
        cons = [j for j in range(n - 50, n) if j != q and species[j] == species[q] and d[q, j] >= 0]
        inter = [j for j in range(n - 50, n) if species[j] != species[q] and d[q, j] >= 0]
        assert r["n_intra"] == len(cons) and r["n_inter"] == len(inter)
        if cons:
            assert r["d_max_intra"] == max(d[q, j] for j in cons)
        if inter:
            assert r["d_min_inter"] == min(d[q, j] for j in inter)


def test_row_extremes_edge_cases():
    """A single row, rows without species names, a species alone in its genus."""
    import taxondna_b200 as t
    ctx = t.Context(0)
    try:
        sym = np.frombuffer(b"ACGTACGTACGT" * 3, dtype=np.uint8).reshape(3, 12).copy()
        sym[1, 0] = ord("T")
        sym[2, 0:2] = np.frombuffer(b"GG", dtype=np.uint8)
        a = t.Alignment(ctx, sym[:1])
        a.set_labels(np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32), np.zeros(1, np.int32))
        a.set_config(0, 1, True)
        r = a.row_extremes()
        assert r["max_intra"][0] == -1 and r["min_inter"][0] == -1 and r["n_intra"][0] == 0 and (r["flags"][0] & t.ROW_SELF_VALID)
        a.close()
        a = t.Alignment(ctx, sym)
        # row 0 and 1: one species; row 2: unnamed, same genus id
        a.set_labels(np.array([5, 5, -1], np.int32), np.array([2, 2, 2], np.int32), np.array([0, 0, 0], np.int32), np.arange(3, dtype=np.int32))
        a.set_config(0, 1, True)
        r = a.row_extremes()
        assert r["flags"][2] & t.EXT_UNNAMED
        assert r["max_intra"][0] == 1 and r["max_intra"][1] == 0 and r["min_inter"][0] == -1  # the unnamed congener is skipped
        assert r["d_max_intra"][0] == 1.0 - 11 / 12 and r["shared_intra"][0] == 12
        a.close()
    finally:
        ctx.close()
