"""SURVEY.md 8(f) row 3 -- SequenceMatrix's two distance callers over many small per-gene alignments:
FindDistances.displayResults (SequenceMatrix/FindDistances.java:123-289) and DisplayDistancesMode.getSortedSequences
(SequenceMatrix/DisplayDistancesMode.java:217-398), host layer (one ragged alignment on the device, one launch) against the
oracle's restatement (per-pair char loops), in all three distance modes."""
import math

import numpy as np
import pytest

import oracle
from oracle import consumers as C

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def H():
    from taxondna_b200 import _host
    yield _host
    _host.Sequence.setMinOverlap(300)
    _host.Sequence.setPairwiseDistanceMethod(0)
    _host.Sequence.ambiguousBasesAllowed(True)


def make_grid(seed, n_taxa=14, genes=(("coi", 658), ("cytb", 420), ("28S", 301), ("ef1a", 97)), missing=0.15):
    """A multi-gene matrix: per gene an alignment of its own length, taxa related through a star phylogeny, IUPAC letters, '?',
    internal and external gaps, cells missing at random."""
    rng = np.random.default_rng(seed)
    acgt = np.frombuffer(b"ACGT", dtype=np.uint8)
    taxa = ["Genus%s species%s" % (chr(97 + i // 4) * 3, chr(97 + i % 4) * 4) for i in range(n_taxa)]
    taxa = [t[0].upper() + t[1:] for t in taxa]
    cells = {}
    for g, L in genes:
        root = rng.choice(acgt, size=L)
        for ti, t in enumerate(taxa):
            if rng.random() < missing:
                continue
            s = root.copy()
            m = rng.random(L) < (0.02 + 0.03 * (ti % 5))
            s[m] = rng.choice(acgt, size=int(m.sum()))
            m = rng.random(L) < 0.03
            s[m] = rng.choice(np.frombuffer(b"RYKMSWBDHVN", dtype=np.uint8), size=int(m.sum()))
            m = rng.random(L) < 0.03
            s[m] = ord("?")
            m = rng.random(L) < 0.02
            s[m] = ord("-")
            lead, trail = int(rng.integers(0, L // 6)), int(rng.integers(0, L // 6))
            s[:lead] = ord("-")
            if trail:
                s[L - trail:] = ord("-")
            cut = L if rng.random() < 0.7 else int(rng.integers(L // 2, L))  # ragged: a shorter sequence
            cells[(g, t)] = ("gi|%d| %s" % (1000 + len(cells), t), s[:cut].tobytes().decode())
    return [g for g, _ in genes], taxa, cells


def both_grids(H, charsets, taxa, cells):
    og = C.OGrid(charsets, taxa, {k: C.OSeq(n, s) for k, (n, s) in cells.items()})
    hg = H.SequenceGrid(charsets, taxa)
    for (g, t), (n, s) in cells.items():
        hg.setSequence(g, t, H.Sequence(n, s))
    return og, hg


@pytest.mark.parametrize("mode,overlap,lo,hi", [(0, 50, 0.0, 0.08), (1, 50, 0.02, 0.12), (2, 50, 0.0, 0.03), (0, 300, -1.0, 0.05), (0, 1, 0.0, 1.0)])
def test_find_distances(H, mode, overlap, lo, hi):
    charsets, taxa, cells = make_grid(3 + mode)
    og, hg = both_grids(H, charsets, taxa, cells)
    H.Sequence.setPairwiseDistanceMethod(mode)
    H.Sequence.setMinOverlap(overlap)
    H.Sequence.ambiguousBasesAllowed(True)
    cfg = C.Config(mode, overlap, True)
    for col in (None, "cytb"):
        etext, eres = C.find_distances(og, cfg, lo, hi, col, threads=8)
        fd = H.FindDistances(hg)
        text = fd.displayResults(lo, hi, col or "")
        assert text == etext
        got = fd.results()
        assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in eres]
        assert np.array_equal(np.array([d for _, _, d in got]).view(np.uint64), np.array([d for _, _, d in eres]).view(np.uint64))
        assert len(eres) > 0


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_display_distances_mode(H, mode):
    charsets, taxa, cells = make_grid(17 + mode, missing=0.2)
    og, hg = both_grids(H, charsets, taxa, cells)
    H.Sequence.setMinOverlap(300)
    H.Sequence.setPairwiseDistanceMethod(0)
    for ref in (taxa[0], taxa[5]):
        for sel in ("coi", "ef1a"):
            exp = C.display_distances(og, mode, ref, sel, threads=8)
            ddm = H.DisplayDistancesMode(hg)
            ddm.setDistanceMethod(mode)
            ddm.activateDisplay(sel)
            assert H.Sequence.getMinOverlap() == 1 and H.Sequence.getPairwiseDistanceMethod() == mode  # DisplayDistancesMode.java:163-164
            order = ddm.getSortedSequences(ref)
            ddm.deactivateDisplay()
            assert H.Sequence.getMinOverlap() == 300 and H.Sequence.getPairwiseDistanceMethod() == 0
            assert ddm.columns() == exp["columns"]
            assert order == exp["taxa"]
            for name in ("distances", "norm_distances"):
                got = np.array(getattr(ddm, "distances" if name == "distances" else "normDistances")(), dtype=np.float64)
                want = np.array(exp[name], dtype=np.float64)
                assert got.shape == want.shape
                assert np.array_equal(got.view(np.uint64), want.view(np.uint64)) or all(
                    (a == b) or (math.isnan(a) and math.isnan(b)) for a, b in zip(got.ravel(), want.ravel())), name
            assert ddm.ranks() == exp["ranks"]
            gs, ws = ddm.scores(), exp["scores"]
            assert all((a == b) or (math.isnan(a) and math.isnan(b)) for a, b in zip(gs, ws))


def test_pairs_between_against_the_matrix():
    """tdna_pairs_between through the C ABI: the pair set equals the matrix's entries in [lo, hi], -1 included when lo <= -1."""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    data = synth.generate("C4", seed=5, genera=4, species=5, reps=5, L=700)
    ctx = t.Context(0)
    aln = t.Alignment(ctx, data["symbols"])
    try:
        for mode, overlap, lo, hi in ((0, 300, 0.0, 0.05), (1, 300, 0.01, 0.2), (2, 600, -1.0, 0.0), (0, 690, -1.0, -1.0)):
            aln.set_config(mode, overlap, True)
            d = aln.matrix()
            ii, jj, dd = aln.pairs_between(lo, hi)
            iu = np.triu_indices(d.shape[0], 1)
            keep = (d[iu] >= lo) & (d[iu] <= hi)
            want = sorted(zip(iu[0][keep].tolist(), iu[1][keep].tolist(), d[iu][keep].view(np.uint64).tolist()))
            got = sorted(zip(ii.tolist(), jj.tolist(), dd.view(np.uint64).tolist()))
            assert got == want
    finally:
        aln.close()
        ctx.close()


def test_edge_cases_of_pairs_between_and_the_grid_callers(H):
    """One sequence, nothing in range, a capacity below the count (two-call protocol), an empty grid, a gene nobody has."""
    import ctypes
    import taxondna_b200 as t
    ctx = t.Context(0)
    try:
        one = t.Alignment(ctx, np.frombuffer(b"ACGTACGTAC", dtype=np.uint8).reshape(1, 10))
        one.set_config(0, 1, True)
        ii, jj, dd = one.pairs_between(-1.0, 1.0)
        assert ii.size == 0
        one.close()
        sym = np.frombuffer(b"AAAAAAAAAA" b"AAAAAAAAAT" b"AAAAAAAATT" b"??????????", dtype=np.uint8).reshape(4, 10)
        a = t.Alignment(ctx, sym)
        a.set_config(0, 1, True)
        assert a.pairs_between(0.5, 0.9)[0].size == 0           # nothing in range
        ii, jj, dd = a.pairs_between(-1.0, -1.0)                 # only the pairs below minOverlap: everything against the all-'?' row
        assert sorted(zip(ii.tolist(), jj.tolist())) == [(0, 3), (1, 3), (2, 3)] and (dd == -1.0).all()
        n = ctypes.c_int64()
        buf_i, buf_j, buf_d = np.full(2, -7, np.int32), np.full(2, -7, np.int32), np.zeros(2)
        t.check(ctx.L.tdna_pairs_between(a.h, 0.0, 1.0, buf_i.ctypes.data, buf_j.ctypes.data, buf_d.ctypes.data, 2, ctypes.byref(n)))
        assert n.value == 3 and (buf_i >= 0).all()             # three pairs match, the first two that arrived were stored
        a.close()
    finally:
        ctx.close()
    H.Sequence.setPairwiseDistanceMethod(0)
    H.Sequence.setMinOverlap(1)
    empty = H.SequenceGrid(["g1"], ["Aus bus"])
    text = H.FindDistances(empty).displayResults(0.0, 0.5, "")
    assert text.startswith("0 sequence pairs found with distances between 0.0% and 50.0%.")
    grid = H.SequenceGrid(["g1", "g2"], ["Aus bus", "Aus cus"])
    grid.setSequence("g1", "Aus bus", H.Sequence("Aus bus", "ACGTACGTAC"))
    text = H.FindDistances(grid).displayResults(0.0, 0.5, "")  # one sequence in all: it is never compared to itself
    assert text.startswith("0 sequence pairs found")
    ddm = H.DisplayDistancesMode(grid)
    ddm.setDistanceMethod(0)
    ddm.activateDisplay("g1")
    order = ddm.getSortedSequences("Aus bus")
    ddm.deactivateDisplay()
    assert order[0] == "Aus bus"
    d = ddm.distances()
    assert d[0] == [H.DisplayDistancesMode.DIST_SEQ_ON_TOP, H.DisplayDistancesMode.DIST_SEQ_NA]   # g1: the reference itself, a taxon without the gene
    assert d[1] == [H.DisplayDistancesMode.DIST_NO_COMPARE_SEQ] * 2                                # g2: the reference taxon has no sequence
    H.Sequence.setMinOverlap(300)
