"""Oracle consumers: the one hand-derivable reference case (C1) + internal sanity."""
import functools
import os
import random

import numpy as np

from oracle import PDM_K2P, PDM_UNCORRECTED
from oracle import consumers as C


def test_c1_fixture_both_no_match(golden_dir):
    # by reading BestMatch.run (BestMatch.java:273-284) both sequences of
    # "Tests/best match/doesnt_add_up_to_100_bug.txt" must come out "No match."
    seqs = C.sort_by_name(C.read_fasta(os.path.join(golden_dir, "c1_best_match_fixture.txt")))
    assert len(seqs) == 2 and {len(s.seq) for s in seqs} == {1050}
    D = C.Dist(seqs, C.Config(PDM_UNCORRECTED, 300, True))
    text, recs, cnt = C.best_match(D, 3.0)
    assert cnt["no_matches"] == 2
    assert text.count("\t\tNo match.\n") == 2
    assert "Sequences:\t2\n" in text
    assert "Sequences with atleast one matching sequence in the data set:\t0\n" in text
    assert "Correct identifications according to \"Best Match\":\t0 (0.0%)" in text
    assert "Sequences without any match closer than 3.0%:\t0 (0.0%)" in text


def test_timsort_equals_stable_sort_for_consistent_comparators():
    rnd = random.Random(5)
    for n in (0, 1, 2, 5, 31, 32, 33, 64, 100, 257, 1000, 3000):
        for trial in range(3):
            kind = trial % 3
            if kind == 0:
                keys = [rnd.randrange(max(1, n // 4)) for _ in range(n)]
            elif kind == 1:  # long runs, triggers galloping
                keys = sorted(rnd.randrange(50) for _ in range(n))
                for _ in range(n // 20):
                    keys[rnd.randrange(n)] = rnd.randrange(50)
            else:
                keys = [n - i for i in range(n)]
            a = list(range(n))
            calls = [0]

            def cmp(x, y):
                calls[0] += 1
                return keys[x] - keys[y]

            C.java_sort(a, cmp)
            assert a == sorted(range(n), key=lambda i: keys[i])


def test_java_number_formatting():
    f = C.java_double_str
    assert f(0.0) == "0.0" and f(3.0) == "3.0" and f(33.33) == "33.33" and f(100.0) == "100.0"
    assert f(0.001) == "0.001" and f(0.0001) == "1.0E-4" and f(1.0e7) == "1.0E7" and f(1234567.0) == "1234567.0"
    assert f(float("nan")) == "NaN" and f(1.5e-5) == "1.5E-5" and f(-0.0) == "-0.0"
    assert C.java_float_str(np.float32(0.1)) == "0.1" and C.java_float_str(np.float32(2.75)) == "2.75"


def test_name_parsing_and_display():
    s = C.OSeq("gi|123| Aus bus", "ACGT")
    assert (s.genus, s.species, s.gi) == ("Aus", "bus", "123") and s.display_name() == "Aus bus (gi:123)"
    s = C.OSeq("Acanthocnema glaucescens gi|6979497|gb|AF181023-1|AF181023", "ACGT")
    # the 3-word pattern swallows "gi" as a subspecies (Sequence.java:690-697) -- reference behaviour
    assert s.species_name() == "Acanthocnema glaucescens" and s.subspecies == "gi"
    assert s.display_name() == "Acanthocnema glaucescens (gi) (gi:6979497)"
    s = C.OSeq("Aus bus cus gi|1|", "ACGT")
    assert s.subspecies == "cus" and s.display_name() == "Aus bus (cus) (gi:1)"
    s = C.OSeq("weird", "ACGT")
    assert s.warning and s.species_name() is None and s.display_name() == "{weird}"
    s = C.OSeq("Aus sp. 1", "ACGT")
    assert not s.warning and s.species_name() == "Aus sp"  # the 2-word pattern wins; the "sp." branch is dead


def test_small_end_to_end_consumers_run():
    rnd = np.random.default_rng(3)
    ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
    base = rnd.choice(ACGT, 400)
    seqs = []
    gid = 0
    for sp in range(6):
        spseq = base.copy()
        idx = rnd.choice(400, 12, replace=False)
        spseq[idx] = rnd.choice(ACGT, 12)
        for rep in range(4):
            s = spseq.copy()
            idx = rnd.choice(400, 2, replace=False)
            s[idx] = rnd.choice(ACGT, 2)
            gid += 1
            seqs.append(C.OSeq("gi|%d| G%s s%s" % (gid, "ab"[sp // 3], "xyzuvw"[sp]), s.tobytes().decode()))
    seqs = C.sort_by_name(seqs)
    for mode in (PDM_UNCORRECTED, PDM_K2P):
        D = C.Dist(seqs, C.Config(mode, 300, True))
        text, recs, cnt = C.best_match(D, 3.0)
        assert cnt["no_matches"] == 0 and text.startswith("Sequences:\t24\n")
        btext, brecs, bc = C.block_analysis(D, 3.0)
        assert bc["noseq"] == 0
        stext, five, intra, inter = C.pairwise_summary(D)
        assert intra.count_valid() == 6 * 6 and inter.count_valid() == 2 * 3 * 16
        clusters, stats = C.cluster(D, 0.03)
        assert sorted(x for c in clusters for x in c) == list(range(24))


HAND = (">gi|1| Aus bus\nAAAAAAAAAA\n>gi|2| Aus bus\nAAAAAAAAAC\n>gi|3| Aus cus\nAAAAAACCCC\n>gi|4| Bus dus\nCCCCCCCCCC\n")


def hand_list():
    seqs = C.sort_by_name(C.read_fasta_text(HAND))
    assert [s.species_name() for s in seqs] == ["Aus bus", "Aus bus", "Aus cus", "Bus dus"]
    return seqs, C.Dist(seqs, C.Config(PDM_UNCORRECTED, 5, True))


def test_extreme_pairwise_by_hand():
    # d(A1,A2) = 0.1, d(A1,B1) = 0.4, d(A2,B1) = 0.3, everything to C1 = 0.6 .. 1.0 (another genus: never "congeneric")
    seqs, D = hand_list()
    assert [D.pairwise(0, 1), D.pairwise(0, 2), D.pairwise(1, 2)] == [1.0 - 9 / 10, 1.0 - 6 / 10, 1.0 - 7 / 10]
    text, recs = C.extreme_pairwise(D)
    assert recs == [(1, 2), (0, 2), (None, 1), (None, None)]
    lines = text.split("\n")
    assert lines[0].startswith("Sequence name\tLargest conspecific match\tDistance\tOverlap\t")
    # 1.0 - 9/10 is 0.09999999999999998 and Settings.percentage truncates (Settings.java:69-87): 9.99, not 10.0
    assert lines[1] == "Aus bus (gi:1)\tAus bus (gi:2)\t9.99\t10\tAus cus (gi:3)\t40.0\t10\t"
    assert lines[3] == "Aus cus (gi:3)\tNo matching conspecific sequence\tN/A\tN/A\tAus bus (gi:2)\t30.0\t10\t"
    assert lines[4] == "Bus dus (gi:4)\tNo matching conspecific sequence\tN/A\tN/A\tNo matching congeneric, interspecific sequence\tN/A\tN/A\t"


def test_distance_analysis_by_hand():
    # per sequence (smallest, mean) congeneric allospecific: A1 (.4, .4), A2 (.3, .3), B1 (.3, .35); species means: bus (.35, .35),
    # cus (.3, .35); genus Aus: smallest (.35 + .3) / 2, mean .35; genus Bus has no comparison: the mean of nothing is NaN, NaN < 0 is
    # false, and Settings.percentage turns NaN into 0.0 ((long) NaN == 0)
    _, D = hand_list()
    text, rows = C.distance_analysis(D)
    assert abs(rows["Aus"][0] - 0.325) < 1e-12 and abs(rows["Aus"][1] - 0.35) < 1e-12
    assert text == "Genus\t\t\tAvg. Avg. Smallest Inter\t\tAvg. Avg. Avg. Inter\nAus\t\t\t32.5\t\t35.0\nBus\t\t\t0.0\t\t0.0\n"


def test_pairwise_explorer_and_query_by_hand():
    seqs, D = hand_list()
    cats = C.pairwise_explorer(D, C.PD_INTRA)  # the one conspecific pair, pushed from both sides, d = 0.1 = "9.5% to 10.0%"
    assert [k for k, _ in cats] == ["9.5% to 10.0% (2 matches)", "Averages"]
    # Aus bus x Aus cus: 0.4 and 0.30000000000000004, both directions -- and the latter falls BETWEEN two categories, whose bounds
    # are (x/1000 + 0.000001, (x+5)/1000] (PairwiseExplorer.java:163-165): the reference lists it nowhere
    cats = C.pairwise_explorer(D, C.PD_INTER)
    assert [k for k, _ in cats] == ["39.5% to 40.0% (2 matches)", "Averages"]
    lines, pos = C.query_sequence(seqs, C.Config(PDM_UNCORRECTED, 5, True), "aaaaa aaaac\n")
    assert pos == [1, 0, 2, 3]
    assert lines[0] == "(0.0000%) Aus bus (gi:2)" and lines[1] == "(10.0000%) Aus bus (gi:1)" and lines[2].startswith("(30.0000%) Aus cus")


def test_consensus_by_hand():
    n = C.normalise
    assert C.get_consensus(n("ACGT"), n("ACGA")) == b"ACGW"          # T | A = W
    assert C.get_consensus(n("AC-T"), n("ACGT")) == b"ACGT"          # a gap gives way to a letter
    assert C.get_consensus(n("A?GT"), n("ACGT")) == b"A?GT"          # missing data is absorbing
    assert C.get_consensus(n("--GT"), n("-CGT")) == b"_CGT"          # external gaps: '_' only where both are gaps
    assert C.get_consensus(n("ACGT"), n("AC")) == b"AC??"            # the shorter sequence has ended: '?'
    assert C.consensus_of_bin([n("AAAA"), n("CAAA"), n("GAAA"), n("TAA-")]) == b"NAAA"


def test_consumer_digests_of_the_reference_fixture(golden_dir):
    """The report text of every restated module on the reference's small Diptera fixture hashes to the committed digests
    (tests/golden/make_consumer_digests.py): the oracle does not drift, and the GPU host layer is held to the same bytes."""
    import gzip
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_consumer_digests", os.path.join(golden_dir, "make_consumer_digests.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    want = json.load(open(os.path.join(golden_dir, "consumer_digests.json")))["diptera_small_nogaps"]
    text = gzip.open(os.path.join(golden_dir, "diptera_small_nogaps.fasta.gz"), "rt", encoding="latin-1").read()
    assert mod.digests(text, PDM_UNCORRECTED) == want["p"]
    assert mod.digests(text, PDM_K2P) == want["k2p"]


def test_sequence_matrix_callers_hand_derived():
    """FindDistances / DisplayDistancesMode restatements on a grid small enough to work out by hand (SequenceMatrix/
    FindDistances.java:196-262, DisplayDistancesMode.java:255-385)."""
    cells = {("g1", "Aus bus"): C.OSeq("Aus bus", "AAAAAAAAAA"), ("g1", "Aus cus"): C.OSeq("Aus cus", "AAAAAAAAAT"),
             ("g1", "Aus dus"): C.OSeq("Aus dus", "AAAAAAAATT"),
             ("g2", "Aus bus"): C.OSeq("Aus bus", "CCCCC"), ("g2", "Aus dus"): C.OSeq("Aus dus", "CCCCG")}
    grid = C.OGrid(["g1", "g2"], ["Aus bus", "Aus cus", "Aus dus"], cells)
    cfg = C.Config(C.PDM_UNCORRECTED, 1, True)
    # rows: g1 bus, cus, dus (0..2), g2 bus, dus (3, 4).  1.0 - 9/10 = 0.09999999999999998 is inside [0, 0.1] and prints as 9.99 %
    # (Settings.percentage truncates twice); every cross-gene pair compares A's with C's over five columns: 1.0
    text, res = C.find_distances(grid, cfg, 0.0, 0.1)
    assert [(a, b) for a, b, _ in res] == [(0, 1), (1, 0), (1, 2), (2, 1)]
    assert all(d == 1.0 - 9 / 10 for _, _, d in res)
    assert text == ("4 sequence pairs found with distances between 0.0% and 10.0%.\nNote that distances were calculated using uncorrected pairwise "
                    "distances and sequence overlaps of less than 1 bp weren't counted.\n\n"
                    "Aus bus (from g1)\tAus cus (from g1)\t9.99%\nAus cus (from g1)\tAus bus (from g1)\t9.99%\n"
                    "Aus cus (from g1)\tAus dus (from g1)\t9.99%\nAus dus (from g1)\tAus cus (from g1)\t9.99%\n")
    text, res = C.find_distances(grid, cfg, 0.15, 0.25, "g2")
    assert [(a, b, d) for a, b, d in res] == [(0, 1, 1.0 - 4 / 5), (1, 0, 1.0 - 4 / 5)]
    # DisplayDistancesMode against "Aus bus", gene g1 selected: g1 = [ON_TOP, 0.1-, 0.2-], g2 = [ON_TOP, SEQ_NA, 0.2-]
    r = C.display_distances(grid, C.PDM_UNCORRECTED, "Aus bus", "g1")
    assert r["columns"] == ["g1", "g2"]
    d1, d2 = 1.0 - 9 / 10, 1.0 - 8 / 10
    # scores over the OTHER genes (g2): bus NaN (no distance), cus NaN, dus (0.2 - 0.2) / (0.2 - 0.2) = NaN: all tie at (long) NaN = 0
    assert r["taxa"] == ["Aus bus", "Aus cus", "Aus dus"]
    assert r["distances"][0] == [C.DIST_SEQ_ON_TOP, d1, d2]
    assert r["distances"][1] == [C.DIST_SEQ_ON_TOP, C.DIST_SEQ_NA, 1.0 - 4 / 5]
    assert r["norm_distances"][0][1:] == [0.0, 1.0]
    assert r["ranks"][0] == [-1024, 0, 1] and r["ranks"][1] == [-1024, -128, 0]
