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
