"""Host-layer logic that needs no device: names, list order, FASTA reading, Settings, Java number
formatting and the TimSort restatement -- each against the oracle's independent restatement."""
import gzip
import os
import random

import numpy as np

from oracle import consumers as C
from taxondna_b200 import _host as H


def test_name_parsing_matches_the_oracle():
    names = ["gi|123| Aus bus", "Acanthocnema glaucescens gi|6979497|gb|AF181023-1|AF181023", "Aus bus cus gi|1|", "weird", "Aus sp. 1",
             "gi|5:12-44 Xus yus (family: Zidae)", "  padded Name here  ", "UPPER lower", "Aa b", "x" * 100, "Genus species gi|77|"]
    for nm in names:
        o, h = C.OSeq(nm, "ACGT"), H.Sequence(nm, "ACGT")
        assert h.getFullName() == o.name
        assert h.getSpeciesName() == o.species_name()
        assert h.getGenusName() == o.genus and h.getSpeciesNameOnly() == o.species and h.getSubspeciesName() == o.subspecies
        assert h.getGI() == o.get_gi() and h.getWarningFlag() == o.warning
        assert h.getDisplayName() == o.display_name()


def test_normalisation_matches_the_oracle():
    rnd = random.Random(1)
    for _ in range(300):
        n = rnd.randrange(0, 40)
        raw = "".join(rnd.choice("ACGTacgtRYKMSWBDHVN--??-[]()") for _ in range(n))
        try:
            exp = C.normalise(raw).decode()
        except ValueError:
            exp = None
        try:
            got = H.Sequence("x", raw).getSequenceRaw()
        except H.SequenceException:
            got = None
        assert got == exp, raw


def test_fasta_and_byname_order(golden_dir):
    for fn in ("diptera_small_nogaps.fasta.gz",):
        text = gzip.open(os.path.join(golden_dir, fn), "rt", encoding="latin-1").read()
        lst = H.SequenceList.fromFastaText(text)
        oseqs = C.read_fasta_text(text)
        assert lst.names() == [s.name for s in oseqs]
        lst.resort(H.SequenceList.SORT_BYNAME)
        assert lst.names() == [s.name for s in C.sort_by_name(oseqs)]
    text = "# comment\n\n>\nACGT\n>A_b c\nAC GT\nuu\n"
    lst = H.SequenceList.fromFastaText(text)
    o = C.read_fasta_text(text)
    assert lst.names() == [s.name for s in o] == ["No name specified in file", "A b c"]
    assert lst.get(1).getSequenceRaw() == o[1].seq.decode() == "ACGTTT"


def test_settings_and_formatting():
    rnd = random.Random(2)
    for _ in range(2000):
        x = rnd.choice([rnd.random(), rnd.random() * 1e-4, rnd.random() * 1e6, rnd.random() * 1e9, float(rnd.randrange(0, 100000)) / 100])
        assert H.javaDouble(x) == C.java_double_str(x), x
        assert H.javaFloat(np.float32(x)) == C.java_float_str(np.float32(x)), x
        y = rnd.random()
        assert H.Settings.percentage(x, y) == C.percentage(x, y)
        assert H.Settings.makeLongFromDouble(x) == C.make_long(x)
        assert H.Settings.identical(x, x + 3e-6) == C.identical(x, x + 3e-6)
    for x in (float("nan"), float("inf"), -0.0, 0.0, 1e-3, 1e7, 9999999.999):
        assert H.javaDouble(x) == C.java_double_str(x)


def test_timsort_both_restatements_agree():
    rnd = random.Random(3)
    for n in (0, 1, 2, 5, 31, 32, 33, 64, 100, 257, 1000, 3000):
        for kind in range(3):
            if kind == 0:
                vals = [rnd.random() * 0.01 for _ in range(n)]
            elif kind == 1:
                vals = sorted(rnd.random() for _ in range(n))
                for _ in range(n // 20):
                    vals[rnd.randrange(n)] = rnd.random()
            else:
                vals = [rnd.randrange(5) * 1e-5 * 0.7 for _ in range(n)]

            def cmp(a, b):  # the reference's non-transitive "equal within 1e-5" comparison
                return int((vals[a] - vals[b]) * 100000)

            a = list(range(n))
            try:
                C.java_sort(a, cmp)
                exp = a
            except C.ComparisonContractError:
                exp = None
            try:
                got = H.javaSort(list(range(n)), cmp)
            except ValueError:
                got = None
            assert got == exp, (n, kind)
