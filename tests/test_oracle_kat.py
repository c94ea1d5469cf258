"""The oracle against every known-answer test the reference holds for this path
(SURVEY.md section 8c): KAT-1, KAT-2 (Sequence.java:1913-1955), KAT-3 + the class
table (Testing.java:40-87), plus the alphabet truth tables the Java source implies."""
import math

import numpy as np

import oracle
from oracle import PDM_K2P, PDM_TRANS_ONLY, PDM_UNCORRECTED, lib, normalise


def test_kat1_shared_length_and_p_distance():
    s1 = normalise("AAAAAAAAAAAAAAAAAAAAAAAAA" + "-" * 75)
    s2 = normalise("----------AAAAATTTTT?????" + "-" * 75)
    assert lib.tdo_shared_length(s1, len(s1), s2, len(s2), PDM_UNCORRECTED) == 10
    assert lib.tdo_pairwise(s1, len(s1), s2, len(s2), PDM_UNCORRECTED, 1, 1) == 0.50


def test_kat2_double_gapped_overlap():
    s1 = normalise("AAA---CCCCCCTTTTTTTTTTTTTTTTTTTTTTTT---TTTTTTTTTTTT")
    s2 = normalise("AAA---CCCCCCTTTTT----------------------------------")
    assert lib.tdo_shared_length(s1, len(s1), s2, len(s2), PDM_UNCORRECTED) == 17


def test_kat3_transversion_only():
    a = normalise("--ATCCCTGGTA")
    b = normalise("ACCGCCCT--CG")
    assert lib.tdo_pairwise(a, len(a), b, len(b), PDM_TRANS_ONLY, 1, 1) == 0.2
    assert lib.tdo_count_transversions(a, len(a), b, len(b)) == 2
    assert lib.tdo_shared_length(a, len(a), b, len(b), PDM_TRANS_ONLY) == 10


def test_purine_pyrimidine_table():
    for ch in "AGR":
        assert lib.tdo_is_purine(ord(ch)) and not lib.tdo_is_pyrimidine(ord(ch))
    for ch in "TCY":
        assert lib.tdo_is_pyrimidine(ord(ch)) and not lib.tdo_is_purine(ord(ch))
    for ch in "KMSWBDHVN":
        assert not lib.tdo_is_pyrimidine(ord(ch)) and not lib.tdo_is_purine(ord(ch))


def test_normalise_create_sequence_kat():
    # Sequence.java:1892-1907: the 28-symbol sequence is accepted, length 28
    s = normalise("----ACTG?WRKYSMBHDVN----????")
    assert len(s) == 28
    assert s == b"____ACTG?WRKYSMBHDVN____????"


def test_normalise_rules():
    assert normalise("acgt") == b"ACGT"
    assert normalise("--?-AC-GT-?--") == b"__?_AC-GT_?__"
    assert normalise("A[CT]G(AG)") == b"AYGR"
    assert normalise("----") == b"____"
    assert normalise("-?-") == b"_?_"
    for bad, code in (("-_AC", -3), ("ACUG", -4), ("ACXG", -5), ("A[C[T]]", -1), ("A[CX]", -2)):
        try:
            normalise(bad)
        except ValueError as e:
            assert e.args[0] == code, (bad, e.args)
        else:
            raise AssertionError(bad)


IUPAC = {"A": 1, "C": 2, "T": 4, "G": 8, "R": 9, "Y": 6, "K": 12, "M": 3, "S": 10, "W": 5,
         "B": 14, "D": 13, "H": 7, "V": 11, "N": 15, "-": 0, "_": 0, "?": 0}


def test_getint_table():
    for ch, m in IUPAC.items():
        assert lib.tdo_getint(ord(ch)) == m
        if m:
            assert chr(lib.tdo_getcode(m)) == ch


def test_identical_truth_table():
    for a, ma in IUPAC.items():
        for b, mb in IUPAC.items():
            got = lib.tdo_identical(ord(a), ord(b), 1)
            if a in "?_" or b in "?_":
                exp = 0
            elif a == "-" or b == "-":
                exp = int(a == b)
            else:
                exp = int((ma & mb) != 0)
            assert got == exp, (a, b)
            got0 = lib.tdo_identical(ord(a), ord(b), 0)
            if a in "?_" or b in "?_":
                exp0 = 0
            elif a == "-" or b == "-":
                exp0 = int(a == b)
            else:
                exp0 = 1  # the always-true test at Sequence.java:1280-1283
            assert got0 == exp0, (a, b)


def test_shared_length_per_symbol_pair():
    U = set("AGRCTY")
    L = set(IUPAC) - set("-_?")
    for a in IUPAC:
        for b in IUPAC:
            sa, sb = a.encode(), b.encode()
            va, vb = a in L or a == "-", b in L or b == "-"
            assert lib.tdo_shared_length(sa, 1, sb, 1, PDM_UNCORRECTED) == int(va and vb)
            assert lib.tdo_shared_length(sa, 1, sb, 1, PDM_K2P) == int(a in L and b in L)
            exp_t = int((a in U and b in U) or (va and vb and (a == "-" or b == "-")))
            assert lib.tdo_shared_length(sa, 1, sb, 1, PDM_TRANS_ONLY) == exp_t, (a, b)


def test_settings():
    assert lib.tdo_make_long(0.123456789) == 12345
    assert lib.tdo_make_long(float("nan")) == 0
    assert lib.tdo_make_long(float("inf")) == 2**63 - 1
    assert lib.tdo_percentage(1, 0) == 0
    assert lib.tdo_percentage(0.03, 1) == 3.0
    assert lib.tdo_percentage(1, 3) == 33.33
    assert lib.tdo_settings_identical(0.000001, 0.0)
    assert not lib.tdo_settings_identical(0.00001, 0.0)


def test_log_matches_libm_within_1ulp():
    rng = np.random.default_rng(7)
    xs = np.concatenate([rng.uniform(1e-6, 1.0, 20000), rng.uniform(1.0, 10.0, 2000),
                         1.0 - 2.0 ** -rng.integers(1, 52, 200), [1.0, 0.5, 2.0, 1e-300, 5e-324, 1e300]])
    worst = 0.0
    exact = 0
    for x in xs:
        a, b = lib.tdo_log(float(x)), math.log(float(x))
        u = abs(a - b) / math.ulp(b) if b != 0 else abs(a - b)
        worst = max(worst, u)
        exact += a == b
    assert worst <= 1.0
    assert exact / len(xs) > 0.90  # fdlibm log is <1 ulp, not correctly rounded
    assert lib.tdo_log(0.0) == -math.inf and math.isnan(lib.tdo_log(-1.0)) and lib.tdo_log(math.inf) == math.inf
    assert lib.tdo_log(1.0) == 0.0


def test_k2p_formula_and_edges():
    a, b = b"AAAACCCCGGGGTTTT", b"AAAACCCTGGGATTTA"  # 1 ts (C>T), 1 ts (G>A), 1 tv (T>A)
    n, ts, tv = (oracle.ctypes.c_int(), oracle.ctypes.c_int(), oracle.ctypes.c_int())
    d = lib.tdo_k2p(a, 16, b, 16, n, ts, tv)
    assert (n.value, ts.value, tv.value) == (16, 2, 1)
    w1 = 1.0 - (2.0 * 2.0) / 16 - 1.0 / 16
    w2 = 1.0 - (2.0 * 1.0) / 16
    assert abs(d - (-0.5 * math.log(w1) - 0.25 * math.log(w2))) < 1e-15
    assert lib.tdo_k2p(a, 16, a, 16, n, ts, tv) == 0.0  # -0.0 folded to +0.0
    assert math.copysign(1.0, lib.tdo_k2p(a, 16, a, 16, n, ts, tv)) == 1.0
    assert math.isnan(lib.tdo_k2p(b"NNNN", 4, b"NNNN", 4, n, ts, tv))  # n == 0
    assert math.isnan(lib.tdo_k2p(b"AAAA", 4, b"CCCC", 4, n, ts, tv))  # w2 < 0
    assert lib.tdo_k2p(b"AAAA", 4, b"GGAA", 4, n, ts, tv) == math.inf  # w1 == 0 -> +Inf
    # R vs A is a transition (same class, different char); R vs R identical
    d = lib.tdo_k2p(b"RRAA", 4, b"ARAA", 4, n, ts, tv)
    assert (n.value, ts.value, tv.value) == (4, 1, 0)
    # K2P gate counts ambiguous letters, n does not (Sequence.java:1457-1471 vs 1520-1528)
    assert lib.tdo_shared_length(b"NNAA", 4, b"NNAA", 4, PDM_K2P) == 4
    lib.tdo_k2p(b"NNAA", 4, b"NNAA", 4, n, ts, tv)
    assert n.value == 2


def test_oracle_reproduces_the_committed_fixture_checksums(golden_dir):
    """tests/golden/checksums.json (make_golden.py; the figures SURVEY.md Appendix B got from an independent numpy restatement):
    the C oracle still produces them on the reference's small Diptera fixture, in all three modes."""
    import hashlib
    import json
    import os

    import numpy as np

    from oracle import consumers as C
    want = json.load(open(os.path.join(golden_dir, "checksums.json")))["diptera_small_nogaps"]
    seqs = C.read_fasta(os.path.join(golden_dir, "diptera_small_nogaps.fasta.gz"))
    rows, lens = oracle.pack_rows([s.seq for s in seqs])
    assert (len(seqs), int(rows.shape[1])) == (want["n"], want["L"])
    iu = np.triu_indices(len(seqs), 1)
    for mode, nm in ((0, "p"), (1, "k2p"), (2, "trans")):
        r = oracle.all_pairs(rows, lens, mode, 300, True, threads=8)
        d = r["dist"][iu]
        got = {"sum_shared": int(r["shared"][iu].sum()), "sum_c1": int(r["c1"][iu].sum()), "sum_c2": int(r["c2"][iu].sum()),
               "sum_c3": int(r["c3"][iu].sum()), "valid_pairs": int((~(d < 0)).sum()), "exact_zero": int((d == 0).sum()),
               "nan": int(np.isnan(d).sum()), "sha256_dist_upper": hashlib.sha256(np.ascontiguousarray(d).tobytes()).hexdigest()}
        for k, v in got.items():
            assert want[nm][k] == v, (nm, k)
