"""The host layer (taxondna_b200._host: the reference's Java API re-hosted over the C ABI) against
the oracle's restatement of the same consumers.  Reads like the reference's own tests would."""
import gzip
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import consumers as C

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def H():
    from taxondna_b200 import _host
    yield _host
    _host.Sequence.setMinOverlap(300)
    _host.Sequence.setPairwiseDistanceMethod(0)
    _host.Sequence.ambiguousBasesAllowed(True)


def configure(H, mode, overlap, ambig=True):
    H.Sequence.setPairwiseDistanceMethod(mode)
    H.Sequence.setMinOverlap(overlap)
    H.Sequence.ambiguousBasesAllowed(ambig)


def test_sequence_kats(H):
    # Sequence.test(), Sequence.java:1881-1955, through the same API
    configure(H, H.Sequence.PDM_UNCORRECTED, 1)
    seq = H.Sequence("weird name", "----ACTG?WRKYSMBHDVN----????")
    assert seq.getLength() == 28
    seq1 = H.Sequence("1", "AAAAAAAAAAAAAAAAAAAAAAAAA" + "-" * 75)
    seq2 = H.Sequence("2", "----------AAAAATTTTT?????" + "-" * 75)
    assert seq1.getSharedLength(seq2) == 10
    assert seq1.getPairwise(seq2) == 0.50
    seq1 = H.Sequence("1", "AAA---CCCCCCTTTTTTTTTTTTTTTTTTTTTTTT---TTTTTTTTTTTT")
    seq2 = H.Sequence("2", "AAA---CCCCCCTTTTT----------------------------------")
    assert seq1.getSharedLength(seq2) == 17
    # Testing.testTransScoring, Testing.java:36-87
    for ch in "AGR":
        assert H.Sequence.isPurine(ch) and not H.Sequence.isPyrimidine(ch)
    for ch in "TCY":
        assert H.Sequence.isPyrimidine(ch) and not H.Sequence.isPurine(ch)
    for ch in "KMSWBDHVN":
        assert not H.Sequence.isPyrimidine(ch) and not H.Sequence.isPurine(ch)
    configure(H, H.Sequence.PDM_TRANS_ONLY, 1)
    a, b = H.Sequence("T1", "--ATCCCTGGTA"), H.Sequence("T2", "ACCGCCCT--CG")
    assert a.getPairwise(b) == 0.2
    assert a.countTransversions(b) == 2 and a.getSharedLength(b) == 10
    with pytest.raises(H.SequenceException):
        H.Sequence("bad", "ACGU")
    configure(H, H.Sequence.PDM_UNCORRECTED, 300)
    assert seq1.getPairwise(seq2) == -1.0  # inadequate overlap sentinel (Sequence.java:1715)


def test_java_formatting_and_settings(H):
    for v in (0.0, 3.0, 33.33, 100.0, 0.001, 0.0001, 1.0e7, 1234567.0, 1.5e-5, 0.1, 2.75, 12.05, 99.99):
        assert H.javaDouble(v) == C.java_double_str(v), v
    for v in (0.1, 2.75, 3.0, 0.33, 1e-4):
        assert H.javaFloat(np.float32(v)) == C.java_float_str(np.float32(v))
    assert H.Settings.percentage(1, 3) == 33.33 and H.Settings.percentage(1, 0) == 0
    assert H.Settings.makeLongFromDouble(float("nan")) == 0


def test_timsort_matches_the_oracle_restatement_even_when_inconsistent(H):
    import random
    rnd = random.Random(3)
    for n in (5, 31, 32, 33, 100, 500, 2000):
        vals = [rnd.random() * 0.01 for _ in range(n)]

        def cmp(a, b):  # the non-transitive comparison the reference uses (|diff| < 1e-5 is "equal")
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
        assert got == exp


def load_lists(H, text):
    lst = H.SequenceList.fromFastaText(text)
    lst.resort(H.SequenceList.SORT_BYNAME)
    oseqs = C.sort_by_name(C.read_fasta_text(text))
    assert lst.names() == [s.name for s in oseqs]
    return lst, oseqs


def run_all_consumers(H, text, mode, overlap, threshold=3.0, expect_host_rows=None):
    configure(H, mode, overlap)
    lst, oseqs = load_lists(H, text)
    D = C.Dist(oseqs, C.Config(mode, overlap, True), threads=8)
    # distances first: the host layer's matrix is the library's, bit for bit the oracle's
    M = lst.matrix()
    assert np.array_equal(M.view(np.uint64), D.dist.view(np.uint64))
    otext, _, _ = C.best_match(D, threshold)
    bm = H.BestMatch(lst)
    bm.setThreshold(threshold)
    assert bm.run() == otext
    fast_rows = bm.rows_resolved_on_host
    bm.force_exact = True
    assert bm.run() == otext
    btext, _, _ = C.block_analysis(D, threshold)
    ba = H.BlockAnalysis(lst)
    ba.setThreshold(threshold)
    assert ba.run() == btext
    stext, five, _, _ = C.pairwise_summary(D)
    ps = H.PairwiseSummary(lst)
    assert ps.run() == stext
    assert np.float32(ps.getFivePercentCutoff()) == five
    for kind in (H.PairwiseDistances.PD_INTRA, H.PairwiseDistances.PD_INTER):  # PairwiseDistances keeps the pair identities
        try:
            exp_pairs, exp_avg = C.pairwise_distances(D, kind)
        except C.ComparisonContractError:
            exp_pairs = None
        if exp_pairs is not None:
            pd = H.PairwiseDistances(lst, kind)
            got = pd.triples()
            assert len(got) == len(exp_pairs) == pd.countValidComparisons()
            assert [(a, b) for a, b, _ in got] == [(a, b) for a, b, _ in exp_pairs]
            assert np.array_equal(np.array([d for _, _, d in got]).view(np.uint64), np.array([d for _, _, d in exp_pairs]).view(np.uint64))
            assert set(pd.getAveragedSequences()) == set(exp_avg)
            for nm, v in exp_avg.items():
                g = pd.getAverageDistance(nm)
                assert (g != g and v != v) or g == v, (nm, g, v)
            assert pd.getAverageDistance("no such sequence") == -1.0
    # the remaining getPairwise consumers
    etext, erecs = C.extreme_pairwise(D)
    ep = H.ExtremePairwise(lst)
    assert ep.run() == etext
    assert ep.rows() == [(-2, -2) if r is None else (-1 if r[0] is None else r[0], -1 if r[1] is None else r[1]) for r in erecs]
    for kind in (H.PairwiseDistances.PD_INTRA, H.PairwiseDistances.PD_INTER):
        try:
            ecats = C.pairwise_explorer(D, kind)
        except C.ComparisonContractError:
            continue
        got = H.PairwiseExplorer(lst).run(kind)
        assert [k for k, _ in got] == [k for k, _ in ecats]
        for (_, g), (_, e) in zip(got, ecats):
            assert [(a, b) for a, b, _ in g] == [(a, b) for a, b, _ in e]
            assert np.array_equal(np.array([d for _, _, d in g]).view(np.uint64), np.array([d for _, _, d in e]).view(np.uint64))
    try:
        dtext, _ = C.distance_analysis(D)
    except C.JavaNPE:
        with pytest.raises(Exception):
            H.DistanceAnalysis(lst).run()
    else:
        assert H.DistanceAnalysis(lst).run() == dtext
    # QuerySequence: a pasted sequence that is not in the list -- a mutated copy of a list member, a short fragment, and whitespace
    rng = np.random.default_rng(5)
    for which, cut in ((0, None), (len(oseqs) // 2, 350)):
        member = oseqs[which].seq
        raw = list((member.decode() if isinstance(member, bytes) else member).replace("_", "-"))
        for pos in rng.integers(0, len(raw), size=12):
            raw[pos] = "ACGT"[int(rng.integers(0, 4))]
        raw = "".join(raw)[:cut]
        pasted = "\n".join(raw[i:i + 60] for i in range(0, len(raw), 60)) + " \n"
        qlines, qpos = C.query_sequence(oseqs, C.Config(mode, overlap, True), pasted, threads=8)
        qs = H.QuerySequence(lst)
        assert qs.query(pasted) == qlines
        assert qs.positions() == qpos
    oc, ostats = C.cluster(D, threshold / 100)
    cl = H.Cluster(lst)
    cl.setThreshold(threshold)
    cl.run()
    assert cl.clusters() == oc
    assert cl.consensusSequences() == [C.consensus_of_bin([oseqs[i].seq for i in b]) for b in oc]  # Cluster.makeConsensusOfBin
    for s, e in zip(cl.stats(), ostats):
        assert (s.size, s.valid_comparisons, s.valid_comparisons_over) == (e["size"], e["valid"], e["over"])
        assert s.largest_pairwise == e["largest"] and s.percentage == e["pct"]
    return fast_rows, lst.count()


def test_c1_fixture_best_match(H, golden_dir):
    text = open(os.path.join(golden_dir, "c1_best_match_fixture.txt"), encoding="latin-1").read()
    configure(H, 0, 300)
    lst, oseqs = load_lists(H, text)
    out = H.BestMatch(lst).run()
    assert out.count("\t\tNo match.\n") == 2 and "Sequences:\t2\n" in out
    assert out == C.best_match(C.Dist(oseqs, C.Config(0, 300, True)), 3.0)[0]


def fasta_subset(path, k):
    text = gzip.open(path, "rt", encoding="latin-1").read()
    recs = text.split(">")[1:k + 1]
    return "".join(">" + r for r in recs)


@pytest.mark.parametrize("mode", [0, 1])
def test_diptera_small_nogaps_all_consumers(H, golden_dir, mode):
    text = gzip.open(os.path.join(golden_dir, "diptera_small_nogaps.fasta.gz"), "rt", encoding="latin-1").read()
    run_all_consumers(H, text, mode, 300)


@pytest.mark.parametrize("mode", [0, 1])
def test_diptera_gappy_subset_all_consumers(H, golden_dir, mode):
    text = fasta_subset(os.path.join(golden_dir, "diptera_coi.fasta.gz"), 420)
    run_all_consumers(H, text, mode, 300)


def synth_fasta(cfg, **kw):
    from taxondna_b200 import synth
    d = synth.generate(cfg, **kw)
    sym = d["symbols"].copy()
    sym[sym == ord("_")] = ord("-")
    return "".join(">%s\n%s\n" % (nm, row.tobytes().decode()) for nm, row in zip(d["names"], sym)), d


@pytest.mark.parametrize("cfg,mode,kw", [("C3", 1, dict(genera=6, species=5, reps=8)), ("C3", 0, dict(genera=6, species=5, reps=8)),
                                         ("C4", 0, dict(genera=5, species=6, reps=8, L=900)), ("C4", 1, dict(genera=5, species=6, reps=8, L=900))])
def test_synthetic_shapes_all_consumers(H, cfg, mode, kw):
    text, d = synth_fasta(cfg, **kw)
    fast_rows, n = run_all_consumers(H, text, mode, 300)
    print("rows resolved on the host: %d of %d" % (fast_rows, n))


def test_generation_order_is_byname_order(H):
    text, d = synth_fasta("tiny")
    lst = H.SequenceList.fromFastaText(text)
    before = lst.names()
    lst.resort(H.SequenceList.SORT_BYNAME)
    assert lst.names() == before
    assert lst.speciesIds() == d["species_id"].tolist()


def test_fasta_through_the_device_parser_gives_the_same_list(H, golden_dir):
    """SequenceList.fromFastaTextOnDevice == fromFastaText: names, normalised sequences, and a file with a "[AG]" group (host path)."""
    for text in (gzip.open(os.path.join(golden_dir, "diptera_coi.fasta.gz"), "rt", encoding="latin-1").read(),
                 synth_fasta("C3", genera=4, species=3, reps=5)[0],
                 ">Aus bus gi|1|\nAC[AG]TAC-T\n>Aus_cus gi|2|\r\n--acgu acgt??\r\n"):
        a = H.SequenceList.fromFastaText(text)
        b = H.SequenceList.fromFastaTextOnDevice(text)
        assert a.count() == b.count() and a.names() == b.names()
        assert a.rawSequences() == b.rawSequences()


@pytest.mark.parametrize("mode,key", [(0, "p"), (1, "k2p")])
def test_reports_hash_to_the_committed_digests(H, golden_dir, mode, key):
    """No oracle in the loop: the host layer's report text on the reference's small Diptera fixture against
    tests/golden/consumer_digests.json (generated by tests/golden/make_consumer_digests.py)."""
    import hashlib
    import json
    want = json.load(open(os.path.join(golden_dir, "consumer_digests.json")))["diptera_small_nogaps"][key]
    text = gzip.open(os.path.join(golden_dir, "diptera_small_nogaps.fasta.gz"), "rt", encoding="latin-1").read()
    configure(H, mode, 300)
    lst = H.SequenceList.fromFastaText(text)
    lst.resort(H.SequenceList.SORT_BYNAME)
    assert lst.count() == want["n"]
    sha = lambda s: hashlib.sha256(s.encode("utf-8")).hexdigest()  # noqa: E731
    bm = H.BestMatch(lst)
    bm.setThreshold(3.0)
    assert sha(bm.run()) == want["best_match"]
    ba = H.BlockAnalysis(lst)
    ba.setThreshold(3.0)
    assert sha(ba.run()) == want["block_analysis"]
    assert sha(H.PairwiseSummary(lst).run()) == want["pairwise_summary"]
    assert sha(H.ExtremePairwise(lst).run()) == want["extreme_pairwise"]
    if want["distance_analysis"] != "NullPointerException":
        assert sha(H.DistanceAnalysis(lst).run()) == want["distance_analysis"]
    cl = H.Cluster(lst)
    cl.setThreshold(3.0)
    cl.run()
    assert sha(json.dumps(cl.clusters())) == want["cluster"]

