"""GPU parity: the CUDA library (through its C ABI) against the CPU oracle, bit for bit."""
import gzip
import hashlib
import json
import os

import numpy as np
import pytest

import sys

import oracle
from oracle import consumers as C

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

SYMS = b"ACGTRYKMSWBDHVN-_?"
MODES = [(0, True), (0, False), (1, True), (2, True)]


@pytest.fixture(scope="module")
def ctx():
    import taxondna_b200 as t
    c = t.Context(0)
    yield c
    c.close()


def u64(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def check_against_oracle(ctx, sym, lens, min_overlap, modes=MODES, budget=None):
    import taxondna_b200 as t
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym, lens)
    try:
        for mode, ambig in modes:
            aln.set_config(mode, min_overlap, ambig)
            d, c = aln.matrix(counts=True)
            ref = oracle.all_pairs(sym, lens if lens is not None else np.full(n, sym.shape[1], np.int32), mode, min_overlap, ambig, threads=8)
            assert np.array_equal(c["shared"], ref["shared"]), (mode, ambig, "shared")
            assert np.array_equal(c["c1"], ref["c1"]), (mode, ambig, "c1")
            assert np.array_equal(c["c2"], ref["c2"]), (mode, ambig, "c2")
            assert np.array_equal(c["c3"], ref["c3"]), (mode, ambig, "c3")
            assert np.array_equal(u64(d), u64(ref["dist"])), (mode, ambig, "dist")
            d2 = aln.matrix()  # the symmetric (upper-triangle + mirror) path
            assert np.array_equal(u64(d2), u64(ref["dist"])), (mode, ambig, "dist-symmetric")
    finally:
        aln.close()


def test_symbol_pair_truth_table(ctx):
    # 18 x 18 symbol pairs, every mode: one-column alignment, row r = symbol r
    sym = np.frombuffer(SYMS, dtype=np.uint8).reshape(-1, 1).copy()
    for mo in (0, 1):
        check_against_oracle(ctx, sym, None, mo)


def test_kats_through_the_abi(ctx):
    import taxondna_b200 as t
    s1 = oracle.normalise("AAAAAAAAAAAAAAAAAAAAAAAAA" + "-" * 75)
    s2 = oracle.normalise("----------AAAAATTTTT?????" + "-" * 75)
    rows, lens = oracle.pack_rows([s1, s2])
    aln = t.Alignment(ctx, rows, lens)
    aln.set_config(t.PDM_UNCORRECTED, 1, True)
    (shared, ident, _, _), d = aln.pair(0, 1)
    assert shared == 10 and d == 0.50  # KAT-1, Sequence.java:1913-1936
    aln.close()
    s1 = oracle.normalise("AAA---CCCCCCTTTTTTTTTTTTTTTTTTTTTTTT---TTTTTTTTTTTT")
    s2 = oracle.normalise("AAA---CCCCCCTTTTT----------------------------------")
    rows, lens = oracle.pack_rows([s1, s2])
    aln = t.Alignment(ctx, rows, lens)
    aln.set_config(t.PDM_UNCORRECTED, 1, True)
    assert aln.pair(0, 1)[0][0] == 17  # KAT-2, Sequence.java:1942-1955
    aln.close()
    a, b = oracle.normalise("--ATCCCTGGTA"), oracle.normalise("ACCGCCCT--CG")
    rows, lens = oracle.pack_rows([a, b])
    aln = t.Alignment(ctx, rows, lens)
    aln.set_config(t.PDM_TRANS_ONLY, 1, True)
    (shared, tv, _, _), d = aln.pair(0, 1)
    assert (shared, tv, d) == (10, 2, 0.2)  # KAT-3, Testing.java:73-87
    aln.close()


def random_alignment(rng, n, L, ragged=False, p_special=0.25):
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    sym = np.tile(base, (n, 1))
    mut = rng.random((n, L)) < 0.15
    sym[mut] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(mut.sum()))
    sp = rng.random((n, L)) < p_special
    sym[sp] = rng.choice(np.frombuffer(SYMS, dtype=np.uint8), size=int(sp.sum()))
    lens = None
    if ragged:
        lens = rng.integers(max(1, L // 2), L + 1, size=n).astype(np.int32)
        lens[0] = L
    return sym, lens


@pytest.mark.parametrize("n,L,ragged", [(1, 5, False), (2, 31, False), (3, 32, False), (65, 33, True), (128, 64, False),
                                        (129, 257, True), (200, 658, False), (300, 700, True), (257, 1539, False)])
def test_random_alignments_all_modes(ctx, n, L, ragged):
    rng = np.random.default_rng(n * 1000 + L)
    sym, lens = random_alignment(rng, n, L, ragged)
    check_against_oracle(ctx, sym, lens, min_overlap=max(1, L // 3))


def test_banded_path_equals_single_band(ctx):
    import taxondna_b200 as t
    rng = np.random.default_rng(99)
    sym, _ = random_alignment(rng, 700, 300)
    aln = t.Alignment(ctx, sym)
    aln.set_config(t.PDM_K2P, 100, True)
    full = aln.matrix()
    ctx.set_memory_budget(2 << 20)  # 2 MiB -> 128-row bands (rectangular tiles, no mirroring)
    try:
        aln.set_config(t.PDM_K2P, 100, True)
        banded = aln.matrix()
    finally:
        ctx.set_memory_budget(64 << 30)
    assert np.array_equal(u64(full), u64(banded))
    aln.set_row_range(256, 600)
    part = aln.matrix()
    assert np.array_equal(u64(part), u64(full[256:600]))
    aln.close()


def test_golden_fixture_checksums(ctx, golden_dir):
    import taxondna_b200 as t
    sums = json.load(open(os.path.join(golden_dir, "checksums.json")))
    for key in ("diptera_small_nogaps", "diptera_coi"):
        seqs = C.read_fasta(os.path.join(golden_dir, key + ".fasta.gz"))
        rows, lens = oracle.pack_rows([s.seq for s in seqs])
        aln = t.Alignment(ctx, rows, lens)
        iu = np.triu_indices(len(seqs), 1)
        for mode, nm in ((0, "p"), (1, "k2p"), (2, "trans")):
            aln.set_config(mode, 300, True)
            d, c = aln.matrix(counts=True)
            e = sums[key][nm]
            assert int(c["shared"][iu].sum()) == e["sum_shared"]
            assert int(c["c1"][iu].sum()) == e["sum_c1"]
            assert int(c["c2"][iu].sum()) == e["sum_c2"] and int(c["c3"][iu].sum()) == e["sum_c3"]
            du = np.ascontiguousarray(d[iu])
            assert int((~(du < 0)).sum()) == e["valid_pairs"] and int((du == 0).sum()) == e["exact_zero"]
            assert hashlib.sha256(du.tobytes()).hexdigest() == e["sha256_dist_upper"], (key, nm)
        aln.close()


def test_row_distances_and_errors(ctx):
    import taxondna_b200 as t
    rng = np.random.default_rng(5)
    sym, lens = random_alignment(rng, 150, 400, ragged=True)
    aln = t.Alignment(ctx, sym, lens)
    aln.set_config(t.PDM_UNCORRECTED, 100, True)
    ref = oracle.all_pairs(sym, lens, 0, 100, True)
    for q in (0, 77, 149):
        d, sh = aln.row_distances(q)
        assert np.array_equal(u64(d), u64(ref["dist"][q])) and np.array_equal(sh, ref["shared"][q])
    aln.close()
    bad = sym.copy()
    bad[3, 7] = ord("X")
    with pytest.raises(t.TdnaError) as ei:
        t.Alignment(ctx, bad)
    assert ei.value.code == 3  # TDNA_E_SYMBOL


def labelled_alignment(rng, n_species, reps, L, unnamed=0, dup=0, gappy=True):
    """Rows grouped by species (list order == BYNAME order), optional unnamed rows and exact duplicates."""
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    rows, sp, gen = [], [], []
    for s in range(n_species):
        spseq = base.copy()
        m = rng.random(L) < 0.04
        spseq[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
        for _ in range(reps):
            x = spseq.copy()
            m = rng.random(L) < 0.01
            x[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
            if gappy:
                m = rng.random(L) < 0.05
                x[m] = rng.choice(np.frombuffer(SYMS, dtype=np.uint8), size=int(m.sum()))
            rows.append(x)
            sp.append(s)
            gen.append(s // 3)
    for k in range(dup):  # exact duplicates of existing rows under ANOTHER species -> ties, mixed blocks
        src = int(rng.integers(0, len(rows)))
        rows.append(rows[src].copy())
        sp.append((sp[src] + 1) % n_species)
        gen.append(((sp[src] + 1) % n_species) // 3)
    for k in range(unnamed):
        src = int(rng.integers(0, len(rows)))
        x = rows[src].copy()
        x[rng.integers(0, L, size=3)] = ord("A")
        rows.append(x)
        sp.append(-1)
        gen.append(10 ** 6)
    sym = np.stack(rows)
    species = np.array(sp, dtype=np.int32)
    genus = np.array(gen, dtype=np.int32)
    n = len(rows)
    return sym, species, genus, species.copy(), np.arange(n, dtype=np.int32)


@pytest.mark.parametrize("mode,n_species,reps,L,unnamed,dup", [(0, 6, 5, 400, 0, 0), (0, 9, 6, 700, 4, 8), (1, 9, 6, 700, 3, 6),
                                                              (2, 5, 4, 330, 2, 2), (0, 40, 10, 658, 5, 20), (1, 40, 10, 658, 5, 20)])
def test_row_summaries(ctx, mode, n_species, reps, L, unnamed, dup):
    import taxondna_b200 as t
    from ref_rows import compare, row_results
    rng = np.random.default_rng(mode * 77 + n_species)
    sym, sp, gen, epi, full = labelled_alignment(rng, n_species, reps, L, unnamed, dup)
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym)
    aln.set_labels(sp, gen, epi, full)
    aln.set_config(mode, 200, True)
    ref = oracle.all_pairs(sym, np.full(n, L, np.int32), mode, 200, True, threads=8)
    exp = row_results(ref["dist"], sp)
    res = aln.best_match()
    compare(res, exp, ref["shared"])
    # banded: same answers when the matrix is streamed through 128-row bands
    ctx.set_memory_budget(1 << 20)
    try:
        aln.set_config(mode, 200, True)
        res2 = aln.best_match()
    finally:
        ctx.set_memory_budget(64 << 30)
    assert res2.tobytes() == res.tobytes()
    aln.close()


def test_class_distances_edges_and_valid_conspecifics(ctx):
    import taxondna_b200 as t
    rng = np.random.default_rng(11)
    sym, sp, gen, epi, full = labelled_alignment(rng, 12, 5, 500, unnamed=3, dup=5)
    n = sym.shape[0]
    full = rng.permutation(n).astype(np.int32)  # arbitrary full-name order
    full[3] = full[4]                           # two rows with EQUAL full names: pushed twice (PairwiseDistribution.java:172)
    aln = t.Alignment(ctx, sym)
    aln.set_labels(sp, gen, epi, full)
    for mode in (0, 1):
        aln.set_config(mode, 150, True)
        D = oracle.all_pairs(sym, np.full(n, 500, np.int32), mode, 150, True, threads=8)["dist"]
        intra, inter = [], []
        for q in range(n):
            for j in range(n):
                if j == q:
                    continue
                f = np.float32(D[q, j])
                if sp[q] >= 0 and sp[j] == sp[q] and not (full[j] < full[q]) and f != np.float32(-1):
                    intra.append(f)
                if gen[j] == gen[q] and epi[j] != epi[q] and epi[q] < epi[j] and f != np.float32(-1):
                    inter.append(f)
        got_intra = np.sort(aln.class_distances(t.PD_INTRA))
        got_inter = np.sort(aln.class_distances(t.PD_INTER))
        assert np.array_equal(got_intra.view(np.uint32), np.sort(np.array(intra, np.float32)).view(np.uint32))
        assert np.array_equal(got_inter.view(np.uint32), np.sort(np.array(inter, np.float32)).view(np.uint32))
        thr = 0.03
        ii, jj, dd = aln.edges_below(thr)
        exp = sorted((i, j, D[i, j]) for i in range(n) for j in range(i + 1, n) if 0 <= D[i, j] <= thr)
        got = sorted(zip(ii.tolist(), jj.tolist(), dd.tolist()))
        assert got == exp
        v = aln.valid_conspecifics()
        expv = [int(sp[i] >= 0 and any(j != i and sp[j] == sp[i] and D[i, j] != -1.0 for j in range(n))) for i in range(n)]
        assert v.tolist() == expv
    aln.close()


def test_pairs_batch(ctx):
    """tdna_pairs: an explicit pair list in one launch (PairwiseDistances / PairwiseExplorer) == the oracle's entries."""
    import taxondna_b200 as t
    rng = np.random.default_rng(21)
    sym, lens = random_alignment(rng, 180, 500, ragged=True)
    aln = t.Alignment(ctx, sym, lens)
    ii = rng.integers(0, 180, size=1000).astype(np.int32)
    jj = rng.integers(0, 180, size=1000).astype(np.int32)
    for mode in (0, 1, 2):
        aln.set_config(mode, 120, True)
        ref = oracle.all_pairs(sym, lens, mode, 120, True, threads=8)
        d, c = aln.pairs(ii, jj, want_counts=True)
        assert np.array_equal(u64(d), u64(ref["dist"][ii, jj]))
        for k in ("shared", "c1", "c2", "c3"):
            assert np.array_equal(c[k], ref[k][ii, jj]), (mode, k)
    with pytest.raises(t.TdnaError):
        aln.pairs(np.array([0, 180], np.int32), np.array([1, 2], np.int32))
    aln.close()


def test_query_distances(ctx):
    """tdna_query_distances: a sequence that is not in the list against every row (QuerySequence.java:113-145) == the
    oracle's row for that sequence appended to the list; queries shorter than, as long as and longer than the alignment."""
    import taxondna_b200 as t
    rng = np.random.default_rng(33)
    n, L = 300, 200
    sym, lens = random_alignment(rng, n, L, ragged=True)
    aln = t.Alignment(ctx, sym, lens)
    for qlen in (0, 1, 150, 200, 260):
        q = np.frombuffer(SYMS, dtype=np.uint8)[rng.integers(0, len(SYMS), size=qlen)].copy()
        width = max(L, qlen)
        rows2 = np.full((n + 1, width), ord("?"), dtype=np.uint8)
        rows2[:n, :L] = sym
        rows2[n, :qlen] = q
        lens2 = np.concatenate([lens, [qlen]]).astype(np.int32)
        for mode, ambig in MODES:
            aln.set_config(mode, 40, ambig)
            ref = oracle.all_pairs(rows2, lens2, mode, 40, ambig, threads=8)
            d, sh, c = aln.query_distances(q, want_counts=True)
            assert np.array_equal(u64(d), u64(ref["dist"][n, :n])), (qlen, mode, ambig)
            assert np.array_equal(sh, ref["shared"][n, :n]), (qlen, mode, ambig)
            for k in ("shared", "c1", "c2", "c3"):
                assert np.array_equal(c[k], ref[k][n, :n]), (qlen, mode, ambig, k)
    with pytest.raises(t.TdnaError):
        aln.query_distances(b"ACGTU")  # U is outside the normalised alphabet (Sequence.java:1727-1763 rejects it)
    aln.close()
