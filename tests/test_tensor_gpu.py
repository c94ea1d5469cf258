"""Count kernel (b): the tcgen05 / TMEM one-hot int8 contraction must reproduce the oracle's integer
counts bit for bit, and the streaming Best Match built on it must equal the popc kernel's."""
import os
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

CLEAN = b"ACGT-?"


@pytest.fixture(scope="module")
def ctx():
    import taxondna_b200 as t
    c = t.Context(0)
    yield c
    c.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    c.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    c.close()


def clean_alignment(rng, n, L, p_special=0.08, ragged=False):
    base = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=L)
    sym = np.tile(base, (n, 1))
    mut = rng.random((n, L)) < 0.12
    sym[mut] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(mut.sum()))
    sp = rng.random((n, L)) < p_special
    sym[sp] = rng.choice(np.frombuffer(CLEAN, dtype=np.uint8), size=int(sp.sum()))
    for r in range(0, n, 7):  # leading / trailing external gaps ('_' after Sequence.changeSequence)
        k = int(rng.integers(0, L // 4))
        sym[r, :k] = ord("_")
    lens = None
    if ragged:
        lens = rng.integers(max(1, L // 2), L + 1, size=n).astype(np.int32)
        lens[0] = L
    return sym, lens


@pytest.mark.parametrize("n,L,ragged", [(2, 40, False), (130, 200, False), (300, 658, True), (257, 1539, False), (140, 2664, True), (520, 330, False)])
def test_tensor_counts_equal_the_oracle(ctx, n, L, ragged):
    import taxondna_b200 as t
    rng = np.random.default_rng(n * 31 + L)
    sym, lens = clean_alignment(rng, n, L, ragged=ragged)
    aln = t.Alignment(ctx, sym, lens)
    aln.set_config(t.PDM_UNCORRECTED, max(1, L // 3), True)
    got = aln.tensor_counts()
    ref = oracle.all_pairs(sym, lens if lens is not None else np.full(n, L, np.int32), 0, max(1, L // 3), True, threads=8)
    iu = np.triu_indices(n, 0)  # the triangle incl. the diagonal is always covered
    assert np.array_equal(got["shared"][iu], ref["shared"][iu]), "shared"
    assert np.array_equal(got["c1"][iu], ref["c1"][iu]), "ident"
    aln.close()


def test_not_eligible_is_reported(ctx):
    import taxondna_b200 as t
    rng = np.random.default_rng(1)
    sym, _ = clean_alignment(rng, 50, 100)
    sym[3, 5] = ord("N")
    aln = t.Alignment(ctx, sym)
    with pytest.raises(t.TdnaError) as ei:
        aln.tensor_counts()
    assert ei.value.code == 4
    aln.close()
    sym, _ = clean_alignment(rng, 50, 100)
    aln = t.Alignment(ctx, sym)
    aln.set_config(t.PDM_K2P, 10, True)
    with pytest.raises(t.TdnaError):
        aln.tensor_counts()
    aln.close()


AMBIG = b"RYKMSWBDHVN"


def sprinkle_ambiguity(rng, sym, rate):
    """IUPAC ambiguity letters everywhere at `rate`; every ninth row is mostly ambiguity letters and '?' so that its overlaps
    straddle minOverlap (kernel (b) only bounds such pairs: the exact counts must settle them)."""
    m = rng.random(sym.shape) < rate
    sym[m] = rng.choice(np.frombuffer(AMBIG, dtype=np.uint8), size=int(m.sum()))
    for r in range(4, sym.shape[0], 9):
        m = rng.random(sym.shape[1]) < 0.55
        sym[r, m] = rng.choice(np.frombuffer(AMBIG + b"????", dtype=np.uint8), size=int(m.sum()))


@pytest.mark.parametrize("shuffle,ragged,L,ambig", [(False, False, 658, 0), (True, True, 500, 0), (False, True, 2664, 0),
                                                    (False, False, 658, 0.03), (True, True, 500, 0.05), (False, True, 2664, 0.04)])
def test_streaming_best_match_on_the_tensor_kernel(ctx, shuffle, ragged, L, ambig):
    import taxondna_b200 as t
    from test_stream_gpu import first_difference
    from test_gpu_parity import labelled_alignment
    rng = np.random.default_rng(17 + L)
    sym, sp, gen, epi, full = labelled_alignment(rng, 36, 9, L, unnamed=4, dup=10, gappy=False)
    n = sym.shape[0]
    m = rng.random(sym.shape) < 0.04
    sym[m] = rng.choice(np.frombuffer(CLEAN, dtype=np.uint8), size=int(m.sum()))
    if ambig:
        sprinkle_ambiguity(rng, sym, ambig)
    lens = rng.integers(L // 2, L + 1, size=n).astype(np.int32) if ragged else None
    if shuffle:
        perm = rng.permutation(n)
        sym, sp, gen, epi = sym[perm], sp[perm], gen[perm], epi[perm]
        if lens is not None:
            lens = lens[perm]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    aln.set_config(t.PDM_UNCORRECTED, L // 2, True)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_POPC)
    a = aln.analyze(best_match=True, intra=True, inter=True, edges=0.03)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_TENSOR)
    b = aln.analyze(best_match=True, intra=True, inter=True, edges=0.03)
    assert aln.last_count_kernel() == t.KERNEL_TENSOR
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    assert first_difference(a["rows"], b["rows"]) is None, first_difference(a["rows"], b["rows"])
    assert np.array_equal(np.sort(a["intra"]).view(np.uint32), np.sort(b["intra"]).view(np.uint32))
    assert np.array_equal(np.sort(a["inter"]).view(np.uint32), np.sort(b["inter"]).view(np.uint32))
    assert sorted(zip(a["edge_i"].tolist(), a["edge_j"].tolist(), a["edge_d"].tolist())) == \
        sorted(zip(b["edge_i"].tolist(), b["edge_j"].tolist(), b["edge_d"].tolist()))
    aln.close()


@pytest.mark.parametrize("shuffle,ragged,L,saturated,ambig", [(False, False, 658, 0, 0), (True, True, 700, 0, 0), (False, False, 400, 25, 0),
                                                              (False, False, 658, 0, 0.04), (True, True, 700, 10, 0.05),
                                                              (False, True, 1500, 0, 0.03), (False, False, 200, 0, 0)])  # L >= 1024: the five-dimension code
def test_k2p_streaming_on_the_tensor_kernel(ctx, shuffle, ragged, L, saturated, ambig):
    """K2P: the accumulator carries (n, ts + tv) -- enough for the conservative filter -- and the exact (n, ts, tv) of the
    queued pairs come from the bit-planes; records, classes and edges must equal the popc kernel's."""
    import taxondna_b200 as t
    from test_stream_gpu import first_difference
    from test_gpu_parity import labelled_alignment
    rng = np.random.default_rng(99 + L)
    sym, sp, gen, epi, full = labelled_alignment(rng, 30, 8, L, unnamed=3, dup=8, gappy=False)
    m = rng.random(sym.shape) < 0.05
    sym[m] = rng.choice(np.frombuffer(CLEAN, dtype=np.uint8), size=int(m.sum()))
    if ambig:  # R and Y are inside the K2P model (purine / pyrimidine), the other ambiguity letters outside it
        sprinkle_ambiguity(rng, sym, ambig)
    if saturated:  # unrelated rows: P + Q ~ 0.75 -> NaN / +Inf / huge distances must be flagged exactly as by the popc kernel
        rnd = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(saturated, L))
        sym = np.concatenate([sym, rnd])
        extra = rng.integers(0, 30, size=saturated).astype(np.int32)
        sp, gen, epi = np.concatenate([sp, extra]), np.concatenate([gen, extra // 3]), np.concatenate([epi, extra])
    n = sym.shape[0]
    full = np.arange(n, dtype=np.int32)
    lens = rng.integers(L // 2, L + 1, size=n).astype(np.int32) if ragged else None
    if shuffle:
        perm = rng.permutation(n)
        sym, sp, gen, epi = sym[perm], sp[perm], gen[perm], epi[perm]
        if lens is not None:
            lens = lens[perm]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    aln.set_config(t.PDM_K2P, L // 3, True)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_POPC)
    a = aln.analyze(best_match=True, intra=True, inter=True, edges=0.03)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_TENSOR)
    b = aln.analyze(best_match=True, intra=True, inter=True, edges=0.03)
    assert aln.last_count_kernel() == t.KERNEL_TENSOR
    # and back to uncorrected p on the same alignment: the operands are re-encoded for the other symbol set
    aln.set_config(t.PDM_UNCORRECTED, L // 3, True)
    bp = aln.best_match()
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_POPC)
    ap = aln.best_match()
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    if saturated:
        assert (a["rows"]["flags"] & t.ROW_NONFINITE).any()
    assert np.array_equal(a["rows"]["flags"], b["rows"]["flags"]) and np.array_equal(a["rows"]["n_valid"], b["rows"]["n_valid"])
    ok = (a["rows"]["flags"] & t.ROW_NONFINITE) == 0
    assert first_difference(a["rows"][ok], b["rows"][ok]) is None, first_difference(a["rows"][ok], b["rows"][ok])
    assert np.array_equal(np.sort(a["intra"]).view(np.uint32), np.sort(b["intra"]).view(np.uint32))
    assert np.array_equal(np.sort(a["inter"]).view(np.uint32), np.sort(b["inter"]).view(np.uint32))
    assert sorted(zip(a["edge_i"].tolist(), a["edge_j"].tolist(), a["edge_d"].tolist())) == \
        sorted(zip(b["edge_i"].tolist(), b["edge_j"].tolist(), b["edge_d"].tolist()))
    assert first_difference(ap, bp) is None, first_difference(ap, bp)
    aln.close()


@pytest.mark.parametrize("n,L", [(1024, 2664), (700, 1300)])
def test_ambiguity_codes_at_scale_fall_back_cleanly(ctx, n, L):
    """An alignment full of IUPAC ambiguity letters ('?', gaps and external gaps too): kernel (b) cannot give exact counts
    (tdna_tensor_counts says so and leaves the device usable) but serves the streaming consumers as a filter, with the
    exact counts of the queued pairs from the bit-planes -- identical to kernel (a) and to the dense path.
    (Regression: the C4 shape, 5,000 x 2,664 with 4 % ambiguity letters, once took the context down in the encoder.)"""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate("C4", names=False, genera=8, species=8, reps=max(1, n // 64), L=L)
    sym = np.ascontiguousarray(d["symbols"][:n])
    n = sym.shape[0]
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    aln = t.Alignment(ctx, sym)
    aln.set_labels(*(d[k][:n] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")))
    aln.set_config(t.PDM_UNCORRECTED, 300, True)
    with pytest.raises(t.TdnaError) as ei:
        aln.tensor_counts()
    assert ei.value.code == 4
    (shared, ident, _, _), dist = aln.pair(0, 1)  # the context is alive
    ref = oracle.all_pairs(sym[:2], np.full(2, L, np.int32), 0, 300, True)
    assert shared == ref["shared"][0, 1] and ident == ref["c1"][0, 1]
    intra = aln.class_distances(t.PD_INTRA)  # AUTO: kernel (b) bounds the pairs with ambiguity letters, the planes settle them
    assert aln.last_count_kernel() == t.KERNEL_TENSOR
    rows_b = aln.best_match()
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_POPC)
    rows_a = aln.best_match()
    intra_a = aln.class_distances(t.PD_INTRA)
    assert aln.last_count_kernel() == t.KERNEL_POPC
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
    dense = aln.class_distances(t.PD_INTRA)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    assert np.array_equal(np.sort(intra), np.sort(dense)) and np.array_equal(np.sort(intra_a), np.sort(dense))
    assert rows_a.tobytes() == rows_b.tobytes()
    aln.close()


@pytest.mark.parametrize("reps,L", [(600, 300), (300, 658)])
def test_species_blocks_wider_than_the_seed_window(ctx, reps, L):
    """Hundreds of conspecifics in a row: a 256-row diagonal block lies inside ONE species, so the seed pass must also visit the
    block with the nearest species of the other id parity (else one of the three minima stays unbounded and every pair of
    the row is queued), and every conspecific closer than the nearest other species is a candidate (the record buffers grow
    on demand).  Kernel (b) == kernel (a) == the dense path."""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate("H", names=False, genera=2, species=3, reps=reps, L=L)
    sym = d["symbols"]
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym)
    aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    aln.set_config(t.PDM_UNCORRECTED, L // 2, True)
    got = {}
    for name, path, k in (("dense", t.PATH_DENSE, t.KERNEL_AUTO), ("popc", t.PATH_STREAM, t.KERNEL_POPC), ("tensor", t.PATH_STREAM, t.KERNEL_TENSOR)):
        ctx.set_option(t.OPT_BEST_MATCH_PATH, path)
        ctx.set_option(t.OPT_COUNT_KERNEL, k)
        got[name] = aln.best_match()
        if name != "dense":
            assert aln.last_count_kernel() == k
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    assert got["tensor"].tobytes() == got["dense"].tobytes() and got["popc"].tobytes() == got["dense"].tobytes()
    assert (got["dense"]["n_valid"] == n - 1).all()
    aln.close()


# ---- transversions only (PDM_TRANS_ONLY) on count kernel (b): six dimensions, exact for every letter (tdna_tc.cuh: EncodeValues) ----
def trans_alignment(rng, n_species, reps, L, ambig, gaps):
    from test_gpu_parity import labelled_alignment
    sym, sp, gen, epi, full = labelled_alignment(rng, n_species, reps, L, unnamed=3, dup=6, gappy=False)
    m = rng.random(sym.shape) < 0.06  # transversions need more divergence than the generator's default to be interesting
    sym[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
    if ambig:
        sprinkle_ambiguity(rng, sym, ambig)
    if gaps:  # internal gaps, '?' and external runs: a letter outside AGRCTY is only "shared" with a gap (Sequence.java:1446-1451, 1472-1478)
        m = rng.random(sym.shape) < gaps
        sym[m] = ord("-")
        m = rng.random(sym.shape) < gaps / 2
        sym[m] = ord("?")
        for r in range(0, sym.shape[0], 5):
            k = int(rng.integers(1, L // 5))
            sym[r, :k] = ord("_")
    return sym, sp, gen, epi, full


@pytest.mark.parametrize("n,L,ambig,gaps", [(130, 200, 0.0, 0.0), (300, 658, 0.04, 0.03), (257, 1539, 0.02, 0.05), (140, 2664, 0.05, 0.02)])
def test_transversion_counts_from_the_tensor_kernel_equal_the_oracle(ctx, n, L, ambig, gaps):
    import taxondna_b200 as t
    rng = np.random.default_rng(n + L)
    sym, *_ = trans_alignment(rng, n // 5, 5, L, ambig, gaps)
    n = sym.shape[0]
    lens = rng.integers(L // 2, L + 1, size=n).astype(np.int32)
    aln = t.Alignment(ctx, sym, lens)
    aln.set_config(t.PDM_TRANS_ONLY, max(1, L // 3), True)
    got = aln.tensor_counts()
    ref = oracle.all_pairs(sym, lens, 2, max(1, L // 3), True, threads=8)
    iu = np.triu_indices(n, 0)
    assert np.array_equal(got["shared"][iu], ref["shared"][iu]), "shared"
    assert np.array_equal(got["c1"][iu], ref["c1"][iu]), "transversions"
    assert aln.tensor_dims() == 6
    aln.close()


@pytest.mark.parametrize("shuffle,ragged,L,ambig,gaps", [(False, False, 658, 0.0, 0.0), (False, False, 658, 0.03, 0.03), (True, True, 500, 0.05, 0.04),
                                                         (False, True, 2664, 0.04, 0.02), (False, False, 250, 0.0, 0.0)])
def test_transversions_only_streaming_on_the_tensor_kernel(ctx, shuffle, ragged, L, ambig, gaps):
    import taxondna_b200 as t
    from test_stream_gpu import first_difference
    rng = np.random.default_rng(23 + L)
    sym, sp, gen, epi, full = trans_alignment(rng, 36, 9, L, ambig, gaps)
    n = sym.shape[0]
    lens = rng.integers(L // 2, L + 1, size=n).astype(np.int32) if ragged else None
    if shuffle:
        perm = rng.permutation(n)
        sym, sp, gen, epi = sym[perm], sp[perm], gen[perm], epi[perm]
        if lens is not None:
            lens = lens[perm]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    aln.set_config(t.PDM_TRANS_ONLY, L // 2, True)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
    dense = aln.best_match()
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_POPC)
    a = aln.analyze(best_match=True, intra=True, inter=True, edges=0.01)
    ha = {k: aln.class_histogram(kl)[0] for k, kl in (("hist_intra", t.PD_INTRA), ("hist_inter", t.PD_INTER))}
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)  # AUTO takes kernel (b) for this mode too
    b = aln.analyze(best_match=True, intra=True, inter=True, edges=0.01)
    assert aln.last_count_kernel() == t.KERNEL_TENSOR
    hb = {k: aln.class_histogram(kl)[0] for k, kl in (("hist_intra", t.PD_INTRA), ("hist_inter", t.PD_INTER))}
    assert aln.last_count_kernel() == t.KERNEL_TENSOR
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    assert first_difference(dense, a["rows"]) is None
    assert first_difference(a["rows"], b["rows"]) is None, first_difference(a["rows"], b["rows"])
    assert np.array_equal(np.sort(a["intra"]).view(np.uint32), np.sort(b["intra"]).view(np.uint32))
    assert np.array_equal(np.sort(a["inter"]).view(np.uint32), np.sort(b["inter"]).view(np.uint32))
    assert sorted(zip(a["edge_i"].tolist(), a["edge_j"].tolist(), a["edge_d"].tolist())) == \
        sorted(zip(b["edge_i"].tolist(), b["edge_j"].tolist(), b["edge_d"].tolist()))
    for k in ("hist_intra", "hist_inter"):
        ea, eb = ha[k], hb[k]
        assert sorted(zip(ea["d"].view(np.uint64).tolist(), ea["count"].tolist())) == sorted(zip(eb["d"].view(np.uint64).tolist(), eb["count"].tolist())), k
    aln.close()


@pytest.mark.parametrize("mode,cfg,n", [(0, "C3", 3000), (1, "Hg", 2600), (2, "C3", 1500)])
def test_cluster_variants_of_the_bulk_launch_give_the_same_bytes(ctx, mode, cfg, n):
    """TDNA_OPT_CLUSTER: pairs of CTAs sharing a b-block by multicast (2, the default), cta_group::2 MMAs (3), clusters of 2 x 2 tiles
    (4).  The variants differ in how operands reach shared memory, not in what is computed (A/B, the pairs win)."""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate(cfg, names=False, scale_n=n, seed=77)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_TENSOR)
    out = {}
    try:
        for cl in (2, 3, 4):
            ctx.set_option(t.OPT_CLUSTER, cl)
            aln = t.Alignment(ctx, d["symbols"])  # (mode 3 encodes the b-side in 128-row panels: a fresh alignment per mode)
            aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
            aln.set_config(mode, 300, True)
            out[cl] = aln.analyze(best_match=True, intra=True, edges=0.02)
            assert aln.last_count_kernel() == t.KERNEL_TENSOR
            aln.close()
    finally:
        ctx.set_option(t.OPT_CLUSTER, 2)
        ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    for cl in (3, 4):
        assert out[cl]["rows"].tobytes() == out[2]["rows"].tobytes(), cl
        assert np.array_equal(np.sort(out[cl]["intra"]).view(np.uint32), np.sort(out[2]["intra"]).view(np.uint32))
        assert sorted(zip(out[cl]["edge_i"].tolist(), out[cl]["edge_j"].tolist())) == sorted(zip(out[2]["edge_i"].tolist(), out[2]["edge_j"].tolist()))
