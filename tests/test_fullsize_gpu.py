"""BASELINE.json's configurations at FULL size (C2, C3, C4, H) through properties that do not need an N x N oracle:
kernel (b) against kernel (a) byte for byte, sampled query rows against the CPU oracle, and invariants of the records
(a minimum is a minimum, an index points at its distance, counts add up).  The oracle only computes the sampled rows."""
import os
import sys

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ctx():
    import taxondna_b200 as t
    c = t.Context(0)
    yield c
    c.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    c.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    c.close()


def u64(a):
    return np.ascontiguousarray(a, dtype=np.float64).view(np.uint64)


def sample_blocks(n, blocks=4, rows=8, seed=0):
    rng = np.random.default_rng(seed)
    starts = sorted(set([0, max(0, n - rows)] + [int(x) for x in rng.integers(0, max(1, n - rows), size=blocks - 2)]))
    return [(s, min(n, s + rows)) for s in starts]


def check_rows_against_oracle(res, sym, species, mode, min_overlap, blocks):
    """Every sampled query row: what the record says is what the oracle's row of distances says."""
    n, L = sym.shape
    lens = np.full(n, L, np.int32)
    named = species >= 0
    for b0, b1 in blocks:
        ref = oracle.all_pairs(sym, lens, mode, min_overlap, True, row_begin=b0, row_end=b1, threads=8)
        for q in range(b0, b1):
            d = ref["dist"][q - b0]
            r = res[q]
            valid = ~(d < 0)
            valid[q] = False
            assert r["n_valid"] == int(valid.sum()), q
            if not valid.any():
                assert r["best"] == -1 and r["d_best"] == -1.0, q
                continue
            finite = valid & np.isfinite(d)
            if (valid & ~np.isfinite(d)).any():
                continue  # NaN / +Inf rows are flagged and resolved on the host: not this test's subject
            assert u64(r["d_best"]) == u64(d[finite].min()), q
            assert u64(d[r["best"]]) == u64(r["d_best"]), q
            assert r["tie_count"] == int((d[finite] == r["d_best"]).sum()) - 1, q
            for field, idx, sel in (("d_con", "con", finite & named & (species == species[q]) & bool(named[q])),
                                    ("d_allo", "allo", finite & named & ((species != species[q]) | (not named[q])))):
                if sel.any():
                    assert u64(r[field]) == u64(d[sel].min()), (q, field)
                    assert u64(d[r[idx]]) == u64(r[field]) and sel[r[idx]], (q, idx)
                    if field == "d_con":
                        assert r["shared_con"] == ref["shared"][q - b0][r["con"]], q
                    else:
                        assert r["shared_allo"] == ref["shared"][q - b0][r["allo"]], q
                else:
                    assert r[field] == -1.0 and r[idx] == -1, (q, field)


def record_invariants(res, n):
    has = res["best"] >= 0
    assert (res["best"][has] < n).all() and (res["best"][has] != np.nonzero(has)[0]).all()
    assert (res["d_best"][has] >= 0).all() | np.isnan(res["d_best"][has]).any()
    for f in ("d_con", "d_allo"):
        m = has & (res[f] >= 0)
        assert (res["d_best"][m] <= res[f][m]).all(), f  # the best match is the closest of all
    assert ((res["n_valid"] >= 0) & (res["n_valid"] <= n - 1)).all()
    assert (res["tie_count"][has] >= 0).all()


def both_kernels(ctx, aln):
    import taxondna_b200 as t
    out = {}
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    for name, k in (("tensor", t.KERNEL_TENSOR), ("popc", t.KERNEL_POPC)):
        ctx.set_option(t.OPT_COUNT_KERNEL, k)
        out[name] = aln.best_match()
        assert aln.last_count_kernel() == k
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    return out


def labelled(ctx, cfg):
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate(cfg, names=False)
    aln = t.Alignment(ctx, d["symbols"])
    aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    return d, aln


def test_h_50000_by_658_streaming_best_match(ctx):
    """The headline configuration: 1,249,975,000 pairs per pass."""
    d, aln = labelled(ctx, "H")
    n = d["symbols"].shape[0]
    assert d["symbols"].shape == (50000, 658)
    aln.set_config(0, 300, True)
    out = both_kernels(ctx, aln)
    assert out["tensor"].tobytes() == out["popc"].tobytes()
    record_invariants(out["tensor"], n)
    # a checksum of checksums: on this clean alignment every pair is valid, so every row counts n - 1 valid columns
    assert int(out["tensor"]["n_valid"].astype(np.int64).sum()) == n * (n - 1)
    check_rows_against_oracle(out["tensor"], d["symbols"], d["species_id"], 0, 300, sample_blocks(n, seed=1))
    # idempotence: the same pass again gives the same bytes
    assert aln.best_match().tobytes() == out["tensor"].tobytes()
    aln.close()


def test_c3_8000_by_658_k2p_best_close_match_and_intraspecific(ctx):
    import taxondna_b200 as t
    d, aln = labelled(ctx, "C3")
    n = d["symbols"].shape[0]
    assert d["symbols"].shape == (8000, 658)
    aln.set_config(1, 300, True)
    out = both_kernels(ctx, aln)
    assert out["tensor"].tobytes() == out["popc"].tobytes()
    record_invariants(out["tensor"], n)
    check_rows_against_oracle(out["tensor"], d["symbols"], d["species_id"], 1, 300, sample_blocks(n, seed=2))
    # the intraspecific distances (PairwiseSummary's percentile input): one pass on either kernel and the dense sweep agree
    got = {}
    for name, path, k in (("tensor", t.PATH_STREAM, t.KERNEL_TENSOR), ("popc", t.PATH_STREAM, t.KERNEL_POPC), ("dense", t.PATH_DENSE, t.KERNEL_AUTO)):
        ctx.set_option(t.OPT_BEST_MATCH_PATH, path)
        ctx.set_option(t.OPT_COUNT_KERNEL, k)
        got[name] = np.sort(aln.class_distances(t.PD_INTRA))
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    sp = d["species_id"]
    expect = sum(int(c) * (int(c) - 1) // 2 for c in np.bincount(sp[sp >= 0]))
    assert len(got["dense"]) == expect
    assert np.array_equal(got["tensor"].view(np.uint32), got["dense"].view(np.uint32)) and np.array_equal(got["popc"].view(np.uint32), got["dense"].view(np.uint32))
    aln.close()


@pytest.mark.parametrize("mode", [0, 1])
def test_c4_5000_by_2664_gappy_ambiguous(ctx, mode):
    """20 % of the sites are '?', gaps (internal and external) or IUPAC ambiguity letters: kernel (b) only bounds most pairs."""
    d, aln = labelled(ctx, "C4")
    n = d["symbols"].shape[0]
    assert d["symbols"].shape == (5000, 2664)
    aln.set_config(mode, 300, True)
    out = both_kernels(ctx, aln)
    assert out["tensor"].tobytes() == out["popc"].tobytes()
    record_invariants(out["tensor"], n)
    check_rows_against_oracle(out["tensor"], d["symbols"], d["species_id"], mode, 300, sample_blocks(n, blocks=3, rows=6, seed=3 + mode))
    aln.close()


def test_c2_2185_by_2664_full_matrix(ctx):
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate("C2", names=False)
    sym = d["symbols"]
    n, L = sym.shape
    assert (n, L) == (2185, 2664)
    aln = t.Alignment(ctx, sym)
    aln.set_config(0, 300, True)
    D = aln.matrix()
    assert D.shape == (n, n)
    assert np.array_equal(u64(D), u64(D.T))                      # symmetric, bit for bit
    diag = np.diagonal(D)
    assert ((diag == 0.0) | (diag == -1.0)).all()                # a sequence is at distance 0 from itself (or has too few sites)
    assert (((D >= 0.0) & (D <= 1.0)) | (D == -1.0)).all()       # a p-distance or the "inadequate overlap" sentinel
    lens = np.full(n, L, np.int32)
    for b0, b1 in sample_blocks(n, blocks=4, rows=6, seed=4):
        ref = oracle.all_pairs(sym, lens, 0, 300, True, row_begin=b0, row_end=b1, threads=8)
        assert np.array_equal(u64(D[b0:b1]), u64(ref["dist"]))
    # encode -> decode: the rows the library was given are the rows it answers for (a single pair straight from the planes)
    (shared, ident, _, _), dist = aln.pair(7, 1234)
    assert u64(dist) == u64(D[7, 1234])
    aln.close()


def test_n_200000_kernel_b_equals_kernel_a(ctx):
    """2 * 10^10 pairs: the tcgen05 kernel's records against the popc kernel's, byte for byte, plus sampled oracle rows."""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate("H", names=False, scale_n=200000)
    n = d["symbols"].shape[0]
    assert n == 200000
    aln = t.Alignment(ctx, d["symbols"])
    aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    aln.set_config(0, 300, True)
    out = both_kernels(ctx, aln)
    assert out["tensor"].tobytes() == out["popc"].tobytes()
    record_invariants(out["tensor"], n)
    assert int(out["tensor"]["n_valid"].astype(np.int64).sum()) == n * (n - 1)
    check_rows_against_oracle(out["tensor"], d["symbols"], d["species_id"], 0, 300, sample_blocks(n, blocks=3, rows=4, seed=5))
    aln.close()


def test_c5_shape_one_million_rows_rows_and_both_histograms_in_one_pass(ctx):
    """BASELINE.json configs[4] on one GPU: 1,000,000 x 658, 5 * 10^11 pairs, streaming Best Match + the intra / congeneric-inter
    histograms from ONE pass, nothing O(pairs) anywhere.  Checked through what does not need an N x N oracle: sampled query rows
    against the oracle, record invariants, and the histograms' totals against the counts the labels imply."""
    import ctypes

    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate("C5", names=False)
    sym = d["symbols"]
    n = sym.shape[0]
    assert sym.shape == (1000000, 658)
    aln = t.Alignment(ctx, sym)
    aln.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    aln.set_config(0, 300, True)
    rows = np.zeros(n, dtype=t.ROW_RESULT_DTYPE)
    hi = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
    he = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
    io = t.Analysis()
    io.want = t.WANT_BEST_MATCH | t.WANT_HIST_INTRA | t.WANT_HIST_INTER
    io.rows = rows.ctypes.data
    io.hist_intra, io.hist_intra_cap = hi.ctypes.data, hi.size
    io.hist_inter, io.hist_inter_cap = he.ctypes.data, he.size
    t.check(ctx.L.tdna_analyze(aln.h, ctypes.byref(io)))
    assert aln.last_count_kernel() == t.KERNEL_TENSOR
    record_invariants(rows, n)
    assert int(rows["n_valid"].astype(np.int64).sum()) == n * (n - 1)  # clean data: every pair is valid
    # 100,000 species of 10: C(10, 2) intraspecific pairs each; 1,000 genera of 100 species: C(100, 2) * 10 * 10 congeneric pairs each
    assert io.hist_intra_pushes == 100000 * 45 and int(hi[: io.n_hist_intra]["count"].sum()) == 100000 * 45
    assert io.hist_inter_pushes == 1000 * 4950 * 100 and int(he[: io.n_hist_inter]["count"].sum()) == 1000 * 4950 * 100
    assert 0 < io.n_hist_intra <= hi.size and 0 < io.n_hist_inter <= he.size
    # a conspecific is each row's best match on this data
    sp = d["species_id"]
    assert (sp[rows["best"]] == sp).mean() > 0.999
    check_rows_against_oracle(rows, sym, sp, 0, 300, [(0, 3), (499999, 500002), (n - 3, n)])
    aln.close()
