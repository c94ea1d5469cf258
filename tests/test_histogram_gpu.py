"""tdna_class_histogram / TDNA_WANT_HIST_*: a PairwiseDistribution class as (distinct fp64 distance, count) pairs.  Expanded and cast
to float the histogram must be the very multiset tdna_class_distances lists -- on both count kernels, in every mode, with ambiguity
letters (deferred class pairs), on the dense path, through a one-rank communicator (the multi-GPU merge), and together with Best Match
rows in one pass."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import taxondna_b200 as t
    c = t.Context(0)
    yield c
    c.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    c.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    c.close()


def expand(entries):
    return np.sort(np.repeat(entries["d"].astype(np.float32), entries["count"]).view(np.uint32))


def make(ctx, cfg, mode, scale_n=None):
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate(cfg, names=False, scale_n=scale_n)
    a = t.Alignment(ctx, d["symbols"])
    a.set_labels(d["species_id"], d["genus_id"], d["epithet_rank"], d["fullname_rank"])
    a.set_config(mode, 300, True)
    return a


@pytest.mark.parametrize("cfg,mode,scale_n", [("C3", 0, None), ("C3", 1, None), ("C3", 2, None), ("Hg", 0, 4000), ("Hg", 1, 4000), ("tiny", 0, None)])
def test_histogram_is_the_listed_multiset(ctx, cfg, mode, scale_n):
    import taxondna_b200 as t
    a = make(ctx, cfg, mode, scale_n)
    for kern in (t.KERNEL_AUTO, t.KERNEL_POPC):
        ctx.set_option(t.OPT_COUNT_KERNEL, kern)
        for klass in (t.PD_INTRA, t.PD_INTER):
            a.set_config(mode, 300, True)
            ref = np.sort(a.class_distances(klass).view(np.uint32))
            a.set_config(mode, 300, True)
            h, pushes = a.class_histogram(klass)
            assert pushes == ref.size == int(h["count"].sum())
            assert len(np.unique(h["d"].view(np.uint64))) == len(h)
            assert np.array_equal(expand(h), ref)
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    # dense path
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
    a.set_config(mode, 300, True)
    ref = np.sort(a.class_distances(t.PD_INTER).view(np.uint32))
    h, pushes = a.class_histogram(t.PD_INTER)
    assert np.array_equal(expand(h), ref)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    a.close()


def test_rows_and_both_histograms_from_one_pass(ctx):
    import taxondna_b200 as t
    a = make(ctx, "C3", 1)
    ref_rows = a.best_match()
    ref_intra = np.sort(a.class_distances(t.PD_INTRA).view(np.uint32))
    ref_inter = np.sort(a.class_distances(t.PD_INTER).view(np.uint32))
    a.set_config(1, 300, True)
    rows = np.zeros(a.n, dtype=t.ROW_RESULT_DTYPE)
    hi = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
    he = np.zeros(1 << 16, dtype=t.HIST_ENTRY_DTYPE)
    io = t.Analysis()
    io.want = t.WANT_BEST_MATCH | t.WANT_HIST_INTRA | t.WANT_HIST_INTER
    io.rows = rows.ctypes.data
    io.hist_intra, io.hist_intra_cap = hi.ctypes.data, hi.size
    io.hist_inter, io.hist_inter_cap = he.ctypes.data, he.size
    launches0 = ctx.launch_count()
    t.check(ctx.L.tdna_analyze(a.h, ctypes.byref(io)))
    assert rows.tobytes() == ref_rows.tobytes()
    assert np.array_equal(expand(hi[: io.n_hist_intra]), ref_intra) and io.hist_intra_pushes == ref_intra.size
    assert np.array_equal(expand(he[: io.n_hist_inter]), ref_inter) and io.hist_inter_pushes == ref_inter.size
    with pytest.raises(t.TdnaError):
        io.want = t.WANT_INTRA | t.WANT_HIST_INTRA
        t.check(ctx.L.tdna_analyze(a.h, ctypes.byref(io)))
    a.close()


def test_a_table_that_fills_up_is_regrown(ctx):
    """Random sequences: nearly every congeneric pair has its own distance -- far more distinct values than a 4,096-slot table."""
    import taxondna_b200 as t
    ctx.set_option(t.OPT_HIST_SLOTS, 1 << 12)
    rng = np.random.default_rng(11)
    n, L = 1500, 2000
    sym = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(n, L))
    lens = rng.integers(1200, L + 1, size=n).astype(np.int32)
    a = t.Alignment(ctx, sym, lens)
    ctx.set_option(t.OPT_HIST_SLOTS, 1 << 18)
    sp = (np.arange(n) // 3).astype(np.int32)
    a.set_labels(sp, np.zeros(n, np.int32), sp, np.arange(n, dtype=np.int32))
    a.set_config(0, 300, True)
    ref = np.sort(a.class_distances(t.PD_INTER).view(np.uint32))
    h, pushes = a.class_histogram(t.PD_INTER)
    assert len(h) > (1 << 14) and pushes == ref.size
    assert np.array_equal(expand(h), ref)
    a.close()


def test_histogram_through_a_one_rank_communicator(ctx):
    import taxondna_b200 as t
    a = make(ctx, "C3", 1)
    ref, pushes = a.class_histogram(t.PD_INTRA)
    try:
        ctx.comm_init(t.Context.comm_unique_id(), 0, 1)
    except t.TdnaError:
        pass  # already has one
    ctx.set_option(t.OPT_MULTI, 2)
    a.set_config(1, 300, True)
    got, p2 = a.class_histogram(t.PD_INTRA)
    ctx.set_option(t.OPT_MULTI, 1)
    assert p2 == pushes and got.tobytes() == ref.tobytes()
    a.close()
