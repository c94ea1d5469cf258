"""tdna_analyze_host: create + labels + configuration + analysis in one call from host buffers; and the chunked upload behind
tdna_alignment_create (8,192+ rows: row chunks on a copy stream, each packed and encoded into count kernel (b)'s operands as it
lands).  The records must be the bytes the separate calls give -- with ambiguity letters that only show up in a late chunk, in K2P
(operands rebuilt for the other mode), with ragged rows, and a symbol outside the alphabet must be reported however late it arrives."""
import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import taxondna_b200 as t
    c = t.Context(0)
    c.set_option(t.OPT_PIPELINE_CHUNKS, 8)  # the pass pipelined with the upload (off by default: it measured slower) is what these tests stress
    yield c
    c.close()


def test_default_is_one_pass_after_the_upload():
    import taxondna_b200 as t
    c = t.Context(0)
    try:
        sym, labels = labelled("H", 17000)
        ref = separate(c, sym, labels, 0, 300)
        got = t.best_match_host(c, sym, labels, 0, 300)
        assert got.tobytes() == ref["rows"].tobytes()
        assert c.pipelined_passes() == 0
    finally:
        c.close()


def labelled(cfg, n, seed=None):
    from taxondna_b200 import synth
    d = synth.generate(cfg, names=False, scale_n=n, seed=seed)
    return d["symbols"], [d[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]


def separate(ctx, sym, labels, mode, mo, lengths=None):
    import taxondna_b200 as t
    a = t.Alignment(ctx, sym, lengths)
    a.set_labels(*labels)
    a.set_config(mode, mo, True)
    r = a.analyze(best_match=True, intra=True, edges=0.01)
    a.close()
    return r


@pytest.mark.parametrize("cfg,n,mode", [("H", 20000, 0), ("H", 16500, 1), ("Hg", 17000, 0), ("Hg", 17000, 1), ("H", 3000, 0), ("H", 41000, 0), ("Hg", 33000, 1)])
def test_one_call_equals_the_separate_calls(ctx, cfg, n, mode):
    import taxondna_b200 as t
    sym, labels = labelled(cfg, n)
    ref = separate(ctx, sym, labels, mode, 300)
    before = ctx.pipelined_passes()
    for _ in range(2):  # the second call starts on what the first one learned about the list (ambiguity letters? internal gaps?)
        got = t.best_match_host(ctx, sym, labels, mode, 300)
        assert got.tobytes() == ref["rows"].tobytes()
    # 16,384+ rows: the pass ran in region launches while the row chunks were still arriving (Hg in p: its internal gaps send the
    # list to the five-dimension code, i.e. to the ordinary pass)
    if n >= 16384 and not (cfg == "Hg" and mode == 0):
        assert ctx.pipelined_passes() >= before + 1
    else:
        assert ctx.pipelined_passes() == before


def test_a_shuffled_list_overflows_the_pipelined_pass_and_still_matches(ctx):
    """Rows in random order: the diagonal blocks hold unrelated rows, the seeded thresholds stay loose, the queue overflows -- the
    pipelined pass gives up and the ordinary pass (larger buffers, kernel (a), the dense path) takes over."""
    import taxondna_b200 as t
    sym, labels = labelled("H", 17000)
    perm = np.random.default_rng(9).permutation(sym.shape[0])
    sym = np.ascontiguousarray(sym[perm])
    labels = [np.ascontiguousarray(l[perm]) for l in labels]
    ref = separate(ctx, sym, labels, 0, 300)
    got = t.best_match_host(ctx, sym, labels, 0, 300)
    assert got.tobytes() == ref["rows"].tobytes()


def test_letters_that_arrive_late_switch_the_instantiation(ctx):
    import taxondna_b200 as t
    sym, labels = labelled("H", 20000)
    sym = sym.copy()
    rng = np.random.default_rng(3)
    late = np.arange(15000, 20000)
    m = rng.random((late.size, sym.shape[1])) < 0.03
    block = sym[late]
    block[m] = rng.choice(np.frombuffer(b"RYKMSWBDHVN", dtype=np.uint8), size=int(m.sum()))
    sym[late] = block
    clean, clean_labels = labelled("H", 17000)
    t.best_match_host(ctx, clean, clean_labels, 0, 300)  # the context now expects a list without ambiguity letters
    a = t.Alignment(ctx, sym)
    a.set_labels(*labels)
    a.set_config(0, 300, True)
    ctx.set_option(t.OPT_PIPELINE_CHUNKS, 0)
    ref = a.best_match()  # (built with the option off so that the expectation stays what the clean list left)
    a.close()
    ctx.set_option(t.OPT_PIPELINE_CHUNKS, 8)
    t.best_match_host(ctx, clean, clean_labels, 0, 300)
    before = ctx.pipelined_passes()
    got = t.best_match_host(ctx, sym, labels, 0, 300)
    assert got.tobytes() == ref.tobytes()
    assert ctx.pipelined_passes() == before  # the plain start was contradicted by the last chunk: the ordinary pass ran
    got = t.best_match_host(ctx, sym, labels, 0, 300)
    assert got.tobytes() == ref.tobytes()
    assert ctx.pipelined_passes() == before + 1  # ... and the next call starts on the GENERAL instantiation


def test_ragged_rows_and_the_returned_alignment(ctx):
    import taxondna_b200 as t
    sym, labels = labelled("H", 18000)
    rng = np.random.default_rng(4)
    lens = rng.integers(500, sym.shape[1] + 1, size=sym.shape[0]).astype(np.int32)
    ref = separate(ctx, sym, labels, 0, 300, lens)
    h, keep = t.host_list(sym, labels, 0, 300, True, lens)
    rows = np.zeros(h.n, dtype=t.ROW_RESULT_DTYPE)
    intra = np.empty(max(1, ref["intra"].size), dtype=np.float32)
    io = t.Analysis()
    io.want = t.WANT_BEST_MATCH | t.WANT_INTRA
    io.rows, io.intra, io.intra_cap = rows.ctypes.data, intra.ctypes.data, intra.size
    out = ctypes.c_void_p()
    t.check(ctx.L.tdna_analyze_host(ctx.h, ctypes.byref(h), ctypes.byref(io), ctypes.byref(out)))
    assert rows.tobytes() == ref["rows"].tobytes()
    assert io.n_intra == ref["intra"].size
    assert np.array_equal(np.sort(intra[: io.n_intra].view(np.uint32)), np.sort(ref["intra"].view(np.uint32)))
    # the alignment it hands back is resident like any other
    a = t.Alignment.__new__(t.Alignment)
    a.L, a.ctx, a.h, a.n, a.ncols, a.row_begin, a.row_end = ctx.L, ctx, out, h.n, h.L, 0, h.n
    assert a.best_match().tobytes() == ref["rows"].tobytes()
    d, c = a.pairs(np.array([0, 17999], np.int32), np.array([1, 5], np.int32), want_counts=True)
    a2 = t.Alignment(ctx, sym, lens)
    d2, c2 = a2.pairs(np.array([0, 17999], np.int32), np.array([1, 5], np.int32), want_counts=True)
    assert d.tobytes() == d2.tobytes() and c.tobytes() == c2.tobytes()
    a.close()
    a2.close()


def test_a_bad_symbol_in_the_last_chunk_is_reported(ctx):
    import taxondna_b200 as t
    sym, labels = labelled("H", 17000)
    sym = sym.copy()
    sym[16990, 100] = ord("!")
    with pytest.raises(t.TdnaError) as ei:
        t.best_match_host(ctx, sym, labels, 0, 300)
    assert ei.value.code == 3
    # and the context is still usable
    sym[16990, 100] = ord("A")
    got = t.best_match_host(ctx, sym, labels, 0, 300)
    assert (got["best"] >= 0).all()
