"""Streaming Best Match (no N x N matrix): the candidate-list path must give byte-identical
tdna_row_result records to the dense two-sweep path and match the oracle-derived expectation, on
friendly input (list sorted by name) and on hostile input (shuffled list, singletons, unnamed rows,
one species only, gappy / ragged rows, K2P saturation), in one part and in several."""
import os
import sys

import numpy as np
import pytest

import oracle
from oracle import consumers as C

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

SYMS = b"ACGTRYKMSWBDHVN-_?"


@pytest.fixture(scope="module")
def ctx():
    import taxondna_b200 as t
    c = t.Context(0)
    yield c
    c.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    c.close()


def both_paths(ctx, aln):
    import taxondna_b200 as t
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    s = aln.best_match()
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
    d = aln.best_match()
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    return s, d


def first_difference(a, b):
    for q in range(len(a)):
        if a[q].tobytes() != b[q].tobytes():
            return q, {k: (a[k][q], b[k][q]) for k in a.dtype.names if a[k][q] != b[k][q] and not (a[k][q] != a[k][q] and b[k][q] != b[k][q])}
    return None


def make(rng, n_species, reps, L, unnamed=0, dup=0, gap=0.05, ragged=False, shuffle=False, singletons=0):
    from test_gpu_parity import labelled_alignment
    sym, sp, gen, epi, full = labelled_alignment(rng, n_species, reps, L, unnamed, dup, gappy=gap > 0)
    n = sym.shape[0]
    if singletons:  # species with exactly one member
        extra = sym[rng.integers(0, n, size=singletons)].copy()
        for x in extra:
            m = rng.random(L) < 0.03
            x[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
        sym = np.concatenate([sym, extra])
        sp = np.concatenate([sp, np.arange(1000, 1000 + singletons, dtype=np.int32)])
        gen = np.concatenate([gen, np.arange(500, 500 + singletons, dtype=np.int32)])
        n = sym.shape[0]
    lens = None
    if ragged:
        lens = rng.integers(L // 2, L + 1, size=n).astype(np.int32)
    if shuffle:
        perm = rng.permutation(n)
        sym, sp, gen = sym[perm], sp[perm], gen[perm]
        if lens is not None:
            lens = lens[perm]
    return sym, lens, sp.astype(np.int32), gen.astype(np.int32), sp.astype(np.int32).copy(), np.arange(n, dtype=np.int32)


CASES = [
    # mode, species, reps, L, unnamed, dup, gap, ragged, shuffle, singletons, min_overlap
    (0, 12, 6, 400, 0, 0, 0.0, False, False, 0, 200),
    (0, 30, 9, 658, 4, 12, 0.05, False, False, 3, 300),
    (0, 30, 9, 658, 4, 12, 0.05, False, True, 3, 300),
    (1, 30, 9, 658, 4, 12, 0.05, False, True, 3, 300),
    (2, 10, 5, 330, 2, 3, 0.05, False, True, 2, 150),
    (0, 20, 7, 700, 3, 5, 0.05, True, True, 4, 420),   # ragged: many pairs fall under minOverlap
    (1, 20, 7, 700, 3, 5, 0.05, True, False, 4, 420),
    (0, 1, 150, 300, 0, 0, 0.0, False, False, 0, 100),  # ONE species: thresholds never tighten
    (0, 2, 90, 300, 0, 0, 0.0, False, True, 0, 100),
    (0, 60, 5, 257, 0, 0, 0.0, False, False, 0, 1),     # 300 rows: three row tiles, partial last tile
]


@pytest.mark.parametrize("mode,nsp,reps,L,unnamed,dup,gap,ragged,shuffle,single,mo", CASES)
def test_stream_equals_dense_and_oracle(ctx, mode, nsp, reps, L, unnamed, dup, gap, ragged, shuffle, single, mo):
    import taxondna_b200 as t
    from ref_rows import compare, row_results
    rng = np.random.default_rng(1000 * mode + nsp * 7 + L)
    sym, lens, sp, gen, epi, full = make(rng, nsp, reps, L, unnamed, dup, gap, ragged, shuffle, single)
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    aln.set_config(mode, mo, True)
    s, d = both_paths(ctx, aln)
    assert first_difference(s, d) is None, first_difference(s, d)
    ref = oracle.all_pairs(sym, lens if lens is not None else np.full(n, L, np.int32), mode, mo, True, threads=8)
    compare(s, row_results(ref["dist"], sp), ref["shared"])
    aln.close()


def test_stream_parts_merge_to_the_single_part_answer(ctx):
    """Three parts (every third tile each), candidates concatenated, counts summed, flags ORed:
    the multi-GPU data flow on one device."""
    import torch
    import taxondna_b200 as t
    rng = np.random.default_rng(77)
    sym, lens, sp, gen, epi, full = make(rng, 40, 8, 500, unnamed=5, dup=10, gap=0.05, ragged=True, shuffle=True, singletons=5)
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    for mode in (0, 1):
        aln.set_config(mode, 260, True)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
        whole = aln.best_match()

        class Dev:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

        rows, cols, ds, inval, flags = [], [], [], None, None
        # diagonal phase of every part first: the per-row minima are MIN-reduced across the parts
        allmin = None
        for part in range(3):
            m = torch.as_tensor(Dev(aln.best_match_seed(part, 3), (3 * n,), "<i8"), device="cuda").clone()
            allmin = m if allmin is None else torch.minimum(allmin, m)
        for part in range(3):
            mp = aln.best_match_seed(part, 3)  # this part's own state again (one device plays all three ranks)
            torch.as_tensor(Dev(mp, (3 * n,), "<i8"), device="cuda").copy_(allmin)
            torch.cuda.synchronize()
            p = aln.best_match_part(part, 3, seeded=True)
            torch.cuda.synchronize()
            k = int(p.n_cand)
            if k:
                rows.append(torch.as_tensor(Dev(p.cand_row, (k,), "<i4"), device="cuda").clone())
                cols.append(torch.as_tensor(Dev(p.cand_col, (k,), "<i4"), device="cuda").clone())
                ds.append(torch.as_tensor(Dev(p.cand_d, (k,), "<f8"), device="cuda").clone())
            iv = torch.as_tensor(Dev(p.invalid_count, (n,), "<i4"), device="cuda").clone()
            fl = torch.as_tensor(Dev(p.flags, (n,), "<i4"), device="cuda").clone()
            inval = iv if inval is None else inval + iv
            flags = fl if flags is None else flags | fl
        # and the unseeded single-call form of a part must agree as well
        p0 = aln.best_match_part(0, 1)
        assert int(p0.n_cand) > 0
        r, c_, d_ = torch.cat(rows), torch.cat(cols), torch.cat(ds)
        merged = aln.best_match_merge(r.data_ptr(), c_.data_ptr(), d_.data_ptr(), r.numel(), inval.data_ptr(), flags.data_ptr())
        assert first_difference(merged, whole) is None, (mode, first_difference(merged, whole))
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    aln.close()


def test_k2p_saturation_flags_rows_like_the_dense_path(ctx, golden_dir):
    """The 326-sequence no-gaps fixture reaches p = 0.607: K2P hits NaN / +Inf there."""
    import taxondna_b200 as t
    seqs = C.read_fasta(os.path.join(golden_dir, "diptera_small_nogaps.fasta.gz"))
    rows, lens = oracle.pack_rows([s.seq for s in seqs])
    n = len(seqs)
    rng = np.random.default_rng(3)
    sp = rng.integers(0, 40, size=n).astype(np.int32)
    aln = t.Alignment(ctx, rows, lens)
    aln.set_labels(sp, (sp // 4).astype(np.int32), sp.copy(), np.arange(n, dtype=np.int32))
    aln.set_config(t.PDM_K2P, 300, True)
    s, d = both_paths(ctx, aln)
    assert np.array_equal(s["flags"], d["flags"]) and np.array_equal(s["n_valid"], d["n_valid"])
    ok = (d["flags"] & t.ROW_NONFINITE) == 0  # flagged rows are re-resolved by the host from tdna_row_distances
    assert first_difference(s[ok], d[ok]) is None, first_difference(s[ok], d[ok])
    aln.close()
    # unrelated random sequences next to related ones: P + Q ~ 0.75, so w1 = 1 - 2P - Q straddles zero -> NaN, +Inf, huge d
    rel = make(rng, 10, 6, 400, gap=0.0)[0]
    rnd = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=(140, 400))
    sym = np.concatenate([rel, rnd])
    n = sym.shape[0]
    sp = rng.integers(0, 25, size=n).astype(np.int32)
    aln = t.Alignment(ctx, sym)
    aln.set_labels(sp, (sp // 4).astype(np.int32), sp.copy(), np.arange(n, dtype=np.int32))
    aln.set_config(t.PDM_K2P, 100, True)
    s, d = both_paths(ctx, aln)
    assert (d["flags"] & t.ROW_NONFINITE).any()
    assert np.array_equal(s["flags"], d["flags"]) and np.array_equal(s["n_valid"], d["n_valid"])
    ok = (d["flags"] & t.ROW_NONFINITE) == 0
    assert first_difference(s[ok], d[ok]) is None, first_difference(s[ok], d[ok])
    aln.close()


def test_real_fixture_gappy_rows(ctx, golden_dir):
    """Diptera COI (1390 x 1539, 55 % gaps): 42 % of the pairs are below minOverlap."""
    import taxondna_b200 as t
    seqs = C.read_fasta(os.path.join(golden_dir, "diptera_coi.fasta.gz"))
    rows, lens = oracle.pack_rows([s.seq for s in seqs])
    lab = C.labels(seqs) if hasattr(C, "labels") else None
    n = len(seqs)
    if lab is None:
        names = [s.species_name() if hasattr(s, "species_name") else None for s in seqs]
        ids, sp = {}, []
        for nm in names:
            sp.append(-1 if nm is None else ids.setdefault(nm, len(ids)))
        sp = np.array(sp, dtype=np.int32)
        lab = (sp, (sp // 3).astype(np.int32), sp.copy(), np.arange(n, dtype=np.int32))
    aln = t.Alignment(ctx, rows, lens)
    aln.set_labels(*lab)
    for mode in (0, 1):
        aln.set_config(mode, 300, True)
        s, d = both_paths(ctx, aln)
        ok = (d["flags"] & t.ROW_NONFINITE) == 0
        assert np.array_equal(s["flags"], d["flags"]) and np.array_equal(s["n_valid"], d["n_valid"])
        assert first_difference(s[ok], d[ok]) is None, (mode, first_difference(s[ok], d[ok]))
    aln.close()


def test_overflow_falls_back_to_dense(ctx):
    """A tiny candidate buffer overflows: AUTO answers through the dense path, STREAM reports TDNA_E_TOO_LARGE."""
    import taxondna_b200 as t
    rng = np.random.default_rng(5)
    sym, lens, sp, gen, epi, full = make(rng, 1, 400, 200, gap=0.0)  # one species: every pair is a candidate
    ctx.set_memory_budget(1 << 20)  # candidate capacity = budget / 32 records
    try:
        aln = t.Alignment(ctx, sym, lens)
        aln.set_labels(sp, gen, epi, full)
        aln.set_config(0, 50, True)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
        with pytest.raises(t.TdnaError) as ei:
            aln.best_match()
        assert ei.value.code == 5
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
        auto = aln.best_match()
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
        dense = aln.best_match()
        assert first_difference(auto, dense) is None
        aln.close()
    finally:
        ctx.set_memory_budget(64 << 30)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)


@pytest.mark.parametrize("mode,shuffle,ragged", [(0, False, False), (0, True, True), (1, True, False), (2, False, True)])
def test_one_pass_analysis_equals_the_dense_functions(ctx, mode, shuffle, ragged):
    """tdna_analyze: Best Match + both PairwiseDistribution classes + Cluster edges from ONE streaming pass."""
    import taxondna_b200 as t
    rng = np.random.default_rng(40 + mode)
    sym, lens, sp, gen, epi, full = make(rng, 24, 7, 520, unnamed=4, dup=8, gap=0.05, ragged=ragged, shuffle=shuffle, singletons=3)
    n = sym.shape[0]
    full = rng.permutation(n).astype(np.int32)
    full[5] = full[9]  # equal full names: the pair is pushed twice if conspecific (PairwiseDistribution.java:172)
    sp[9] = sp[5]
    gen[9] = gen[5]
    epi = sp.copy()
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    aln.set_config(mode, 260, True)
    thr = 0.03
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
    d_rows = aln.best_match()
    d_intra, d_inter = np.sort(aln.class_distances(t.PD_INTRA)), np.sort(aln.class_distances(t.PD_INTER))
    d_edges = sorted(zip(*[x.tolist() for x in aln.edges_below(thr)]))
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
    out = aln.analyze(best_match=True, intra=True, inter=True, edges=thr)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    assert len(d_intra) > 0 and len(d_inter) > 0 and len(d_edges) > 0
    assert first_difference(out["rows"], d_rows) is None, first_difference(out["rows"], d_rows)
    assert np.array_equal(np.sort(out["intra"]).view(np.uint32), d_intra.view(np.uint32))
    assert np.array_equal(np.sort(out["inter"]).view(np.uint32), d_inter.view(np.uint32))
    assert sorted(zip(out["edge_i"].tolist(), out["edge_j"].tolist(), out["edge_d"].tolist())) == d_edges
    # the single-purpose entry points ride the same pass
    assert np.array_equal(np.sort(aln.class_distances(t.PD_INTRA)).view(np.uint32), d_intra.view(np.uint32))
    assert sorted(zip(*[x.tolist() for x in aln.edges_below(thr)])) == d_edges
    aln.close()


def test_packed_parts_exchange_on_one_device(ctx):
    """The glue-free multi-GPU data flow (tdna_best_match_part_packed -> NCCL -> tdna_best_match_merge_gathered), three ranks
    played by one device: seeds MIN-reduced, packed segments placed side by side, states summed."""
    import torch
    import taxondna_b200 as t
    rng = np.random.default_rng(123)
    sym, lens, sp, gen, epi, full = make(rng, 36, 8, 500, unnamed=4, dup=8, gap=0.05, ragged=True, shuffle=False, singletons=4)
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)

    class Dev:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

    for mode, kern in ((0, t.KERNEL_POPC), (1, t.KERNEL_POPC), (0, t.KERNEL_AUTO)):
        aln.set_config(mode, 260, True)
        ctx.set_option(t.OPT_COUNT_KERNEL, kern)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
        whole = aln.best_match()
        world, seg = 3, 32 * n
        allmin = None
        for part in range(world):
            m = torch.as_tensor(Dev(aln.best_match_seed(part, world), (3 * n,), "<i8"), device="cuda").clone()
            allmin = m if allmin is None else torch.minimum(allmin, m)
        allrec = torch.zeros((world, seg, 2), dtype=torch.int64, device="cuda")
        counts = torch.zeros(world, dtype=torch.int64, device="cuda")
        state = torch.zeros(n, dtype=torch.int64, device="cuda")
        for part in range(world):
            mp = aln.best_match_seed(part, world)
            torch.as_tensor(Dev(mp, (3 * n,), "<i8"), device="cuda").copy_(allmin)
            st = torch.zeros(n, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            aln.best_match_part_packed(part, world, True, allrec[part].data_ptr(), seg, counts[part:].data_ptr(), st.data_ptr())
            state += st
        torch.cuda.synchronize()
        out = np.zeros(n, dtype=t.ROW_RESULT_DTYPE)
        aln.best_match_merge_gathered(allrec.data_ptr(), world, seg, counts.data_ptr(), state.data_ptr(), out.ctypes.data)
        assert first_difference(out, whole) is None, (mode, kern, first_difference(out, whole))
        with pytest.raises(t.TdnaError) as ei:  # a segment that is too small is reported, not overrun
            aln.best_match_part_packed(0, 1, False, allrec.data_ptr(), 8, counts.data_ptr(), state.data_ptr())
        assert ei.value.code == 5
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    aln.close()


def test_auto_falls_back_from_kernel_b_to_kernel_a_on_a_shuffled_list(ctx):
    """A list in random order: the diagonal blocks hold unrelated rows, so kernel (b)'s static thresholds are loose and its queue
    overflows (small budget here); AUTO retries on kernel (a), whose thresholds tighten while it runs, and the records still
    equal the dense path's."""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    data = synth.generate("C3", names=False, genera=40, species=10, reps=10)
    rng = np.random.default_rng(8)
    perm = rng.permutation(data["symbols"].shape[0])
    sym = data["symbols"][perm]
    lab = [data[k][perm] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
    ctx.set_memory_budget(24 << 20)  # queue capacity = budget / 24 records (~ 1 M), candidates budget / 32
    try:
        aln = t.Alignment(ctx, sym)
        aln.set_labels(*lab)
        aln.set_config(0, 300, True)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
        ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
        auto = aln.best_match()
        used = aln.last_count_kernel()
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
        ctx.set_memory_budget(64 << 30)
        dense = aln.best_match()
        assert first_difference(auto, dense) is None, first_difference(auto, dense)
        print("shuffled list, AUTO ended on count kernel", used)
        aln.close()
    finally:
        ctx.set_memory_budget(64 << 30)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)


@pytest.mark.parametrize("world", [2, 3, 5])
def test_routed_parts_exchange_on_one_device(ctx, world):
    """The row-sharded multi-GPU data flow (tdna_best_match_part_routed -> all-to-all -> tdna_best_match_merge_routed -> all-gather
    of the row slices), the ranks played by one device: every part finalises only the rows it owns; the slices side by side
    are the one-GPU result byte for byte."""
    import torch
    import taxondna_b200 as t
    rng = np.random.default_rng(321 + world)
    sym, lens, sp, gen, epi, full = make(rng, 40, 9, 500, unnamed=4, dup=8, gap=0.05, ragged=True, shuffle=False, singletons=4)
    n = sym.shape[0]
    aln = t.Alignment(ctx, sym, lens)
    aln.set_labels(sp, gen, epi, full)
    rpp = aln.rows_per_part(world)
    assert rpp % 128 == 0 and rpp * world >= n and rpp == t.load().tdna_rows_per_part(n, world)

    class Dev:
        def __init__(self, ptr, shape, typestr):
            self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}

    for mode, kern in ((0, t.KERNEL_POPC), (1, t.KERNEL_AUTO), (0, t.KERNEL_AUTO)):
        aln.set_config(mode, 260, True)
        ctx.set_option(t.OPT_COUNT_KERNEL, kern)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_STREAM)
        whole = aln.best_match()
        seg = 32 * n
        allmin = None
        for part in range(world):
            m = torch.as_tensor(Dev(aln.best_match_seed(part, world), (3 * n,), "<i8"), device="cuda").clone()
            allmin = m if allmin is None else torch.minimum(allmin, m)
        sent = torch.zeros((world, world, seg, 2), dtype=torch.int64, device="cuda")     # [source][destination]
        sent_counts = torch.zeros((world, world), dtype=torch.int64, device="cuda")
        state = torch.zeros(n, dtype=torch.int64, device="cuda")
        for part in range(world):
            mp = aln.best_match_seed(part, world)
            torch.as_tensor(Dev(mp, (3 * n,), "<i8"), device="cuda").copy_(allmin)
            st = torch.zeros(n, dtype=torch.int64, device="cuda")
            torch.cuda.synchronize()
            aln.best_match_part_routed(part, world, True, sent[part].data_ptr(), seg, sent_counts[part].data_ptr(), st.data_ptr())
            state += st
        recv = sent.transpose(0, 1).contiguous()            # the all-to-all: [destination][source]
        recv_counts = sent_counts.t().contiguous()
        torch.cuda.synchronize()
        out = np.zeros(world * rpp, dtype=t.ROW_RESULT_DTYPE)
        for part in range(world):
            sl = out[part * rpp:(part + 1) * rpp]
            aln.best_match_merge_routed(part, world, recv[part].data_ptr(), seg, recv_counts[part].data_ptr(), state.data_ptr(), sl.ctypes.data)
        assert first_difference(out[:n], whole) is None, (mode, kern, first_difference(out[:n], whole))
        assert not out[n:].tobytes().strip(b"\0")  # the padding past the end of the list is zero
        with pytest.raises(t.TdnaError) as ei:  # a segment that is too small is reported, not overrun
            aln.best_match_part_routed(0, world, False, sent.data_ptr(), 8, sent_counts.data_ptr(), state.data_ptr())
        assert ei.value.code == 5
    ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
    ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
    aln.close()



@pytest.mark.parametrize("n,L", [(1, 1), (1, 40), (2, 1), (2, 33), (3, 129), (129, 5)])
def test_tiny_alignments_through_every_path(ctx, n, L):
    """The smallest lists (one sequence, one column, one row past a tile) through the streaming pass on both count kernels, the
    dense path, the one-pass analysis and the foreign-query entry point: same bytes everywhere, nothing out of range."""
    import taxondna_b200 as t
    rng = np.random.default_rng(n * 100 + L)
    sym = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=(n, L))].copy()
    sp = (np.arange(n) // 2).astype(np.int32)
    aln = t.Alignment(ctx, sym)
    aln.set_labels(sp, (sp // 2).astype(np.int32), sp, np.arange(n, dtype=np.int32))
    for mode in (0, 1):
        aln.set_config(mode, 1, True)
        got = {}
        for name, path, kern in (("dense", t.PATH_DENSE, t.KERNEL_AUTO), ("popc", t.PATH_STREAM, t.KERNEL_POPC), ("tensor", t.PATH_STREAM, t.KERNEL_TENSOR)):
            ctx.set_option(t.OPT_BEST_MATCH_PATH, path)
            ctx.set_option(t.OPT_COUNT_KERNEL, kern)
            got[name] = aln.analyze(best_match=True, intra=True, inter=True, edges=0.5)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
        ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
        for name in ("popc", "tensor"):
            assert first_difference(got[name]["rows"], got["dense"]["rows"]) is None, (mode, name)
            for k in ("intra", "inter"):
                assert np.array_equal(np.sort(got[name][k]).view(np.uint32), np.sort(got["dense"][k]).view(np.uint32)), (mode, name, k)
            assert sorted(zip(got[name]["edge_i"].tolist(), got[name]["edge_j"].tolist())) == sorted(zip(got["dense"]["edge_i"].tolist(), got["dense"]["edge_j"].tolist()))
        ref = oracle.all_pairs(sym, np.full(n, L, np.int32), mode, 1, True)
        d, sh = aln.query_distances(sym[0])
        assert np.array_equal(d.view(np.uint64), ref["dist"][0].view(np.uint64)) and np.array_equal(sh, ref["shared"][0])
    aln.close()
