"""FASTA on the device (tdna_fasta_parse): the record table and the normalised symbol matrix against a line-by-line restatement of
FastaFile.appendFromFile (Common/DNA/formats/FastaFile.java:144-262) + the oracle's Sequence.changeSequence normaliser."""
import gzip
import os
import re
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
    c.close()


def java_lines(text):
    """BufferedReader.readLine: lines end at \\n, \\r or \\r\\n; a trailing terminator does not start another line."""
    lines = re.split(r"\r\n|\n|\r", text)
    if lines and lines[-1] == "":
        lines.pop()
    return lines


def reference_records(text):
    """[(header text after '>', raw sequence with whitespace removed)] as FastaFile.appendFromFile collects them."""
    out, name, seq = [], None, ""
    for line in java_lines(text):
        if re.match(r"^[ \t\n\x0b\f\r]*$", line) or re.match(r"^[ \t\n\x0b\f\r]*#.*$", line):
            continue
        if line.startswith(">"):
            if name is not None:
                out.append((name, seq))
            name, seq = line[1:], ""
            continue
        seq += re.sub(r"[ \t\n\x0b\f\r]+", "", line)
    if name is not None:
        out.append((name, seq))
    return out


def expected(records):
    """per record: (normalised bytes or None, flags)"""
    import taxondna_b200 as t
    exp = []
    for _, raw in records:
        fl = 0
        if "U" in raw or "u" in raw:
            fl |= t.FASTA_URACIL
            raw = raw.replace("U", "T").replace("u", "t")
        if re.search(r"[\[\]()]", raw):
            fl |= t.FASTA_NEEDS_HOST
            if re.search(r"[^A-Za-z\-?\[\]()]", raw) or "_" in raw:
                fl |= t.FASTA_INVALID
            exp.append((None, fl))
            continue
        try:
            exp.append((oracle.normalise(raw), fl))
        except ValueError:
            exp.append((None, fl | t.FASTA_INVALID))
    return exp


def check_text(ctx, text, matrix=True):
    import taxondna_b200 as t
    f = t.Fasta(ctx, text)
    recs = reference_records(text)
    exp = expected(recs)
    assert f.n == len(recs)
    assert f.headers() == [h for h, _ in recs]
    off, hl, sl, fl = f.records()
    assert sl.tolist() == [len(raw) for _, raw in recs]
    assert f.width == max([len(raw) for _, raw in recs] + [0])
    sym = f.symbols()
    for r, (norm, flags) in enumerate(exp):
        assert fl[r] & (t.FASTA_INVALID | t.FASTA_URACIL) == flags & (t.FASTA_INVALID | t.FASTA_URACIL) or (flags & t.FASTA_NEEDS_HOST), (r, fl[r], flags)
        assert bool(fl[r] & t.FASTA_NEEDS_HOST) == bool(flags & t.FASTA_NEEDS_HOST), r
        if norm is not None:
            assert sym[r, :len(norm)].tobytes() == norm, (r, sym[r, :len(norm)].tobytes()[:80], norm[:80])
            assert (sym[r, len(norm):] == ord("?")).all(), r
    if matrix and f.n >= 2 and all(norm is not None for norm, _ in exp):
        rows, lens = oracle.pack_rows([norm for norm, _ in exp])
        a = f.alignment()
        b = t.Alignment(ctx, rows, lens)
        for mode in (0, 1):
            a.set_config(mode, 50, True)
            b.set_config(mode, 50, True)
            assert a.matrix().tobytes() == b.matrix().tobytes()
        a.close()
        b.close()
    elif f.n and not all(norm is not None for norm, _ in exp):
        with pytest.raises(t.TdnaError):
            f.alignment()
    f.close()
    return recs


def test_reference_fixture(ctx, golden_dir):
    text = gzip.open(os.path.join(golden_dir, "diptera_coi.fasta.gz"), "rt", encoding="latin-1").read()
    recs = check_text(ctx, text)
    assert len(recs) > 100
    # and the same records through the oracle's own reader (names with '_' -> ' ' are the host layer's business)
    from oracle import consumers as C
    assert [s.seq for s in C.read_fasta_text(text)] == [oracle.normalise(raw.replace("U", "T").replace("u", "t")) for _, raw in recs]


def test_line_structure_corner_cases(ctx):
    text = ("ACGT stray sequence before any name\n"
            "# a comment\n"
            ">first one\r\n"
            "acgtnnry--ACGT\r\n"
            "   \r\n"
            "  ac gt\tac\r\n"
            "   # indented comment\n"
            ">second_with_gaps and ? (gi|12|)\n"
            "--?-ACGT-A?-?--\n"
            ">\n"
            "ACGU acgu\n"
            ">all gaps\n"
            "--??--\n"
            ">empty sequence\n"
            ">old mac line ends\rAC\rGT\r"
            ">   \n"
            "TTTT\n"
            " >not a name: an indented line is a sequence line\n"
            ">last one without newline\nACGTRYKMSWBDHVN-?acgt")
    check_text(ctx, text, matrix=False)


def test_flagged_records(ctx):
    import taxondna_b200 as t
    # an interior '_' passes the reference's checks (the strokes never reach it, isValid accepts it); one that a stroke meets does not
    text = (">ok\nACGTACGT\n>group\nAC[AG]T(CT)A\n>bad\nACGTXACGT\n>interior underscore\nAC_GT\n>digit\nAC1GT\n>ok2\nACGTAC-T\n"
            ">leading underscore\n-?_ACGT\n>trailing underscore\nACGT_?-\n")
    check_text(ctx, text, matrix=False)
    f = t.Fasta(ctx, text)
    _, _, _, fl = f.records()
    assert fl.tolist() == [0, t.FASTA_NEEDS_HOST, t.FASTA_INVALID, 0, t.FASTA_INVALID, 0, t.FASTA_INVALID, t.FASTA_INVALID]
    f.close()


def test_empty_and_nameless_texts(ctx):
    import taxondna_b200 as t
    for text in ("", "\n\n", "ACGT\nACGT\n", "# only a comment"):
        f = t.Fasta(ctx, text)
        assert f.n == 0 and f.width == 0
        with pytest.raises(t.TdnaError):
            f.alignment()
        f.close()


def test_large_synthetic_file(ctx):
    """8,000 x 658 as 60-column FASTA (5.6 MB): every record, then the distance matrix of a slice."""
    import taxondna_b200 as t
    from taxondna_b200 import synth
    d = synth.generate("C3")
    sym = d["symbols"].copy()
    sym[sym == ord("_")] = ord("-")
    parts = []
    for nm, row in zip(d["names"], sym):
        s = row.tobytes().decode()
        parts.append(">" + nm + "\n" + "\n".join(s[i:i + 60] for i in range(0, len(s), 60)) + "\n")
    text = "".join(parts)
    f = t.Fasta(ctx, text)
    assert f.n == sym.shape[0] and f.width == sym.shape[1]
    assert f.headers() == list(d["names"])
    assert np.array_equal(f.symbols(), d["symbols"])
    a = f.alignment()
    b = t.Alignment(ctx, d["symbols"])
    a.set_config(1, 300, True)
    b.set_config(1, 300, True)
    assert np.array_equal(a.row_distances(17)[0].view(np.uint64), b.row_distances(17)[0].view(np.uint64))
    a.close()
    b.close()
    f.close()


def test_consensus_of_bins_equals_the_reference_fold(ctx):
    """tdna_consensus (a column-wise closed form) against the reference's own procedure: the left fold of Sequence.getConsensus with
    a re-normalisation after every step (Cluster.java:123-138), on rows with gaps, external gaps, '?', ambiguity letters and
    unequal lengths; bins of one to forty members."""
    import taxondna_b200 as t
    from oracle import consumers as C
    rng = np.random.default_rng(77)
    n, L = 160, 240
    alphabet = np.frombuffer(b"ACGTACGTACGTACGTRYKMSWBDHVN---??", dtype=np.uint8)
    base = alphabet[rng.integers(0, 16, size=L)]
    raws = []
    for i in range(n):
        row = base.copy()
        m = rng.random(L) < 0.15
        row[m] = alphabet[rng.integers(0, len(alphabet), size=int(m.sum()))]
        k0, k1 = int(rng.integers(0, 30)), int(rng.integers(0, 30))
        if i % 3 == 0:
            row[:k0] = ord("-")
        if i % 4 == 0:
            row[L - k1:] = ord("-")
        if i % 5 == 0:
            row[:k0 // 2] = ord("?")
        ln = L if i % 7 else int(rng.integers(L // 2, L))
        raws.append(row[:ln].tobytes().decode())
    seqs = [oracle.normalise(r) for r in raws]
    rows, lens = oracle.pack_rows(seqs)
    aln = t.Alignment(ctx, rows, lens)
    bins = [[int(i)] for i in (0, 5, 21)]
    for size in (2, 3, 7, 40):
        for _ in range(6):
            bins.append([int(x) for x in rng.choice(n, size=size, replace=False)])
    bins.append(list(range(n)))
    out, out_len = aln.consensus(bins)
    for k, b in enumerate(bins):
        ref = C.consensus_of_bin([seqs[i] for i in b])
        assert out_len[k] == len(ref), (k, b)
        assert out[k, :len(ref)].tobytes() == ref, (k, b, out[k, :len(ref)].tobytes()[:60], ref[:60])
        assert (out[k, len(ref):] == ord("?")).all()
    with pytest.raises(t.TdnaError):
        aln.consensus([[0, n]])
    aln.close()
