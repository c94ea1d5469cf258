"""Count kernel (b)'s integer code, checked on the CPU: the algebra the int8 one-hot contraction relies on, and a
simulated accumulator decoded the way the epilogue decodes it, against direct counts."""
import ctypes

import numpy as np
import pytest


def code_for(L, dims):
    import taxondna_b200 as t
    lib = t.load()
    code = (ctypes.c_int32 * 4)()
    W = ctypes.c_uint32()
    rc = lib.tdna_tensor_code(L, dims, code, ctypes.byref(W))
    return rc, list(code), int(W.value)


@pytest.mark.parametrize("dims", [4, 5])
@pytest.mark.parametrize("L", [1, 200, 255, 256, 658, 1023, 1024, 2664, 4095])
def test_code_algebra_and_decoding(L, dims):
    rc, (a, b, c, d), W = code_for(L, dims)
    if rc != 0:
        # four dimensions only have a code at W = 1024 (the library then uses five dimensions with the gap dimension left empty)
        assert dims == 4 and L >= 1024
        return
    assert W > L and W & (W - 1) == 0
    assert a * c == -(W - 1)
    assert a * d + b * c + dims * b * d == W
    for v in (b, a + b, d, c + d):
        assert -128 <= v <= 127  # every operand byte is an int8
    # one site: row symbol x, column symbol y (dims = "contributes nothing": a zero vector)
    for x in range(dims + 1):
        for y in range(dims + 1):
            u = np.zeros(dims, np.int64) if x == dims else np.full(dims, b, np.int64) + a * np.eye(dims, dtype=np.int64)[x]
            w = np.zeros(dims, np.int64) if y == dims else np.full(dims, d, np.int64) + c * np.eye(dims, dtype=np.int64)[y]
            dot = int(u @ w)
            expect = 0 if (x == dims or y == dims) else (1 if x == y else W)
            assert dot == expect, (x, y, dot)
    # a whole pair of rows: the int32 accumulator decodes into (ident, mism) by mask and shift
    rng = np.random.default_rng(L * 7 + dims)
    xs = rng.integers(0, dims + 1, size=L)
    ys = np.where(rng.random(L) < 0.7, xs, rng.integers(0, dims + 1, size=L))
    both = (xs < dims) & (ys < dims)
    ident = int((both & (xs == ys)).sum())
    mism = int((both & (xs != ys)).sum())
    acc = ident * 1 + mism * W
    assert acc < 2 ** 31
    shift = W.bit_length() - 1
    assert acc & (W - 1) == ident and acc >> shift == mism


def test_no_int8_code_beyond_4095_columns():
    rc, _, _ = code_for(4096, 5)
    assert rc != 0


SYMS = "ACGTRYKMSWBDHVN-_?"


def tables_for(mode, L, five_dims=False):
    import taxondna_b200 as t
    lib = t.load()
    ta = (ctypes.c_int8 * 48)()
    tb = (ctypes.c_int8 * 48)()
    dims, W = ctypes.c_int32(), ctypes.c_uint32()
    rc = lib.tdna_tensor_tables(mode, 1, L, 1 if five_dims else 0, ta, tb, ctypes.byref(dims), ctypes.byref(W))
    assert rc == 0
    return np.array(list(ta), np.int64).reshape(6, 8), np.array(list(tb), np.int64).reshape(6, 8), dims.value, int(W.value)


def symbol_code(ch, mode, five_dims):
    """tdna_tc.cuh: symbol_code, restated."""
    if mode == 2:
        return 0 if ch in "AGR" else 1 if ch in "CTY" else 2 if ch == "-" else 3 if ch in "KMSWBDHVN" else 5
    if ch in "ACGT":
        return "ACGT".index(ch)
    return 4 if (ch == "-" and mode == 0 and five_dims) else 5


@pytest.mark.parametrize("mode,five_dims,L", [(2, False, 200), (2, False, 658), (2, False, 2664), (0, False, 658), (0, True, 658), (0, True, 2664), (1, False, 658), (1, False, 2664)])
def test_operand_tables_against_the_oracles_per_site_counts(mode, five_dims, L):
    """One site, every pair of the 18 symbols: row vector . column vector must be what the accumulator is decoded as --
    ident + W * mism for the pairs the one-hot codes express, (shared - transversions) + W * transversions for EVERY pair in
    transversions-only mode -- with the counts taken from the oracle's restatement of Sequence.java's loops."""
    import oracle
    ta, tb, dims, W = tables_for(mode, L, five_dims)
    assert W > L and W & (W - 1) == 0 and dims <= 6
    assert (np.abs(ta) <= 127).all() and (np.abs(tb) <= 127).all()
    rows = np.frombuffer(SYMS.encode(), dtype=np.uint8).reshape(len(SYMS), 1)
    ref = oracle.all_pairs(rows, np.ones(len(SYMS), np.int32), mode, 0, True)
    for i, x in enumerate(SYMS):
        for j, y in enumerate(SYMS):
            cx, cy = symbol_code(x, mode, five_dims), symbol_code(y, mode, five_dims)
            dot = int(ta[cx][:dims] @ tb[cy][:dims]) if cx < 6 and cy < 6 else 0
            shared, c1 = int(ref["shared"][i, j]), int(ref["c1"][i, j])
            if mode == 2:  # exact for every letter
                assert dot == (shared - c1) + W * c1, (x, y, dot, shared, c1)
                continue
            expressed = (x in "ACGT" or (x == "-" and mode == 0 and five_dims)) and (y in "ACGT" or (y == "-" and mode == 0 and five_dims))
            if not expressed:
                assert dot == 0, (x, y)  # contributes nothing: such pairs are only bounded
            elif mode == 0:
                assert dot == c1 + W * (shared - c1), (x, y, dot, shared, c1)   # ident + W * mismatches
            else:  # K2P: c1 = n (both in AGRCTY); same = identical bases
                same = 1 if x == y else 0
                assert c1 == 1 and dot == same + W * (c1 - same), (x, y, dot)
