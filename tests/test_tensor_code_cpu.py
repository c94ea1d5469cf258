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
