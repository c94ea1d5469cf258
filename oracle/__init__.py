"""CPU oracle for the TaxonDNA pairwise-distance path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
(``taxondna_b200``) never does; see DESIGN.md "Oracle".

``oracle.lib`` is a ctypes view of ``oracle/libtdna_oracle.so`` (built from
``tdna_oracle.c`` by ``oracle/Makefile``); ``oracle.consumers`` restates the
reference's consumers (SortedSequenceList, BestMatch, BlockAnalysis,
PairwiseDistribution, PairwiseSummary, Cluster) in pure Python on top of it.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libtdna_oracle.so")

PDM_UNCORRECTED, PDM_K2P, PDM_TRANS_ONLY = 0, 1, 2


def build(force=False):
    src = os.path.join(_HERE, "tdna_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def _load():
    build()
    L = ctypes.CDLL(_SO)
    c_int, c_dbl, c_p = ctypes.c_int, ctypes.c_double, ctypes.c_void_p
    cs = ctypes.c_char_p
    L.tdo_is_valid.argtypes = [c_int]
    L.tdo_is_ambiguous.argtypes = [c_int]
    L.tdo_getint.argtypes = [c_int]
    L.tdo_getcode.argtypes = [c_int]
    L.tdo_is_purine.argtypes = [c_int]
    L.tdo_is_pyrimidine.argtypes = [c_int]
    L.tdo_identical.argtypes = [c_int, c_int, c_int]
    L.tdo_normalise.argtypes = [cs, c_int, c_p, ctypes.POINTER(c_int)]
    L.tdo_shared_length.argtypes = [cs, c_int, cs, c_int, c_int]
    L.tdo_count_identical.argtypes = [cs, c_int, cs, c_int, c_int]
    L.tdo_count_transversions.argtypes = [cs, c_int, cs, c_int]
    L.tdo_k2p.argtypes = [cs, c_int, cs, c_int] + [ctypes.POINTER(c_int)] * 3
    L.tdo_k2p.restype = c_dbl
    L.tdo_pairwise.argtypes = [cs, c_int, cs, c_int, c_int, c_int, c_int]
    L.tdo_pairwise.restype = c_dbl
    L.tdo_log.argtypes = [c_dbl]
    L.tdo_log.restype = c_dbl
    L.tdo_set_use_libm_log.argtypes = [c_int]
    L.tdo_make_long.argtypes = [c_dbl]
    L.tdo_make_long.restype = ctypes.c_int64
    L.tdo_round_off.argtypes = [c_dbl]
    L.tdo_round_off.restype = c_dbl
    L.tdo_percentage.argtypes = [c_dbl, c_dbl]
    L.tdo_percentage.restype = c_dbl
    L.tdo_settings_identical.argtypes = [c_dbl, c_dbl]
    L.tdo_all_pairs.argtypes = [c_p, c_p] + [c_int] * 9 + [c_p] * 5
    L.tdo_all_pairs.restype = None
    L.tdo_bench_pairs.argtypes = [c_p, c_p] + [c_int] * 8 + [c_p]
    L.tdo_bench_pairs.restype = c_dbl
    return L


lib = _load()


def normalise(raw):
    """Sequence(name, raw) normalisation -> bytes, or raises ValueError(code)."""
    if isinstance(raw, str):
        raw = raw.encode("latin-1")
    out = ctypes.create_string_buffer(len(raw) + 1)
    n = ctypes.c_int(0)
    rc = lib.tdo_normalise(raw, len(raw), out, ctypes.byref(n))
    if rc != 0:
        raise ValueError(rc)
    return out.raw[: n.value]


def pack_rows(seqs):
    """list of normalised byte strings -> (uint8 [N, stride] '?'-padded, int32 lens)."""
    n = len(seqs)
    stride = max([len(s) for s in seqs] + [1])
    rows = np.full((n, stride), ord("?"), dtype=np.uint8)
    lens = np.zeros(n, dtype=np.int32)
    for i, s in enumerate(seqs):
        rows[i, : len(s)] = np.frombuffer(s, dtype=np.uint8)
        lens[i] = len(s)
    return rows, lens


def all_pairs(rows, lens, mode, min_overlap, ambig=True, row_begin=0, row_end=None, upper_only=False,
              threads=1, counts=True):
    """Distances (and counts) for rows [row_begin,row_end) x all columns.

    Returns dict(dist, shared, c1, c2, c3); see tdo_all_pairs."""
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n, stride = rows.shape
    if row_end is None:
        row_end = n
    m = row_end - row_begin
    dist = np.full((m, n), -1.0, dtype=np.float64)
    out = {"dist": dist}
    ptrs = [dist.ctypes.data]
    for k in ("shared", "c1", "c2", "c3"):
        if counts:
            out[k] = np.zeros((m, n), dtype=np.int32)
            ptrs.append(out[k].ctypes.data)
        else:
            ptrs.append(None)
    lib.tdo_all_pairs(rows.ctypes.data, lens.ctypes.data, n, stride, mode, min_overlap, int(ambig),
                      row_begin, row_end, int(upper_only), threads, *ptrs)
    return out


def bench_pairs(rows, lens, mode, min_overlap, ambig=True, row_begin=0, row_end=None, threads=1):
    rows = np.ascontiguousarray(rows, dtype=np.uint8)
    lens = np.ascontiguousarray(lens, dtype=np.int32)
    n, stride = rows.shape
    if row_end is None:
        row_end = n
    cnt = ctypes.c_int64(0)
    s = lib.tdo_bench_pairs(rows.ctypes.data, lens.ctypes.data, n, stride, mode, min_overlap, int(ambig),
                            row_begin, row_end, threads, ctypes.byref(cnt))
    return s, cnt.value
