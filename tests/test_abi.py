"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every
symbol include/taxondna_cuda.h declares, lays its structs out as the binding expects, and refuses
to compute without a device (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import taxondna_b200 as t
from taxondna_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "taxondna_cuda.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tdna_[a-z0-9_]+)\s*\(", src)) - {"tdna_progress_fn"})


def test_library_exports_every_declared_symbol():
    L = t.load()
    names = header_functions()
    assert len(names) >= 20
    for nm in names:
        assert hasattr(L, nm), "libtaxondna_cuda.so does not export %s" % nm
    assert sorted(_lib.EXPORTS) == names
    assert L.tdna_abi_version() == 2


def test_struct_layouts_match_the_header():
    assert ctypes.sizeof(_lib.Counts) == 16 and _lib.COUNTS_DTYPE.itemsize == 16
    d = _lib.ROW_RESULT_DTYPE
    assert d.itemsize == 120
    assert [d.fields[k][1] for k in ("d_best", "d_con", "d_allo", "d_other", "best", "shared_con", "tie_count", "block_count", "near_best", "n_valid", "flags")] == \
        [0, 8, 16, 24, 32, 48, 56, 72, 76, 104, 108]


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    with pytest.raises(t.TdnaError) as ei:
        t.Context(0)
    assert ei.value.code == 1 and "no CPU fallback" in str(ei.value)


def test_host_layer_loads_and_refuses_to_compute_without_a_gpu():
    import torch
    from taxondna_b200 import _host as H
    a, b = H.Sequence("1", "ACGT"), H.Sequence("2", "ACGA")
    assert a.getLength() == 4 and a.getSequenceRaw() == "ACGT"
    if not torch.cuda.is_available():
        with pytest.raises(H.EngineError):
            a.getPairwise(b)


def test_rows_per_part_is_pure_host_arithmetic():
    """Owner-of-row arithmetic of the routed multi-GPU merge: slices are multiples of 128 rows and cover the list."""
    import taxondna_b200 as t
    lib = t.load()
    for n in (1, 127, 128, 129, 2185, 50000, 141500, 1000000):
        for parts in (1, 2, 3, 4, 8, 64):
            rpp = lib.tdna_rows_per_part(n, parts)
            assert rpp % 128 == 0 and rpp * parts >= n and (rpp - 128) * parts < n + 128 * parts
    assert lib.tdna_rows_per_part(0, 4) == 0
