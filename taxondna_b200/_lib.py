"""ctypes binding of the C ABI in include/taxondna_cuda.h (the same entry points the Java host
layer binds through Panama FFM -- see INTEGRATION.md).  No CPU fallback: if the CUDA library is
missing or no GPU is visible, every compute path raises."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TDNA_LIB") or os.path.join(_HERE, "libtaxondna_cuda.so")  # TDNA_LIB: A/B of library builds

PDM_UNCORRECTED, PDM_K2P, PDM_TRANS_ONLY = 0, 1, 2
PD_INTRA, PD_INTER = 0, 1
ROW_SELF_VALID, ROW_NONFINITE = 1, 2
OPT_BEST_MATCH_PATH, OPT_TILE_COLUMNS, OPT_COUNT_KERNEL, OPT_MULTI, OPT_HIST_SLOTS, OPT_SEGMENT_ITEMS, OPT_PIPELINE_CHUNKS, OPT_CLUSTER, OPT_UNIT_ITEMS = 1, 2, 3, 4, 5, 6, 7, 8, 9
COMM_ID_BYTES = 128
KERNEL_AUTO, KERNEL_POPC, KERNEL_TENSOR = 0, 1, 2
PATH_AUTO, PATH_DENSE, PATH_STREAM = 0, 1, 2


class TdnaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("taxondna_cuda error %d: %s" % (code, msg))
        self.code = code


E_CUDA, E_ARG, E_SYMBOL, E_STATE, E_TOO_LARGE, E_CANCELLED = 1, 2, 3, 4, 5, 6


class Counts(ctypes.Structure):
    _fields_ = [("shared", ctypes.c_int32), ("c1", ctypes.c_int32), ("c2", ctypes.c_int32), ("c3", ctypes.c_int32)]


WANT_BEST_MATCH, WANT_INTRA, WANT_INTER, WANT_EDGES, WANT_HIST_INTRA, WANT_HIST_INTER = 1, 2, 4, 8, 16, 32
HIST_ENTRY_DTYPE = np.dtype([("d", "<f8"), ("count", "<i8")])
EXTREME_RESULT_DTYPE = np.dtype([("d_max_intra", "<f8"), ("d_min_inter", "<f8"), ("sum_inter", "<f8"), ("max_intra", "<i4"), ("min_inter", "<i4"),
                                 ("shared_intra", "<i4"), ("shared_inter", "<i4"), ("n_intra", "<i4"), ("n_inter", "<i4"), ("near_intra", "<i4"),
                                 ("near_inter", "<i4"), ("near_order", "<i4"), ("flags", "<i4")])
EXT_UNNAMED, EXT_TOO_MANY = 16, 32


class Analysis(ctypes.Structure):
    """tdna_analysis: request / result of one streaming pass (tdna_analyze)."""
    _fields_ = [("want", ctypes.c_int32), ("reserved", ctypes.c_int32), ("rows", ctypes.c_void_p),
                ("intra", ctypes.c_void_p), ("intra_cap", ctypes.c_int64), ("inter", ctypes.c_void_p), ("inter_cap", ctypes.c_int64),
                ("edge_threshold", ctypes.c_double), ("edge_i", ctypes.c_void_p), ("edge_j", ctypes.c_void_p),
                ("edge_d", ctypes.c_void_p), ("edge_cap", ctypes.c_int64),
                ("n_intra", ctypes.c_int64), ("n_inter", ctypes.c_int64), ("n_edges", ctypes.c_int64),
                ("n_intra_local", ctypes.c_int64), ("n_inter_local", ctypes.c_int64), ("n_edges_local", ctypes.c_int64),
                ("hist_intra", ctypes.c_void_p), ("hist_inter", ctypes.c_void_p), ("hist_intra_cap", ctypes.c_int64), ("hist_inter_cap", ctypes.c_int64),
                ("n_hist_intra", ctypes.c_int64), ("n_hist_inter", ctypes.c_int64), ("hist_intra_pushes", ctypes.c_int64), ("hist_inter_pushes", ctypes.c_int64)]


class HostList(ctypes.Structure):
    """tdna_host_list: a list that is not on the device yet (tdna_analyze_host)."""
    _fields_ = [("symbols", ctypes.c_void_p), ("n", ctypes.c_int64), ("L", ctypes.c_int32), ("mode", ctypes.c_int32),
                ("row_stride", ctypes.c_int64), ("lengths", ctypes.c_void_p),
                ("species_id", ctypes.c_void_p), ("genus_id", ctypes.c_void_p), ("epithet_rank", ctypes.c_void_p), ("fullname_rank", ctypes.c_void_p),
                ("min_overlap", ctypes.c_int32), ("ambiguous_allowed", ctypes.c_int32)]


class StreamPart(ctypes.Structure):
    """tdna_stream_part: device pointers to one part's streaming Best-Match output."""
    _fields_ = [("cand_row", ctypes.c_void_p), ("cand_col", ctypes.c_void_p), ("cand_d", ctypes.c_void_p),
                ("n_cand", ctypes.c_int64), ("invalid_count", ctypes.c_void_p), ("flags", ctypes.c_void_p)]


COUNTS_DTYPE = np.dtype([("shared", "<i4"), ("c1", "<i4"), ("c2", "<i4"), ("c3", "<i4")])

ROW_RESULT_DTYPE = np.dtype([
    ("d_best", "<f8"), ("d_con", "<f8"), ("d_allo", "<f8"), ("d_other", "<f8"),
    ("best", "<i4"), ("con", "<i4"), ("allo", "<i4"), ("other", "<i4"),
    ("shared_con", "<i4"), ("shared_allo", "<i4"),
    ("tie_count", "<i4"), ("tie_species_min", "<i4"), ("tie_species_max", "<i4"), ("tie_unnamed", "<i4"),
    ("block_count", "<i4"),
    ("near_best", "<i4"), ("near_con", "<i4"), ("near_allo", "<i4"), ("near_other", "<i4"),
    ("unnamed_at_con", "<i4"), ("unnamed_at_allo", "<i4"), ("unnamed_at_other", "<i4"),
    ("n_valid", "<i4"), ("flags", "<i4"), ("reserved", "<i4"), ("reserved2", "<i4"),
], align=True)
assert ROW_RESULT_DTYPE.itemsize == 120, ROW_RESULT_DTYPE.itemsize

# every symbol include/taxondna_cuda.h declares (tests/test_abi.py checks the .so exports them all)
EXPORTS = [
    "tdna_abi_version", "tdna_init", "tdna_shutdown", "tdna_last_error", "tdna_stream", "tdna_set_memory_budget",
    "tdna_launch_count", "tdna_alignment_create", "tdna_alignment_destroy", "tdna_alignment_set_labels",
    "tdna_set_config", "tdna_set_row_range", "tdna_pair", "tdna_pairs", "tdna_row_distances", "tdna_query_distances", "tdna_matrix", "tdna_matrix_compute",
    "tdna_best_match", "tdna_best_match_compute", "tdna_class_distances", "tdna_edges_below", "tdna_pairs_between", "tdna_row_extremes",
    "tdna_valid_conspecifics", "tdna_set_progress", "tdna_cancel", "tdna_time_matrix_kernel", "tdna_measure_int_peaks",
    "tdna_analyze", "tdna_tensor_counts", "tdna_last_count_kernel", "tdna_best_match_seed", "tdna_best_match_part", "tdna_best_match_merge", "tdna_best_match_part_packed", "tdna_best_match_merge_gathered", "tdna_rows_per_part", "tdna_tensor_code", "tdna_consensus", "tdna_fasta_parse", "tdna_fasta_destroy", "tdna_fasta_count", "tdna_fasta_width", "tdna_fasta_records", "tdna_fasta_symbols", "tdna_alignment_create_from_fasta", "tdna_best_match_part_routed", "tdna_best_match_merge_routed", "tdna_set_option", "tdna_last_tile_pass_ms",
    "tdna_tensor_dims", "tdna_tensor_tables", "tdna_last_pass_stats", "tdna_pipelined_passes", "tdna_host_call_phase_ms", "tdna_multi_phase_ms", "tdna_analyze_host", "tdna_class_histogram", "tdna_comm_unique_id", "tdna_comm_init", "tdna_comm_destroy", "tdna_comm_info", "tdna_init_multi",
]

PROGRESS_FN = ctypes.CFUNCTYPE(None, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p)

_lib = None
_nccl = None


def preload_nccl():
    """Python hosts usually carry a pip-installed NCCL (nvidia-nccl-cu12, what torch links against).  The library binds NCCL with
    dlopen("libnccl.so.2") at the first tdna_comm_* call and would take the SYSTEM copy if none is loaded yet -- after which a
    later `import torch` would be bound to that (older) copy too and fail on its newer symbols.  So load the pip copy first."""
    global _nccl
    if _nccl is not None:
        return
    _nccl = False
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia.nccl")
        for loc in (spec.submodule_search_locations if spec else []):
            path = os.path.join(loc, "lib", "libnccl.so.2")
            if os.path.exists(path):
                _nccl = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
                return
    except Exception:  # noqa: BLE001 -- no pip NCCL: the system copy is what there is
        pass


def load():
    """Loads libtaxondna_cuda.so (built in-tree by __graft_entry__.build / csrc/Makefile)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TdnaError(-1, "CUDA library not built: %s is missing (run `python -c 'import __graft_entry__ as g; g.build()'`); "
                            "there is no CPU fallback" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    L.tdna_abi_version.restype = i32
    L.tdna_last_error.restype = ctypes.c_char_p
    L.tdna_init.argtypes = [i32, ctypes.POINTER(vp)]
    L.tdna_shutdown.argtypes = [vp]
    L.tdna_stream.argtypes = [vp]
    L.tdna_stream.restype = vp
    L.tdna_set_memory_budget.argtypes = [vp, i64]
    L.tdna_launch_count.argtypes = [vp]
    L.tdna_launch_count.restype = i64
    L.tdna_alignment_create.argtypes = [vp, vp, i64, i32, i64, vp, ctypes.POINTER(vp)]
    L.tdna_alignment_destroy.argtypes = [vp]
    L.tdna_alignment_set_labels.argtypes = [vp, vp, vp, vp, vp]
    L.tdna_set_config.argtypes = [vp, i32, i32, i32]
    L.tdna_set_row_range.argtypes = [vp, i64, i64]
    L.tdna_pair.argtypes = [vp, i64, i64, ctypes.POINTER(Counts), ctypes.POINTER(dbl)]
    L.tdna_pairs.argtypes = [vp, vp, vp, i64, vp, vp]
    if hasattr(L, "tdna_tensor_code"):
        L.tdna_tensor_code.argtypes = [i32, i32, vp, vp]
    if hasattr(L, "tdna_consensus"):
        L.tdna_consensus.argtypes = [vp, vp, vp, i64, vp, i64, vp]
    if hasattr(L, "tdna_fasta_parse"):
        L.tdna_fasta_parse.argtypes = [vp, vp, i64, ctypes.POINTER(vp)]
        L.tdna_fasta_destroy.argtypes = [vp]
        L.tdna_fasta_count.argtypes = [vp]
        L.tdna_fasta_count.restype = i64
        L.tdna_fasta_width.argtypes = [vp]
        L.tdna_fasta_records.argtypes = [vp, vp, vp, vp, vp]
        L.tdna_fasta_symbols.argtypes = [vp, vp, i64]
        L.tdna_alignment_create_from_fasta.argtypes = [vp, vp, ctypes.POINTER(vp)]
    if hasattr(L, "tdna_rows_per_part"):  # absent from older builds loaded through TDNA_LIB for an A/B
        L.tdna_rows_per_part.argtypes = [i64, i32]
        L.tdna_rows_per_part.restype = i64
        L.tdna_best_match_part_routed.argtypes = [vp, i32, i32, i32, vp, i64, vp, vp]
        L.tdna_best_match_merge_routed.argtypes = [vp, i32, i32, vp, i64, vp, vp, vp]
    L.tdna_row_distances.argtypes = [vp, i64, vp, vp]
    L.tdna_query_distances.argtypes = [vp, vp, i32, vp, vp, vp]
    L.tdna_matrix.argtypes = [vp, vp, vp]
    L.tdna_matrix_compute.argtypes = [vp, ctypes.POINTER(vp)]
    L.tdna_best_match.argtypes = [vp, vp]
    L.tdna_best_match_compute.argtypes = [vp, ctypes.POINTER(vp)]
    L.tdna_class_distances.argtypes = [vp, i32, vp, i64, ctypes.POINTER(i64)]
    L.tdna_edges_below.argtypes = [vp, dbl, vp, vp, vp, i64, ctypes.POINTER(i64)]
    L.tdna_row_extremes.argtypes = [vp, vp]
    L.tdna_pairs_between.argtypes = [vp, dbl, dbl, vp, vp, vp, i64, ctypes.POINTER(i64)]
    L.tdna_class_histogram.argtypes = [vp, i32, vp, i64, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.tdna_valid_conspecifics.argtypes = [vp, vp]
    L.tdna_set_progress.argtypes = [vp, PROGRESS_FN, vp]
    L.tdna_cancel.argtypes = [vp]
    L.tdna_time_matrix_kernel.argtypes = [vp, i32, ctypes.POINTER(ctypes.c_float)]
    L.tdna_measure_int_peaks.argtypes = [vp, ctypes.POINTER(dbl), ctypes.POINTER(dbl)]
    L.tdna_tensor_counts.argtypes = [vp, vp]
    L.tdna_last_count_kernel.argtypes = [vp]
    L.tdna_tensor_dims.argtypes = [vp]
    L.tdna_host_call_phase_ms.argtypes = [vp, vp, i32, ctypes.POINTER(i32)]
    L.tdna_pipelined_passes.argtypes = [vp]
    L.tdna_pipelined_passes.restype = i64
    L.tdna_last_pass_stats.argtypes = [vp, ctypes.POINTER(i64), ctypes.POINTER(i64)]
    L.tdna_multi_phase_ms.argtypes = [vp, vp, i32, ctypes.POINTER(i32)]
    L.tdna_analyze.argtypes = [vp, ctypes.POINTER(Analysis)]
    L.tdna_analyze_host.argtypes = [vp, ctypes.POINTER(HostList), ctypes.POINTER(Analysis), vp]
    L.tdna_best_match_seed.argtypes = [vp, i32, i32, ctypes.POINTER(vp)]
    L.tdna_best_match_part.argtypes = [vp, i32, i32, i32, ctypes.POINTER(StreamPart)]
    L.tdna_best_match_merge.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp]
    L.tdna_best_match_part_packed.argtypes = [vp, i32, i32, i32, vp, i64, vp, vp]
    L.tdna_best_match_merge_gathered.argtypes = [vp, vp, i32, i64, vp, vp, vp]
    L.tdna_set_option.argtypes = [vp, i32, i64]
    L.tdna_comm_unique_id.argtypes = [vp]
    L.tdna_comm_init.argtypes = [vp, vp, i32, i32]
    L.tdna_comm_destroy.argtypes = [vp]
    L.tdna_comm_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32)]
    L.tdna_init_multi.argtypes = [vp, i32, vp]
    L.tdna_last_tile_pass_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise TdnaError(rc, load().tdna_last_error().decode("utf-8", "replace"))


class Context:
    """tdna_ctx: one per GPU per process."""

    def __init__(self, device=0, handle=None):
        self.L = load()
        self.h = ctypes.c_void_p()
        if handle is not None:
            self.h = ctypes.c_void_p(handle)  # a context tdna_init_multi created
        else:
            check(self.L.tdna_init(device, ctypes.byref(self.h)))
        self.device = device
        self._cb = None

    @staticmethod
    def init_multi(devices):
        """One process, several GPUs (the one-JVM layout): one context per device, joined in one communicator.  Collective
        calls are then made from one thread per context."""
        L = load()
        preload_nccl()
        devs = (ctypes.c_int32 * len(devices))(*devices)
        out = (ctypes.c_void_p * len(devices))()
        check(L.tdna_init_multi(devs, len(devices), out))
        return [Context(d, handle=h) for d, h in zip(devices, out)]

    @staticmethod
    def comm_unique_id():
        preload_nccl()
        buf = ctypes.create_string_buffer(COMM_ID_BYTES)
        check(load().tdna_comm_unique_id(buf))
        return buf.raw

    def comm_init(self, unique_id, rank, nranks):
        """Collective over the nranks contexts (one per GPU): joins the NCCL communicator the library runs its multi-GPU passes on."""
        assert len(unique_id) == COMM_ID_BYTES
        preload_nccl()
        check(self.L.tdna_comm_init(self.h, ctypes.c_char_p(unique_id), int(rank), int(nranks)))

    def comm_init_torch(self, dist, device):
        """comm_init with the id broadcast through an existing torch.distributed group (control plane only)."""
        import torch
        rank, world = dist.get_rank(), dist.get_world_size()
        t = torch.zeros(COMM_ID_BYTES, dtype=torch.uint8, device=device)
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(Context.comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        self.comm_init(bytes(t.cpu().numpy().tobytes()), rank, world)

    MULTI_PHASES = ("seed", "allreduce_min", "bulk+resolve", "route+exchange", "merge_own_rows", "allgather_rows")

    def multi_phase_ms(self):
        """Device time of the phases of the most recent multi-GPU pass: {name: ms}."""
        ms = (ctypes.c_float * 12)()
        k = ctypes.c_int32()
        check(self.L.tdna_multi_phase_ms(self.h, ms, 12, ctypes.byref(k)))
        return {nm: float(ms[i]) for i, nm in zip(range(k.value), self.MULTI_PHASES)}

    def comm_destroy(self):
        check(self.L.tdna_comm_destroy(self.h))

    def comm_info(self):
        r, n = ctypes.c_int32(), ctypes.c_int32()
        check(self.L.tdna_comm_info(self.h, ctypes.byref(r), ctypes.byref(n)))
        return r.value, n.value

    def close(self):
        if self.h:
            self.L.tdna_shutdown(self.h)
            self.h = ctypes.c_void_p()

    def stream_ptr(self):
        return self.L.tdna_stream(self.h)

    def set_memory_budget(self, nbytes):
        check(self.L.tdna_set_memory_budget(self.h, int(nbytes)))

    def launch_count(self):
        return int(self.L.tdna_launch_count(self.h))

    def host_call_phase_ms(self):
        ms = (ctypes.c_double * 12)()
        k = ctypes.c_int32()
        check(self.L.tdna_host_call_phase_ms(self.h, ms, 12, ctypes.byref(k)))
        names = ("labels+plan", "buffers", "uploads_enqueued", "first_chunk_verdict", "chunks_enqueued", "upload_synced", "pass_done", "analysis_done", "returned")
        return {nm: round(ms[i], 4) for i, nm in zip(range(k.value), names)}

    def pipelined_passes(self):
        return int(self.L.tdna_pipelined_passes(self.h))

    def set_progress(self, fn):
        self._cb = PROGRESS_FN(lambda done, total, user: fn(done, total)) if fn else PROGRESS_FN()
        check(self.L.tdna_set_progress(self.h, self._cb, None))

    def cancel(self):
        check(self.L.tdna_cancel(self.h))

    def set_option(self, key, value):
        check(self.L.tdna_set_option(self.h, int(key), int(value)))

    def last_tile_pass_ms(self):
        ms = ctypes.c_float()
        check(self.L.tdna_last_tile_pass_ms(self.h, ctypes.byref(ms)))
        return ms.value

    def measure_int_peaks(self):
        a, b = ctypes.c_double(), ctypes.c_double()
        check(self.L.tdna_measure_int_peaks(self.h, ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value


def host_list(symbols, labels, mode, min_overlap, ambiguous_allowed=True, lengths=None):
    """HostList over numpy arrays (kept alive by the returned tuple: (HostList, keepalive))."""
    symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
    labs = [np.ascontiguousarray(a, dtype=np.int32) for a in labels] if labels is not None else [None] * 4
    lens = np.ascontiguousarray(lengths, dtype=np.int32) if lengths is not None else None
    h = HostList()
    h.symbols, h.n, h.L, h.mode, h.row_stride = symbols.ctypes.data, symbols.shape[0], symbols.shape[1], int(mode), symbols.strides[0]
    h.lengths = lens.ctypes.data if lens is not None else None
    if labels is not None:
        h.species_id, h.genus_id, h.epithet_rank, h.fullname_rank = [a.ctypes.data for a in labs]
    h.min_overlap, h.ambiguous_allowed = int(min_overlap), 1 if ambiguous_allowed else 0
    return h, (symbols, labs, lens)


def best_match_host(ctx, symbols, labels, mode, min_overlap, ambiguous_allowed=True, lengths=None, out=None):
    """Best Match rows of a list that is not on the device yet, in one call (tdna_analyze_host: the upload is pipelined with the pass)."""
    h, keep = host_list(symbols, labels, mode, min_overlap, ambiguous_allowed, lengths)
    r = out if out is not None else np.zeros(h.n, dtype=ROW_RESULT_DTYPE)
    io = Analysis()
    io.want = WANT_BEST_MATCH
    io.rows = r.ctypes.data
    check(ctx.L.tdna_analyze_host(ctx.h, ctypes.byref(h), ctypes.byref(io), None))
    return r


class Alignment:
    """tdna_aln over a normalised symbol matrix (numpy uint8 [n, L])."""

    def __init__(self, ctx, symbols, lengths=None):
        self.ctx, self.L = ctx, ctx.L
        symbols = np.ascontiguousarray(symbols, dtype=np.uint8)
        self.n, self.ncols = symbols.shape
        self.h = ctypes.c_void_p()
        lp = None
        if lengths is not None:
            lengths = np.ascontiguousarray(lengths, dtype=np.int32)
            lp = lengths.ctypes.data
        check(self.L.tdna_alignment_create(ctx.h, symbols.ctypes.data, self.n, self.ncols, symbols.strides[0], lp,
                                           ctypes.byref(self.h)))
        self.row_begin, self.row_end = 0, self.n

    def close(self):
        if self.h:
            self.L.tdna_alignment_destroy(self.h)
            self.h = ctypes.c_void_p()

    def set_labels(self, species_id, genus_id, epithet_rank, fullname_rank):
        arrs = [np.ascontiguousarray(a, dtype=np.int32) for a in (species_id, genus_id, epithet_rank, fullname_rank)]
        assert all(a.shape == (self.n,) for a in arrs)
        check(self.L.tdna_alignment_set_labels(self.h, *[a.ctypes.data for a in arrs]))

    def set_config(self, mode, min_overlap, ambiguous_allowed=True):
        check(self.L.tdna_set_config(self.h, int(mode), int(min_overlap), 1 if ambiguous_allowed else 0))

    def set_row_range(self, b, e):
        check(self.L.tdna_set_row_range(self.h, int(b), int(e)))
        self.row_begin, self.row_end = int(b), int(e)

    def pair(self, i, j):
        c, d = Counts(), ctypes.c_double()
        check(self.L.tdna_pair(self.h, i, j, ctypes.byref(c), ctypes.byref(d)))
        return (c.shared, c.c1, c.c2, c.c3), d.value

    def pairs(self, ii, jj, want_counts=False):
        """A batch of pairs in one launch: returns d (and the counts)."""
        ii = np.ascontiguousarray(ii, dtype=np.int32)
        jj = np.ascontiguousarray(jj, dtype=np.int32)
        assert ii.shape == jj.shape and ii.ndim == 1
        d = np.empty(len(ii), dtype=np.float64)
        c = np.empty(len(ii), dtype=COUNTS_DTYPE) if want_counts else None
        check(self.L.tdna_pairs(self.h, ii.ctypes.data, jj.ctypes.data, len(ii), c.ctypes.data if want_counts else None, d.ctypes.data))
        return (d, c) if want_counts else d

    def row_distances(self, q, want_shared=True):
        d = np.empty(self.n, dtype=np.float64)
        sh = np.empty(self.n, dtype=np.int32) if want_shared else None
        check(self.L.tdna_row_distances(self.h, q, d.ctypes.data, sh.ctypes.data if want_shared else None))
        return d, sh

    def consensus(self, bins):
        """Consensus sequence of every bin (list of lists of row indices): (uint8 [nbins, L] '?'-padded, int32 lengths)."""
        members = np.ascontiguousarray(np.concatenate([np.asarray(b, dtype=np.int32) for b in bins]) if bins else np.zeros(0, np.int32))
        offsets = np.zeros(len(bins) + 1, dtype=np.int64)
        offsets[1:] = np.cumsum([len(b) for b in bins])
        out = np.full((len(bins), self.ncols), ord("?"), dtype=np.uint8)
        lens = np.zeros(len(bins), dtype=np.int32)
        check(self.L.tdna_consensus(self.h, members.ctypes.data, offsets.ctypes.data, len(bins), out.ctypes.data, self.ncols, lens.ctypes.data))
        return out, lens

    def query_distances(self, symbols, want_counts=False):
        """A normalised sequence that is not a row of the alignment against every row (QuerySequence.java:113-145)."""
        q = np.ascontiguousarray(np.frombuffer(symbols, dtype=np.uint8) if isinstance(symbols, (bytes, bytearray)) else symbols, dtype=np.uint8)
        d = np.empty(self.n, dtype=np.float64)
        sh = np.empty(self.n, dtype=np.int32)
        c = np.empty(self.n, dtype=COUNTS_DTYPE) if want_counts else None
        check(self.L.tdna_query_distances(self.h, q.ctypes.data, int(q.size), d.ctypes.data, sh.ctypes.data, c.ctypes.data if want_counts else None))
        return (d, sh, c) if want_counts else (d, sh)

    def matrix(self, counts=False, out=None):
        rows = self.row_end - self.row_begin
        d = out if out is not None else np.empty((rows, self.n), dtype=np.float64)
        c = np.empty((rows, self.n), dtype=COUNTS_DTYPE) if counts else None
        check(self.L.tdna_matrix(self.h, d.ctypes.data, c.ctypes.data if counts else None))
        return (d, c) if counts else d

    def matrix_compute(self):
        p = ctypes.c_void_p()
        check(self.L.tdna_matrix_compute(self.h, ctypes.byref(p)))
        return p.value

    def best_match(self, out=None):
        r = out if out is not None else np.zeros(self.n, dtype=ROW_RESULT_DTYPE)
        check(self.L.tdna_best_match(self.h, r.ctypes.data))
        return r

    def best_match_into(self, ptr):
        """ptr: host or device address of n tdna_row_result records."""
        check(self.L.tdna_best_match(self.h, ptr))

    def best_match_compute(self):
        p = ctypes.c_void_p()
        check(self.L.tdna_best_match_compute(self.h, ctypes.byref(p)))
        return p.value

    def analyze(self, best_match=False, intra=False, inter=False, edges=None):
        """One pass, several analyses (tdna_analyze).  Two calls when class distances / edges are wanted:
        the first sizes the outputs.  Returns a dict of numpy arrays."""
        want = (WANT_BEST_MATCH if best_match else 0) | (WANT_INTRA if intra else 0) | (WANT_INTER if inter else 0) | \
               (WANT_EDGES if edges is not None else 0)
        io = Analysis()
        io.want = want
        io.edge_threshold = float(edges) if edges is not None else 0.0
        out = {}
        if want & ~WANT_BEST_MATCH:  # size the variable-length outputs
            io.want = want & ~WANT_BEST_MATCH
            check(self.L.tdna_analyze(self.h, ctypes.byref(io)))
            io.want = want
        if best_match:
            out["rows"] = np.zeros(self.n, dtype=ROW_RESULT_DTYPE)
            io.rows = out["rows"].ctypes.data
        if intra:
            out["intra"] = np.empty(max(1, io.n_intra), dtype=np.float32)
            io.intra, io.intra_cap = out["intra"].ctypes.data, io.n_intra
        if inter:
            out["inter"] = np.empty(max(1, io.n_inter), dtype=np.float32)
            io.inter, io.inter_cap = out["inter"].ctypes.data, io.n_inter
        if edges is not None:
            k = max(1, io.n_edges)
            out["edge_i"], out["edge_j"], out["edge_d"] = np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.float64)
            io.edge_i, io.edge_j, io.edge_d, io.edge_cap = out["edge_i"].ctypes.data, out["edge_j"].ctypes.data, out["edge_d"].ctypes.data, io.n_edges
        ni, nn, ne = io.n_intra, io.n_inter, io.n_edges
        check(self.L.tdna_analyze(self.h, ctypes.byref(io)))
        assert (not intra or io.n_intra == ni) and (not inter or io.n_inter == nn) and (edges is None or io.n_edges == ne)
        # with a communicator the variable-length lists hold THIS rank's share (n_*_local entries); the totals are in n_*
        if intra:
            out["intra"] = out["intra"][:io.n_intra_local]
        if inter:
            out["inter"] = out["inter"][:io.n_inter_local]
        if edges is not None:
            k = io.n_edges_local
            out["edge_i"], out["edge_j"], out["edge_d"] = out["edge_i"][:k], out["edge_j"][:k], out["edge_d"][:k]
        out["totals"] = (io.n_intra, io.n_inter, io.n_edges)
        return out

    def last_count_kernel(self):
        return int(self.L.tdna_last_count_kernel(self.h))

    def last_pass_stats(self):
        c, q = ctypes.c_int64(), ctypes.c_int64()
        check(self.L.tdna_last_pass_stats(self.h, ctypes.byref(c), ctypes.byref(q)))
        return {"candidates": c.value, "queued_pairs": q.value}

    def tensor_dims(self):
        return int(self.L.tdna_tensor_dims(self.h))

    def tensor_counts(self):
        """(shared, ident) of every pair from count kernel (b); pairs outside the triangle's tiles are -1."""
        c = np.empty((self.n, self.n), dtype=COUNTS_DTYPE)
        check(self.L.tdna_tensor_counts(self.h, c.ctypes.data))
        return c

    def best_match_seed(self, part, nparts):
        """Diagonal blocks of this part; returns the device address of the 3n uint64 per-row minima."""
        p = ctypes.c_void_p()
        check(self.L.tdna_best_match_seed(self.h, int(part), int(nparts), ctypes.byref(p)))
        return p.value

    def best_match_part(self, part, nparts, seeded=False):
        """Streaming pass over every nparts-th tile of the triangle; returns a StreamPart of DEVICE pointers."""
        sp = StreamPart()
        check(self.L.tdna_best_match_part(self.h, int(part), int(nparts), 1 if seeded else 0, ctypes.byref(sp)))
        return sp

    def best_match_part_packed(self, part, nparts, seeded, rec_ptr, seg_stride, count_ptr, state_ptr):
        """The part's outputs packed into caller-provided DEVICE buffers (addresses), ready for NCCL."""
        check(self.L.tdna_best_match_part_packed(self.h, int(part), int(nparts), 1 if seeded else 0, rec_ptr, int(seg_stride), count_ptr, state_ptr))

    def rows_per_part(self, nparts):
        return int(self.L.tdna_rows_per_part(self.n, int(nparts)))

    def best_match_part_routed(self, part, nparts, seeded, rec_ptr, seg_stride, counts_ptr, state_ptr):
        check(self.L.tdna_best_match_part_routed(self.h, int(part), int(nparts), 1 if seeded else 0, rec_ptr, int(seg_stride), counts_ptr, state_ptr))

    def best_match_merge_routed(self, part, nparts, rec_ptr, seg_stride, seg_count_ptr, state_ptr, out_slice_ptr=None):
        check(self.L.tdna_best_match_merge_routed(self.h, int(part), int(nparts), rec_ptr, int(seg_stride), seg_count_ptr, state_ptr, out_slice_ptr))

    def best_match_merge_gathered(self, rec_ptr, nseg, seg_stride, seg_count_ptr, state_ptr, out_ptr=None):
        check(self.L.tdna_best_match_merge_gathered(self.h, rec_ptr, int(nseg), int(seg_stride), seg_count_ptr, state_ptr, out_ptr))

    def best_match_merge(self, cand_row, cand_col, cand_d, n_cand, invalid_count, flags, out=None):
        """All inputs are device addresses (ints); returns the n row records (host)."""
        r = out if out is not None else np.zeros(self.n, dtype=ROW_RESULT_DTYPE)
        check(self.L.tdna_best_match_merge(self.h, cand_row, cand_col, cand_d, int(n_cand), invalid_count, flags, r.ctypes.data))
        return r

    def class_distances(self, klass):
        n = ctypes.c_int64()
        check(self.L.tdna_class_distances(self.h, klass, None, 0, ctypes.byref(n)))
        out = np.empty(max(n.value, 1), dtype=np.float32)
        m = ctypes.c_int64()
        check(self.L.tdna_class_distances(self.h, klass, out.ctypes.data, n.value, ctypes.byref(m)))
        assert m.value == n.value
        return out[: n.value]

    def class_histogram(self, klass):
        """(entries sorted by d, pushes): the distinct distances of a PairwiseDistribution class with their counts (tdna_class_histogram)."""
        n, pushes = ctypes.c_int64(), ctypes.c_int64()
        cap = 1 << 16
        while True:
            out = np.zeros(cap, dtype=HIST_ENTRY_DTYPE)
            check(self.L.tdna_class_histogram(self.h, klass, out.ctypes.data, cap, ctypes.byref(n), ctypes.byref(pushes)))
            if n.value <= cap:
                break
            cap = int(n.value)
        out = out[: n.value]
        return out[np.argsort(out["d"], kind="stable")], pushes.value

    def edges_below(self, t):
        n = ctypes.c_int64()
        check(self.L.tdna_edges_below(self.h, float(t), None, None, None, 0, ctypes.byref(n)))
        k = max(n.value, 1)
        ii, jj, dd = np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.float64)
        m = ctypes.c_int64()
        check(self.L.tdna_edges_below(self.h, float(t), ii.ctypes.data, jj.ctypes.data, dd.ctypes.data, n.value, ctypes.byref(m)))
        return ii[: n.value], jj[: n.value], dd[: n.value]

    def row_extremes(self):
        """tdna_row_extremes: one record per row (EXTREME_RESULT_DTYPE)."""
        out = np.zeros(self.n, dtype=EXTREME_RESULT_DTYPE)
        check(self.L.tdna_row_extremes(self.h, out.ctypes.data))
        return out

    def pairs_between(self, lo, hi):
        """All unordered pairs i < j with lo <= d <= hi (FindDistances.java:227), unsorted: (ii, jj, d)."""
        n = ctypes.c_int64()
        check(self.L.tdna_pairs_between(self.h, float(lo), float(hi), None, None, None, 0, ctypes.byref(n)))
        k = max(n.value, 1)
        ii, jj, dd = np.empty(k, np.int32), np.empty(k, np.int32), np.empty(k, np.float64)
        m = ctypes.c_int64()
        check(self.L.tdna_pairs_between(self.h, float(lo), float(hi), ii.ctypes.data, jj.ctypes.data, dd.ctypes.data, n.value, ctypes.byref(m)))
        return ii[: n.value], jj[: n.value], dd[: n.value]

    def valid_conspecifics(self):
        v = np.zeros(self.n, dtype=np.uint8)
        check(self.L.tdna_valid_conspecifics(self.h, v.ctypes.data))
        return v

    def time_matrix_kernel(self, reps=10):
        ms = ctypes.c_float()
        check(self.L.tdna_time_matrix_kernel(self.h, reps, ctypes.byref(ms)))
        return ms.value


FASTA_URACIL, FASTA_INVALID, FASTA_NEEDS_HOST = 1, 2, 4


class Fasta:
    """FASTA text parsed on the device (tdna_fasta_parse): record table + the normalised symbol matrix, device resident."""

    def __init__(self, ctx, text):
        self.L, self.ctx = ctx.L, ctx
        data = np.frombuffer(text if isinstance(text, (bytes, bytearray)) else text.encode("latin-1"), dtype=np.uint8)
        self.text = data
        h = ctypes.c_void_p()
        check(self.L.tdna_fasta_parse(ctx.h, data.ctypes.data if data.size else b"", int(data.size), ctypes.byref(h)))
        self.h = h
        self.n = int(self.L.tdna_fasta_count(self.h))
        self.width = int(self.L.tdna_fasta_width(self.h))

    def records(self):
        off = np.zeros(self.n, dtype=np.int64)
        hl = np.zeros(self.n, dtype=np.int32)
        sl = np.zeros(self.n, dtype=np.int32)
        fl = np.zeros(self.n, dtype=np.int32)
        check(self.L.tdna_fasta_records(self.h, off.ctypes.data, hl.ctypes.data, sl.ctypes.data, fl.ctypes.data))
        return off, hl, sl, fl

    def headers(self):
        off, hl, _, _ = self.records()
        raw = self.text.tobytes()
        return [raw[o:o + k].decode("latin-1") for o, k in zip(off.tolist(), hl.tolist())]

    def symbols(self):
        out = np.full((self.n, max(self.width, 1)), ord("?"), dtype=np.uint8)
        if self.n and self.width:
            check(self.L.tdna_fasta_symbols(self.h, out.ctypes.data, out.shape[1]))
        return out

    def alignment(self):
        """An Alignment straight from the device-resident records (no host round trip of the symbols)."""
        h = ctypes.c_void_p()
        check(self.L.tdna_alignment_create_from_fasta(self.ctx.h, self.h, ctypes.byref(h)))
        a = Alignment.__new__(Alignment)
        a.L, a.ctx, a.h, a.n, a.ncols = self.L, self.ctx, h, self.n, self.width
        a.row_begin, a.row_end = 0, self.n
        return a

    def close(self):
        if self.h:
            self.L.tdna_fasta_destroy(self.h)
            self.h = None

