#!/usr/bin/env python
"""bench.py -- the measurement contract.

  python bench.py --gpus N --steps K --warmup W [--workload C2|C3|H|C4] [--impl reference]

A "step" is one pass of the hot path over one synthetic alignment of a BASELINE.json shape:
  H (default): 50,000 x 658, p: per-query Best Match summaries -- the shape BASELINE.json's north_star quotes
      its target on ("... at 50,000 x 658 bp ..."); fits one GPU
  C2 (configs[1]): 2,185 x 2,664, uncorrected p: full fp64 pairwise matrix + the two Pairwise Summary
      classes (intra / congeneric inter distance extraction)
  C3: 8,000 x 658, K2P: per-query Best Close Match summaries + intraspecific distances (threshold percentile)
metric = unordered pairs i<j per second (Gpairs/s).  `value`: inputs (bit-planes, labels) resident
in HBM, results left in HBM, each step timed with CUDA events on the library's stream, L2 flushed
between steps.  `e2e`: the same work through the C ABI with HOST buffers (symbols in, results out),
host<->device copies inside the timed region.

N > 1 (torchrun, one rank per GPU): every rank holds the whole packed alignment (each uploads 1/N of the rows, the library
all-gathers them over NVLink).  Best-Match steps are ONE library call per rank: inside it (tdna_comm_init, NCCL on the library's
stream) the ranks take units of 128 tile slots of the triangle cyclically, all-reduce the seeded per-row minima, route the candidate
records to the rank that owns the row, and all-gather the row records; matrix steps split the query rows into bands (no
collective).  --scaling strong (default): the named shape itself on N GPUs; the other mode (weak: the alignment grows as sqrt(N)
so that pairs per GPU stay fixed) is measured too and reported under an extra key.  After the timed region rank 0 recomputes the
result on ONE GPU and compares it byte for byte with the all-gathered N-GPU result: "parity_vs_1gpu".

--impl reference: the reference's CPU algorithm (oracle/ C restatement of Sequence.getPairwise --
the reference is Java and no JVM exists in this image) on the host cores, bounded sample.
"""
import argparse
import ctypes
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

_OUT = sys.stdout

WORKLOADS = {
    "C2": dict(config="C2", mode=0, kind="matrix+summary", desc="synthetic multi-gene alignment, uncorrected p: full pairwise matrix + Pairwise Summary classes"),
    "C3": dict(config="C3", mode=1, kind="bestmatch+intra", desc="synthetic COI, K2P: Best Close Match + intraspecific distances for the 95th-percentile threshold"),
    "C4": dict(config="C4", mode=0, kind="matrix+summary", desc="gappy/ambiguous alignment (20% ?/-/IUPAC), uncorrected p: matrix + Pairwise Summary classes"),
    "H": dict(config="H", mode=0, kind="bestmatch", desc="synthetic COI, uncorrected p: streaming per-query Best Match summaries"),
    # H's shape with C4's noise (8 % '?', 8 % '-', 4 % IUPAC letters): no tile is free of gaps / ambiguity letters (the GENERAL instantiation)
    "Hg": dict(config="Hg", mode=0, kind="bestmatch", desc="synthetic COI with 20% ?/-/IUPAC noise, uncorrected p: streaming per-query Best Match summaries"),
    "HgK": dict(config="Hg", mode=1, kind="bestmatch", desc="synthetic COI with 20% ?/-/IUPAC noise, K2P: streaming per-query Best Match summaries"),
    # transversions only (PDM_TRANS_ONLY): six K-bytes per site on count kernel (b), exact for every letter
    "Ht": dict(config="H", mode=2, kind="bestmatch", desc="synthetic COI, transversions only: streaming per-query Best Match summaries"),
    "Hgt": dict(config="Hg", mode=2, kind="bestmatch", desc="synthetic COI with 20% ?/-/IUPAC noise, transversions only: streaming per-query Best Match summaries"),
}
POPC_PER_PAIR_WORD = {0: 2, 1: 4, 2: 2}  # C = 2 (p), 4 (K2P), 2 (trans-only)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="H", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong = the named shape itself on N GPUs (the north star's own shapes); weak = n grows as sqrt(N)")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the run of the other scaling mode (extra key)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="target CPU time of the cpu_baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--kernel", default="auto", choices=["auto", "popc", "tensor"], help="count kernel of the streaming pass")
    return ap.parse_args()


class ClockSampler:
    """SM clock + throttle reasons while the timed region runs (B200_PROFILING.md recipe): NVML polled every 10 ms from a thread
    (the region is tens of milliseconds; nvidia-smi -lms 100 would see it twice -- and a 2 ms period cost the multi-rank step 6 %
    through the interpreter lock), nvidia-smi when NVML is unavailable."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []
        self.nvml, self.handle, self.samples, self.bits, self.stop_flag = None, None, [], 0, False

    def start(self):
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:
                uuid = str(torch.cuda.get_device_properties(self.index).uuid)
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode() if not uuid.startswith("GPU-") else uuid.encode())
            except Exception:  # noqa: BLE001
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:  # noqa: BLE001
            self.nvml = None
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self.stop_flag:
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                try:
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:  # noqa: BLE001
                    self.bits |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.010)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag = True
            self.thread.join(timeout=1)
            reasons = sorted(name for bit, name in self.REASONS if self.bits & bit)
            return {"sm_mhz": float(np.median(self.samples)) if self.samples else None, "sm_max_mhz": self.max_mhz, "reasons": reasons,
                    "samples": len(self.samples), "source": "nvml, 10 ms period"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 100"}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def cpu_reference_leg(data, mode, seconds, threads):
    """Times the oracle's getPairwiseNoBuffer restatement on a bounded sample of the same workload."""
    import oracle
    sym = data["symbols"]
    n, L = sym.shape
    lens = np.full(n, L, dtype=np.int32)
    # calibrate on a few rows, then size the sample for ~`seconds` of wall time
    t0 = time.perf_counter()
    _, cnt = oracle.bench_pairs(sym, lens, mode, 300, True, 0, min(n, 2 * max(1, threads)), threads)
    dt = max(time.perf_counter() - t0, 1e-4)
    rate = cnt / dt
    want = rate * seconds
    rows = 1
    acc = 0
    while rows < n and acc < want:
        acc += n - rows
        rows += 1
    t0 = time.perf_counter()
    _, cnt = oracle.bench_pairs(sym, lens, mode, 300, True, 0, rows, threads)
    dt = time.perf_counter() - t0
    return cnt / dt, cnt, dt, rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from taxondna_b200 import synth
    wl = WORKLOADS[args.workload]
    data = synth.generate(wl["config"], names=False)
    n, L = data["symbols"].shape
    threads = os.cpu_count() or 1
    per_step = max(2.0, min(20.0, 90.0 / max(1, args.steps + args.warmup)))
    for _ in range(args.warmup):
        cpu_reference_leg(data, wl["mode"], per_step, threads)
    rates, pairs, secs = [], 0, 0.0
    for _ in range(args.steps):
        r, cnt, dt, rows = cpu_reference_leg(data, wl["mode"], per_step, threads)
        rates.append(r)
        pairs += cnt
        secs += dt
    v = pairs / secs / 1e9
    line = {
        "impl": "reference", "metric": "pairwise distances/s", "value": v, "unit": "Gpairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / max(1, args.steps) * 1e3, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "int32 counts + f64 distance", "data": "synthetic",
        "config": {"workload": "%s x %s bp: %s" % (format(n, ","), format(L, ","), wl["desc"]), "n": n, "L": L, "count_kernel": "cpu (per-pair char loops)",
                   "mode": ["uncorrected", "K2P", "trans-only"][wl["mode"]], "min_overlap": 300, "pairs": n * (n - 1) // 2, "step": wl["kind"],
                   "l2": "n/a (host cores)", "parallelism": "%d host threads, a sample of the query rows per step" % threads},
        "cpu_baseline": {"value": v, "unit": "Gpairs/s", "cores": threads, "kind": "port",
                         "sample": "first %d query rows x all later columns per step (%d pairs/step), C restatement of Sequence.getPairwiseNoBuffer (no JVM in image), %d pthreads" % (rows, pairs // max(1, args.steps), threads)},
        "e2e": {"value": v, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_OUT, flush=True)


class _DevBuf:
    """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def int8_peak_live(torch, dev):
    """A cuBLASLt int8 GEMM (torch._int_mm, 8192^3) timed here: the tensor-pipe denominator (MEASURED_PEAKS.json has no int8 line)."""
    a8 = torch.randint(-3, 3, (8192, 8192), dtype=torch.int8, device=dev)
    b8 = torch.randint(-3, 3, (8192, 8192), dtype=torch.int8, device=dev)
    for _ in range(3):
        torch._int_mm(a8, b8)
    best = 1e9
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch._int_mm(a8, b8)
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a8, b8
    return 2 * 8192 ** 3 / (best * 1e-3)


class Bench:
    """One process per GPU.  Every data-plane step of the N > 1 path happens inside the C ABI (tdna_comm_init: NCCL on the library's
    stream); torch.distributed is the control plane only (the unique-id broadcast, barriers, the max over ranks of the timings)."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist

        import taxondna_b200 as t
        self.args, self.torch, self.dist, self.t = args, torch, dist, t
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
        if os.environ.get("TDNA_BENCH_TORCH_THREADS"):
            torch.set_num_threads(int(os.environ["TDNA_BENCH_TORCH_THREADS"]))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.ctx = t.Context(self.local)
        self.ctx.set_option(t.OPT_COUNT_KERNEL, {"auto": t.KERNEL_AUTO, "popc": t.KERNEL_POPC, "tensor": t.KERNEL_TENSOR}[args.kernel])
        if self.world > 1:
            self.ctx.comm_init_torch(dist, self.dev)  # from here on whole-range library calls are collective over the ranks
        if os.environ.get("TDNA_BENCH_UNIT"):
            self.ctx.set_option(t.OPT_UNIT_ITEMS, int(os.environ["TDNA_BENCH_UNIT"]))
        self.stream = torch.cuda.ExternalStream(self.ctx.stream_ptr(), device=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)  # > 126 MB L2
        self.L_ = t.load()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        tm = self.torch.tensor([x], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(tm, op=self.dist.ReduceOp.MAX)
        return float(tm.item())

    def close(self):
        self.flush = None
        gc.collect()
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.ctx.comm_destroy()
        self.ctx.close()
        if self.world > 1:
            self.dist.destroy_process_group()

    def measure(self, workload, scale_n=None, steps=None, warmup=None, full=True):
        """Times `steps` steps of the workload.  full: also the roofline's live peak, the e2e leg, the parity check against one GPU
        and (N = 1) the CPU baseline; otherwise only `value` and `e2e`."""
        from taxondna_b200 import synth
        torch, dist, t, args = self.torch, self.dist, self.t, self.args
        world, rank, dev, ctx, stream, L_ = self.world, self.rank, self.dev, self.ctx, self.stream, self.L_
        steps = steps or args.steps
        warmup = max(warmup if warmup is not None else args.warmup, 3)
        wl = WORKLOADS[workload]
        mode, kind = wl["mode"], wl["kind"]
        streaming = kind.startswith("bestmatch")
        data = synth.generate(wl["config"], names=False, scale_n=scale_n)
        sym = data["symbols"]
        n, L = sym.shape
        pairs = n * (n - 1) // 2
        W32 = (L + 31) // 32
        c64 = ctypes.c_int64()
        RS = t.ROW_RESULT_DTYPE.itemsize
        sym_pinned = torch.from_numpy(sym).pin_memory()
        labels = [data[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
        aln = t.Alignment(ctx, sym_pinned.numpy())
        aln.set_labels(*labels)
        aln.set_config(mode, 300, True)

        # how the pair space is shared between the ranks
        #  streaming kinds: inside the library (unit-cyclic tile parts, NCCL exchange, row-sharded merge, all-gathered records)
        #  matrix kinds:    row bands (each rank owns rows [b0, b1) x all columns); no collective
        if world > 1 and not streaming:
            per = -(-n // world)
            per = -(-per // 128) * 128
            b0, b1 = min(n, rank * per), min(n, (rank + 1) * per)
            aln.set_row_range(b0, b1)
        else:
            b0, b1 = 0, n
        rows_mine = b1 - b0

        n_intra = n_inter = 0
        if "summary" in kind or "intra" in kind:
            io = t.Analysis()
            io.want = t.WANT_INTRA | (t.WANT_INTER if "summary" in kind else 0)
            t.check(L_.tdna_analyze(aln.h, ctypes.byref(io)))  # sizes of this rank's lists (they depend on the labels only)
            n_intra, n_inter = int(io.n_intra_local), int(io.n_inter_local)
        d_intra = torch.empty(max(1, n_intra), dtype=torch.float32, device=dev)
        d_inter = torch.empty(max(1, n_inter), dtype=torch.float32, device=dev)

        def step_device():
            aln.set_config(mode, 300, True)  # invalidates every cached result: each step recomputes everything
            if kind.startswith("matrix"):
                aln.matrix_compute()
                t.check(L_.tdna_class_distances(aln.h, t.PD_INTRA, d_intra.data_ptr(), n_intra, ctypes.byref(c64)))
                t.check(L_.tdna_class_distances(aln.h, t.PD_INTER, d_inter.data_ptr(), n_inter, ctypes.byref(c64)))
            elif "intra" in kind:
                # Best Close Match + the intraspecific distances for the percentile from ONE streaming pass (N > 1: the ranks share it)
                io = t.Analysis()
                io.want = t.WANT_BEST_MATCH | t.WANT_INTRA
                io.intra, io.intra_cap = d_intra.data_ptr(), n_intra
                t.check(L_.tdna_analyze(aln.h, ctypes.byref(io)))
            else:
                aln.best_match_compute()  # all n row records left in HBM (on every rank)

        for _ in range(warmup):
            step_device()
        self.barrier()
        sampler = ClockSampler(self.local)
        if rank == 0 and full and not os.environ.get("TDNA_BENCH_NOSAMPLER"):
            sampler.start()
        launches0 = ctx.launch_count()
        step_ms, tile_ms, phase_ms = [], [], []
        t_region0 = time.perf_counter()
        for _ in range(steps):
            self.flush.fill_(1)  # L2 flush between timed iterations (on torch's stream; joined below)
            torch.cuda.synchronize()
            if world > 1:
                # the untimed flush ends at different moments on different ranks; without this the ranks that finish it first would
                # time their wait for the others inside the step's first collective
                dist.barrier()
                torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            with torch.cuda.stream(stream):
                e0.record(stream)
                step_device()
                e1.record(stream)
            e1.synchronize()
            torch.cuda.synchronize()
            step_ms.append(e0.elapsed_time(e1))
            if streaming:
                tile_ms.append(ctx.last_tile_pass_ms())  # the count-tile launches of THIS step (CUDA events on the library's stream)
                if world > 1:
                    phase_ms.append(ctx.multi_phase_ms())
        self.barrier()
        t_region = time.perf_counter() - t_region0
        launches = ctx.launch_count() - launches0
        clocks = sampler.stop() if (rank == 0 and full and not os.environ.get("TDNA_BENCH_NOSAMPLER")) else None
        ms_per_step = self.max_over_ranks(float(sum(step_ms))) / steps
        value = pairs / (ms_per_step * 1e-3) / 1e9
        used_tensor = streaming and aln.last_count_kernel() == t.KERNEL_TENSOR
        pass_stats = aln.last_pass_stats() if streaming else None
        out = {"n": n, "L": L, "pairs": pairs, "mode": mode, "kind": kind, "desc": wl["desc"], "value": value, "ms_per_step": ms_per_step,
               "steps": steps, "warmup": warmup, "launches": int(launches), "clocks": clocks, "region_wall_s": t_region,
               "used_tensor": used_tensor, "rows_mine": rows_mine, "b0": b0, "b1": b1, "pass_stats": pass_stats}

        # ---- the result of the N-GPU pass against ONE GPU's, byte for byte ------------------------------------------
        if world > 1 and streaming:
            import hashlib
            p = aln.best_match_compute()
            multi = torch.as_tensor(_DevBuf(p, n * RS, "|u1"), device=dev).cpu().numpy().tobytes()
            digest = hashlib.sha256(multi).hexdigest()
            digests = [None] * world
            dist.all_gather_object(digests, digest)
            same = None
            if rank == 0:
                ctx.set_option(t.OPT_MULTI, 0)  # this rank alone, the whole triangle
                aln.set_config(mode, 300, True)
                one = aln.best_match().tobytes()
                ctx.set_option(t.OPT_MULTI, 1)
                same = bool(one == multi) and all(d == digest for d in digests)
            flag = [same]
            dist.broadcast_object_list(flag, 0)
            out["parity_vs_1gpu"] = bool(flag[0])
            out["result_sha256"] = digest
            self.barrier()

        # ---- the dominant kernel: the count-tile pass, timed live with CUDA events on the library's stream ----
        if streaming:
            kern_ms = float(np.mean(tile_ms))
            kern_pairs = pairs / world  # this rank's share of the triangle's tiles
        else:
            kern_ms = aln.time_matrix_kernel(max(3, min(20, steps)))
            kern_pairs = pairs if world == 1 else rows_mine * n  # symmetric triangle vs rectangular band
        out["kernel_ms"] = kern_ms
        out["tensor_dims"] = aln.tensor_dims() if used_tensor else 0
        if phase_ms:  # rank 0's device time per phase of the multi-GPU pass, mean over the timed steps
            out["phases_ms"] = {k: float(np.mean([p[k] for p in phase_ms if k in p])) for k in phase_ms[0]}
        if full:
            peaks, peak_src = measured_peaks()
            prof = {}
            pj = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(pj):
                prof = json.load(open(pj))
            if used_tensor:
                # kernel (b): int8 one-hot contraction on tcgen05.  Algorithmic work = 2 * K * L int8 ops per pair with the K symbol
                # dimensions per site the kernel's code uses (5 for p, 4 for K2P below 1,024 columns).  Peaks: a cuBLASLt
                # int8 GEMM timed in this run, and twice MEASURED_PEAKS.json's bf16 burst figure beside it.
                dims = aln.tensor_dims()
                ops = kern_pairs * 2 * dims * L
                twice_bf16 = 2 * peaks.get("bf16_tflops", 1590.0) * 1e12
                try:
                    int8_peak, int8_src = int8_peak_live(torch, dev), "measured live: torch._int_mm int8 8192^3, best of 10 (cuBLASLt; MEASURED_PEAKS.json has no int8 figure)"
                except Exception as ex:  # noqa: BLE001
                    int8_peak, int8_src = twice_bf16, "2 x MEASURED_PEAKS.json bf16_tflops (%s; torch._int_mm unavailable: %s)" % (peak_src, type(ex).__name__)
                achieved = ops / (kern_ms * 1e-3)
                P_bytes = 2 * n * dims * L  # a-side + b-side one-hot operands
                out["roofline"] = {
                    "bound": "tensor", "kernel": "tc::tile_kernel (tcgen05.mma kind::i8, TMEM accumulators)",
                    "achieved": achieved / 1e12, "peak": int8_peak / 1e12, "unit": "TOP/s (int8)", "frac": achieved / int8_peak,
                    "peak_source": int8_src, "peak_2x_bf16_burst": twice_bf16 / 1e12, "frac_of_2x_bf16_burst": achieved / twice_bf16,
                    "kernel_ms": kern_ms, "kernel_gpairs_per_s": kern_pairs / (kern_ms * 1e-3) / 1e9,
                    "algorithmic_int8_ops_per_pair": 2 * dims * L, "symbol_dimensions": dims,
                    "traffic": prof.get(workload + "_tensor", {}).get("dram_bytes_per_launch"),
                    "hbm": {"algorithmic_bytes_per_launch": int(P_bytes), "achieved_gbs": P_bytes / (kern_ms * 1e-3) / 1e9,
                            "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src},
                }
            else:
                popc_alg = kern_pairs * POPC_PER_PAIR_WORD[mode] * W32
                popc_peak, lop3_peak = ctx.measure_int_peaks()
                achieved = popc_alg / (kern_ms * 1e-3)
                P_planes = {0: 6, 1: 5, 2: 4}[mode]
                plane_bytes = n * P_planes * W32 * 4
                out["roofline"] = {
                    "bound": "int-issue: POPC on the XU pipe (16 lanes/clk/SM); neither HBM nor tensor bound -- see DESIGN.md",
                    "kernel": "count_tile_kernel" + ("<EPI_STREAM>" if streaming else "<EPI_MATRIX>"),
                    "achieved": achieved / 1e12, "peak": popc_peak / 1e12, "unit": "Tpopc32/s", "frac": achieved / popc_peak,
                    "peak_source": "measured live by tdna_measure_int_peaks: dependent-free POPC stream on all SMs (MEASURED_PEAKS.json has no integer-pipe figure); LOP3 stream %.2f T/s (two LOP3 per result)" % (lop3_peak / 1e12),
                    "kernel_ms": kern_ms, "kernel_gpairs_per_s": kern_pairs / (kern_ms * 1e-3) / 1e9,
                    "algorithmic_popc32_per_pair": POPC_PER_PAIR_WORD[mode] * W32,
                    "traffic": prof.get(workload, {}).get("dram_bytes_per_launch"),
                    "hbm": {"algorithmic_bytes_per_launch": int(plane_bytes if streaming else plane_bytes + rows_mine * n * 8),
                            "achieved_gbs": (plane_bytes if streaming else plane_bytes + rows_mine * n * 8) / (kern_ms * 1e-3) / 1e9,
                            "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src},
                }

        # ---- e2e: the same step through the C ABI with HOST buffers ---------------------------------
        e2e_steps = max(1, min(steps, 10))
        if kind.startswith("matrix"):
            out_host = torch.empty((rows_mine, n), dtype=torch.float64).pin_memory()
            d2h_rank = out_host.numel() * 8 + 4 * (n_intra + n_inter)
        else:
            out_host = torch.empty((n, RS), dtype=torch.uint8).pin_memory()
            d2h_rank = n * RS + 4 * n_intra
        h_intra = torch.empty(max(1, n_intra), dtype=torch.float32).pin_memory()
        h_inter = torch.empty(max(1, n_inter), dtype=torch.float32).pin_memory()
        # N > 1: every rank uploads 1 / N of the symbol rows (the library all-gathers them over NVLink) and the whole label arrays,
        # and reads back its own copy of the result
        h2d = sym.nbytes + world * 4 * 4 * n
        d2h = sum(self._gather_int(d2h_rank))

        e2e_marks = []
        one_call = streaming and kind == "bestmatch"  # tdna_analyze_host: create + labels + config + Best Match in ONE library call
        hl = t.host_list(sym_pinned.numpy(), labels, mode, 300, True) if one_call else None

        def step_e2e_one_call():
            io = t.Analysis()
            io.want = t.WANT_BEST_MATCH
            io.rows = out_host.data_ptr()
            t.check(L_.tdna_analyze_host(ctx.h, ctypes.byref(hl[0]), ctypes.byref(io), None))

        def step_e2e():
            m = [time.perf_counter()]
            a2 = t.Alignment(ctx, sym_pinned.numpy())  # H2D of the symbols (N > 1: this rank's 1/N of the rows + all-gather) + pack
            m.append(time.perf_counter())
            a2.set_labels(*labels)
            m.append(time.perf_counter())
            a2.set_config(mode, 300, True)
            if world > 1 and not streaming:
                a2.set_row_range(b0, b1)
            m.append(time.perf_counter())
            if kind.startswith("matrix"):
                t.check(L_.tdna_matrix(a2.h, out_host.data_ptr(), None))
                t.check(L_.tdna_class_distances(a2.h, t.PD_INTRA, h_intra.data_ptr(), n_intra, ctypes.byref(c64)))
                t.check(L_.tdna_class_distances(a2.h, t.PD_INTER, h_inter.data_ptr(), n_inter, ctypes.byref(c64)))
            elif "intra" in kind:
                io = t.Analysis()
                io.want = t.WANT_BEST_MATCH | t.WANT_INTRA
                io.rows = out_host.data_ptr()
                io.intra, io.intra_cap = h_intra.data_ptr(), n_intra
                t.check(L_.tdna_analyze(a2.h, ctypes.byref(io)))
            else:
                a2.best_match_into(out_host.data_ptr())
            m.append(time.perf_counter())
            a2.close()
            m.append(time.perf_counter())
            e2e_marks.append([1e3 * (y - x) for x, y in zip(m, m[1:])])

        for _ in range(2):  # untimed: the first calls grow the library's memory pool and its buffer-size hints
            step_e2e()
        e2e_marks.clear()
        self.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            step_e2e()
        self.barrier()
        e2e_s = self.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
        sep = {"value": pairs / e2e_s / 1e9, "unit": "Gpairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": e2e_s * 1e3, "steps": e2e_steps, "api": "tdna_alignment_create + _set_labels + tdna_set_config + analysis call + _destroy",
               "host_phases_ms_this_rank": dict(zip(("create(h2d+pack)", "labels", "config", "analysis(+d2h)", "destroy"),
                                                    [float(x) for x in np.mean(np.array(e2e_marks), axis=0)]))}
        out["e2e"] = sep
        if one_call:
            # the same step as ONE call (the upload pipelined with the pass on one GPU); checked against the separate calls' bytes
            check = out_host.numpy().tobytes()
            step_e2e_one_call()
            same = out_host.numpy().tobytes() == check
            step_e2e_one_call()
            self.barrier()
            one_steps = int(os.environ.get("TDNA_BENCH_E2E_STEPS", max(e2e_steps, steps)))  # all K steps: a call is a few milliseconds
            per_call = []
            t0 = time.perf_counter()
            for _ in range(one_steps):
                t1 = time.perf_counter()
                step_e2e_one_call()
                per_call.append((time.perf_counter() - t1) * 1e3)
                if os.environ.get("TDNA_BENCH_DEBUG") and per_call[-1] > 4.5:
                    print("slow one-call step: %.2f ms %s tile pass %.3f ms" % (per_call[-1], ctx.host_call_phase_ms(), ctx.last_tile_pass_ms()), file=sys.stderr, flush=True)
            self.barrier()
            s1 = self.max_over_ranks((time.perf_counter() - t0) / one_steps)
            out["e2e"] = {"value": pairs / s1 / 1e9, "unit": "Gpairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                          "ms_per_step": s1 * 1e3, "steps": one_steps, "api": "tdna_analyze_host (one call: host list in, row records out)",
                          "ms_per_call_this_rank": {"min": min(per_call), "median": sorted(per_call)[len(per_call) // 2], "max": max(per_call)},
                          "same_bytes_as_separate_calls": bool(same), "separate_calls": sep}

        out["cpu_baseline"] = None
        if full and rank == 0 and world == 1 and not args.no_cpu_baseline:
            import shutil
            threads = os.cpu_count() or 1
            v, cnt, dt, rows = cpu_reference_leg(data, mode, args.cpu_seconds, threads)
            v1, cnt1, dt1, rows1 = cpu_reference_leg(data, mode, min(4.0, args.cpu_seconds), 1)
            out["cpu_baseline"] = {
                "value": v / 1e9, "unit": "Gpairs/s", "cores": threads, "kind": "port",
                "sample": "first %d query rows x all later columns of the same alignment: %d pairs in %.1f s; C restatement of the reference's Sequence.getPairwiseNoBuffer (oracle/tdna_oracle.c)" % (rows, cnt, dt),
                "single_thread_value": v1 / 1e9,
                # the reference is Java; its sources cannot travel to this box (they may not be copied into the repo), so even a JVM
                # found here could not run it -- recorded for the record
                "jvm": shutil.which("java") or "absent"}
        # torch releases pinned blocks that were used on the library's stream by recording an event on that stream: let them go
        # while the stream still exists
        out_host = h_intra = h_inter = sym_pinned = None  # noqa: F841
        gc.collect()
        torch.cuda.synchronize()
        aln.close()
        return out

    def _gather_int(self, v):
        if self.world == 1:
            return [v]
        vals = [None] * self.world
        self.dist.all_gather_object(vals, int(v))
        return vals


def main():
    args = parse()
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ and args.impl != "reference":
        # asked for N GPUs outside torchrun: become the launch the driver uses (one rank per GPU over NCCL, rendezvous on 127.0.0.1)
        os.execvp(sys.executable, [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
                                   "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 500),
                                   os.path.abspath(__file__)] + sys.argv[1:])
    # the contract is ONE JSON line on stdout: keep the real stdout for it and send everything else
    # (NCCL's version banner, library chatter) to stderr
    global _OUT
    _OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        return run_reference(args)

    from taxondna_b200 import synth
    b = Bench(args)
    world, rank = b.world, b.rank
    wl = WORKLOADS[args.workload]
    streaming = wl["kind"].startswith("bestmatch")
    base = synth.CONFIGS[wl["config"]]
    n1 = base["genera"] * base["species"] * base["reps"]
    # the headline run.  --scaling strong (default): the named shape itself on N GPUs; weak: the alignment grows as sqrt(N) so that
    # the pairs per GPU stay fixed
    weak_n = n1 * (world ** 0.5) if world > 1 else None
    r = b.measure(args.workload, scale_n=weak_n if (args.scaling == "weak" and world > 1) else None)
    extras = {}
    if world > 1 and streaming and not args.no_extras:
        # the other scaling mode of the same workload as an extra key (value + e2e + parity only)
        other = "weak" if args.scaling == "strong" else "strong"
        ro = b.measure(args.workload, scale_n=weak_n if other == "weak" else None, steps=max(3, min(args.steps, 10)), warmup=3, full=False)
        extras[other + "_scaling"] = {"n": ro["n"], "pairs": ro["pairs"], "value": ro["value"], "unit": "Gpairs/s", "ms_per_step": ro["ms_per_step"],
                                      "kernel_ms": ro["kernel_ms"], "e2e": ro["e2e"], "parity_vs_1gpu": ro.get("parity_vs_1gpu"),
                                      "note": "pairs per GPU fixed (n grows as sqrt(N))" if other == "weak" else "the named shape itself on N GPUs"}
    if world == 1 and args.workload == "H" and not args.no_extras:
        # the headline shape on data that is NOT best case, as extra keys of the same line: C4's noise (20 % '?' / '-' / IUPAC letters: the
        # GENERAL instantiation of count kernel (b)) in p and K2P, and transversions only (the exact six-dimension code)
        peak = (r.get("roofline") or {}).get("peak")
        ow = {}
        for name in ("Hg", "HgK", "Ht"):
            ro = b.measure(name, steps=max(3, min(args.steps, 5)), warmup=3, full=False)
            e = {"workload": ro["desc"], "value": ro["value"], "unit": "Gpairs/s", "ms_per_step": ro["ms_per_step"], "kernel_ms": ro["kernel_ms"],
                 "count_kernel": "tensor" if ro["used_tensor"] else "popc", "symbol_dimensions": ro["tensor_dims"], "e2e": ro["e2e"]["value"],
                 "pass_stats": ro.get("pass_stats")}
            if peak and ro["used_tensor"]:
                e["roofline_frac"] = ro["pairs"] * 2 * ro["tensor_dims"] * ro["L"] / (ro["kernel_ms"] * 1e-3) / (peak * 1e12)
            ow[name] = e
        extras["other_workloads"] = ow
    if rank == 0:
        if streaming:
            par = "single GPU, whole triangle"
            if world > 1:
                par = ("inside the C ABI (tdna_comm_init): every %d-th unit of 128 tile slots of the triangle per rank; NCCL on the library's stream: "
                       "all_reduce(min) of the seeded per-row minima, ONE grouped send/recv that carries the candidate records, the per-row state and the status "
                       "words to the rank that owns the row, all_gather of the row records each rank finalised; one host sync per step" % world)
        else:
            par = "row bands, %d rank(s), rows [%d,%d) on rank 0" % (world, r["b0"], r["b1"])
        n, L, mode = r["n"], r["L"], r["mode"]
        line = {
            "metric": "pairwise distances/s", "value": r["value"], "unit": "Gpairs/s", "n_gpus": world, "steps": r["steps"],
            "warmup": r["warmup"], "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": ("int8 one-hot -> int32 counts (tcgen05) -> f64 distance" if r["used_tensor"] else "u32 bit-planes -> int32 counts -> f64 distance"), "data": "synthetic",
            "config": {"workload": "%s x %s bp: %s" % (format(n, ","), format(L, ","), r["desc"]), "n": n, "L": L,
                       "count_kernel": ("tensor" if r["used_tensor"] else "popc"), "mode": ["uncorrected", "K2P", "trans-only"][mode], "min_overlap": 300,
                       "pairs": r["pairs"], "step": r["kind"], "l2": "flushed between timed steps (256 MiB write)", "parallelism": par},
            "roofline": r.get("roofline"), "cpu_baseline": r["cpu_baseline"], "e2e": r["e2e"], "gpu_launches": r["launches"], "clocks": r["clocks"],
            "region_wall_s": r["region_wall_s"],
        }
        if r.get("pass_stats"):
            line["config"]["pass_stats_rank0"] = r["pass_stats"]
        if "phases_ms" in r:
            line["phases_ms_rank0"] = r["phases_ms"]
        if "parity_vs_1gpu" in r:
            line["parity_vs_1gpu"] = r["parity_vs_1gpu"]
            line["result_sha256"] = r["result_sha256"]
        line.update(extras)
        print(json.dumps(line), file=_OUT, flush=True)
    bad = ("parity_vs_1gpu" in r and not r["parity_vs_1gpu"]) or any(v.get("parity_vs_1gpu") is False for v in extras.values())
    b.close()
    if bad:
        raise SystemExit("the N-GPU result differs from one GPU's")


if __name__ == "__main__":
    main()
