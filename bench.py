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

N > 1 (torchrun, one rank per GPU): every rank holds the whole packed alignment.  Best-Match steps
share the triangle's tile list in N contiguous slices (tdna_best_match_seed / _part_packed) and
exchange their candidate records / per-row counts once per step over NCCL, then merge; matrix steps
split the query rows into bands (no collective).  Weak scaling: the alignment grows as sqrt(N) so
pairs per GPU stay fixed.

--impl reference: the reference's CPU algorithm (oracle/ C restatement of Sequence.getPairwise --
the reference is Java and no JVM exists in this image) on the host cores, bounded sample.
"""
import argparse
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
    "C2": dict(config="C2", mode=0, kind="matrix+summary", desc="2,185 x 2,664 bp synthetic multi-gene alignment, uncorrected p: full pairwise matrix + Pairwise Summary classes"),
    "C3": dict(config="C3", mode=1, kind="bestmatch+intra", desc="8,000 x 658 bp synthetic COI, K2P: Best Close Match + intraspecific distances for the 95th-percentile threshold"),
    "C4": dict(config="C4", mode=0, kind="matrix+summary", desc="5,000 x 2,664 bp gappy/ambiguous alignment (20% ?/-/IUPAC), uncorrected p: matrix + Pairwise Summary classes"),
    "H": dict(config="H", mode=0, kind="bestmatch", desc="50,000 x 658 bp synthetic COI, uncorrected p: streaming per-query Best Match summaries"),
}
POPC_PER_PAIR_WORD = {0: 2, 1: 4, 2: 2}  # C = 2 (p), 4 (K2P), 2 (trans-only)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="H", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
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
        "config": {"workload": wl["desc"], "n": n, "L": L, "mode": wl["mode"]},
        "cpu_baseline": {"value": v, "unit": "Gpairs/s", "cores": threads, "kind": "port",
                         "sample": "first %d query rows x all later columns per step (%d pairs/step), C restatement of Sequence.getPairwiseNoBuffer (no JVM in image), %d pthreads" % (rows, pairs // max(1, args.steps), threads)},
        "e2e": {"value": v, "unit": "Gpairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_OUT, flush=True)


class _DevBuf:
    """Zero-copy torch view of library-owned device memory (__cuda_array_interface__)."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": typestr, "data": (ptr, False), "version": 2}


def exchange_parts(torch, dist, dev, part, n, world):
    """The multi-GPU exchange step of the streaming path (taxondna_b200.shard.exchange_candidates) on
    zero-copy views of the library's device buffers."""
    from taxondna_b200 import shard
    k = int(part.n_cand)
    if k:
        rows = torch.as_tensor(_DevBuf(part.cand_row, k, "<i4"), device=dev)
        cols = torch.as_tensor(_DevBuf(part.cand_col, k, "<i4"), device=dev)
        d = torch.as_tensor(_DevBuf(part.cand_d, k, "<f8"), device=dev)
    else:
        rows = torch.zeros(0, dtype=torch.int32, device=dev)
        cols = torch.zeros(0, dtype=torch.int32, device=dev)
        d = torch.zeros(0, dtype=torch.float64, device=dev)
    inval = torch.as_tensor(_DevBuf(part.invalid_count, n, "<i4"), device=dev)
    flags = torch.as_tensor(_DevBuf(part.flags, n, "<i4"), device=dev)
    return shard.exchange_candidates(rows, cols, d, inval, flags)


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

    import ctypes

    import torch
    import torch.distributed as dist

    import taxondna_b200 as t
    from taxondna_b200 import synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    wl = WORKLOADS[args.workload]
    mode = wl["mode"]
    kind = wl["kind"]
    streaming = kind.startswith("bestmatch")
    base = synth.CONFIGS[wl["config"]]
    n1 = base["genera"] * base["species"] * base["reps"]
    scale_n = None
    if world > 1 and args.scaling == "weak":
        scale_n = n1 * (world ** 0.5)  # pairs per GPU stay fixed
    data = synth.generate(wl["config"], names=False, scale_n=scale_n)
    sym = data["symbols"]
    n, L = sym.shape
    pairs = n * (n - 1) // 2
    W32 = (L + 31) // 32

    ctx = t.Context(local)
    ctx.set_option(t.OPT_COUNT_KERNEL, {"auto": t.KERNEL_AUTO, "popc": t.KERNEL_POPC, "tensor": t.KERNEL_TENSOR}[args.kernel])
    stream = torch.cuda.ExternalStream(ctx.stream_ptr(), device=dev)
    sym_pinned = torch.from_numpy(sym).pin_memory()
    labels = [data[k] for k in ("species_id", "genus_id", "epithet_rank", "fullname_rank")]
    aln = t.Alignment(ctx, sym_pinned.numpy())
    aln.set_labels(*labels)
    aln.set_config(mode, 300, True)
    L_ = t.load()
    c64 = ctypes.c_int64()
    RS = t.ROW_RESULT_DTYPE.itemsize

    # ---- how the pair space is shared between ranks ------------------------------------------
    #  streaming kinds: every rank takes every world-th tile of the triangle (tdna_best_match_part),
    #                   then one NCCL exchange of the candidate records / per-row counts, then merge.
    #  matrix kinds:    row bands (each rank owns rows [b0,b1) x all columns); no collective.
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
        L_.tdna_class_distances(aln.h, t.PD_INTRA, None, 0, ctypes.byref(c64))  # sizes depend on the labels only
        n_intra = c64.value
        if "summary" in kind:
            L_.tdna_class_distances(aln.h, t.PD_INTER, None, 0, ctypes.byref(c64))
            n_inter = c64.value
    d_intra = torch.empty(max(1, n_intra), dtype=torch.float32, device=dev)
    d_inter = torch.empty(max(1, n_inter), dtype=torch.float32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    tile_ms = []
    cand_counts = []
    # exchange buffers of the multi-GPU streaming path: one segment of `seg` packed records per rank
    # Exchange of the candidate records at N > 1 (TDNA_BENCH_EXCHANGE overrides the choice):
    #  "gathered": one segment per rank, all_gather, every rank merges all n rows            (least latency: 2 GPUs)
    #  "routed":   one segment per (source, destination), all_to_all to the rank that owns the row, every rank finalises its own
    #              rows_per_part rows, all_gather of the row records                          (least volume and merge work: 4+ GPUs)
    exchange = os.environ.get("TDNA_BENCH_EXCHANGE", "gathered" if world <= 2 else "routed")
    routed = exchange == "routed"
    seg = max(1 << 14, (48 * n) // max(world * world, 1)) if routed else max(1 << 16, (32 * n) // max(world, 1))
    xb = None
    if world > 1 and streaming and routed:
        rpp = aln.rows_per_part(world)
        xb = {"rec": torch.empty((world, seg, 2), dtype=torch.int64, device=dev), "counts": torch.zeros(world, dtype=torch.int64, device=dev),
              "recv": torch.empty((world, seg, 2), dtype=torch.int64, device=dev), "recv_counts": torch.zeros(world, dtype=torch.int64, device=dev),
              "state": torch.empty(n, dtype=torch.int64, device=dev),
              "slice": torch.empty(rpp * RS, dtype=torch.uint8, device=dev), "rows": torch.empty(world * rpp * RS, dtype=torch.uint8, device=dev)}
    elif world > 1 and streaming:
        xb = {"rec": torch.empty((seg, 2), dtype=torch.int64, device=dev), "count": torch.zeros(1, dtype=torch.int64, device=dev),
              "counts": torch.zeros(world, dtype=torch.int64, device=dev), "state": torch.empty(n, dtype=torch.int64, device=dev),
              "allrec": torch.empty((world, seg, 2), dtype=torch.int64, device=dev)}

    def best_match_step(a, out=None):
        """Per-query Best Match summaries for ALL rows, left on the device (or copied to the pinned host tensor `out`)."""
        out_ptr = out.data_ptr() if out is not None else None
        if world == 1:
            if out_ptr is None:
                a.best_match_compute()
            else:
                a.best_match_into(out_ptr)
            return
        mins = a.best_match_seed(rank, world)  # this rank's diagonal blocks -> per-row minima
        with torch.cuda.stream(stream):
            mt = torch.as_tensor(_DevBuf(mins, 3 * n, "<i8"), device=dev)
            dist.all_reduce(mt, op=dist.ReduceOp.MIN)  # thresholds from ALL list neighbours on every rank
            stream.synchronize()
        if not routed:
            # the part writes packed records / count / per-row state straight into these buffers; NCCL takes them as they are
            a.best_match_part_packed(rank, world, True, xb["rec"].data_ptr(), seg, xb["count"].data_ptr(), xb["state"].data_ptr())
            with torch.cuda.stream(stream):
                dist.all_gather_into_tensor(xb["counts"], xb["count"])
                dist.all_gather_into_tensor(xb["allrec"].view(-1), xb["rec"].view(-1))
                dist.all_reduce(xb["state"], op=dist.ReduceOp.SUM)
                stream.synchronize()
            a.best_match_merge_gathered(xb["allrec"].data_ptr(), world, seg, xb["counts"].data_ptr(), xb["state"].data_ptr(), out_ptr)
            return
        a.best_match_part_routed(rank, world, True, xb["rec"].data_ptr(), seg, xb["counts"].data_ptr(), xb["state"].data_ptr())
        with torch.cuda.stream(stream):
            dist.all_to_all_single(xb["recv_counts"], xb["counts"])
            dist.all_to_all_single(xb["recv"].view(-1), xb["rec"].view(-1))
            dist.all_reduce(xb["state"], op=dist.ReduceOp.SUM)
            stream.synchronize()
        a.best_match_merge_routed(rank, world, xb["recv"].data_ptr(), seg, xb["recv_counts"].data_ptr(), xb["state"].data_ptr(), xb["slice"].data_ptr())
        with torch.cuda.stream(stream):
            dist.all_gather_into_tensor(xb["rows"], xb["slice"])  # every rank holds all n row records (120 B each)
            if out is not None:
                out.view(torch.uint8).view(-1)[: n * RS].copy_(xb["rows"][: n * RS], non_blocking=True)
            stream.synchronize()

    def step_device():
        aln.set_config(mode, 300, True)  # invalidates every cached result: each step recomputes everything
        if kind.startswith("matrix"):
            aln.matrix_compute()
            t.check(L_.tdna_class_distances(aln.h, t.PD_INTRA, d_intra.data_ptr(), n_intra, ctypes.byref(c64)))
            t.check(L_.tdna_class_distances(aln.h, t.PD_INTER, d_inter.data_ptr(), n_inter, ctypes.byref(c64)))
        elif "intra" in kind and world == 1:
            # Best Close Match + the intraspecific distances for the percentile from ONE streaming pass
            io = t.Analysis()
            io.want = t.WANT_BEST_MATCH | t.WANT_INTRA
            io.intra, io.intra_cap = d_intra.data_ptr(), n_intra
            t.check(L_.tdna_analyze(aln.h, ctypes.byref(io)))
        else:
            best_match_step(aln)
            if "intra" in kind:
                t.check(L_.tdna_class_distances(aln.h, t.PD_INTRA, d_intra.data_ptr(), n_intra, ctypes.byref(c64)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launch_count()
    step_ms = []
    t_region0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)  # L2 flush between timed iterations (on torch's stream; joined below)
        torch.cuda.synchronize()
        if world > 1:
            # the untimed flush ends at different moments on different ranks; without this the ranks that finish it first would time
            # their wait for the others inside the step's first collective
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
            tile_ms.append(ctx.last_tile_pass_ms())  # the count-tile pass of THIS step (CUDA events on the library's stream)
    barrier()
    if xb is not None:
        cand_counts.append(int(xb["counts"].sum().item()) if routed else int(xb["count"].item()))
    t_region = time.perf_counter() - t_region0
    launches = ctx.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    total_ms = float(sum(step_ms))
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    total_ms = float(tm.item())
    ms_per_step = total_ms / args.steps
    value = pairs / (ms_per_step * 1e-3) / 1e9

    # ---- the dominant kernel: the count-tile pass, timed live with CUDA events on the library's stream ----
    if streaming:
        kern_ms = float(np.mean(tile_ms))
        kern_pairs = pairs / world  # every world-th tile of the triangle
    else:
        kern_ms = aln.time_matrix_kernel(max(3, min(20, args.steps)))
        kern_pairs = pairs if world == 1 else rows_mine * n  # symmetric triangle vs rectangular band
    peaks, peak_src = measured_peaks()
    prof = {}
    pj = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(pj):
        prof = json.load(open(pj))
    used_tensor = streaming and aln.last_count_kernel() == t.KERNEL_TENSOR
    if used_tensor:
        # kernel (b): int8 one-hot contraction on tcgen05.  Algorithmic work = 2 * K * L int8 ops per pair with the K = 5 (p) or
        # 4 (K2P) symbol dimensions per site this kernel uses; peak = a cuBLASLt int8 GEMM (torch._int_mm, 8192^3)
        # measured here, else twice MEASURED_PEAKS.json's bf16 burst figure.
        dims = 4 if mode == 1 else 5  # K2P ignores gap sites: four symbol dimensions
        ops = kern_pairs * 2 * dims * L
        int8_peak, int8_src = None, None
        try:
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
            int8_peak, int8_src = 2 * 8192 ** 3 / (best * 1e-3), "measured live: torch._int_mm int8 8192^3, best of 10 (cuBLASLt; MEASURED_PEAKS.json has no int8 figure)"
            del a8, b8
        except Exception as ex:  # noqa: BLE001
            int8_peak, int8_src = 2 * peaks.get("bf16_tflops", 1590.0) * 1e12, "2 x MEASURED_PEAKS.json bf16_tflops (%s; torch._int_mm unavailable: %s)" % (peak_src, type(ex).__name__)
        achieved = ops / (kern_ms * 1e-3)
        P_bytes = 2 * n * dims * L  # a-side + b-side one-hot operands
        roofline = {
            "bound": "tensor", "kernel": "tc::tile_kernel (tcgen05.mma kind::i8, TMEM accumulators)",
            "achieved": achieved / 1e12, "peak": int8_peak / 1e12, "unit": "TOP/s (int8)", "frac": achieved / int8_peak,
            "peak_source": int8_src, "kernel_ms": kern_ms, "kernel_gpairs_per_s": kern_pairs / (kern_ms * 1e-3) / 1e9,
            "algorithmic_int8_ops_per_pair": 2 * dims * L,
            "traffic": prof.get(args.workload + "_tensor", {}).get("dram_bytes_per_launch"),
            "hbm": {"algorithmic_bytes_per_launch": int(P_bytes), "achieved_gbs": P_bytes / (kern_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src},
        }
    else:
        popc_alg = kern_pairs * POPC_PER_PAIR_WORD[mode] * W32
        popc_peak, lop3_peak = ctx.measure_int_peaks()
        achieved = popc_alg / (kern_ms * 1e-3)
        P_planes = {0: 6, 1: 5, 2: 4}[mode]
        plane_bytes = n * P_planes * W32 * 4
        roofline = {
            "bound": "int-issue: POPC on the XU pipe (16 lanes/clk/SM); neither HBM nor tensor bound -- see DESIGN.md",
            "kernel": "count_tile_kernel" + ("<EPI_STREAM>" if streaming else "<EPI_MATRIX>"),
            "achieved": achieved / 1e12, "peak": popc_peak / 1e12, "unit": "Tpopc32/s", "frac": achieved / popc_peak,
            "peak_source": "measured live by tdna_measure_int_peaks: dependent-free POPC stream on all SMs (MEASURED_PEAKS.json has no integer-pipe figure); LOP3 stream %.2f T/s (two LOP3 per result)" % (lop3_peak / 1e12),
            "kernel_ms": kern_ms, "kernel_gpairs_per_s": kern_pairs / (kern_ms * 1e-3) / 1e9,
            "algorithmic_popc32_per_pair": POPC_PER_PAIR_WORD[mode] * W32,
            "traffic": prof.get(args.workload, {}).get("dram_bytes_per_launch"),
            "hbm": {"algorithmic_bytes_per_launch": int(plane_bytes if streaming else plane_bytes + rows_mine * n * 8),
                    "achieved_gbs": (plane_bytes if streaming else plane_bytes + rows_mine * n * 8) / (kern_ms * 1e-3) / 1e9,
                    "peak_gbs": peaks.get("hbm_gbs"), "peak_source": peak_src},
        }

    # ---- e2e: the same step through the C ABI with HOST buffers ---------------------------------
    e2e_steps = max(1, min(args.steps, 5))
    h2d = sym.nbytes + 4 * 4 * n
    if kind.startswith("matrix"):
        out_host = torch.empty((rows_mine, n), dtype=torch.float64).pin_memory()
        d2h = out_host.numel() * 8 + 4 * (n_intra + n_inter)
    else:
        out_host = torch.empty((n, RS), dtype=torch.uint8).pin_memory()
        d2h = n * RS + 4 * n_intra
    h_intra = torch.empty(max(1, n_intra), dtype=torch.float32).pin_memory()
    h_inter = torch.empty(max(1, n_inter), dtype=torch.float32).pin_memory()

    def step_e2e():
        a2 = t.Alignment(ctx, sym_pinned.numpy())  # H2D of the symbols + pack
        a2.set_labels(*labels)
        a2.set_config(mode, 300, True)
        if world > 1 and not streaming:
            a2.set_row_range(b0, b1)
        if kind.startswith("matrix"):
            t.check(L_.tdna_matrix(a2.h, out_host.data_ptr(), None))
            t.check(L_.tdna_class_distances(a2.h, t.PD_INTRA, h_intra.data_ptr(), n_intra, ctypes.byref(c64)))
            t.check(L_.tdna_class_distances(a2.h, t.PD_INTER, h_inter.data_ptr(), n_inter, ctypes.byref(c64)))
        elif "intra" in kind and world == 1:
            io = t.Analysis()
            io.want = t.WANT_BEST_MATCH | t.WANT_INTRA
            io.rows = out_host.data_ptr()
            io.intra, io.intra_cap = h_intra.data_ptr(), n_intra
            t.check(L_.tdna_analyze(a2.h, ctypes.byref(io)))
        else:
            best_match_step(a2, out_host)
            if "intra" in kind:
                t.check(L_.tdna_class_distances(a2.h, t.PD_INTRA, h_intra.data_ptr(), n_intra, ctypes.byref(c64)))
        a2.close()

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        step_e2e()
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    e2e = {"value": pairs / e2e_s / 1e9, "unit": "Gpairs/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
           "ms_per_step": e2e_s * 1e3, "steps": e2e_steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        v, cnt, dt, rows = cpu_reference_leg(data, mode, args.cpu_seconds, threads)
        v1, cnt1, dt1, rows1 = cpu_reference_leg(data, mode, min(4.0, args.cpu_seconds), 1)
        cpu = {"value": v / 1e9, "unit": "Gpairs/s", "cores": threads, "kind": "port",
               "sample": "first %d query rows x all later columns of the same alignment: %d pairs in %.1f s; C restatement of the reference's Sequence.getPairwiseNoBuffer (oracle/tdna_oracle.c; the reference is Java and no JVM exists in the image)" % (rows, cnt, dt),
               "single_thread_value": v1 / 1e9}

    if rank == 0:
        if streaming:
            par = "single GPU, whole triangle"
            if world > 1 and routed:
                par = "every %d-th unit of 128 tile slots of the triangle per rank; NCCL: all_reduce(min) of the seeded minima, all_to_all of the candidate records (to the rank that owns the row), all_reduce(sum) of the per-row state, all_gather of the row records each rank finalised" % world
            elif world > 1:
                par = "every %d-th unit of 128 tile slots of the triangle per rank; NCCL: all_reduce(min) of the seeded minima, all_gather of the candidate records, all_reduce(sum) of the per-row state; merge on every rank" % world
        else:
            par = "row bands, %d rank(s), rows [%d,%d) on rank 0" % (world, b0, b1)
        line = {
            "metric": "pairwise distances/s", "value": value, "unit": "Gpairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": ("int8 one-hot -> int32 counts (tcgen05) -> f64 distance" if used_tensor else "u32 bit-planes -> int32 counts -> f64 distance"), "data": "synthetic",
            "config": {"workload": wl["desc"], "n": n, "L": L, "count_kernel": ("tensor" if used_tensor else "popc"), "mode": ["uncorrected", "K2P", "trans-only"][mode], "min_overlap": 300,
                       "pairs": pairs, "step": kind, "l2": "flushed between timed steps (256 MiB write)", "parallelism": par},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "region_wall_s": t_region,
        }
        if cand_counts:
            line["config"]["candidates_per_rank_per_step"] = int(np.mean(cand_counts))
        print(json.dumps(line), file=_OUT, flush=True)
    # torch releases pinned blocks that were used on the library's stream by recording an event on that stream: let them go
    # while the stream still exists
    out_host = h_intra = h_inter = xb = None  # noqa: F841
    gc.collect()
    torch.cuda.synchronize()
    aln.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
