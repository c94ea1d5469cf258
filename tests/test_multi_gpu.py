"""The multi-GPU pass behind the C ABI (tdna_comm_init / tdna_init_multi; include/taxondna_cuda.h "several GPUs of one box").

On a one-GPU box the whole choreography -- seed, MIN all-reduce, bulk, device-driven routing, grouped send / recv, SUM all-reduce of
the state, row-sharded merge, all-gather -- runs through a ONE-rank NCCL communicator (TDNA_OPT_MULTI = 2) and must reproduce the
single-GPU bytes.  With two or more devices visible the same is checked across real ranks: as processes (torchrun, the bench's
layout) and as threads of one process (tdna_init_multi, the one-JVM layout)."""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def device_count():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("cfg,mode,scale_n", [("C3", 0, None), ("C3", 1, None), ("Hg", 0, 5000), ("Hg", 1, 5000), ("tiny", 0, None)])
def test_one_rank_communicator_reproduces_the_single_gpu_bytes(cfg, mode, scale_n):
    import taxondna_b200 as t
    from taxondna_b200 import synth
    ctx = t.Context(0)
    try:
        ctx.comm_init(t.Context.comm_unique_id(), 0, 1)
        data = synth.generate(cfg, names=False, scale_n=scale_n)
        aln = t.Alignment(ctx, data["symbols"])
        aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
        aln.set_config(mode, 300, True)
        one = aln.analyze(best_match=True, intra=True, inter=True, edges=0.01)
        ctx.set_option(t.OPT_MULTI, 2)  # the collective path although there is one rank
        aln.set_config(mode, 300, True)
        multi = aln.analyze(best_match=True, intra=True, inter=True, edges=0.01)
        assert multi["rows"].tobytes() == one["rows"].tobytes()
        for k in ("intra", "inter"):
            assert np.array_equal(np.sort(multi[k].view(np.uint32)), np.sort(one[k].view(np.uint32)))
        assert sorted(zip(multi["edge_i"].tolist(), multi["edge_j"].tolist(), multi["edge_d"].view(np.int64).tolist())) == \
            sorted(zip(one["edge_i"].tolist(), one["edge_j"].tolist(), one["edge_d"].view(np.int64).tolist()))
        # both count kernels, and the dense fallback of the collective path
        for kern in (t.KERNEL_POPC, t.KERNEL_TENSOR):
            ctx.set_option(t.OPT_COUNT_KERNEL, kern)
            aln.set_config(mode, 300, True)
            assert aln.best_match().tobytes() == one["rows"].tobytes()
        ctx.set_option(t.OPT_COUNT_KERNEL, t.KERNEL_AUTO)
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
        aln.set_config(mode, 300, True)
        assert aln.best_match().tobytes() == one["rows"].tobytes()
        aln.close()
    finally:
        ctx.close()


def test_overflowing_segments_are_regrown_by_every_rank_alike():
    """One species of 3,000 rows: every conspecific is a candidate, far more than the default exchange segment holds."""
    import taxondna_b200 as t
    rng = np.random.default_rng(5)
    root = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=400)
    sym = np.repeat(root[None, :], 3000, axis=0).copy()
    m = rng.random(sym.shape) < 0.02
    sym[m] = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=int(m.sum()))
    sp = np.zeros(3000, dtype=np.int32)
    sp[2000:] = 1
    ctx = t.Context(0)
    try:
        ctx.comm_init(t.Context.comm_unique_id(), 0, 1)
        aln = t.Alignment(ctx, sym)
        aln.set_labels(sp, sp, sp, np.arange(3000, dtype=np.int32))
        aln.set_config(0, 100, True)
        one = aln.best_match()
        ctx.set_option(t.OPT_MULTI, 2)
        aln.set_config(0, 100, True)
        assert aln.best_match().tobytes() == one.tobytes()
        aln.close()
    finally:
        ctx.close()


def test_two_ranks_as_processes():
    if device_count() < 2:
        pytest.skip("needs two GPUs")
    env = dict(os.environ)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29611", os.path.join(ROOT, "tests", "multi_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MULTI_WORKER OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]


def test_two_ranks_as_threads_of_one_process():
    """The one-JVM layout: tdna_init_multi makes the contexts and their communicator; one thread per context makes the calls."""
    if device_count() < 2:
        pytest.skip("needs two GPUs")
    import taxondna_b200 as t
    from taxondna_b200 import synth
    data = synth.generate("C3", names=False)
    ctxs = t.Context.init_multi([0, 1])
    out, errs = [None, None], []

    def work(k):
        try:
            c = ctxs[k]
            aln = t.Alignment(c, data["symbols"])
            aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
            aln.set_config(1, 300, True)
            out[k] = aln.best_match().tobytes()
            aln.close()
        except Exception as e:  # noqa: BLE001
            errs.append(repr(e))

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join(timeout=600)
    assert not errs, errs
    ctxs[0].set_option(t.OPT_MULTI, 0)
    aln = t.Alignment(ctxs[0], data["symbols"])
    aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
    aln.set_config(1, 300, True)
    one = aln.best_match().tobytes()
    aln.close()
    assert out[0] == one and out[1] == one
    for c in ctxs:
        c.close()
