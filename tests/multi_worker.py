"""Worker of tests/test_multi_gpu.py: one rank of a torchrun launch (one process per GPU).  Every data-plane step happens inside
the C ABI (tdna_comm_init); torch.distributed only carries the communicator id and the verdicts."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist

    import taxondna_b200 as t
    from taxondna_b200 import synth
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = t.Context(local)
    ctx.comm_init_torch(dist, dev)
    assert ctx.comm_info() == (rank, world)
    failures = []
    cases = [("C3", 0, None), ("C3", 1, None), ("H", 0, 12000), ("Hg", 0, 6000), ("Hg", 1, 6000)]
    for cfg, mode, scale_n in cases:
        data = synth.generate(cfg, names=False, scale_n=scale_n)
        sym = data["symbols"]
        n = sym.shape[0]
        aln = t.Alignment(ctx, sym)  # N > 1: sharded upload + all-gather inside the call
        aln.set_labels(data["species_id"], data["genus_id"], data["epithet_rank"], data["fullname_rank"])
        aln.set_config(mode, 300, True)
        multi = aln.analyze(best_match=True, intra=True, inter=True, edges=0.01)
        ctx.set_option(t.OPT_MULTI, 0)
        aln.set_config(mode, 300, True)
        one = aln.analyze(best_match=True, intra=True, inter=True, edges=0.01)
        ctx.set_option(t.OPT_MULTI, 1)
        if multi["rows"].tobytes() != one["rows"].tobytes():
            failures.append("%s mode %d: rank %d rows differ from one GPU's" % (cfg, mode, rank))
        # variable-length outputs: this rank's share; the union over the ranks is the one-GPU multiset
        for key, dt in (("intra", np.float32), ("inter", np.float32)):
            parts = [None] * world
            dist.all_gather_object(parts, multi[key].tobytes())
            allv = np.sort(np.concatenate([np.frombuffer(p, dtype=dt) for p in parts]).view(np.uint32))
            if not np.array_equal(allv, np.sort(one[key].view(np.uint32))):
                failures.append("%s mode %d: %s multiset differs (%d vs %d)" % (cfg, mode, key, allv.size, one[key].size))
        parts = [None] * world
        e = np.stack([multi["edge_i"].astype(np.int64), multi["edge_j"].astype(np.int64), multi["edge_d"].view(np.int64)], axis=1)
        dist.all_gather_object(parts, e.tobytes())
        alle = np.concatenate([np.frombuffer(p, dtype=np.int64).reshape(-1, 3) for p in parts])
        e1 = np.stack([one["edge_i"].astype(np.int64), one["edge_j"].astype(np.int64), one["edge_d"].view(np.int64)], axis=1)
        key = lambda a: a[np.lexsort((a[:, 2], a[:, 1], a[:, 0]))]
        if alle.shape != e1.shape or not np.array_equal(key(alle), key(e1)):
            failures.append("%s mode %d: edges differ (%d vs %d)" % (cfg, mode, len(alle), len(e1)))
        # dense fallback on several GPUs
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_DENSE)
        aln.set_config(mode, 300, True)
        dense = aln.best_match()
        ctx.set_option(t.OPT_BEST_MATCH_PATH, t.PATH_AUTO)
        if dense.tobytes() != one["rows"].tobytes():
            failures.append("%s mode %d: rank %d dense multi-GPU rows differ" % (cfg, mode, rank))
        aln.close()
    allf = [None] * world
    dist.all_gather_object(allf, failures)
    if rank == 0:
        flat = [f for fs in allf for f in fs]
        print("MULTI_WORKER " + ("OK" if not flat else "FAIL: " + "; ".join(flat)), flush=True)
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
