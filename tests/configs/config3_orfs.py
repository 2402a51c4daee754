#!/usr/bin/env python
"""tests/configs/config3_orfs.py -- configs[3]: a metagenome-scale batch of synthetic ORFs (default 1e8 proteins, ~3.2e10 residues)
sharded across the ranks with the signature table replicated.  Each rank generates its shard ON THE DEVICE, one million
proteins at a time (the 30 GB of residues never cross PCIe), and runs every batch through the C ABI (`kg_batch_run`).  Not
the driver's bench line; prints one JSON line on rank 0: proteins/s and lookups/s over the whole job (max over ranks), and a
bit-exact parity check of deterministic samples (the first proteins of the shard and a random batch) against the CPU oracle.

    python tests/configs/config3_orfs.py --orfs 100000000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/configs/config3_orfs.py --gpus 8
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools import kg_synth as synth  # noqa: E402
from tools import kg_benchlib as bl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--orfs", type=int, default=100_000_000, help="proteins in the whole job (all ranks)")
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--families", type=int, default=2_000_000)
    ap.add_argument("--sigs", type=int, default=200_000_000)
    ap.add_argument("--parity", type=int, default=10_000, help="proteins per parity sample (0 = skip)")
    a = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import kmergutsjava_b200 as kg
    ctx = kg.Context(local)
    u = synth.Universe(n_families=a.families)
    dk, dp, nsig = bl.synth_signatures(ctx, u, a.sigs)
    table = ctx.table_from_device_entries(dk, dp, nsig)
    params = kg.default_params()
    per = (a.orfs + world - 1) // world  # this rank's slice of the job's protein indices
    first, count = rank * per, max(0, min(per, a.orfs - rank * per))
    nb = (count + a.batch - 1) // a.batch

    def gen(b):
        n = min(a.batch, count - b * a.batch)
        ds, do, total = bl.synth_proteins(ctx, u, first + b * a.batch, n, seed=3)
        return ds, do, total, n

    # warm-up (buffer pools, clocks)
    ds, do, total, n = gen(0)
    bt = ctx.batch_from_device(kg.MODE_AA, ds, do, n, total)
    for _ in range(3):
        ctx.run_batch(table, bt, params).free()
    bt.free()
    bl.device_free(ds)
    bl.device_free(do)

    def sync_all():
        if dist is not None:
            import torch
            torch.cuda.synchronize()
            dist.barrier()

    sync_all()
    run_s = 0.0
    dev_ms = 0.0
    lookups = hits = calls = residues = 0
    t_job = time.perf_counter()
    for b in range(nb):
        ds, do, total, n = gen(b)
        bt = ctx.batch_from_device(kg.MODE_AA, ds, do, n, total)
        t0 = time.perf_counter()
        r = ctx.run_batch(table, bt, params)  # returns when the records are on the device and the counters are back
        run_s += time.perf_counter() - t0
        st = r.stats
        dev_ms += st.ms_device
        lookups += st.num_kmers
        hits += st.num_hits
        calls += st.num_calls
        residues += total
        r.free()
        bt.free()
        bl.device_free(ds)
        bl.device_free(do)
    sync_all()
    job_s = time.perf_counter() - t_job

    def reduce(x, op):
        if dist is None:
            return x
        import torch
        t = torch.tensor([float(x)], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    run_max = reduce(run_s, "MAX")
    job_max = reduce(job_s, "MAX")
    tot_lookups = reduce(lookups, "SUM")
    tot_hits = reduce(hits, "SUM")
    tot_calls = reduce(calls, "SUM")
    tot_res = reduce(residues, "SUM")
    out = {"workload": f"configs[3]: {a.orfs} synthetic ORFs over {world} GPU(s), {nsig} signatures replicated",
           "n_gpus": world, "batches_per_rank": nb, "proteins_per_s": a.orfs / run_max, "lookups_per_s": tot_lookups / run_max,
           "run_seconds": round(run_max, 3), "job_seconds_incl_generation": round(job_max, 3),
           "residues": int(tot_res), "lookups": int(tot_lookups), "hits": int(tot_hits), "calls": int(tot_calls),
           "rank0_device_ms_per_batch": round(dev_ms / nb, 3), "scaling": "strong (fixed 1e8-ORF job split over ranks)"}
    if a.parity and rank == 0:
        from oracle import kgo
        from tests.parity import assert_same
        kgo.build()
        img = bl.synth_reference_image(ctx, dk, dp, nsig, 3 * nsig + 1)
        otable = kgo.Table(borrow=img)
        rng = np.random.default_rng(3)
        picks = [("first of shard", first), ("random", first + int(rng.integers(0, max(count - a.parity, 1))))]
        notes = []
        for what, start in picks:
            n = min(a.parity, count)
            ds, do, total = bl.synth_proteins(ctx, u, start, n, seed=3)
            off = bl.to_host(ctx, do, 8 * (n + 1)).view(np.uint64).copy()
            sb = bl.to_host(ctx, ds, int(off[-1]))
            ref = kgo.run(otable, kgo.make_params(aa=True), sb, off, kgo.DIRECT_PROBE, threads=os.cpu_count() or 1)
            g = ctx.run(table, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
            assert_same(g, ref, what=f"configs[3] {what}")
            notes.append(f"{what} ({start}..+{n}): {len(ref.hits)} hits, {len(ref.calls)} calls")
            g.free()
            bl.device_free(ds)
            bl.device_free(do)
        out["parity"] = "bit-exact vs the CPU oracle: " + "; ".join(notes)
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
