#!/usr/bin/env python
"""tests/configs/config4_sharded.py -- configs[4]: hash-sharded signature table across N B200 (default: 1.4 M families per GPU, ~2 B
signatures over 8 GPUs, 250 M per shard; the same code bench.py runs as its `configs4` leg: bench_legs.configs4), 1 M synthetic proteins per rank, k-mers exchanged over NVLink (NCCL send/recv
inside the C-ABI library; torch.distributed is only the launcher's plumbing: communicator id, barriers, reductions).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/configs/config4_sharded.py --gpus 8

Not the driver's bench line.  Prints one JSON line on rank 0: whole-job lookups/s (max over ranks), the phase times of
rank 0, interconnect bytes, and two full-size parity properties:
  * hits: every rank scans the proteins of EVERY rank against its own shard with the naive one-thread-per-position kernel;
    hit counts and (position, payload) checksums add up over the shards (all-reduce) to what the sharded run reported for
    each rank's proteins -- an independent path through the lookups (no routing, no prefilter, no exchange);
  * calls: the hits of the first --sample proteins go through the CPU oracle's run FSM; calls and OTU counts must match.
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
import bench_legs as legs  # noqa: E402
from bench import dist_setup, log  # noqa: E402


def local_main(a):
    """R virtual ranks on one GPU: the kernels see the same bin / owner structure as an R-GPU run (for ncu)."""
    import kmergutsjava_b200 as kg
    R = a.local_ranks
    ctx = kg.Context(0)
    u = synth.Universe(n_families=a.families or 1_250_000, sig_keep_per_1024=a.keep)
    tables = []
    for r in range(R):
        dk, dp, n = bl.synth_signatures_sharded(ctx, u, r, R)
        tables.append(ctx.table_from_device_entries_sharded(dk, dp, n, r, R))
        bl.device_free(dk)
        bl.device_free(dp)
    comms = kg.Comm.local([ctx] * R)
    batches = []
    for r in range(R):
        ds, do, total = bl.synth_proteins(ctx, u, r * a.proteins, a.proteins, seed=1)
        batches.append(ctx.batch_from_device(kg.MODE_AA, ds, do, a.proteins, total))
    params = kg.default_params()
    for _ in range(a.warmup):
        for res in kg.run_sharded_local(comms, tables, batches, params):
            res.free()
    t0 = time.perf_counter()
    lookups = 0
    for _ in range(a.steps):
        for res in kg.run_sharded_local(comms, tables, batches, params):
            lookups += res.stats.num_kmers
            res.free()
    dt = time.perf_counter() - t0
    st = comms[0].stats
    print(json.dumps({"workload": f"{R} virtual ranks on one GPU, {a.proteins} proteins each", "ms_per_step_all_ranks": 1e3 * dt / a.steps,
                      "lookups_per_s": lookups / dt,
                      "rank0_phase_ms": {k: round(getattr(st, "ms_" + k), 3) for k in ("route", "keys", "answer", "replies", "merge", "total")}}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--families", type=int, default=0, help="0 = 1.41 M per GPU (about 250 M signatures per shard, 2.0e9 on 8 GPUs)")
    ap.add_argument("--keep", type=int, default=legs.C4_KEEP, help="signature density, per 1024 consensus windows")
    ap.add_argument("--proteins", type=int, default=1_000_000, help="per rank")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--sample", type=int, default=2000)
    ap.add_argument("--no-check", action="store_true")
    ap.add_argument("--local-ranks", type=int, default=0,
                    help="profiling aid: R virtual ranks in ONE process on one GPU (peer copies instead of NCCL); prints phase times only")
    a = ap.parse_args()
    if a.local_ranks:
        return local_main(a)
    world, rank, local, torch, dist = dist_setup()
    import kmergutsjava_b200 as kg
    plumb = legs.Plumbing(torch, dist, rank, world, local)
    ctx = kg.Context(local)
    out = legs.configs4(kg, ctx, plumb, proteins=a.proteins, steps=a.steps, warmup=a.warmup, families=a.families, keep=a.keep,
                        check=not a.no_check, sample=a.sample, log=log)
    if rank == 0:
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
