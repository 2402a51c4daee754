#!/usr/bin/env python
"""tests/configs/config3_orfs.py -- configs[3] at full length: a metagenome-scale batch of synthetic ORFs (default 1e8 proteins,
~3.2e10 residues) sharded across the ranks with the signature table replicated (the same code bench.py runs as its `configs3`
leg: bench_legs.configs3).  Each rank generates its shard ON THE DEVICE, one million proteins at a time, and runs every batch
through the C ABI.  Prints one JSON line on rank 0: proteins/s and lookups/s over the whole job (max over ranks) and a bit-exact
parity check of deterministic samples against the CPU oracle.

    python tests/configs/config3_orfs.py --orfs 100000000
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/configs/config3_orfs.py --gpus 8
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools import kg_synth as synth  # noqa: E402
from tools import kg_benchlib as bl  # noqa: E402
import bench_legs as legs  # noqa: E402
from bench import dist_setup, log  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--orfs", type=int, default=100_000_000, help="proteins in the whole job (all ranks)")
    ap.add_argument("--batch", type=int, default=1_000_000)
    ap.add_argument("--families", type=int, default=2_000_000)
    ap.add_argument("--sigs", type=int, default=200_000_000)
    ap.add_argument("--parity", type=int, default=10_000, help="proteins per parity sample (0 = skip)")
    a = ap.parse_args()
    world, rank, local, torch, dist = dist_setup()
    import kmergutsjava_b200 as kg
    plumb = legs.Plumbing(torch, dist, rank, world, local)
    ctx = kg.Context(local)
    u = synth.Universe(n_families=a.families)
    dk, dp, nsig = bl.synth_signatures(ctx, u, a.sigs)
    table = ctx.table_from_device_entries(dk, dp, nsig)
    otable = None
    if a.parity and rank == 0:
        from oracle import kgo
        kgo.build()
        otable = kgo.Table(borrow=bl.synth_reference_image(ctx, dk, dp, nsig, 3 * nsig + 1))
    out = legs.configs3(kg, ctx, table, u, plumb, orfs=a.orfs, batch=a.batch, otable=otable, parity=a.parity,
                        threads=os.cpu_count() or 1, log=log)
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
