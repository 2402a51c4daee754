#!/usr/bin/env python
"""tests/configs/config2_dna.py -- configs[2] at full length: 6-frame contig mode, G synthetic genomes of L bp against the
200M-signature table (the same code bench.py runs as its `configs2` leg: bench_legs.configs2).  Prints one JSON line with stage
times, Mbp/s, lookups/s and a bit-exact parity check of --parity-genomes genomes against the CPU oracle.

    python tests/configs/config2_dna.py --parity-genomes 50
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 tests/configs/config2_dna.py
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
    ap.add_argument("--genomes", type=int, default=50)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--families", type=int, default=2_000_000)
    ap.add_argument("--sigs", type=int, default=200_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--parity-genomes", type=int, default=2)
    ap.add_argument("--gpus", type=int, default=1)
    a = ap.parse_args()
    world, rank, local, torch, dist = dist_setup()
    import kmergutsjava_b200 as kg
    plumb = legs.Plumbing(torch, dist, rank, world, local)
    ctx = kg.Context(local)
    u = synth.Universe(n_families=a.families)
    dk, dp, nsig = bl.synth_signatures(ctx, u, a.sigs)
    table = ctx.table_from_device_entries(dk, dp, nsig)
    otable = None
    if a.parity_genomes and rank == 0:
        from oracle import kgo
        kgo.build()
        otable = kgo.Table(borrow=bl.synth_reference_image(ctx, dk, dp, nsig, 3 * nsig + 1))
    out = legs.configs2(kg, ctx, table, u, plumb, genomes=a.genomes, length=a.length, steps=a.steps, otable=otable,
                        parity_genomes=a.parity_genomes, threads=os.cpu_count() or 1, log=log)
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
