#!/usr/bin/env python
"""tests/configs/config2_dna.py -- configs[2]: 6-frame contig mode, N synthetic G-Mbp genomes against the 200M-signature table.
Not the driver's bench line (that is bench.py / configs[1]); prints one JSON line with stage times, Mbp/s, lookups/s and a
bit-exact parity check of the first genomes against the CPU oracle."""
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
    ap.add_argument("--genomes", type=int, default=50)
    ap.add_argument("--length", type=int, default=5_000_000)
    ap.add_argument("--families", type=int, default=2_000_000)
    ap.add_argument("--sigs", type=int, default=200_000_000)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--parity-genomes", type=int, default=2)
    a = ap.parse_args()
    import kmergutsjava_b200 as kg
    ctx = kg.Context(0)
    u = synth.Universe(n_families=a.families)
    dk, dp, nsig = bl.synth_signatures(ctx, u, a.sigs)
    table = ctx.table_from_device_entries(dk, dp, nsig)
    ds, do, total = bl.synth_genomes(ctx, u, a.genomes, a.length, seed=2)
    batch = ctx.batch_from_device(kg.MODE_DNA, ds, do, a.genomes, total)
    params = kg.default_params()
    for _ in range(3):
        ctx.run_batch(table, batch, params).free()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        r = ctx.run_batch(table, batch, params)
        st = r.stats
        r.free()
    dt = (time.perf_counter() - t0) / a.steps
    out = {"workload": f"configs[2]: {a.genomes} x {a.length} bp, {nsig} signatures", "ms_per_step": dt * 1e3,
           "mbp_per_s": total / dt / 1e6, "lookups_per_s": st.num_kmers / dt, "positions": int(st.num_positions),
           "lookups": int(st.num_kmers), "hits": int(st.num_hits), "calls": int(st.num_calls),
           "stage_ms": {"prepare(translate)": round(st.ms_prepare, 3), "probe": round(st.ms_probe, 3), "group(fsm)": round(st.ms_group, 3),
                        "device": round(st.ms_device, 3)}}
    # end to end through kg_run: pinned host buffers in, host records out (H2D + six-frame pipeline + D2H, sliced)
    try:
        import torch
        h_seq = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
        h_off = torch.empty(a.genomes + 1, dtype=torch.int64, pin_memory=True)
        kg._check(bl.lib().kg_device_to_host(ctx._h, h_seq.data_ptr(), ds, total))
        kg._check(bl.lib().kg_device_to_host(ctx._h, h_off.data_ptr(), do, 8 * (a.genomes + 1)))
        for _ in range(3):
            ctx.run_ptr(table, kg.MODE_DNA, h_seq.data_ptr(), h_off.data_ptr(), a.genomes, params).free()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            ctx.run_ptr(table, kg.MODE_DNA, h_seq.data_ptr(), h_off.data_ptr(), a.genomes, params).free()
        edt = (time.perf_counter() - t0) / a.steps
        out["e2e"] = {"ms_per_step": edt * 1e3, "mbp_per_s": total / edt / 1e6, "h2d_bytes_per_step": int(total + 8 * (a.genomes + 1))}
    except ImportError:
        pass
    if a.parity_genomes:
        from oracle import kgo
        from tests.parity import assert_same
        kgo.build()
        img = bl.synth_reference_image(ctx, dk, dp, nsig, 3 * nsig + 1)
        n = min(a.parity_genomes, a.genomes)
        off = bl.to_host(ctx, do, 8 * (n + 1)).view(np.uint64).copy()
        sb = bl.to_host(ctx, ds, int(off[-1]))
        t0 = time.time()
        ref = kgo.run(kgo.Table(borrow=img), kgo.make_params(aa=False), sb, off, kgo.STREAM_JOIN, threads=n)
        cpu_s = time.time() - t0
        g = ctx.run(table, kg.MODE_DNA, sb, off, kg.default_params(emit_hits=1))
        assert_same(g, ref, what="configs[2] sample")
        out["parity"] = f"bit-exact on the first {n} genomes: {len(ref.hits)} hits, {len(ref.calls)} calls"
        out["cpu_port"] = {"lookups_per_s": ref.num_kmers / cpu_s, "threads": n, "seconds": round(cpu_s, 2)}
        g.free()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
