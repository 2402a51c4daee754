#!/usr/bin/env python
"""tests/configs/config4_sharded.py -- configs[4]: hash-sharded signature table across N B200 (default: 10 M families, ~2 B
signatures over 8 GPUs, 250 M per shard), 1 M synthetic proteins per rank, k-mers exchanged over NVLink (NCCL send/recv
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
    ap.add_argument("--families", type=int, default=0, help="0 = 1.25 M per GPU (about 250 M signatures per shard)")
    ap.add_argument("--keep", type=int, default=700, help="signature density, per 1024 consensus windows")
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
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    families = a.families or 1_250_000 * world
    # more point-to-point channels for the library's exchanges (NCCL reads this once per process, before torch's communicator)
    os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "64")
    os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "64")
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import kmergutsjava_b200 as kg

    def log(msg):
        if rank == 0:
            print(f"[config4 +{time.time() - T0:6.1f}s] {msg}", file=sys.stderr, flush=True)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def reduce(x, op, dtype=torch.float64):
        if dist is None:
            return x
        t = torch.tensor(x, dtype=dtype, device=f"cuda:{local}")
        dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return t.tolist()

    T0 = time.time()
    ctx = kg.Context(local)
    u = synth.Universe(n_families=families, sig_keep_per_1024=a.keep)
    t0 = time.time()
    dk, dp, nsig = bl.synth_signatures_sharded(ctx, u, rank, world)
    t1 = time.time()
    table = ctx.table_from_device_entries_sharded(dk, dp, nsig, rank, world)
    bl.device_free(dk)
    bl.device_free(dp)
    t2 = time.time()
    ti = table.info
    log(f"shard built: {nsig} signatures, {ti.device_bytes / 1e9:.2f} GB ({t1 - t0:.1f} s generate, {t2 - t1:.1f} s build)")
    uid = [kg.Comm.unique_id() if (rank == 0 and world > 1) else None]
    if dist is not None:
        dist.broadcast_object_list(uid, src=0)
    comm = kg.Comm(ctx, rank, world, uid[0])
    ds, do, total = bl.synth_proteins(ctx, u, rank * a.proteins, a.proteins, seed=1)
    batch = ctx.batch_from_device(kg.MODE_AA, ds, do, a.proteins, total)
    params = kg.default_params()
    for _ in range(max(a.warmup, 3)):
        comm.run(table, batch, params).free()
    log("warm")
    barrier()
    t0 = time.perf_counter()
    lookups = 0
    phases = np.zeros(6)
    sent = 0
    for _ in range(a.steps):
        r = comm.run(table, batch, params)
        st = r.stats
        lookups += st.num_kmers
        ss = comm.stats
        phases += [ss.ms_route, ss.ms_keys, ss.ms_answer, ss.ms_replies, ss.ms_merge, ss.ms_total]
        sent += ss.bytes_sent
        r.free()
    barrier()
    dt = time.perf_counter() - t0
    dt_max = reduce([dt], "MAX")[0] if dist is not None else dt
    tot_lookups = reduce([float(lookups)], "SUM")[0] if dist is not None else float(lookups)
    tot_sigs = reduce([float(nsig)], "SUM")[0] if dist is not None else float(nsig)
    tot_sent = reduce([float(sent)], "SUM")[0] if dist is not None else float(sent)
    log(f"timed: {tot_lookups / dt_max:.3e} lookups/s")
    phases /= a.steps
    out = {"workload": f"configs[4]: hash-sharded table, {int(tot_sigs)} signatures over {world} GPU(s), {a.proteins} proteins per rank",
           "n_gpus": world, "steps": a.steps, "ms_per_step": 1e3 * dt_max / a.steps, "lookups_per_s": tot_lookups / dt_max,
           "proteins_per_s": a.proteins * world * a.steps / dt_max, "signatures": int(tot_sigs),
           "shard_bytes_rank0": int(ti.device_bytes), "scaling": "weak (proteins and signatures per GPU fixed)",
           "rank0_phase_ms": dict(zip(["route", "keys_exchange", "answer", "replies_exchange", "merge_and_fsm", "total_host"],
                                      [round(float(x), 3) for x in phases])),
           "interconnect_bytes_per_step": int(tot_sent / a.steps),
           "interconnect_GBps_per_gpu_during_exchange": None}
    out["chunks"] = int(comm.stats.chunks)
    if out["chunks"] > 1:
        out["rank0_phase_ms_note"] = "the step runs in chunks whose exchanges overlap the kernels: route .. replies_exchange are END times since the start of the call"
    ex_ms = phases[1] + phases[3]
    if world > 1 and ex_ms > 0 and out["chunks"] == 1:
        out["interconnect_GBps_per_gpu_during_exchange"] = round(tot_sent / a.steps / world / (ex_ms * 1e-3) / 1e9, 1)

    if not a.no_check:
        res = comm.run(table, batch, kg.default_params(emit_hits=1))
        hits, calls, otus = res.hits, res.calls, res.otus
        my = (len(hits), bl.hits_checksum(ctx, hits, do), int(res.stats.num_kmers))
        contrib = np.zeros((world, 2), dtype=np.uint64)
        valid = 0
        for s in range(world):  # proteins of rank s against MY shard
            if s == rank:
                ds2, do2, tot2 = ds, do, total
            else:
                ds2, do2, tot2 = bl.synth_proteins(ctx, u, s * a.proteins, a.proteins, seed=1)
            v, h, ck = bl.naive_scan_aa(ctx, table, ds2, do2, a.proteins, tot2)
            contrib[s] = (h, ck)
            if s == rank:
                valid = v
            else:
                bl.device_free(ds2)
                bl.device_free(do2)
        tot = np.array(reduce(contrib.view(np.int64).tolist(), "SUM", torch.int64), dtype=np.int64).view(np.uint64) if dist is not None else contrib
        ok = int(tot[rank][0]) == my[0] and int(tot[rank][1]) == my[1] and valid == my[2]
        all_ok = reduce([1.0 if ok else 0.0], "MIN")[0] if dist is not None else float(ok)
        if not ok:
            print(f"rank {rank}: sharded run {my} vs naive scan over all shards {tuple(int(x) for x in tot[rank])}, valid {valid}", file=sys.stderr)
        assert all_ok == 1.0, "hash-sharded run disagrees with the naive scan over all shards"
        tot_hits = reduce([float(my[0])], "SUM")[0] if dist is not None else float(my[0])
        out["parity_hits"] = (f"every rank: hits, lookups and (position, payload) checksum equal the naive scan summed over all {world} shards "
                              f"({int(tot_hits)} hits in total)")
        # calls of the first proteins through the CPU oracle's FSM
        from oracle import kgo
        kgo.build()
        n = min(a.sample, a.proteins)
        hs = hits[hits["seq"] < n]
        cs = calls[calls["seq"] < n]
        bounds_h = np.searchsorted(hs["seq"], np.arange(n + 1))
        bounds_c = np.searchsorted(cs["seq"], np.arange(n + 1))
        oparams = kgo.make_params(aa=True)
        ncalls = 0
        for i in range(n):
            h = hs[bounds_h[i]:bounds_h[i + 1]]
            oh = np.zeros(len(h), dtype=kgo.HIT_DTYPE)
            for f in oh.dtype.names:
                oh[f] = h[f]
            oc, oo = kgo.gather_hits(oparams, oh)
            c = cs[bounds_c[i]:bounds_c[i + 1]]
            assert len(oc) == len(c), f"rank {rank} protein {i}: {len(c)} calls vs oracle {len(oc)}"
            for f in ("start", "end", "count", "fI"):
                assert np.array_equal(oc[f].astype(np.int64), c[f].astype(np.int64)), f"rank {rank} protein {i}: {f}"
            assert np.array_equal(oc["weighted"].view(np.uint32), c["weighted"].view(np.uint32))
            k = int(oo["n"][0])
            assert k == int(otus["n"][i]) and np.array_equal(oo["count"][0][:k], otus["count"][i][:k]) and np.array_equal(oo["oI"][0][:k], otus["oI"][i][:k])
            ncalls += len(c)
        out["parity_calls"] = f"rank 0..{world - 1}: first {n} proteins of each rank through the oracle FSM, bit-exact (rank 0: {ncalls} calls)"
        res.free()
    if rank == 0:
        print(json.dumps(out))
    barrier()
    comm.free()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
