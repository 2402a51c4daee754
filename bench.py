#!/usr/bin/env python
"""bench.py -- headline benchmark of the 8-mer signature lookup + function calling path (BASELINE.json configs[1]):
synthetic 200M-signature table, 1M random-length synthetic proteins, protein mode, per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

A step = one pass of the hot path (patch -> encode+probe -> run FSM -> CALL/OTU records) over one batch of 1M proteins.
`value` = lookups/s with the proteins already resident in HBM; `e2e` = the same through the C-ABI call with pinned HOST
buffers (H2D of the sequences and D2H of calls + OTU counts inside the timed region): kg_run_packed_aa on the 5-bit residue
codes the library's ingest produces (kg_pack_aa), with the call on raw characters (kg_run) reported beside it.
Inputs (table 8.6 GB + 0.3 GB of residues per step) are far larger than the 126 MB L2, so no explicit L2 flush is done.

After the headline the same process runs reduced-time legs of BASELINE.json configs[2], [3] and [4] (bench_legs.py) and adds
their results -- with their parity checks -- to the JSON line as `configs2`, `configs3`, `configs4`; at N > 1 the configs[4]
leg goes through the library's own NCCL communicator (hash-sharded table, k-mers exchanged over NVLink).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from tools import kg_synth as synth  # noqa: E402
from tools import kg_benchlib as bl  # noqa: E402

BYTES_PER_LOOKUP_AA = 33.0  # one 32-byte table sector + 1 residue byte
WORKLOAD = "configs[1]: synthetic 200M-signature table, 1M random-length synthetic proteins, protein mode"

T_START = time.time()


def log(msg):
    print(f"[bench {time.time() - T_START:7.1f}s] {msg}", file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--families", type=int, default=2_000_000)
    ap.add_argument("--sigs", type=int, default=200_000_000)
    ap.add_argument("--proteins", type=int, default=1_000_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="proteins in the CPU baseline sample (0 = the whole step)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e2", action="store_true", help="skip the two-thread variant of the end-to-end measurement")
    ap.add_argument("--no-legs", action="store_true", help="skip the configs[2] / [3] / [4] legs")
    ap.add_argument("--no-table-load", action="store_true", help="skip the reference-format table load measurement")
    ap.add_argument("--legs", default="2,3,4", help="which legs to run")
    ap.add_argument("--orfs", type=int, default=100_000_000, help="configs[3]: proteins in the whole job")
    ap.add_argument("--c4-families", type=int, default=0, help="configs[4]: families (0 = 1.41 M per GPU, 2.0e9 signatures on 8 GPUs)")
    return ap.parse_args()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.windows = index, [], None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def window(self, t0, t1):
        """A timed region (wall-clock): only samples taken inside such windows are reported."""
        self.windows.append((t0, t1))

    def stop(self):
        if self.proc:
            time.sleep(0.06)
            self.proc.terminate()
        rows = [r[1:] for r in self.rows if len(r) >= 7 and any(a - 0.05 <= r[0] <= b + 0.05 for a, b in self.windows)]
        scope = "timed regions"
        if len(rows) < 2:   # very short runs: fall back to everything sampled while the GPU was busy (warm-up included)
            rows, scope = [r[1:] for r in self.rows if len(r) >= 7], "whole run incl. warm-up"
        sm = [int(r[0]) for r in rows if r[0].isdigit()]
        mx = [int(r[1]) for r in rows if r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in rows for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": int(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "scope": scope}


def measured_peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"], "measured"
    except Exception:
        return 6650.0, "fallback"


def dist_setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:  # more point-to-point channels for the library's exchanges in the configs[4] leg (NCCL reads these once per process)
        os.environ.setdefault("NCCL_MIN_P2P_NCHANNELS", "64")
        os.environ.setdefault("NCCL_MAX_P2P_NCHANNELS", "64")
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local, torch, dist


def workload_config(args, nsig, residues):
    """The keys that define the workload: identical in both arms (the reference arm emits exactly this dict)."""
    return {"workload": WORKLOAD, "signatures": int(nsig), "families": args.families, "proteins_per_gpu": args.proteins,
            "residues_per_gpu": int(residues), "table_replicated": True}


def build_inputs(kg, ctx, args, rank):
    """Synthetic universe -> GPU table + this rank's proteins (device resident)."""
    u = synth.Universe(n_families=args.families)
    t0 = time.time()
    dk, dp, nsig = bl.synth_signatures(ctx, u, args.sigs)
    t1 = time.time()
    log(f"signatures generated: {nsig}")
    table = ctx.table_from_device_entries(dk, dp, nsig)
    t2 = time.time()
    log("table built")
    ds, do, total = bl.synth_proteins(ctx, u, rank * args.proteins, args.proteins, seed=1)
    t3 = time.time()
    log(f"proteins generated: {total} residues")
    timing = {"gen_signatures_s": round(t1 - t0, 2), "build_table_s": round(t2 - t1, 2), "gen_proteins_s": round(t3 - t2, 2)}
    return u, (dk, dp, nsig), table, (ds, do, total), timing


def sample_host(ctx, ds, do, nsample):
    off = bl.to_host(ctx, do, 8 * (nsample + 1)).view(np.uint64).copy()
    sb = bl.to_host(ctx, ds, int(off[-1]))
    return sb, off


def timed_calls(plumb, clocks, steps, warmup, call):
    """W untimed + K timed calls bracketed by barrier + synchronize; returns (seconds max over ranks, lookups of this rank, last stats)."""
    for _ in range(warmup):
        call().free()
    plumb.barrier()
    w0 = time.time()
    t0 = time.perf_counter()
    lookups, st = 0, None
    for _ in range(steps):
        r = call()
        st = r.stats
        lookups += st.num_kmers
        r.free()
    plumb.barrier()
    dt = plumb.reduce([time.perf_counter() - t0], "MAX")[0]
    clocks.window(w0, time.time())
    return dt, lookups, st


def run_ours(args):
    world, rank, local, torch, dist = dist_setup()
    import kmergutsjava_b200 as kg
    import bench_legs as legs
    plumb = legs.Plumbing(torch, dist, rank, world, local)

    def guarded(name, leg):
        """The legs come after the headline measurement: at N = 1 a leg that fails (a parity check included) is reported in
        its own object instead of costing the whole JSON line.  At N > 1 a rank cannot leave a collective leg alone, so there
        the failure stays fatal for the job."""
        try:
            return leg()
        except (Exception, SystemExit) as e:
            if world > 1:
                raise
            log(f"{name} leg FAILED: {e!r}")
            return {"error": f"{type(e).__name__}: {e}"[:2000]}

    ctx = kg.Context(local)
    u, (dk, dp, nsig), table, (ds, do, total), prep = build_inputs(kg, ctx, args, rank)
    ti = table.info
    params = kg.default_params()
    batch = ctx.batch_from_device(kg.MODE_AA, ds, do, args.proteins, total)
    warm = max(args.warmup, 3)

    # ---- device-resident throughput ----
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(warm):
        ctx.run_batch(table, batch, params).free()
    log("warm-up done")
    def timed_device(many):
        plumb.barrier()
        w0 = time.time()
        t0 = time.perf_counter()
        probe_ms, dev_ms, lookups, launches, st = [], [], 0, 0, None
        if many:   # two batches in flight: the run FSM of step i overlaps the probe of step i+1
            rs = ctx.run_batches(table, [batch] * args.steps, params)
        else:
            rs = (ctx.run_batch(table, batch, params) for _ in range(args.steps))
        for r in rs:
            st = r.stats
            probe_ms.append(st.ms_probe)
            dev_ms.append(st.ms_device)
            lookups += st.num_kmers
            launches += st.num_launches
            r.free()
        plumb.barrier()
        dt = plumb.reduce([time.perf_counter() - t0], "MAX")[0]
        clocks.window(w0, time.time())
        return dt, probe_ms, dev_ms, lookups, launches, st

    dt, probe_ms, dev_ms, lookups, launches, st = timed_device(False)
    total_lookups = plumb.reduce([float(lookups)])[0]
    value = total_lookups / dt
    log(f"device-resident: {value:.3e} lookups/s")
    # the same K steps through kg_batch_submit / kg_batch_collect (two batches in flight: the run FSM of step i overlaps the probe
    # of step i+1).  Reported beside `value`, not as it: the overlapped FSM competes with k_probe for HBM and L2 (k_probe 4.65 ->
    # 5.3 ms), so the step only gains ~2 % and the kernel's own roofline is cleaner measured one call at a time.
    for r in ctx.run_batches(table, [batch] * 3, params):
        r.free()
    dt2, probe_ms2, _, lookups2, _, _ = timed_device(True)
    two_in_flight = {"value": plumb.reduce([float(lookups2)])[0] / dt2, "unit": "lookups/s", "ms_per_step": 1e3 * dt2 / args.steps,
                     "k_probe_ms": round(float(np.mean(probe_ms2)), 4), "call": "kg_batch_submit / kg_batch_collect, two batches in flight"}

    # ---- end to end through the C ABI with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        h_seq = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
        h_off = torch.empty(args.proteins + 1, dtype=torch.int64, pin_memory=True)
        kg._check(bl.lib().kg_device_to_host(ctx._h, h_seq.data_ptr(), ds, total))
        kg._check(bl.lib().kg_device_to_host(ctx._h, h_off.data_ptr(), do, 8 * (args.proteins + 1)))
        # the library's ingest form: toAminoAcidOff codes, 8 per 5 bytes (kg_pack_aa; packing is host work OUTSIDE the call,
        # like FASTA parsing -- its rate is reported as pack_GBps)
        h_goff = torch.empty(args.proteins + 1, dtype=torch.int64, pin_memory=True)
        kg._check(kg.lib().kg_pack_aa(h_seq.data_ptr(), h_off.data_ptr(), args.proteins, None, h_goff.data_ptr(), 1))
        ngroups = int(h_goff[-1])
        h_pk = torch.empty(5 * ngroups + 64, dtype=torch.uint8, pin_memory=True)
        pack_threads = min(os.cpu_count() or 1, 16)
        tp = time.perf_counter()
        kg._check(kg.lib().kg_pack_aa(h_seq.data_ptr(), h_off.data_ptr(), args.proteins, h_pk.data_ptr(), h_goff.data_ptr(), pack_threads))
        pack_s = time.perf_counter() - tp

        def d2h_bytes(s):   # what the call copies back per step: dense calls + a count byte per protein (+ the used OTU pairs, below)
            return int(s.num_calls * kg.CALL_DTYPE.itemsize + args.proteins)

        def pk_call():
            return ctx.run_packed_aa_ptr(table, h_pk.data_ptr(), h_goff.data_ptr(), args.proteins, params)

        def raw_call():
            return ctx.run_ptr(table, kg.MODE_AA, h_seq.data_ptr(), h_off.data_ptr(), args.proteins, params)

        r0 = pk_call()
        n_entries = len(r0.otus_compact[1])
        r0.free()
        edt, e_lookups, s2 = timed_calls(plumb, clocks, args.steps, warm, pk_call)
        e2e = {"value": plumb.reduce([float(e_lookups)])[0] / edt, "unit": "lookups/s",
               "h2d_bytes_per_step": int(5 * ngroups + 8 * (args.proteins + 1)),
               "d2h_bytes_per_step": d2h_bytes(s2) + 8 * n_entries, "ms_per_step": 1e3 * edt / args.steps,
               "call": "kg_run_packed_aa: 5-bit residue codes (kg_pack_aa) in, calls + compact OTU counts out",
               "pack_GBps": round(total / pack_s / 1e9, 2), "pack_threads": pack_threads}
        log(f"e2e (packed): {e2e['value']:.3e} lookups/s")
        # Ceiling of the host link: the SAME bytes in and out per step and nothing else (one cudaMemcpyAsync each way from / to
        # pinned memory, both directions at once, all ranks together) -- what the end-to-end call could reach if every kernel
        # were free.  At N = 8 all ranks share one host's memory system; e2e is to be read against this number.
        d_in = torch.empty(e2e["h2d_bytes_per_step"], dtype=torch.uint8, device=f"cuda:{local}")
        d_out = torch.empty(e2e["d2h_bytes_per_step"], dtype=torch.uint8, device=f"cuda:{local}")
        h_in = torch.empty(e2e["h2d_bytes_per_step"], dtype=torch.uint8, pin_memory=True)
        h_out = torch.empty(e2e["d2h_bytes_per_step"], dtype=torch.uint8, pin_memory=True)
        s_in, s_out = torch.cuda.Stream(local), torch.cuda.Stream(local)

        def copy_step():
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
            s_in.synchronize()
            s_out.synchronize()

        for _ in range(3):
            copy_step()
        plumb.barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            copy_step()
        plumb.barrier()
        cdt = plumb.reduce([time.perf_counter() - t0], "MAX")[0]
        e2e["copy_only_ms_per_step"] = 1e3 * cdt / args.steps
        e2e["copy_only_GBps_all_ranks"] = round(world * (e2e["h2d_bytes_per_step"] + e2e["d2h_bytes_per_step"]) * args.steps / cdt / 1e9, 1)
        e2e["e2e_over_copy_only"] = round(e2e["ms_per_step"] / e2e["copy_only_ms_per_step"], 3)
        del d_in, d_out, h_in, h_out
        rdt, r_lookups, s3 = timed_calls(plumb, clocks, args.steps, warm, raw_call)
        e2e["raw_bytes_call"] = {"value": plumb.reduce([float(r_lookups)])[0] / rdt, "unit": "lookups/s", "ms_per_step": 1e3 * rdt / args.steps,
                                 "h2d_bytes_per_step": int(total + 8 * (args.proteins + 1)), "d2h_bytes_per_step": d2h_bytes(s3) + 8 * n_entries,
                                 "call": "kg_run: one byte per residue, as FastaCallback.nextEntry receives them (KGJ:780)"}
        # The same packed call from TWO host threads, each with its own context (streams, scratch, pools) and its own
        # pinned buffers, alternating steps: the copies of one call overlap the kernels of the other.  Reported next to
        # `e2e` (which stays the single synchronous call); every step still moves its inputs in and its records out.
        if not args.no_e2e2:
            ctx2 = kg.Context(local)
            table.attach(ctx2)
            h_pk2 = torch.empty(5 * ngroups + 64, dtype=torch.uint8, pin_memory=True)
            h_pk2.copy_(h_pk)
            lanes = [(ctx, h_pk), (ctx2, h_pk2)]
            counts = [0, 0]

            def lane(k, nsteps):
                c, hp = lanes[k]
                for _ in range(nsteps):
                    r2 = c.run_packed_aa_ptr(table, hp.data_ptr(), h_goff.data_ptr(), args.proteins, params)
                    counts[k] += r2.stats.num_kmers
                    r2.free()

            def both(nsteps):
                th = [threading.Thread(target=lane, args=(k, nsteps // 2 + (k < nsteps % 2))) for k in range(2)]
                for t in th:
                    t.start()
                for t in th:
                    t.join()

            both(6)
            counts[:] = [0, 0]
            plumb.barrier()
            w0 = time.time()
            t0 = time.perf_counter()
            both(args.steps)
            plumb.barrier()
            edt2 = plumb.reduce([time.perf_counter() - t0], "MAX")[0]
            clocks.window(w0, time.time())
            e2e["two_threads"] = {"value": plumb.reduce([float(sum(counts))])[0] / edt2, "unit": "lookups/s",
                                  "ms_per_step": 1e3 * edt2 / args.steps,
                                  "note": "two host threads, one context each, alternate the steps (same bytes in and out per step)"}
            ctx2.close()
            del h_pk2
        del h_seq, h_off, h_pk, h_goff
        log("e2e done")
    clk = clocks.stop()

    # ---- rooflines ----
    hbm_peak, peak_src = measured_peaks()
    probe_s = float(np.mean(probe_ms)) * 1e-3
    lookups_per_step = lookups / args.steps
    achieved_gbs = lookups_per_step * BYTES_PER_LOOKUP_AA / probe_s / 1e9
    r_probe = 0.0
    if rank == 0:
        for tpb, infl in ((256, 4), (256, 8), (512, 4), (1024, 2), (128, 8)):
            r_probe = max(r_probe, bl.probe_roofline_table(ctx, table, 1 << 28, tpb, infl))
    traffic, traffic_src = None, None
    for name in ("r02_traffic.json", "r01_traffic.json"):  # dram__bytes_read.sum + dram__bytes_write.sum of the probe kernel, `ncu --set full`
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
            if tj["signatures"] == int(nsig) and tj["proteins"] == args.proteins:
                traffic, traffic_src = tj["k_probe_dram_bytes_per_launch"], tj["source"]
                break
        except Exception:
            pass
    cascade = st.ms_filter > 0
    roofline = {"bound": "hbm", "achieved": round(achieved_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved_gbs / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": lookups_per_step * BYTES_PER_LOOKUP_AA, "peak_source": peak_src,
                "kernel": ("k_probe_half<0> + k_probe_half<1> (two passes, one per half of the key space)" if os.environ.get("KG_FILTER_HALVES", "0") not in ("", "0")
                           else "k_filter + k_refilter + k_probe2 (probe cascade, three launches)" if cascade else "k_probe"),
                "kernel_ms": round(probe_s * 1e3, 4),
                "bytes_per_lookup": BYTES_PER_LOOKUP_AA, "lookups_per_launch": lookups_per_step,
                "probe_roofline_sectors_per_s": r_probe,
                "frac_of_probe_roofline": round(lookups_per_step / probe_s / r_probe, 4) if r_probe else None,
                "kernel_share_of_step": round(float(np.mean(probe_ms)) / float(np.mean(dev_ms)), 4)}

    # ---- CPU baseline + parity on the whole step (rank 0, N=1 only) ----
    cpu = parity = full = table_load = None
    otable = None
    threads = os.cpu_count() or 1
    want_cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    if want_cpu:
        from oracle import kgo
        from tests.parity import assert_same
        kgo.build()
        num_slots = 3 * nsig + 1  # load 1/3: at 1/2 the reference hash (key % numSigs) clusters so badly that no prime near 2n avoids running off the end
        t0 = time.time()
        log("roofline done; building the reference-format image for the CPU baseline")
        img = bl.synth_reference_image(ctx, dk, dp, nsig, num_slots)
        num_slots = int(img[:8].view(np.int64)[0])
        otable = kgo.Table(borrow=img)
        t_img = time.time() - t0
        log("image on host")
        nsample = args.cpu_sample or args.proteins
        sb, off = sample_host(ctx, ds, do, nsample)
        t0 = time.time()
        ref = kgo.run(otable, kgo.make_params(aa=True), sb, off, kgo.STREAM_JOIN, threads=threads)
        secs = time.time() - t0
        log(f"cpu baseline (reference algorithm, {threads} shards) done in {secs:.1f}s")
        cpu = {"value": ref.num_kmers / secs, "unit": "lookups/s", "cores": threads, "kind": "port",
               "sample": f"first {nsample} proteins of the step ({ref.num_kmers} lookups) against the full table in the "
                         f"reference's own 24-byte-slot format ({num_slots} slots); reference algorithm (comparator sort + "
                         f"one pass over the table stream, KGJ:944-1034) run as {threads} independent single-threaded shards; "
                         f"{secs:.1f} s; C port of the Java (no JVM in this image)",
               "image_build_s": round(t_img, 1)}
        # parity of the same sample: GPU hits / calls / OTU counts vs the oracle's, bit for bit -- through BOTH host calls
        g = ctx.run(table, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
        assert_same(g, ref, what="bench sample (kg_run)")
        g.free()
        pk, goff = kg.pack_aa(sb, off, threads=min(threads, 16))
        g = ctx.run_packed_aa(table, pk, goff, kg.default_params(emit_hits=1))
        assert_same(g, ref, what="bench sample (kg_run_packed_aa)")
        g.free()
        del pk, goff
        parity = (f"bit-exact on the {nsample}-protein sample, kg_run and kg_run_packed_aa: {len(ref.hits)} hits, "
                  f"{len(ref.calls)} calls, OTU counts")
        # SURVEY 8(d) asks for two more CPU rows: the reference algorithm on ONE thread (what KmerGutsJava really is: no Thread /
        # executor anywhere in KGJ) and a best-effort direct-probe variant on all host cores
        n1 = max(1000, nsample // 16)
        off1 = off[:n1 + 1].copy()
        t0 = time.time()
        r1 = kgo.run(otable, kgo.make_params(aa=True), sb[:int(off1[-1])], off1, kgo.STREAM_JOIN, threads=1)
        s1 = time.time() - t0
        cpu["single_thread"] = {"value": r1.num_kmers / s1, "unit": "lookups/s", "cores": 1,
                                "sample": f"first {n1} proteins ({r1.num_kmers} lookups), same table; the one pass over the "
                                          f"{num_slots}-slot table stream is paid whatever the batch size (KGJ:992-999); {s1:.1f} s"}
        t0 = time.time()
        r2 = kgo.run(otable, kgo.make_params(aa=True), sb, off, kgo.DIRECT_PROBE, threads=threads)
        s2_ = time.time() - t0
        assert len(r2.calls) == len(ref.calls) and r2.num_kmers == ref.num_kmers
        cpu["direct_probe_all_cores"] = {"value": r2.num_kmers / s2_, "unit": "lookups/s", "cores": threads,
                                         "sample": f"the same {nsample} proteins, in-memory linear probing of the same 24-byte-slot "
                                                   f"table instead of the stream join (not the reference's algorithm: best-effort CPU); {s2_:.1f} s"}
        log("cpu rows done")
        del ref, r1, r2, sb, off
        if not args.no_table_load:
            table_load = guarded("table_load", lambda: legs.table_load(kg, ctx, img, table, batch, params, log=log))
            log("table load leg done")

        # size-independent property at FULL size: lookups, hits and a checksum over every hit's (position, payload) must
        # equal those of the naive one-thread-per-position kernel (no prefilter, no queue, byte-wise reads, full table lookup)
        valid, nh, chk = bl.naive_scan_aa(ctx, table, ds, do, args.proteins, total)
        g = ctx.run_batch(table, batch, kg.default_params(emit_hits=1))
        gs = g.stats
        ok = (gs.num_kmers, gs.num_hits) == (valid, nh) and bl.hits_checksum(ctx, g.hits, do) == chk
        g.free()
        if not ok:
            raise SystemExit("full-size cross-check against the naive kernel FAILED")
        full = f"lookups ({valid}), hits ({nh}) and hit checksum equal the naive kernel's on all {args.proteins} proteins"
        log("full-size cross-check done")

    # ---- the other BASELINE.json configs, reduced-time legs (bench_legs.py) ----
    leg_out = {}
    which = set() if args.no_legs else set(args.legs.split(","))
    lsteps = max(3, min(args.steps, 10))
    if "2" in which:
        leg_out["configs2"] = guarded("configs2", lambda: legs.configs2(kg, ctx, table, u, plumb, steps=lsteps, otable=otable,
                                                                        parity_genomes=50 if otable is not None else 0,
                                                                        threads=threads, log=log))
        log("configs2 leg done")
    if "3" in which:
        leg_out["configs3"] = guarded("configs3", lambda: legs.configs3(kg, ctx, table, u, plumb, orfs=args.orfs, otable=otable,
                                                                        threads=threads, log=log))
        log("configs3 leg done")
    stats_keep = {"hits": int(st.num_hits), "calls": int(st.num_calls), "prepare": st.ms_prepare, "probe": st.ms_probe, "group": st.ms_group,
                  "device": st.ms_device, "filter": st.ms_filter, "refilter": st.ms_refilter, "lines": st.ms_lines,
                  "s1": int(st.num_survivors1), "s2": int(st.num_survivors2)}
    batch.free()
    table.free()
    otable = None
    for p in (dk, dp, ds, do):
        bl.device_free(p)
    if "4" in which:
        leg_out["configs4"] = guarded("configs4", lambda: legs.configs4(kg, ctx, plumb, proteins=args.proteins, steps=lsteps,
                                                                        families=args.c4_families, log=log))
        log("configs4 leg done")

    if rank == 0:
        out = {
            "metric": "8-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": world, "steps": args.steps,
            "warmup": warm, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": workload_config(args, nsig, total),
            "detail": {"l2": "inputs larger than L2 (table %.2f GB, residues %.2f GB per step); no flush" % (ti.device_bytes / 1e9, total / 1e9),
                       "table_buckets": int(ti.num_buckets), "table_flagged_buckets": int(ti.flagged_buckets),
                       "hits_per_step": stats_keep["hits"], "calls_per_step": stats_keep["calls"]},
            "proteins_per_s": float(args.proteins * args.steps * world) / dt,
            "two_in_flight": two_in_flight,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "parity_full_size": full, "prep": prep, "table_load": table_load,
            "stage_ms": {"prepare": round(stats_keep["prepare"], 4), "probe": round(stats_keep["probe"], 4),
                         "group": round(stats_keep["group"], 4), "device_total": round(stats_keep["device"], 4)},
        }
        if cascade:
            out["stage_ms"].update({"probe_filter": round(stats_keep["filter"], 4), "probe_refilter": round(stats_keep["refilter"], 4),
                                    "probe_lines": round(stats_keep["lines"], 4), "survivors_filter1": stats_keep["s1"],
                                    "survivors_filter2": stats_keep["s2"]})
        out.update(leg_out)
        print(json.dumps(out))
    ctx.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU algorithm (C port: no JVM here) on the host cores: same config / metric / unit, and every step is the
    WHOLE 1M-protein batch of configs[1].  This process loads NOTHING of the product: the synthetic inputs are written in the
    reference's own formats by a separate process (tools/gen_reference_inputs.py -- the generators run on the GPU) and only
    read back here."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import shutil
    import tempfile
    from oracle import kgo
    from bench_legs import _scratch_dir
    kgo.build()
    need = 24 * (3 * args.sigs + 1000) + 400 * args.proteins
    scratch = tempfile.mkdtemp(prefix="kg_ref_", dir=_scratch_dir(need) or tempfile.gettempdir())
    try:
        gen = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_reference_inputs.py"), "--out", scratch,
                              "--families", str(args.families), "--sigs", str(args.sigs), "--proteins", str(args.proteins),
                              "--device", os.environ.get("LOCAL_RANK", "0")], capture_output=True, text=True)
        if gen.returncode != 0:
            raise SystemExit("input generation failed:\n" + gen.stderr[-2000:])
        meta = json.loads(gen.stdout.strip().splitlines()[-1])
        nsig, num_slots, total = meta["signatures"], meta["num_slots"], meta["residues"]
        img = np.memmap(os.path.join(scratch, "table.img"), dtype=np.uint8, mode="r")   # tmpfs pages: no second copy
        otable = kgo.Table(borrow=img)
        sb = np.fromfile(os.path.join(scratch, "seq.bin"), dtype=np.uint8)
        off = np.fromfile(os.path.join(scratch, "off.bin"), dtype=np.uint64)
        threads = os.cpu_count() or 1
        nsample = args.cpu_sample or args.proteins
        if nsample < args.proteins:
            off = off[:nsample + 1].copy()
            sb = sb[:int(off[-1])]
        _run_reference_timed(args, kgo, otable, sb, off, nsig, num_slots, total, threads, nsample)
    finally:
        shutil.rmtree(scratch, ignore_errors=True)


def _run_reference_timed(args, kgo, otable, sb, off, nsig, num_slots, total, threads, nsample):
    def step():
        return kgo.run(otable, kgo.make_params(aa=True), sb, off, kgo.STREAM_JOIN, threads=threads)

    warm = min(args.warmup, 1)   # CPU code has nothing to warm beyond the page cache: one untimed step is plenty
    for _ in range(warm):
        step()
    t0 = time.perf_counter()
    lookups = 0
    for _ in range(args.steps):
        lookups += step().num_kmers
    dt = time.perf_counter() - t0
    value = lookups / dt
    sample = (f"each step = {'all' if nsample == args.proteins else 'the first'} {nsample} proteins of configs[1] "
              f"({lookups // args.steps} lookups) against the full {nsig}-signature table in the reference's 24-byte-slot format ({num_slots} slots)")
    print(json.dumps({
        "impl": "reference", "metric": "8-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": warm, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": workload_config(args, nsig, total),
        "cpu_baseline": {"value": value, "unit": "lookups/s", "cores": threads, "kind": "port", "sample": sample,
                         "note": "reference algorithm (comparator sort + one pass over the table stream, KGJ:944-1034) as "
                                 f"{threads} independent single-threaded shards; C port of the Java (no JVM in this image)"},
        "e2e": {"value": value, "unit": "lookups/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "proteins_per_s": nsample * args.steps / dt,
    }))


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
