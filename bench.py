#!/usr/bin/env python
"""bench.py -- headline benchmark of the 8-mer signature lookup + function calling path (BASELINE.json configs[1]):
synthetic 200M-signature table, 1M random-length synthetic proteins, protein mode, per GPU.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU)
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

A step = one pass of the hot path (patch -> encode+probe -> gather -> run FSM -> CALL/OTU records) over one batch of
1M proteins.  `value` = lookups/s with the proteins already resident in HBM; `e2e` = the same through the C-ABI call
kg_run() with pinned HOST buffers (H2D of the sequences and D2H of calls + OTU counts inside the timed region).
Inputs (table 1.8 GB + 0.3 GB of residues per step) are far larger than the 126 MB L2, so no explicit L2 flush is done.
"""
import argparse
import ctypes as C
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
    ap.add_argument("--cpu-sample", type=int, default=0, help="proteins in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-e2e2", action="store_true", help="skip the two-thread variant of the end-to-end measurement")
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


def dist_setup(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return world, rank, local, torch, dist


def all_max(torch, dist, x, local):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def all_sum(torch, dist, x, local):
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def barrier(torch, dist, local):
    torch.cuda.synchronize(local)
    if dist is not None:
        dist.barrier()
    torch.cuda.synchronize(local)


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


def cpu_reference_run(kgo, otable, sb, off, threads):
    t0 = time.time()
    ref = kgo.run(otable, kgo.make_params(aa=True), sb, off, kgo.STREAM_JOIN, threads=threads)
    return ref, time.time() - t0


def sample_host(kg, ctx, ds, do, nsample):
    off = bl.to_host(ctx, do, 8 * (nsample + 1)).view(np.uint64).copy()
    sb = bl.to_host(ctx, ds, int(off[-1]))
    return sb, off


def run_ours(args):
    world, rank, local, torch, dist = dist_setup(args)
    import kmergutsjava_b200 as kg
    ctx = kg.Context(local)
    u, (dk, dp, nsig), table, (ds, do, total), prep = build_inputs(kg, ctx, args, rank)
    ti = table.info
    params = kg.default_params()
    batch = ctx.batch_from_device(kg.MODE_AA, ds, do, args.proteins, total)

    # ---- device-resident throughput ----
    clocks = ClockSampler(local)
    clocks.start()
    for _ in range(max(args.warmup, 3)):
        ctx.run_batch(table, batch, params).free()
        log("warm-up step done")
    barrier(torch, dist, local)
    w0 = time.time()
    t0 = time.perf_counter()
    probe_ms, dev_ms, lookups, launches, st = [], [], 0, 0, None
    for _ in range(args.steps):
        r = ctx.run_batch(table, batch, params)
        st = r.stats
        probe_ms.append(st.ms_probe)
        dev_ms.append(st.ms_device)
        lookups += st.num_kmers
        launches += st.num_launches
        r.free()
    barrier(torch, dist, local)
    dt = all_max(torch, dist, time.perf_counter() - t0, local)
    clocks.window(w0, time.time())
    total_lookups = all_sum(torch, dist, float(lookups), local)
    total_proteins = float(args.proteins * args.steps * world)
    value = total_lookups / dt
    log(f"device-resident: {value:.3e} lookups/s")

    # ---- end to end through kg_run with pinned host buffers ----
    e2e = None
    if not args.no_e2e:
        h_seq = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
        h_off = torch.empty(args.proteins + 1, dtype=torch.int64, pin_memory=True)
        kg._check(bl.lib().kg_device_to_host(ctx._h, h_seq.data_ptr(), ds, total))
        kg._check(bl.lib().kg_device_to_host(ctx._h, h_off.data_ptr(), do, 8 * (args.proteins + 1)))
        d2h = 0
        for _ in range(max(args.warmup, 3)):
            ctx.run_ptr(table, kg.MODE_AA, h_seq.data_ptr(), h_off.data_ptr(), args.proteins, params).free()
        barrier(torch, dist, local)
        w0 = time.time()
        t0 = time.perf_counter()
        e_lookups = 0
        for _ in range(args.steps):
            r = ctx.run_ptr(table, kg.MODE_AA, h_seq.data_ptr(), h_off.data_ptr(), args.proteins, params)
            s2 = r.stats
            e_lookups += s2.num_kmers
            d2h = s2.num_calls * kg.CALL_DTYPE.itemsize + args.proteins * kg.OTU_DTYPE.itemsize + 64
            r.free()
        barrier(torch, dist, local)
        edt = all_max(torch, dist, time.perf_counter() - t0, local)
        clocks.window(w0, time.time())
        e2e = {"value": all_sum(torch, dist, float(e_lookups), local) / edt, "unit": "lookups/s",
               "h2d_bytes_per_step": int(total + 8 * (args.proteins + 1)), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * edt / args.steps}

        # The same end-to-end call from TWO host threads, each with its own context (streams, scratch, pools) and its own
        # pinned buffers, alternating steps: the copies of one call overlap the kernels of the other.  Reported next to
        # `e2e` (which stays the single synchronous call); every step still moves its inputs in and its records out.
        if not args.no_e2e2:
            import threading
            ctx2 = kg.Context(local)
            table.attach(ctx2)
            h_seq2 = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
            h_seq2.copy_(h_seq)
            lanes = [(ctx, h_seq), (ctx2, h_seq2)]
            counts = [0, 0]

            def lane(k, nsteps):
                c, hs = lanes[k]
                for _ in range(nsteps):
                    r2 = c.run_ptr(table, kg.MODE_AA, hs.data_ptr(), h_off.data_ptr(), args.proteins, params)
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
            barrier(torch, dist, local)
            w0 = time.time()
            t0 = time.perf_counter()
            both(args.steps)
            barrier(torch, dist, local)
            edt2 = all_max(torch, dist, time.perf_counter() - t0, local)
            clocks.window(w0, time.time())
            e2e["two_threads"] = {"value": all_sum(torch, dist, float(sum(counts)), local) / edt2, "unit": "lookups/s",
                                  "ms_per_step": 1e3 * edt2 / args.steps,
                                  "note": "two host threads, one context each, alternate the steps (same bytes in and out per step)"}
            ctx2.close()

    clk = clocks.stop()
    log("e2e done")
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
    try:    # dram__bytes_read.sum + dram__bytes_write.sum of k_probe from the committed `ncu --set full` capture of this workload
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj["signatures"] == int(nsig) and tj["proteins"] == args.proteins:
            traffic, traffic_src = tj["k_probe_dram_bytes_per_launch"], tj["source"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": round(achieved_gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved_gbs / hbm_peak, 4), "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": lookups_per_step * BYTES_PER_LOOKUP_AA, "peak_source": peak_src,
                "kernel": "k_probe", "kernel_ms": round(probe_s * 1e3, 4),
                "bytes_per_lookup": BYTES_PER_LOOKUP_AA, "lookups_per_launch": lookups_per_step,
                "probe_roofline_sectors_per_s": r_probe,
                "frac_of_probe_roofline": round(lookups_per_step / probe_s / r_probe, 4) if r_probe else None,
                "kernel_share_of_step": round(float(np.mean(probe_ms)) / float(np.mean(dev_ms)), 4)}

    # ---- CPU baseline + parity on the sample (rank 0, N=1 only) ----
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from oracle import kgo
        kgo.build()
        threads = os.cpu_count() or 1
        num_slots = 3 * nsig + 1  # load 1/3: at 1/2 the reference hash (key % numSigs) clusters so badly that no prime near 2n avoids running off the end
        t0 = time.time()
        log("roofline done; building the reference-format image for the CPU baseline")
        img = bl.synth_reference_image(ctx, dk, dp, nsig, num_slots)
        num_slots = int(img[:8].view(np.int64)[0])
        otable = kgo.Table(borrow=img)
        log("image on host")
        t_img = time.time() - t0
        nsample = args.cpu_sample or min(args.proteins, 64000 * threads)
        sb, off = sample_host(kg, ctx, ds, do, nsample)
        ref, secs = cpu_reference_run(kgo, otable, sb, off, threads)
        log(f"cpu baseline done in {secs:.1f}s")
        cpu = {"value": ref.num_kmers / secs, "unit": "lookups/s", "cores": threads, "kind": "port",
               "sample": f"first {nsample} proteins of the step ({ref.num_kmers} lookups) against the full table in the "
                         f"reference's own 24-byte-slot format ({num_slots} slots); reference algorithm (comparator sort + "
                         f"one pass over the table stream, KGJ:944-1034) run as {threads} independent single-threaded shards; "
                         f"{secs:.1f} s; C port of the Java (no JVM in this image)",
               "image_build_s": round(t_img, 1)}
        # parity of the same sample: GPU calls / OTU counts vs the oracle's, bit for bit
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from tests.parity import assert_same
        g = ctx.run(table, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
        assert_same(g, ref, what="bench sample")
        parity = f"bit-exact on the {nsample}-protein sample: {len(ref.hits)} hits, {len(ref.calls)} calls"
        g.free()

    # ---- size-independent property at FULL size: lookups, hits and a checksum over every hit's (position, payload) must
    # equal those of the naive one-thread-per-position kernel (no prefilter, no queue, byte-wise reads, full table lookup)
    full = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        valid, nh, chk = bl.naive_scan_aa(ctx, table, ds, do, args.proteins, total)
        g = ctx.run_batch(table, batch, kg.default_params(emit_hits=1))
        gs = g.stats
        ok = (gs.num_kmers, gs.num_hits) == (valid, nh) and bl.hits_checksum(ctx, g.hits, do) == chk
        g.free()
        if not ok:
            raise SystemExit("full-size cross-check against the naive kernel FAILED")
        full = f"lookups ({valid}), hits ({nh}) and hit checksum equal the naive kernel's on all {args.proteins} proteins"
        log("full-size cross-check done")

    if rank == 0:
        out = {
            "metric": "8-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
            "config": {"workload": "configs[1]: synthetic 200M-signature table, 1M random-length synthetic proteins, protein mode",
                       "signatures": int(nsig), "families": args.families, "proteins_per_gpu": args.proteins,
                       "residues_per_gpu": int(total), "table_replicated": True,
                       "l2": "inputs larger than L2 (table %.2f GB, residues %.2f GB per step); no flush" % (ti.device_bytes / 1e9, total / 1e9),
                       "table_buckets": int(ti.num_buckets), "table_flagged_buckets": int(ti.flagged_buckets),
                       "hits_per_step": int(st.num_hits), "calls_per_step": int(st.num_calls)},
            "proteins_per_s": total_proteins / dt,
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "parity_full_size": full, "prep": prep,
            "stage_ms": {"prepare": round(st.ms_prepare, 4), "probe": round(st.ms_probe, 4), "group": round(st.ms_group, 4),
                         "device_total": round(st.ms_device, 4), "probe_filter": round(st.ms_filter, 4),
                         "probe_refilter": round(st.ms_refilter, 4), "probe_lines": round(st.ms_lines, 4),
                         "survivors_filter1": int(st.num_survivors1), "survivors_filter2": int(st.num_survivors2)},
        }
        print(json.dumps(out))
    batch.free()
    table.free()
    for p in (dk, dp, ds, do):
        bl.device_free(p)
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU algorithm (C port: no JVM here) on the host cores, same config / metric / unit.  The GPU is
    used only OUTSIDE the timed region, to generate the synthetic inputs in the reference's own table format."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import kmergutsjava_b200 as kg
    from oracle import kgo
    kgo.build()
    ctx = kg.Context(int(os.environ.get("LOCAL_RANK", "0")))
    u = synth.Universe(n_families=args.families)
    dk, dp, nsig = bl.synth_signatures(ctx, u, args.sigs)
    num_slots = 3 * nsig + 1  # load 1/3: at 1/2 the reference hash (key % numSigs) clusters so badly that no prime near 2n avoids running off the end
    img = bl.synth_reference_image(ctx, dk, dp, nsig, num_slots)
    num_slots = int(img[:8].view(np.int64)[0])
    bl.device_free(dk)
    bl.device_free(dp)
    otable = kgo.Table(borrow=img)
    threads = os.cpu_count() or 1
    nsample = args.cpu_sample or min(args.proteins, 16000 * threads)
    ds, do, total = bl.synth_proteins(ctx, u, 0, nsample, seed=1)
    sb, off = sample_host(kg, ctx, ds, do, nsample)
    bl.device_free(ds)
    bl.device_free(do)
    ctx.close()
    for _ in range(args.warmup):
        cpu_reference_run(kgo, otable, sb, off, threads)
    t0 = time.perf_counter()
    lookups = 0
    for _ in range(args.steps):
        ref, _ = cpu_reference_run(kgo, otable, sb, off, threads)
        lookups += ref.num_kmers
    dt = time.perf_counter() - t0
    value = lookups / dt
    sample = (f"each step = first {nsample} proteins of configs[1] ({lookups // args.steps} lookups) against the full "
              f"{nsig}-signature table in the reference's 24-byte-slot format ({num_slots} slots)")
    print(json.dumps({
        "impl": "reference", "metric": "8-mer lookups/sec", "value": value, "unit": "lookups/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
        "config": {"workload": "configs[1]: synthetic 200M-signature table, 1M random-length synthetic proteins, protein mode",
                   "signatures": int(nsig), "families": args.families, "proteins_per_step": nsample},
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
