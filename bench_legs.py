"""bench_legs.py -- part of bench.py: the BASELINE.json configs[2], [3] and [4] workloads as functions, so that bench.py can run reduced-time
legs of them after the headline (configs[1]) and print their results in the same JSON line, and so that the stand-alone
scripts under tests/configs/ run exactly the same code at full length.

Every leg returns a dict (ms per step, lookups/s, stage times, a parity string).  Parity always goes through the CPU oracle
(oracle/ is test infrastructure: bench.py is one of the places allowed to run it, as the checker), except for the
full-size property of configs[4] (hits of every rank against every shard with the naive kernel, summed over the ranks).
"""
import ctypes as C
import os
import sys
import time

import numpy as np

from tools import kg_benchlib as bl
from tools import kg_synth as synth

ROOT = os.path.dirname(os.path.abspath(__file__))


class Plumbing:
    """What a leg needs from the launcher: rank / world, a barrier and reductions (torch.distributed over NCCL when N > 1)."""

    def __init__(self, torch=None, dist=None, rank=0, world=1, local=0):
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local

    def barrier(self):
        if self.torch is not None:
            self.torch.cuda.synchronize(self.local)
        if self.dist is not None:
            self.dist.barrier()
            self.torch.cuda.synchronize(self.local)

    def reduce(self, values, op="SUM", integer=False):
        """All-reduce of a list of numbers (float64, or int64 when integer)."""
        if self.dist is None:
            return list(values)
        t = self.torch.tensor(values, dtype=self.torch.int64 if integer else self.torch.float64, device=f"cuda:{self.local}")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return t.tolist()

    def bcast_obj(self, obj):
        if self.dist is None:
            return obj
        box = [obj]
        self.dist.broadcast_object_list(box, src=0)
        return box[0]


def _pinned_copy(kg, ctx, torch, d_ptr, nbytes, dtype=None, pad=0):
    t = torch.empty(nbytes + pad, dtype=torch.uint8, pin_memory=True)
    kg._check(bl.lib().kg_device_to_host(ctx._h, t.data_ptr(), d_ptr, nbytes))
    return t


# ---------------------------------------------------------------------------------------------------------------------
# loading the reference-format kmer.table.mem_map[.gz] (kg_table_load_file: the host only moves bytes, the
# slots are parsed on the device).  The 200M-signature image the CPU arm uses is written to a scratch file, loaded back and
# the loaded table must answer the headline batch exactly like the table built from the device-side generator.
# ---------------------------------------------------------------------------------------------------------------------
def _scratch_dir(nbytes):
    for d in (os.environ.get("KG_BENCH_TMP"), "/dev/shm", "/tmp"):
        if d and os.path.isdir(d):
            st = os.statvfs(d)
            if st.f_bavail * st.f_frsize > nbytes + (1 << 30):
                return d
    return None


def table_load(kg, ctx, img, table, batch, params, gz_bytes=256 << 20, log=lambda m: None):
    import gzip
    import tempfile
    out = {"what": "kg_table_load_file on the reference's own format (24-byte header + 24-byte slots, KGJ:933-935, 995-999)"}
    nbytes = int(img.nbytes)
    d = _scratch_dir(nbytes)
    if d is None:
        out["skipped"] = "no scratch directory with %.1f GB free" % (nbytes / 1e9)
        return out
    tmp = tempfile.mkdtemp(prefix="kg_table_", dir=d)
    plain = os.path.join(tmp, "kmer.table.mem_map")
    try:
        t0 = time.time()
        with open(plain, "wb") as f:
            step = 256 << 20
            for o in range(0, nbytes, step):
                f.write(memoryview(img[o:o + step]))
        out["write_s"] = round(time.time() - t0, 2)
        log(f"table image written to {plain} ({nbytes / 1e9:.1f} GB, {out['write_s']} s)")
        ref = ctx.run_batch(table, batch, params)
        want = (ref.stats.num_kmers, ref.stats.num_hits, ref.stats.num_calls, int(table.info.num_signatures))
        ref.free()
        secs = []
        for _ in range(2):   # the second load reads the page cache only
            t0 = time.time()
            t2 = ctx.load_table_file(plain)
            secs.append(time.time() - t0)
            got = ctx.run_batch(t2, batch, params)
            have = (got.stats.num_kmers, got.stats.num_hits, got.stats.num_calls, int(t2.info.num_signatures))
            info = t2.info
            got.free()
            t2.free()
            if have != want:
                raise SystemExit(f"table loaded from the file answers differently: {have} vs {want}")
        table.attach(ctx)   # the persisting-L2 window goes back to the bench's own table
        out.update({"file_bytes": nbytes, "slots": int(info.num_slots), "signatures": int(info.num_signatures),
                    "unreachable": int(info.num_unreachable), "tail_run": int(info.tail_run),
                    "load_s": [round(x, 3) for x in secs], "table_load_s": round(min(secs), 3), "GBps": round(nbytes / min(secs) / 1e9, 2),
                    "parity": "the loaded table answers the headline batch like the generated one (lookups, hits, calls, signatures)"})
        # .gz: inflate (one thread, zlib) is the bound; a prefix of the image with a patched slot count keeps this short
        nslots = min((gz_bytes - 24) // 24, (nbytes - 24) // 24)
        small = np.empty(24 + 24 * nslots, dtype=np.uint8)
        small[:] = img[:small.size]
        small[:8].view(np.int64)[0] = nslots
        gzp = os.path.join(tmp, "small.mem_map.gz")
        with gzip.open(gzp, "wb", compresslevel=1) as f:
            f.write(memoryview(small))
        t0 = time.time()
        t3 = ctx.load_table_file(gzp)
        gz_s = time.time() - t0
        t4 = ctx.table_from_image(small)
        same = (int(t3.info.num_signatures), int(t3.info.num_unreachable), int(t3.info.tail_run)) == \
               (int(t4.info.num_signatures), int(t4.info.num_unreachable), int(t4.info.tail_run))
        t3.free()
        t4.free()
        table.attach(ctx)
        if not same:
            raise SystemExit(".gz load and in-memory image load disagree")
        out["gz"] = {"inflated_bytes": int(small.size), "file_bytes": os.path.getsize(gzp), "load_s": round(gz_s, 3),
                     "inflated_GBps": round(small.size / gz_s / 1e9, 3), "note": "one zlib inflate thread feeds the same device-side parser"}
    finally:
        for fn in os.listdir(tmp):
            os.remove(os.path.join(tmp, fn))
        os.rmdir(tmp)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# configs[2]: 6-frame contig mode, G synthetic genomes of L bp against the replicated table
# ---------------------------------------------------------------------------------------------------------------------
def configs2(kg, ctx, table, u, plumb, genomes=50, length=5_000_000, steps=5, warmup=3, e2e=True, otable=None,
             parity_genomes=0, threads=1, log=lambda m: None):
    """The job's genomes are dealt to the ranks in contiguous blocks (contigs are independent units); every rank
    runs its block device-resident; rank 0 checks `parity_genomes` of ITS genomes against the oracle (needs otable)."""
    rank, world = plumb.rank, plumb.world
    per = (genomes + world - 1) // world
    g0, g1 = min(rank * per, genomes), min((rank + 1) * per, genomes)
    mine = g1 - g0
    params = kg.default_params()
    out = {"workload": f"configs[2]: {genomes} synthetic genomes x {length} bp, 6-frame mode, {table.info.num_signatures} signatures "
                       f"(table replicated, genomes dealt to {world} GPU(s))"}
    if mine == 0:
        ds = do = batch = None
        total = 0
    else:
        ds, do, total = bl.synth_genomes_range(ctx, u, g0, mine, length, seed=2)
        batch = ctx.batch_from_device(kg.MODE_DNA, ds, do, mine, total)
        for _ in range(warmup):
            ctx.run_batch(table, batch, params).free()
    plumb.barrier()
    t0 = time.perf_counter()
    st = None
    for _ in range(steps):
        if mine:
            r = ctx.run_batch(table, batch, params)
            st = r.stats
            r.free()
    plumb.barrier()
    dt = plumb.reduce([time.perf_counter() - t0], "MAX")[0] / steps
    lookups, hits, calls, pos = plumb.reduce([float(st.num_kmers if st else 0), float(st.num_hits if st else 0),
                                              float(st.num_calls if st else 0), float(st.num_positions if st else 0)])
    out.update({"n_gpus": world, "steps": steps, "ms_per_step": dt * 1e3, "mbp_per_s": genomes * length / dt / 1e6,
                "lookups_per_s": lookups / dt, "positions": int(pos), "lookups": int(lookups), "hits": int(hits), "calls": int(calls)})
    if st is not None and rank == 0:
        out["rank0_stage_ms"] = {"translate": round(st.ms_prepare, 3), "probe": round(st.ms_probe, 3), "group": round(st.ms_group, 3),
                                 "device": round(st.ms_device, 3), "probe_filter": round(st.ms_filter, 3),
                                 "probe_refilter": round(st.ms_refilter, 3), "probe_lines": round(st.ms_lines, 3)}
    if e2e and plumb.torch is not None:
        torch = plumb.torch
        if mine:
            h_seq = _pinned_copy(kg, ctx, torch, ds, total, pad=64)
            h_off = _pinned_copy(kg, ctx, torch, do, 8 * (mine + 1))
            for _ in range(warmup):
                ctx.run_ptr(table, kg.MODE_DNA, h_seq.data_ptr(), h_off.data_ptr(), mine, params).free()
        plumb.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            if mine:
                ctx.run_ptr(table, kg.MODE_DNA, h_seq.data_ptr(), h_off.data_ptr(), mine, params).free()
        plumb.barrier()
        edt = plumb.reduce([time.perf_counter() - t0], "MAX")[0] / steps
        raw = {"ms_per_step": edt * 1e3, "mbp_per_s": genomes * length / edt / 1e6, "lookups_per_s": lookups / edt,
               "h2d_bytes_per_step_per_gpu": int(total + 8 * (mine + 1)), "call": "kg_run: one byte per nucleotide"}
        # the library's ingest form: 2 bits per nucleotide + the positions of the non-ACGTU characters (kg_pack_dna; the packing
        # is host work outside the call, like FASTA parsing)
        pk_bytes = 0
        if mine:
            sb_np = np.frombuffer((C.c_uint8 * total).from_address(h_seq.data_ptr()), dtype=np.uint8)
            off_np = np.frombuffer((C.c_uint64 * (mine + 1)).from_address(h_off.data_ptr()), dtype=np.uint64)
            pk, boff, exc = kg.pack_dna(sb_np, off_np, threads=min(os.cpu_count() or 1, 16))
            h_pk = torch.empty(pk.nbytes, dtype=torch.uint8, pin_memory=True)
            h_pk.numpy()[:] = pk
            h_boff = torch.from_numpy(boff.view(np.int64)).pin_memory()
            h_exc = torch.from_numpy(np.ascontiguousarray(exc if len(exc) else np.zeros(1, np.uint64)).view(np.int64)).pin_memory()
            pk_bytes = int(boff[-1]) + 16 * (mine + 1) + 8 * len(exc)

            def pk_call():
                return ctx.run_packed_dna_ptr(table, h_pk.data_ptr(), h_off.data_ptr(), h_boff.data_ptr(), h_exc.data_ptr() if len(exc) else None,
                                              len(exc), mine, params)
            chk = pk_call()
            if st is not None and (chk.stats.num_kmers, chk.stats.num_hits, chk.stats.num_calls) != (st.num_kmers, st.num_hits, st.num_calls):
                raise SystemExit("configs2: kg_run_packed_dna disagrees with the device-resident run")
            chk.free()
            for _ in range(warmup):
                pk_call().free()
        plumb.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            if mine:
                pk_call().free()
        plumb.barrier()
        pdt = plumb.reduce([time.perf_counter() - t0], "MAX")[0] / steps
        out["e2e"] = {"ms_per_step": pdt * 1e3, "mbp_per_s": genomes * length / pdt / 1e6, "lookups_per_s": lookups / pdt,
                      "h2d_bytes_per_step_per_gpu": pk_bytes, "call": "kg_run_packed_dna: 2-bit nucleotides + exception list in, calls + compact OTU counts out",
                      "raw_bytes_call": raw}
    if parity_genomes and rank == 0 and otable is not None and mine:
        from oracle import kgo
        from tests.parity import assert_same
        n = min(parity_genomes, mine)
        off_all = bl.to_host(ctx, do, 8 * (n + 1)).view(np.uint64).copy()
        nh = nc = 0
        cpu_s = 0.0
        ref_kmers = 0
        block = max(1, min(threads, 10))  # a block of genomes at a time bounds the host memory of the two hit lists
        for b0 in range(0, n, block):
            b1 = min(b0 + block, n)
            off = off_all[b0:b1 + 1] - off_all[b0]
            sb = bl.to_host_at(ctx, ds, int(off_all[b0]), int(off[-1]))
            t0 = time.time()
            ref = kgo.run(otable, kgo.make_params(aa=False), sb, off, kgo.STREAM_JOIN, threads=min(threads, b1 - b0))
            cpu_s += time.time() - t0
            ref_kmers += ref.num_kmers
            g = ctx.run(table, kg.MODE_DNA, sb, off, kg.default_params(emit_hits=1))
            assert_same(g, ref, what=f"configs[2] genomes {b0}..{b1 - 1}")
            nh += len(ref.hits)
            nc += len(ref.calls)
            g.free()
            log(f"configs2 parity: genomes {b0}..{b1 - 1} ok")
        out["parity"] = f"bit-exact vs the CPU oracle on {n} of {genomes} genomes (hits, calls, OTU counts): {nh} hits, {nc} calls"
        out["cpu_port"] = {"lookups_per_s": ref_kmers / cpu_s, "threads": min(threads, block), "seconds": round(cpu_s, 2)}
    if batch is not None:
        batch.free()
        bl.device_free(ds)
        bl.device_free(do)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# configs[3]: a metagenome-scale batch of synthetic ORFs, sharded across the ranks, table replicated
# ---------------------------------------------------------------------------------------------------------------------
def configs3(kg, ctx, table, u, plumb, orfs=100_000_000, batch=1_000_000, otable=None, parity=10_000, threads=1, log=lambda m: None):
    """A fixed job (strong scaling): rank r generates proteins [r*per, (r+1)*per) ON THE DEVICE, one batch at a time (the
    ~30 GB of residues never cross PCIe), and runs every batch through kg_batch_run."""
    rank, world = plumb.rank, plumb.world
    params = kg.default_params()
    per = (orfs + world - 1) // world
    first, count = rank * per, max(0, min(per, orfs - rank * per))
    nb = (count + batch - 1) // batch

    def gen(b):
        n = min(batch, count - b * batch)
        ds, do, total = bl.synth_proteins(ctx, u, first + b * batch, n, seed=3)
        return ds, do, total, n

    if nb:
        ds, do, total, n = gen(0)
        bt = ctx.batch_from_device(kg.MODE_AA, ds, do, n, total)
        for _ in range(3):
            ctx.run_batch(table, bt, params).free()
        bt.free()
        bl.device_free(ds)
        bl.device_free(do)
    plumb.barrier()
    run_s = dev_ms = 0.0
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
    plumb.barrier()
    job_s = time.perf_counter() - t_job
    run_max, job_max = plumb.reduce([run_s, job_s], "MAX")
    tl, th, tc, tr = plumb.reduce([float(lookups), float(hits), float(calls), float(residues)])
    out = {"workload": f"configs[3]: {orfs} synthetic ORFs over {world} GPU(s), {table.info.num_signatures} signatures replicated",
           "n_gpus": world, "batches_per_rank": nb, "proteins_per_s": orfs / run_max, "lookups_per_s": tl / run_max,
           "run_seconds": round(run_max, 3), "job_seconds_incl_generation": round(job_max, 3), "residues": int(tr), "lookups": int(tl),
           "hits": int(th), "calls": int(tc), "rank0_device_ms_per_batch": round(dev_ms / max(nb, 1), 3),
           "scaling": "strong (a fixed job split over the ranks)"}
    if parity and rank == 0 and otable is not None and count:
        from oracle import kgo
        from tests.parity import assert_same
        rng = np.random.default_rng(3)
        picks = [("first of shard", first), ("random", first + int(rng.integers(0, max(count - parity, 1))))]
        notes = []
        for what, start in picks:
            n = min(parity, count)
            ds, do, total = bl.synth_proteins(ctx, u, start, n, seed=3)
            off = bl.to_host(ctx, do, 8 * (n + 1)).view(np.uint64).copy()
            sb = bl.to_host(ctx, ds, int(off[-1]))
            ref = kgo.run(otable, kgo.make_params(aa=True), sb, off, kgo.DIRECT_PROBE, threads=threads)
            g = ctx.run(table, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
            assert_same(g, ref, what=f"configs[3] {what}")
            notes.append(f"{what} ({start}..+{n}): {len(ref.hits)} hits, {len(ref.calls)} calls")
            g.free()
            bl.device_free(ds)
            bl.device_free(do)
        out["parity"] = "bit-exact vs the CPU oracle: " + "; ".join(notes)
    return out


# ---------------------------------------------------------------------------------------------------------------------
# configs[4]: hash-sharded table across the ranks, k-mers exchanged over NVLink inside the library (kmerguts_shard.h)
# ---------------------------------------------------------------------------------------------------------------------
C4_FAMILIES_PER_GPU = 1_410_000   # x 8 GPUs at keep = 700/1024: 2.005e9 distinct signatures (BASELINE configs[4]: 2 B)
C4_KEEP = 700


def configs4(kg, ctx, plumb, proteins=1_000_000, steps=10, warmup=3, families=0, keep=C4_KEEP, check=True, sample=2000,
             log=lambda m: None):
    """Weak scaling: every rank holds the signatures kg_shard_owner gives it (~250 M) and brings `proteins` sequences of its
    own.  N > 1 goes through kg_comm_init (an NCCL communicator inside the library); N = 1 has no interconnect."""
    torch, dist, rank, world = plumb.torch, plumb.dist, plumb.rank, plumb.world
    families = families or C4_FAMILIES_PER_GPU * world
    u = synth.Universe(n_families=families, sig_keep_per_1024=keep)
    t0 = time.time()
    dk, dp, nsig = bl.synth_signatures_sharded(ctx, u, rank, world)
    t1 = time.time()
    table = ctx.table_from_device_entries_sharded(dk, dp, nsig, rank, world)
    bl.device_free(dk)
    bl.device_free(dp)
    t2 = time.time()
    ti = table.info
    log(f"configs4: shard built, {nsig} signatures, {ti.device_bytes / 1e9:.2f} GB ({t1 - t0:.1f} s generate, {t2 - t1:.1f} s build)")
    uid = plumb.bcast_obj(kg.Comm.unique_id() if (rank == 0 and world > 1) else None)
    comm = kg.Comm(ctx, rank, world, uid)
    ds, do, total = bl.synth_proteins(ctx, u, rank * proteins, proteins, seed=1)
    batch = ctx.batch_from_device(kg.MODE_AA, ds, do, proteins, total)
    params = kg.default_params()
    for _ in range(max(warmup, 3)):
        comm.run(table, batch, params).free()
    plumb.barrier()
    t0 = time.perf_counter()
    lookups = 0
    phases = np.zeros(6)
    sent = 0
    for _ in range(steps):
        r = comm.run(table, batch, params)
        lookups += r.stats.num_kmers
        ss = comm.stats
        phases += [ss.ms_route, ss.ms_keys, ss.ms_answer, ss.ms_replies, ss.ms_merge, ss.ms_total]
        sent += ss.bytes_sent
        r.free()
    plumb.barrier()
    dt_max = plumb.reduce([time.perf_counter() - t0], "MAX")[0]
    tot_lookups, tot_sigs, tot_sent = plumb.reduce([float(lookups), float(nsig), float(sent)])
    phases /= steps
    out = {"workload": f"configs[4]: hash-sharded table, {int(tot_sigs)} signatures over {world} GPU(s), {proteins} proteins per rank",
           "n_gpus": world, "steps": steps, "ms_per_step": 1e3 * dt_max / steps, "lookups_per_s": tot_lookups / dt_max,
           "proteins_per_s": proteins * world * steps / dt_max, "signatures": int(tot_sigs), "shard_bytes_rank0": int(ti.device_bytes),
           "transport": "none" if world == 1 else
                        ("nccl send/recv of staged bins" if os.environ.get("KG_SHARD_TRANSPORT", "direct") != "direct" else
                         "direct: k_route / k_answer store into the peers' buffers (CUDA IPC over NVLink), in-band counts and flags; NCCL only for the bootstrap"),
           "scaling": "weak (proteins and signatures per GPU fixed)",
           "rank0_phase_ms": dict(zip(["route", "keys_exchange", "answer", "replies_exchange", "merge_and_fsm", "total_host"],
                                      [round(float(x), 3) for x in phases])),
           "interconnect_bytes_per_step": int(tot_sent / steps), "chunks": int(comm.stats.chunks)}
    if out["chunks"] > 1:
        out["rank0_phase_ms_note"] = ("the step runs in chunks whose exchanges overlap the kernels: route .. replies_exchange are END "
                                      "times since the start of the call")
    if check:
        res = comm.run(table, batch, kg.default_params(emit_hits=1))
        hits, calls, otus = res.hits, res.calls, res.otus
        my = (len(hits), bl.hits_checksum(ctx, hits, do), int(res.stats.num_kmers))
        contrib = np.zeros((world, 2), dtype=np.uint64)
        valid = 0
        for s in range(world):  # proteins of rank s against MY shard, the most naive way
            if s == rank:
                ds2, do2, tot2 = ds, do, total
            else:
                ds2, do2, tot2 = bl.synth_proteins(ctx, u, s * proteins, proteins, seed=1)
            v, h, ck = bl.naive_scan_aa(ctx, table, ds2, do2, proteins, tot2)
            contrib[s] = (h, ck)
            if s == rank:
                valid = v
            else:
                bl.device_free(ds2)
                bl.device_free(do2)
        tot = np.array(plumb.reduce(contrib.view(np.int64).ravel().tolist(), "SUM", integer=True), dtype=np.int64).view(np.uint64).reshape(world, 2)
        ok = int(tot[rank][0]) == my[0] and int(tot[rank][1]) == my[1] and valid == my[2]
        all_ok = plumb.reduce([1.0 if ok else 0.0], "MIN")[0]
        if not ok:
            print(f"rank {rank}: sharded run {my} vs naive scan over all shards {tuple(int(x) for x in tot[rank])}, valid {valid}", file=sys.stderr)
        if all_ok != 1.0:
            raise SystemExit("configs[4]: the hash-sharded run disagrees with the naive scan over all shards")
        tot_hits = plumb.reduce([float(my[0])])[0]
        out["parity_hits"] = (f"every rank: hits, lookups and (position, payload) checksum equal the naive scan summed over all {world} "
                              f"shards ({int(tot_hits)} hits in total)")
        from oracle import kgo
        kgo.build()
        n = min(sample, proteins)
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
    plumb.barrier()
    batch.free()
    comm.free()
    table.free()
    bl.device_free(ds)
    bl.device_free(do)
    return out
