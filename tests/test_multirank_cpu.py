"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard planning, the timing/counter reductions bench.py
performs, and the result gather.  The per-rank 'device work' is stood in for by the CPU oracle on that rank's shard --
the point is that sharding by sequence with a replicated table reproduces the single-rank answer exactly."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmp):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import kgo
    from tools import kg_synth as synth, shard
    u = synth.Universe(n_families=80, seed=0x4B47000B)
    img = synth.build_table_image(*u.signatures())
    seqs = u.proteins(60, seed=9)
    sb, off = kgo.concat(seqs)
    cuts = shard.balanced_cuts(off, world)
    a, b = int(cuts[rank]), int(cuts[rank + 1])
    loc_off = (off[a:b + 1] - off[a]).astype(np.uint64)
    res = kgo.run(kgo.Table(data=img), kgo.make_params(aa=True), sb[int(off[a]):int(off[b])], loc_off)
    # reductions exactly as bench.py does them: SUM of lookups, MAX of time
    t = torch.tensor([float(res.num_kmers), float(len(res.calls))], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    tm = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(tm, op=dist.ReduceOp.MAX)
    gathered = [None] * world
    dist.all_gather_object(gathered, res.calls)
    if rank == 0:
        full = kgo.run(kgo.Table(data=img), kgo.make_params(aa=True), sb, off)
        merged = shard.merge_records(gathered, cuts)
        ok = (t[0].item() == full.num_kmers and t[1].item() == len(full.calls) and tm.item() == float(world)
              and all(np.array_equal(merged[f], full.calls[f]) for f in full.calls.dtype.names))
        open(os.path.join(tmp, "ok"), "w").write("1" if ok else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_rank(tmp_path, oracle):
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"


def test_balanced_cuts_properties():
    from tools import shard
    rng = np.random.default_rng(0)
    lens = rng.integers(0, 500, size=1000)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    for world in (1, 2, 4, 8):
        cuts = shard.balanced_cuts(off, world)
        assert cuts[0] == 0 and cuts[-1] == 1000 and np.all(np.diff(cuts) >= 0) and len(cuts) == world + 1
        sizes = np.diff(off[cuts].astype(np.int64))
        assert sizes.sum() == int(off[-1]) and sizes.max() - sizes.min() <= 2 * 500
    assert shard.weak_shard(3, 1000) == (3000, 1000)


def _shard_worker(rank, world, port, tmp):
    """The hash-sharded table's protocol (include/kmerguts_shard.h) restated on the CPU over gloo: route every valid 8-mer
    to kg_shard_owner(key), answer from the owner's shard with {index in the asker's bin, payload} for the hits only,
    put the replies back at their positions, run the FSM.  Must equal the single-rank oracle on the full table."""
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import kmergutsjava_b200 as kg
    from oracle import kgo
    from tools import kg_synth as synth
    u = synth.Universe(n_families=80, seed=0x4B47000C)
    keys, otu, avg, fi, wt = u.signatures()
    own = np.array([kg.shard_owner(int(k), world) for k in keys])
    mine = own == rank
    shard = {int(k): (int(o), int(a), int(f), float(w)) for k, o, a, f, w in zip(keys[mine], otu[mine], avg[mine], fi[mine], wt[mine])}
    seqs = u.proteins(40, seed=13)
    my_seqs = seqs[len(seqs) * rank // world: len(seqs) * (rank + 1) // world]
    # route: (sequence, position, key) of every window the reference enumerates (aa mode: i < len - 8, KGJ:912)
    bins = [[] for _ in range(world)]
    where = [[] for _ in range(world)]
    for si, s in enumerate(my_seqs):
        wk = synth.window_keys(synth.aa_codes(s))[:max(len(s) - 8, 0)]
        for pos, k in enumerate(wk):
            if k >= 0:
                o = kg.shard_owner(int(k), world)
                bins[o].append(int(k))
                where[o].append((si, pos))
    got = [None] * world
    dist.all_gather_object(got, bins)                  # keys exchange: got[src][dst]
    replies = [[(i, shard[k]) for i, k in enumerate(got[src][rank]) if k in shard] for src in range(world)]
    back = [None] * world
    dist.all_gather_object(back, replies)              # replies exchange: back[owner][asker]
    hits = []
    for owner in range(world):
        for idx, (o, a, f, w) in back[owner][rank]:
            si, pos = where[owner][idx]
            hits.append((si, 0, pos, o, a, f, w))
    hits.sort()
    hits = np.array(hits, dtype=kgo.HIT_DTYPE) if hits else np.zeros(0, dtype=kgo.HIT_DTYPE)
    calls = []
    for si in range(len(my_seqs)):
        c, _ = kgo.gather_hits(kgo.make_params(aa=True), hits[hits["seq"] == si])
        calls += [(si, int(x["start"]), int(x["end"]), int(x["count"]), int(x["fI"]), float(x["weighted"])) for x in c]
    sb, off = kgo.concat(my_seqs)
    img = synth.build_table_image(keys, otu, avg, fi, wt)
    full = kgo.run(kgo.Table(data=img), kgo.make_params(aa=True), sb, off, kgo.DIRECT_PROBE)
    want = [(int(x["seq"]), int(x["start"]), int(x["end"]), int(x["count"]), int(x["fI"]), float(x["weighted"])) for x in full.calls]
    ok = calls == want and len(hits) == len(full.hits) and np.array_equal(hits["pos"], full.hits["pos"]) and len(want) > 3
    sent = torch.tensor([float(sum(len(b) for b in bins)), float(sum(len(got[s][rank]) for s in range(world))), float(ok)], dtype=torch.float64)
    dist.all_reduce(sent, op=dist.ReduceOp.SUM)
    if rank == 0:
        open(os.path.join(tmp, "ok"), "w").write("1" if sent[0].item() == sent[1].item() and sent[2].item() == world else "0")
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_hash_sharded_protocol(tmp_path, oracle):
    port = 29900 + os.getpid() % 300
    mp.spawn(_shard_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert open(tmp_path / "ok").read() == "1"
