"""GPU tests of the hash-sharded table (include/kmerguts_shard.h, BASELINE.json configs[4]): every rank holds the keys
that hash to it, k-mers travel to their owner and the hits travel back.  Results must equal the CPU oracle (and the
replicated-table path) bit for bit, for any number of ranks.  Ranks live in one process here (peer copies, one GPU is
enough); test_nccl_two_processes drives the NCCL transport when the box has two GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.parity import assert_same
from tools import kg_synth as synth
from tools import kg_benchlib as bl  # noqa: E402

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def kg():
    import kmergutsjava_b200 as kg
    return kg


@pytest.fixture(scope="module")
def ctx(kg):
    c = kg.Context(0)
    yield c
    c.close()


@pytest.fixture(autouse=True, params=["direct", "staged"])
def transport(request, monkeypatch):
    """Both exchanges: `direct` = k_route / k_answer store straight into the peer's buffer (the default), `staged` = bins that
    are moved afterwards (peer copies here, NCCL send/recv between processes)."""
    if request.param == "staged":
        monkeypatch.setenv("KG_SHARD_TRANSPORT", "copy")
    else:
        monkeypatch.delenv("KG_SHARD_TRANSPORT", raising=False)
    return request.param


@pytest.fixture(scope="module")
def universe():
    u = synth.Universe(n_families=300, seed=0x4B470003)
    keys, otu, avg, fi, wt = u.signatures()
    return u, synth.build_table_image(keys, otu, avg, fi, wt), keys


def split(seqs, nranks, cuts=None):
    """Contiguous shares of the sequences, one per rank (a share may be empty)."""
    n = len(seqs)
    cuts = cuts or [round(n * r / nranks) for r in range(nranks + 1)]
    return [seqs[cuts[r]:cuts[r + 1]] for r in range(nranks)]


def run_local(kg, ctx, oracle, img, shares, mode, params, nranks):
    comms = kg.Comm.local([ctx] * nranks)
    tables = [ctx.table_from_image_sharded(img, r, nranks) for r in range(nranks)]
    batches = []
    for sh in shares:
        sb, off = oracle.concat(sh)
        batches.append(ctx.upload(mode, sb, off))
    results = kg.run_sharded_local(comms, tables, batches, params)
    stats = [c.stats for c in comms]
    return comms, tables, batches, results, stats


def free_all(*groups):
    for g in groups:
        for x in g:
            x.free()


@pytest.mark.parametrize("nranks", [2, 3, 8])
def test_shards_partition_the_table(kg, ctx, universe, nranks):
    u, img, keys = universe
    tables = [ctx.table_from_image_sharded(img, r, nranks) for r in range(nranks)]
    sizes = [t.info.num_signatures for t in tables]
    want = np.bincount([kg.shard_owner(int(k), nranks) for k in keys], minlength=nranks)
    assert sizes == list(want) and sum(sizes) == len(keys)
    free_all(tables)


def test_sharded_generator_is_a_partition(kg, ctx):
    """kg_synth_signatures_sharded: the shards' signatures are disjoint, owned by their rank, and their union is the
    unsharded generator's output (same payloads)."""
    u = synth.Universe(n_families=2000, seed=0x4B470009)
    dk, dp, n = bl.synth_signatures(ctx, u, 0)
    keys = bl.to_host(ctx, dk, 8 * n).view(np.uint64).copy()
    pay = bl.to_host(ctx, dp, 16 * n).view(np.uint32).reshape(n, 4).copy()
    whole = {int(k): tuple(p) for k, p in zip(keys, pay)}
    bl.device_free(dk)
    bl.device_free(dp)
    nranks, seen = 3, {}
    for r in range(nranks):
        dk, dp, m = bl.synth_signatures_sharded(ctx, u, r, nranks)
        ks = bl.to_host(ctx, dk, 8 * m).view(np.uint64).copy()
        ps = bl.to_host(ctx, dp, 16 * m).view(np.uint32).reshape(m, 4).copy()
        assert all(kg.shard_owner(int(k), nranks) == r for k in ks[:500])
        for k, p in zip(ks, ps):
            assert int(k) not in seen
            seen[int(k)] = tuple(p)
        t = ctx.table_from_device_entries_sharded(dk, dp, m, r, nranks)
        assert t.info.num_signatures == m
        t.free()
        bl.device_free(dk)
        bl.device_free(dp)
    assert seen == whole and len(whole) > 100000


@pytest.mark.parametrize("nranks", [1, 2, 3, 5])
@pytest.mark.parametrize("flags", [dict(), dict(order_constraint=1), dict(min_hits=3, max_gap=50, min_weighted_hits=2)])
def test_aa_parity_sharded(kg, ctx, oracle, universe, nranks, flags):
    u, img, _ = universe
    seqs = u.proteins(600, seed=77) + [b"", b"A", b"ACDEFGHI", b"ACDEFGHIK", b"acdefghiklmnp", b"ACDEFGHIKXLMNPQRSTVWY", b""]
    shares = split(seqs, nranks)
    params = kg.default_params(emit_hits=1, **flags)
    comms, tables, batches, results, stats = run_local(kg, ctx, oracle, img, shares, kg.MODE_AA, params, nranks)
    otable = oracle.Table(data=img)
    total_calls = 0
    for r in range(nranks):
        sb, off = oracle.concat(shares[r])
        ref = oracle.run(otable, oracle.make_params(aa=True, **flags), sb, off, oracle.DIRECT_PROBE)
        assert_same(results[r], ref, what=f"aa shard {r}/{nranks} {flags}")
        total_calls += len(ref.calls)
        assert stats[r].keys_sent == ref.num_kmers and stats[r].replies_received == len(ref.hits)
    assert total_calls > 20
    assert sum(s.keys_sent for s in stats) == sum(s.keys_received for s in stats)
    assert sum(s.replies_sent for s in stats) == sum(s.replies_received for s in stats)
    if nranks > 1:
        assert sum(s.keys_remote for s in stats) > 0 and sum(s.bytes_sent for s in stats) > 0
    free_all(results, batches, tables, comms)


@pytest.mark.parametrize("nranks", [2, 4])
def test_dna_parity_sharded(kg, ctx, oracle, universe, nranks):
    u, img, _ = universe
    seqs = [synth.genome(u, 30000, seed=31, index=i) for i in range(3)] + [
        b"", b"AC", b"ATG", synth.genome(u, 3001, seed=32), synth.genome(u, 4097 * 3, seed=34), b"acgtnACGTNryk" * 40]
    shares = split(seqs, nranks)
    params = kg.default_params(emit_hits=1)
    comms, tables, batches, results, _ = run_local(kg, ctx, oracle, img, shares, kg.MODE_DNA, params, nranks)
    otable = oracle.Table(data=img)
    for r in range(nranks):
        sb, off = oracle.concat(shares[r])
        ref = oracle.run(otable, oracle.make_params(aa=False), sb, off, oracle.DIRECT_PROBE)
        assert_same(results[r], ref, what=f"dna shard {r}/{nranks}")
    free_all(results, batches, tables, comms)


def same_records(a, b, what):
    for name in ("hits", "calls", "otus"):
        x, y = getattr(a, name), getattr(b, name)
        assert x.tobytes() == y.tobytes(), f"{what}: {name} differ ({len(x)} vs {len(y)})"
    assert a.stats.num_kmers == b.stats.num_kmers and a.stats.num_hits == b.stats.num_hits


def test_sharded_equals_replicated(kg, ctx, oracle, universe):
    """A larger batch (many tiles per bin, uneven shares, one empty rank): every rank's records are byte-identical to what
    the replicated table gives for the same sequences."""
    u, img, _ = universe
    seqs = u.proteins(20000, seed=5)
    nranks = 4
    shares = split(seqs, nranks, cuts=[0, 9000, 9000, 12000, 20000])
    params = kg.default_params(emit_hits=1)
    comms, tables, batches, results, stats = run_local(kg, ctx, oracle, img, shares, kg.MODE_AA, params, nranks)
    full = ctx.table_from_image(img)
    for r in range(nranks):
        sb, off = oracle.concat(shares[r])
        rep = ctx.run(full, kg.MODE_AA, sb, off, params)
        same_records(results[r], rep, f"rank {r}")
        rep.free()
    assert results[1].stats.num_sequences == 0 and stats[1].keys_sent == 0 and stats[1].keys_received > 0
    # a second step on the same communicators (buffers and capacities are reused)
    again = kg.run_sharded_local(comms, tables, batches, params)
    for r in range(nranks):
        same_records(again[r], results[r], f"second step, rank {r}")
    free_all(again, results, batches, tables, comms)
    full.free()


def test_skewed_bins(kg, ctx, oracle):
    """Low-complexity input: every window of a homopolymer is the same key, so one owner's bin takes the whole batch and
    the first-guess capacity overflows; the run repeats the routing pass with the exact size."""
    keys = np.array([0, 1, 20 ** 8 - 1], np.uint64)  # AAAAAAAA, AAAAAAAC, YYYYYYYY
    img = synth.build_table_image(keys, [1, 2, 3], [10, 20, 30], [7, 7, 9], np.array([0.5, 1.0, 2.0], np.float32))
    seqs = [b"A" * 60000, b"Y" * 30000 + b"C", b"ACDEFGHIKLMNPQRSTVWY" * 50, b"A" * 9 + b"C"]
    nranks = 4
    shares = [seqs[:1], seqs[1:2], seqs[2:], []]
    params = kg.default_params(emit_hits=1)
    comms, tables, batches, results, stats = run_local(kg, ctx, oracle, img, shares, kg.MODE_AA, params, nranks)
    otable = oracle.Table(data=img)
    for r in range(nranks):
        sb, off = oracle.concat(shares[r])
        ref = oracle.run(otable, oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
        assert_same(results[r], ref, what=f"skewed shard {r}")
    assert len(results[0].hits) == 60000 - 8  # the last window of a protein is never enumerated (KGJ:912)
    free_all(results, batches, tables, comms)


def test_shard_misuse_is_an_error(kg, ctx, oracle, universe):
    u, img, _ = universe
    sb, off = oracle.concat(u.proteins(10, seed=1))
    shard = ctx.table_from_image_sharded(img, 0, 2)
    full = ctx.table_from_image(img)
    with pytest.raises(kg.KgError, match="shard 0 of 2"):
        ctx.run(shard, kg.MODE_AA, sb, off, kg.default_params())
    comm = kg.Comm(ctx, 0, 1)
    b = ctx.upload(kg.MODE_AA, sb, off)
    with pytest.raises(kg.KgError, match="shard 0 of 2"):
        comm.run(shard, b, kg.default_params())
    r = comm.run(full, b, kg.default_params(emit_hits=1))  # one rank: the whole path without an interconnect
    rep = ctx.run(full, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
    same_records(r, rep, "single-rank communicator")
    with pytest.raises(kg.KgError):
        ctx.table_from_image_sharded(img, 2, 2)
    with pytest.raises(kg.KgError):
        kg.Comm(ctx, 0, 17)
    free_all([r, rep, b, comm, shard, full])


def _gpu_count():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _nccl_two_processes(kg, transport):
    """The NCCL transport: two processes, one GPU each, the id handed over through a file.  Only DEFINED as a test on a box
    with two GPUs: NCCL refuses two ranks on one device ("Duplicate GPU detected"), and two processes that wait for each
    other's kernels on one GPU are ruled out by the profiling guide (Xid 109), so a one-GPU box covers the exchange with
    virtual ranks in one process (above) and bench.py's configs4 leg covers the NCCL transport at N >= 2."""
    import tempfile
    env = dict(os.environ)
    if env.get("KG_SHARD_TRANSPORT") == "copy":
        env["KG_SHARD_TRANSPORT"] = "nccl"
    variants = [env]
    if transport == "direct":   # the peers' buffers cannot be mapped (simulated): every rank falls back to the staged transport
        variants.append(dict(env, KG_SHARD_FAIL_IPC="1"))
    for env in variants:
        _two_processes_once(env)


def _two_processes_once(env):
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        procs = [subprocess.Popen([sys.executable, os.path.join(ROOT, "tests", "shard_nccl_worker.py"), str(r), "2", d],
                                  cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
        outs = []
        for p in procs:
            try:
                out, _ = p.communicate(timeout=300)
            except subprocess.TimeoutExpired:
                for q in procs:
                    q.kill()
                raise
            outs.append(out)
        for r, p in enumerate(procs):
            assert p.returncode == 0, f"rank {r}:\n{outs[r][-3000:]}"
            assert "OK" in outs[r]


if _gpu_count() >= 2:
    test_nccl_two_processes = _nccl_two_processes
