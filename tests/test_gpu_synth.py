"""The CUDA generators of the synthetic universe produce the very bytes of the numpy definition (tools/kg_synth.py),
and the device-written reference-format table image is a valid kmer.table.mem_map for the CPU oracle."""
import numpy as np
import pytest

from tests.parity import assert_same
from tools import kg_synth as synth
from tools import kg_benchlib as bl  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def kg():
    import kmergutsjava_b200 as kg
    return kg


@pytest.fixture(scope="module")
def ctx(kg):
    c = kg.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("max_sigs", [0, 5000])
def test_signatures_match_numpy(kg, ctx, max_sigs):
    u = synth.Universe(n_families=150, seed=0x4B470009)
    keys, otu, avg, fi, wt = u.signatures(max_sigs or None)
    dk, dp, n = bl.synth_signatures(ctx, u, max_sigs)
    assert n == len(keys)
    gk = bl.to_host(ctx, dk, 8 * n).view(np.uint64)
    gp = bl.to_host(ctx, dp, 16 * n).view(np.int32).reshape(n, 4)
    assert np.array_equal(gk.astype(np.int64), keys)
    assert np.array_equal(gp[:, 0], otu) and np.array_equal(gp[:, 1], avg) and np.array_equal(gp[:, 2], fi)
    assert np.array_equal(gp[:, 3].view(np.float32), wt)
    bl.device_free(dk)
    bl.device_free(dp)


def test_proteins_match_numpy(kg, ctx):
    u = synth.Universe(n_families=150, seed=0x4B470009)
    want = u.proteins(60, seed=7, first=1000)
    ds, do, total = bl.synth_proteins(ctx, u, 1000, 60, 7)
    off = bl.to_host(ctx, do, 8 * 61).view(np.uint64)
    sb = bl.to_host(ctx, ds, total)
    assert list(off) == list(np.concatenate([[0], np.cumsum([len(w) for w in want])]))
    assert sb.tobytes() == b"".join(want)
    bl.device_free(ds)
    bl.device_free(do)


def test_device_pipeline_on_device_generated_inputs(kg, ctx, oracle):
    """Generators -> kg_table_from_device_entries -> kg_batch_from_device -> kg_batch_run, checked against the oracle
    reading the device-written reference-format image."""
    u = synth.Universe(n_families=400, seed=0x4B47000A)
    dk, dp, n = bl.synth_signatures(ctx, u, 0)
    table = ctx.table_from_device_entries(dk, dp, n)
    img = bl.synth_reference_image(ctx, dk, dp, n, 3 * n + 1)
    ent = np.frombuffer(img, dtype=synth.ENTRY_DTYPE, offset=24)
    assert int((ent["which"] <= synth.MAX_ENCODED).sum()) == n and ent["which"][-1] > synth.MAX_ENCODED
    ds, do, total = bl.synth_proteins(ctx, u, 0, 2000, 3)
    off = bl.to_host(ctx, do, 8 * 2001).view(np.uint64).copy()
    sb = bl.to_host(ctx, ds, total).copy()
    batch = ctx.batch_from_device(kg.MODE_AA, ds, do, 2000, total)
    res = ctx.run_batch(table, batch, kg.default_params(emit_hits=1))
    ref = oracle.run(oracle.Table(borrow=img), oracle.make_params(aa=True), sb, off, oracle.STREAM_JOIN, threads=4)
    assert len(ref.calls) > 500
    assert_same(res, ref, what="device inputs")
    # the loader accepts the device-written image too and gives the same answers
    t2 = ctx.table_from_image(img.tobytes())
    assert t2.info.num_signatures == n
    res2 = ctx.run(t2, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
    assert_same(res2, ref, what="device image through the loader")
    for x in (res, res2, batch, table, t2):
        x.free()
    for p in (dk, dp, ds, do):
        bl.device_free(p)


def test_probe_roofline_runs(kg, ctx):
    r = bl.probe_roofline(ctx, 256 << 20, 1 << 24, 256, 4)
    assert r > 1e9


def test_full_size_properties_against_naive_kernel(kg, ctx):
    """Size-independent parity at a size the CPU oracle would take minutes for: the pipeline's lookup count, hit count and
    an order-independent checksum over (position, payload) of every hit must equal those of the naive one-thread-per-
    position kernel (byte-wise residue reads, no prefilter, no queue, full table lookup).  Also: running twice changes nothing."""
    u = synth.Universe(n_families=40000, seed=0x4B47000C)
    dk, dp, n = bl.synth_signatures(ctx, u, 0)
    table = ctx.table_from_device_entries(dk, dp, n)
    nprot = 100000
    ds, do, total = bl.synth_proteins(ctx, u, 0, nprot, 5)
    valid, hits, chk = bl.naive_scan_aa(ctx, table, ds, do, nprot, total)
    assert valid > 0.9 * total - 9 * nprot and hits > 0.05 * valid
    batch = ctx.batch_from_device(kg.MODE_AA, ds, do, nprot, total)
    first = None
    for _ in range(2):
        res = ctx.run_batch(table, batch, kg.default_params(emit_hits=1))
        st = res.stats
        assert (st.num_kmers, st.num_hits) == (valid, hits)
        h = res.hits
        assert len(h) == hits
        assert bl.hits_checksum(ctx, h, do) == chk
        key = (h["seq"].astype(np.int64) << 32) | h["pos"].astype(np.int64)
        assert np.all(np.diff(key) > 0)                       # sorted by (sequence, position), no duplicates
        calls = res.calls
        ck = (calls["seq"].astype(np.int64) << 32) | calls["start"].astype(np.int64)
        assert np.all(np.diff(ck) > 0) and np.all(calls["count"] >= 5) and np.all(calls["end"] >= calls["start"] + 7)
        sig = (len(calls), int(calls["count"].sum()), float(calls["weighted"].astype(np.float64).sum()), int(res.otus["n"].sum()))
        assert first is None or sig == first                  # idempotent
        first = sig
        res.free()
    # the naive scan does not care whether the batch has been patched (last residue -> 0) or not
    assert bl.naive_scan_aa(ctx, table, ds, do, nprot, total) == (valid, hits, chk)
    batch.free()
    table.free()
    for p in (dk, dp, ds, do):
        bl.device_free(p)
