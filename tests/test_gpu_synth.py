"""The CUDA generators of the synthetic universe produce the very bytes of the numpy definition (tools/kg_synth.py),
and the device-written reference-format table image is a valid kmer.table.mem_map for the CPU oracle."""
import numpy as np
import pytest

from tests.parity import assert_same
from tools import kg_synth as synth

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
    dk, dp, n = kg.synth_signatures(ctx, u, max_sigs)
    assert n == len(keys)
    gk = ctx.to_host(dk, 8 * n).view(np.uint64)
    gp = ctx.to_host(dp, 16 * n).view(np.int32).reshape(n, 4)
    assert np.array_equal(gk.astype(np.int64), keys)
    assert np.array_equal(gp[:, 0], otu) and np.array_equal(gp[:, 1], avg) and np.array_equal(gp[:, 2], fi)
    assert np.array_equal(gp[:, 3].view(np.float32), wt)
    kg.device_free(dk)
    kg.device_free(dp)


def test_proteins_match_numpy(kg, ctx):
    u = synth.Universe(n_families=150, seed=0x4B470009)
    want = u.proteins(60, seed=7, first=1000)
    ds, do, total = kg.synth_proteins(ctx, u, 1000, 60, 7)
    off = ctx.to_host(do, 8 * 61).view(np.uint64)
    sb = ctx.to_host(ds, total)
    assert list(off) == list(np.concatenate([[0], np.cumsum([len(w) for w in want])]))
    assert sb.tobytes() == b"".join(want)
    kg.device_free(ds)
    kg.device_free(do)


def test_device_pipeline_on_device_generated_inputs(kg, ctx, oracle):
    """Generators -> kg_table_from_device_entries -> kg_batch_from_device -> kg_batch_run, checked against the oracle
    reading the device-written reference-format image."""
    u = synth.Universe(n_families=400, seed=0x4B47000A)
    dk, dp, n = kg.synth_signatures(ctx, u, 0)
    table = ctx.table_from_device_entries(dk, dp, n)
    img = kg.synth_reference_image(ctx, dk, dp, n, 3 * n + 1)
    ent = np.frombuffer(img, dtype=synth.ENTRY_DTYPE, offset=24)
    assert int((ent["which"] <= synth.MAX_ENCODED).sum()) == n and ent["which"][-1] > synth.MAX_ENCODED
    ds, do, total = kg.synth_proteins(ctx, u, 0, 2000, 3)
    off = ctx.to_host(do, 8 * 2001).view(np.uint64).copy()
    sb = ctx.to_host(ds, total).copy()
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
        kg.device_free(p)


def test_probe_roofline_runs(kg, ctx):
    r = ctx.probe_roofline(256 << 20, 1 << 24, 256, 4)
    assert r > 1e9
