"""Shared helpers of the GPU parity tests: compare a kmergutsjava_b200.Result with an oracle Result, bit for bit."""
import numpy as np


def assert_same(res, ref, check_hits=True, what=""):
    calls, otus = res.calls, res.otus
    if check_hits:
        hits = res.hits
        assert len(hits) == len(ref.hits), f"{what}: {len(hits)} hits vs oracle {len(ref.hits)}"
        for f in ("seq", "sf", "pos", "oI", "avg", "fI"):
            assert np.array_equal(hits[f].astype(np.int64), ref.hits[f].astype(np.int64)), f"{what}: hit field {f}"
        assert np.array_equal(hits["wt"].view(np.uint32), ref.hits["wt"].view(np.uint32)), f"{what}: hit weights"
    assert len(calls) == len(ref.calls), f"{what}: {len(calls)} calls vs oracle {len(ref.calls)}"
    for f in ("seq", "sf", "start", "end", "count", "fI", "hits_before"):
        assert np.array_equal(calls[f].astype(np.int64), ref.calls[f].astype(np.int64)), f"{what}: call field {f}"
    # fp32 sum in list order: the reference's own additions, so bit-exact (tolerance would be 1e-6 relative)
    assert np.array_equal(calls["weighted"].view(np.uint32), ref.calls["weighted"].view(np.uint32)), f"{what}: weighted"
    assert len(otus) == len(ref.otus)
    assert np.array_equal(otus["n"], ref.otus["n"]), f"{what}: otu n"
    for j in range(5):
        m = otus["n"] > j
        assert np.array_equal(otus["count"][m, j], ref.otus["count"][m, j]), f"{what}: otu count[{j}]"
        assert np.array_equal(otus["oI"][m, j], ref.otus["oI"][m, j]), f"{what}: otu oI[{j}]"
    assert res.stats.num_kmers == ref.num_kmers, f"{what}: lookups {res.stats.num_kmers} vs {ref.num_kmers}"
