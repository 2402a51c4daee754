"""first call on a fresh context vs later calls vs the oracle (debug aid)"""
import sys
sys.path.insert(0, ".")
import numpy as np
import kmergutsjava_b200 as kg
from tools import kg_synth as synth
from tools import kg_benchlib as bl
from oracle import kgo
ctx0 = kg.Context(0)
u = synth.Universe(n_families=20000)
dk, dp, nsig = bl.synth_signatures(ctx0, u, 2000000)
img = bl.synth_reference_image(ctx0, dk, dp, nsig, 3 * nsig + 1)
N = 60000
ds, do, total = bl.synth_proteins(ctx0, u, 0, N, seed=1)
off = bl.to_host(ctx0, do, 8 * (N + 1)).view(np.uint64).copy()
sb = bl.to_host(ctx0, ds, int(off[-1]))
pk, goff = kg.pack_aa(sb, off, threads=4)
ref = kgo.run(kgo.Table(borrow=img), kgo.make_params(aa=True), sb, off, kgo.DIRECT_PROBE, threads=8)
print("oracle calls", len(ref.calls), "hits", len(ref.hits))
params = kg.default_params(emit_hits=1)
for mode in ("raw", "packed"):
    ctx = kg.Context(0)
    table = ctx.table_from_device_entries(dk, dp, nsig)
    for it in range(3):
        r = ctx.run(table, kg.MODE_AA, sb, off, params) if mode == "raw" else ctx.run_packed_aa(table, pk, goff, params)
        c, h = r.calls, r.hits
        print(mode, it, "calls", len(c), "hits", len(h), "kmers", r.stats.num_kmers, ref.num_kmers)
        if len(c) != len(ref.calls):
            a = set(zip(c["seq"].tolist(), c["start"].tolist(), c["fI"].tolist()))
            b = set(zip(ref.calls["seq"].tolist(), ref.calls["start"].tolist(), ref.calls["fI"].tolist()))
            print("  only gpu", sorted(a - b)[:5], "only oracle", sorted(b - a)[:5])
            miss = sorted(b - a)[:1] + sorted(a - b)[:1]
            for (sq, stt, fi) in miss:
                print("  seq", sq, "len", int(off[sq + 1] - off[sq]), "slice-pos", int(off[sq]))
                hh = h[h["seq"] == sq]; rr = ref.hits[ref.hits["seq"] == sq]
                print("  gpu hits", len(hh), "oracle hits", len(rr))
        r.free()
    table.free(); ctx.close()
print("done")
