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
params = kg.default_params(emit_hits=1)
ctx = kg.Context(0)
table = ctx.table_from_device_entries(dk, dp, nsig)
r = ctx.run_packed_aa(table, pk, goff, params)
h = r.hits
gc = np.bincount(h["seq"], minlength=N); oc = np.bincount(ref.hits["seq"], minlength=N)
bad = np.nonzero(gc != oc)[0]
print("affected seqs", len(bad), bad[:10], bad[-10:] if len(bad) else "")
for s in list(bad[:3]) + list(bad[-3:]):
    hh = h[h["seq"] == s]["pos"]; rr = ref.hits[ref.hits["seq"] == s]["pos"]
    print("seq", s, "len", int(off[s+1]-off[s]), "group start", int(goff[s]), "stream pos", int(goff[s])*8, "gpu", hh.tolist()[:12], "oracle", rr.tolist()[:12])
print("total groups", int(goff[-1]))
