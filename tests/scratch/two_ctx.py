"""two host threads, one context each, packed end-to-end calls on a small workload (debug aid)"""
import sys, threading
sys.path.insert(0, ".")
import numpy as np
import kmergutsjava_b200 as kg
from tools import kg_synth as synth
from tools import kg_benchlib as bl
ctx = kg.Context(0)
u = synth.Universe(n_families=20000)
dk, dp, nsig = bl.synth_signatures(ctx, u, 2000000)
table = ctx.table_from_device_entries(dk, dp, nsig)
N = 60000
ds, do, total = bl.synth_proteins(ctx, u, 0, N, seed=1)
off = bl.to_host(ctx, do, 8 * (N + 1)).view(np.uint64).copy()
sb = bl.to_host(ctx, ds, int(off[-1]))
pk, goff = kg.pack_aa(sb, off, threads=4)
params = kg.default_params()
ctx2 = kg.Context(0)
table.attach(ctx2)
def lane(c):
    for _ in range(4):
        r = c.run_packed_aa(table, pk, goff, params)
        print("calls", r.stats.num_calls, flush=True)
        r.free()
th = [threading.Thread(target=lane, args=(c,)) for c in (ctx, ctx2)]
[t.start() for t in th]; [t.join() for t in th]
print("OK")
