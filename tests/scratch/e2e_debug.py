"""per-slice timeline of one end-to-end call (KG_DEBUG=1), configs[1]"""
import os, sys
sys.path.insert(0, ".")
os.environ["KG_DEBUG"] = "1"
import numpy as np, torch
import kmergutsjava_b200 as kg
from tools import kg_synth as synth, kg_benchlib as bl
ctx = kg.Context(0)
u = synth.Universe(n_families=2_000_000)
dk, dp, nsig = bl.synth_signatures(ctx, u, 200_000_000)
table = ctx.table_from_device_entries(dk, dp, nsig)
N = 1_000_000
ds, do, total = bl.synth_proteins(ctx, u, 0, N, seed=1)
h_seq = torch.empty(total + 64, dtype=torch.uint8, pin_memory=True)
h_off = torch.empty(N + 1, dtype=torch.int64, pin_memory=True)
kg._check(bl.lib().kg_device_to_host(ctx._h, h_seq.data_ptr(), ds, total))
kg._check(bl.lib().kg_device_to_host(ctx._h, h_off.data_ptr(), do, 8 * (N + 1)))
h_goff = torch.empty(N + 1, dtype=torch.int64, pin_memory=True)
kg._check(kg.lib().kg_pack_aa(h_seq.data_ptr(), h_off.data_ptr(), N, None, h_goff.data_ptr(), 1))
ng = int(h_goff[-1])
h_pk = torch.empty(5 * ng + 64, dtype=torch.uint8, pin_memory=True)
kg._check(kg.lib().kg_pack_aa(h_seq.data_ptr(), h_off.data_ptr(), N, h_pk.data_ptr(), h_goff.data_ptr(), 16))
p = kg.default_params()
for i in range(4):
    print("---- call", i, file=sys.stderr, flush=True)
    ctx.run_packed_aa_ptr(table, h_pk.data_ptr(), h_goff.data_ptr(), N, p).free()
