import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, '.')
import kmergutsjava_b200 as kg
from tools import kg_synth as synth
fam, sigs = int(sys.argv[1]), int(sys.argv[2])
ctx = kg.Context(0)
u = synth.Universe(n_families=fam)
t = time.time(); dk, dp, n = kg.synth_signatures(ctx, u, sigs); print("sigs", n, time.time() - t, flush=True)
ns, dimg, disp = C.c_uint64(), C.c_void_p(), C.c_double()
t = time.time()
kg._check(kg.lib().kg_synth_reference_image(ctx._h, dk, dp, n, int(float(sys.argv[3]) * n) + 1, C.byref(ns), C.byref(dimg), C.byref(disp)))
print("image on device", ns.value, time.time() - t, flush=True)
t = time.time(); img = ctx.to_host(dimg.value, 24 + 24 * ns.value); print("to host", img.nbytes / 1e9, "GB", time.time() - t, flush=True)
