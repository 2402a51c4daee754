#!/usr/bin/env python
"""tools/gen_reference_inputs.py -- writes the synthetic inputs of BASELINE.json configs[1] in the REFERENCE's own formats
into a directory, for bench.py's `--impl reference` arm:

    table.img   the signatures as a kmer.table.mem_map image (24-byte header + 24-byte slots, KGJ:933-935, 995-999)
    seq.bin     the proteins' residues, concatenated (one byte per residue)
    off.bin     uint64 offsets (n + 1)

It runs as a SEPARATE process: the generators live on the GPU (tools/benchlib, which needs the product library's context),
and the reference arm's own process must load nothing of the product -- it only reads these files.  Prints one JSON line."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--families", type=int, default=2_000_000)
    ap.add_argument("--sigs", type=int, default=200_000_000)
    ap.add_argument("--proteins", type=int, default=1_000_000)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args()
    import kmergutsjava_b200 as kg
    from tools import kg_benchlib as bl
    from tools import kg_synth as synth
    ctx = kg.Context(a.device)
    u = synth.Universe(n_families=a.families)
    dk, dp, nsig = bl.synth_signatures(ctx, u, a.sigs)
    # load 1/3: at 1/2 the reference hash (key % numSigs) clusters so badly that no prime near 2n avoids running off the end
    img = bl.synth_reference_image(ctx, dk, dp, nsig, 3 * nsig + 1)
    num_slots = int(img[:8].view(np.int64)[0])
    bl.device_free(dk)
    bl.device_free(dp)
    with open(os.path.join(a.out, "table.img"), "wb") as f:
        step = 256 << 20
        for o in range(0, img.nbytes, step):
            f.write(memoryview(img[o:o + step]))
    del img
    ds, do, total = bl.synth_proteins(ctx, u, 0, a.proteins, seed=1)
    off = bl.to_host(ctx, do, 8 * (a.proteins + 1)).view(np.uint64).copy()
    sb = bl.to_host(ctx, ds, int(off[-1]))
    sb.tofile(os.path.join(a.out, "seq.bin"))
    off.tofile(os.path.join(a.out, "off.bin"))
    bl.device_free(ds)
    bl.device_free(do)
    ctx.close()
    print(json.dumps({"signatures": int(nsig), "num_slots": num_slots, "residues": int(total), "proteins": a.proteins}))


if __name__ == "__main__":
    main()
