#!/usr/bin/env python
"""tests/java_pin/make_fsm_vectors.py <checkout of rsutormin/KmerGutsJava> -- FSM vectors whose expected output was PRINTED BY THE
REFERENCE'S OWN SOURCE (gatherHits / processSetOfHits / tabulateOtuDataForContig, KGJ:385-524, executed through
tests/java_pin/j2py.py), written to tests/golden/java_fsm_vectors.json in the layout of the hand-traced fsm_kats.json -- except
that a call's weighted score is the text Java's %f printed.  The CPU suite replays them through the oracle
(tests/test_oracle_kats.py), the GPU suite through the whole device path (tests/test_gpu_parity.py)."""
import io
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
OUT = os.path.join(ROOT, "tests", "golden", "java_fsm_vectors.json")


def cases():
    """seeded adversarial hit lists: sticky-but-switching function indices, singletons, gaps around max_gap, OTU churn over more
    than five indices, order-constraint offsets that sometimes break the diagonal, dyadic and non-dyadic fp32 weights"""
    rng = np.random.default_rng(20240)
    for case in range(64):
        n = int(rng.integers(1, 90))
        if case % 3 == 0:
            pos = np.unique(np.cumsum(rng.choice([1, 1, 2, 5, 190, 200, 201, 230], size=n)))
        else:
            pos = np.sort(rng.choice(2500, size=n, replace=False))
        n = len(pos)
        fI = rng.integers(1, 4, size=n)
        for i in range(1, n):
            if rng.random() < 0.75:
                fI[i] = fI[i - 1]
        oI = rng.integers(0, 9, size=n) if case % 2 else rng.choice([3, 3, 3, 5, 8], size=n)
        avg = 3000 - pos + rng.integers(-25, 25, size=n)
        wt = (rng.integers(1, 600, size=n) / 256.0).astype(np.float32) if case % 4 else rng.random(n).astype(np.float32)
        params = {"min_hits": int(rng.integers(2, 7)), "max_gap": int(rng.choice([0, 5, 50, 200, 1000])),
                  "min_weighted_hits": int(rng.integers(0, 4)), "order_constraint": int(rng.integers(2))}
        yield f"J{case:02d}", params, [[int(p), int(f), int(o), float(w), int(a)] for p, f, o, w, a in zip(pos, fI, oI, wt, avg)]


def java_output(mod, rt, params, hits):
    k = mod.KmerGutsJava()
    k.minHits, k.maxGap, k.minWeightedHits, k.orderConstraint = params["min_hits"], params["max_gap"], params["min_weighted_hits"], bool(params["order_constraint"])
    lst = rt.ArrayList()
    for pos, fi, oi, wt, avg in hits:
        x = mod.Hit()
        x.from0InProt, x.fI, x.oI, x.functionWt, x.avgOffFromEnd = pos, fi, oi, float(np.float32(wt)), avg
        lst.add(x)
    buf = io.StringIO()
    pw = rt.PrintWriter(buf)
    otu = rt.ArrayList()
    k.gatherHits(0, "+", 0, lst, rt.ArrayList(["F%d" % i for i in range(8)]), otu, pw)
    k.tabulateOtuDataForContig("v", 0, otu, pw)
    calls, otus = [], []
    for line in buf.getvalue().splitlines():
        f = line.split("\t")
        if f[0] == "CALL":
            calls.append([int(f[1]), int(f[2]), int(f[3]), int(f[4]), f[6]])
        elif f[0] == "OTU-COUNTS":
            otus = [[int(c), int(o)] for c, o in (x.split("-") for x in f[2:])]
    return calls, otus


def main():
    import tempfile
    import transliterated_pin as tp
    with tempfile.TemporaryDirectory() as d:
        mod, java_sha = tp.load_reference(os.path.join(sys.argv[1], tp.JAVA_REL), os.path.join(d, "kgj_transliterated.py"))
        rt = sys.modules["j2py_runtime"]
        out = []
        for name, params, hits in cases():
            calls, otu = java_output(mod, rt, params, hits)
            out.append({"name": name, "params": params, "hits": hits, "calls": calls, "otu": otu})
    doc = {"what": "expected output printed by the reference's own source (see tests/java_pin/make_fsm_vectors.py); hits are "
                   "[pos, fI, oI, weight, avgOffFromEnd], calls are [start, end, count, fI, weighted as printed by %f], otu is [[count, oI] ...]",
           "java_source_sha256": java_sha, "vectors": out}
    with open(OUT, "w") as f:
        json.dump(doc, f, separators=(",", ":"))
    print(len(out), "vectors,", sum(len(v["calls"]) for v in out), "calls,", os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
