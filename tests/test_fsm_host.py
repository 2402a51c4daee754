"""The product's run state machines on the CPU.  kmergutsjava_b200/csrc/kg_fsm.cuh -- the header the CUDA kernels k_fsm,
k_fsm_seg and k_otu_fold are compiled from (KgFsm, KgFsmSeg, the closed-form kg_otu_update_n) -- is compiled for the host by
tests/fsm_host/fsm_host.cpp and driven with thousands of adversarial hit lists: against the oracle, against the vectors the
reference's own Java source printed (tests/golden/java_fsm_vectors.json), against the hand-traced KATs and, where the reference
checkout is present, against the transliterated Java source directly.  Both paths: one FSM per sequence, and the segment path
(containers cut at gaps > max_gap, OTU runs folded afterwards) -- whose correctness rests on the argument in kg_fsm.cuh that
the open run is always empty after such a gap.  This is NOT the GPU binary (tests/test_gpu_parity.py runs the same cases on the
device); it is the same source through another compiler, which is what lets the CPU suite fuzz it this hard."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
CUDA_INC = os.path.join(os.environ.get("CUDA_HOME", "/usr/local/cuda"), "include")
pytestmark = pytest.mark.skipif(shutil.which("g++") is None or not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h")),
                                reason="needs g++ and the CUDA headers")


class HostCall(C.Structure):
    _fields_ = [("container", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("count", C.c_int32), ("fI", C.c_int32),
                ("weighted", C.c_float), ("hits_before", C.c_int32)]


@pytest.fixture(scope="module")
def fsm(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("fsm_host") / "libfsm_host.so")
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-Wall", "-Wno-unknown-pragmas",
                    "-I", CUDA_INC, "-o", so, os.path.join(ROOT, "tests", "fsm_host", "fsm_host.cpp")], check=True)
    lib = C.CDLL(so)
    i32p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_float)
    lib.fsm_host_run.restype = C.c_int
    lib.fsm_host_run.argtypes = [C.c_int] * 4 + [C.c_float, C.c_int, C.c_int, i32p, i32p, i32p, i32p, i32p, f32p,
                                 C.POINTER(HostCall), C.c_int, i32p, i32p, i32p]

    def run(path, flags, ncontainers, hits):
        """hits: structured array with container, pos, fI, avg, oI, wt -> (calls as tuples, otu as [[count, oI] ...])"""
        n = len(hits)
        cols = {k: np.ascontiguousarray(hits[k], dtype=np.int32) for k in ("container", "pos", "fI", "avg", "oI")}
        wt = np.ascontiguousarray(hits["wt"], dtype=np.float32)
        calls = (HostCall * 8192)()
        on = C.c_int32()
        oc, oo = (C.c_int32 * 5)(), (C.c_int32 * 5)()
        nc = lib.fsm_host_run(path, flags.get("min_hits", 5), flags.get("max_gap", 200), int(flags.get("order_constraint", 0)),
                              float(flags.get("min_weighted_hits", 0)), ncontainers, n,
                              *(cols[k].ctypes.data_as(i32p) for k in ("container", "pos", "fI", "avg", "oI")), wt.ctypes.data_as(f32p),
                              calls, 8192, C.byref(on), oc, oo)
        assert nc <= 8192
        got = [(c.container, c.start, c.end, c.count, c.fI, np.float32(c.weighted).tobytes(), c.hits_before) for c in calls[:nc]]
        return got, [[oc[j], oo[j]] for j in range(on.value)]
    return run


HITS = np.dtype([("container", "<i4"), ("pos", "<i4"), ("fI", "<i4"), ("avg", "<i4"), ("oI", "<i4"), ("wt", "<f4")])


def _oracle(oracle, flags, ncontainers, hits):
    calls, otu = [], None
    for k in range(ncontainers):
        h = hits[hits["container"] == k]
        rec = np.zeros(len(h), dtype=oracle.HIT_DTYPE)
        rec["pos"], rec["fI"], rec["oI"], rec["avg"], rec["wt"] = h["pos"], h["fI"], h["oI"], h["avg"], h["wt"]
        cs, otu = oracle.gather_hits(oracle.make_params(aa=True, **flags), rec, otu=otu, max_calls=1 << 14)
        calls += [(k, int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]), np.float32(c["weighted"]).tobytes(), int(c["hits_before"]))
                  for c in cs]
    if otu is None:
        return calls, []
    return calls, [[int(otu["count"][0][j]), int(otu["oI"][0][j])] for j in range(int(otu["n"][0]))]


def _random_container(rng, style):
    n = int(rng.integers(0, 160))
    if style == 0:
        pos = np.sort(rng.choice(4000, size=n, replace=False)) if n else np.zeros(0, int)
    elif style == 1:
        pos = np.unique(np.cumsum(rng.choice([1, 1, 2, 5, 190, 200, 201, 230], size=n))) if n else np.zeros(0, int)
    else:   # long dense runs with rare gaps right at the limit
        pos = np.unique(np.cumsum(rng.choice([1, 1, 1, 3, 49, 50, 51, 199, 200, 201], size=n, p=[.3, .2, .15, .15, .03, .03, .03, .04, .04, .03]))) if n else np.zeros(0, int)
    n = len(pos)
    fI = rng.integers(1, 5, size=n)
    for i in range(1, n):
        if rng.random() < 0.7:
            fI[i] = fI[i - 1]
    oI = rng.integers(0, 9, size=n) if rng.random() < 0.6 else rng.choice([3, 3, 3, 5, 8], size=n)
    avg = (5000 - pos + rng.integers(-25, 25, size=n)) if n else np.zeros(0, int)
    wt = (rng.integers(1, 600, size=n) / 256.0).astype(np.float32) if rng.random() < 0.5 else rng.random(n).astype(np.float32)
    return pos, fI, avg, oI, wt


@pytest.mark.parametrize("seed", range(8))
def test_product_fsm_header_vs_oracle_fuzz(fsm, oracle, seed):
    """250 sequences per seed, one or six containers each, random flags: both product paths == the oracle, including the
    fp32 weighted sums bit for bit, hits_before (the -d interleave) and the OTU buffer carried across the six frames"""
    rng = np.random.default_rng(7000 + seed)
    ncalls = 0
    for case in range(250):
        ncont = 1 if case % 2 else 6
        parts = []
        for k in range(ncont):
            pos, fI, avg, oI, wt = _random_container(rng, case % 3)
            h = np.zeros(len(pos), dtype=HITS)
            h["container"], h["pos"], h["fI"], h["avg"], h["oI"], h["wt"] = k, pos, fI, avg, oI, wt
            parts.append(h)
        hits = np.concatenate(parts)
        flags = dict(order_constraint=bool(rng.integers(2)), min_hits=int(rng.integers(2, 7)), min_weighted_hits=int(rng.integers(0, 4)),
                     max_gap=int(rng.choice([0, 5, 50, 200, 1000])))
        want = _oracle(oracle, flags, ncont, hits)
        for path in (0, 1):
            assert fsm(path, flags, ncont, hits) == want, (seed, case, path, flags)
        ncalls += len(want[0])
    assert ncalls > 300


def test_product_fsm_header_on_the_java_printed_vectors_and_kats(fsm, oracle):
    vectors = json.load(open(os.path.join(GOLD, "java_fsm_vectors.json")))["vectors"]
    kats = json.load(open(os.path.join(GOLD, "fsm_kats.json")))
    for vec in vectors + kats:
        h = np.zeros(len(vec["hits"]), dtype=HITS)
        for i, (pos, fI, oI, wt, avg) in enumerate(vec["hits"]):
            h[i] = (0, pos, fI, avg, oI, wt)
        for path in (0, 1):
            calls, otu = fsm(path, vec["params"], 1, h)
            got = [[c[1], c[2], c[3], c[4], oracle.java_format_f(float(np.frombuffer(c[5], np.float32)[0]))] for c in calls]
            want = [[a, b, c, d, w if isinstance(w, str) else oracle.java_format_f(float(np.float32(w)))] for a, b, c, d, w in vec["calls"]]
            assert got == want and otu == vec["otu"], (vec["name"], path)


def test_product_fsm_header_cap_and_wraparound(fsm, oracle):
    """Q9: 39 998 hits in an open run, the pair-switch test still running on hits that were not appended; and positions near
    2^31 where Java's int arithmetic in the gap test wraps (KGJ:477)"""
    n = 41000
    h = np.zeros(n + 12, dtype=HITS)
    h["pos"] = np.arange(n + 12)
    h["fI"][:n], h["fI"][n:n + 2], h["fI"][n + 2:] = 7, 9, 7
    h["oI"] = np.arange(n + 12) % 3
    h["avg"] = 50000 - np.arange(n + 12)
    h["wt"] = np.float32(0.5)
    for flags in (dict(), dict(order_constraint=True)):
        want = _oracle(oracle, flags, 1, h)
        assert want[0][0][3] == 39998
        for path in (0, 1):
            assert fsm(path, flags, 1, h) == want
    big = np.zeros(14, dtype=HITS)
    big["pos"] = [2147483000, 2147483100, 2147483200, 2147483300, 2147483400, 2147483500, 2147483600, 2147483640, 2147483641, 2147483642,
                  2147483643, 2147483644, 2147483645, 2147483646]
    big["fI"], big["oI"], big["avg"], big["wt"] = 3, 1, 10, np.float32(1.0)
    for flags in (dict(max_gap=200), dict(max_gap=1000), dict(max_gap=2147483647)):
        want = _oracle(oracle, flags, 1, big)
        for path in (0, 1):
            assert fsm(path, flags, 1, big) == want, flags


REF_JAVA = os.path.join(os.environ.get("KG_REFERENCE", "/root/reference"), "lib", "src", "kmergutsjava", "KmerGutsJava.java")


@pytest.mark.skipif(not os.path.exists(REF_JAVA), reason="the reference checkout is not on this box")
def test_product_fsm_header_vs_the_java_source(fsm, tmp_path):
    """No oracle in between: the product's FSM source against the reference's gatherHits / processSetOfHits executed through
    tests/java_pin/j2py.py, six containers sharing one OTU buffer as processSeq does (KGJ:538-558)"""
    import importlib.util
    import io
    import sys
    spec = importlib.util.spec_from_file_location("transliterated_pin", os.path.join(ROOT, "tests", "java_pin", "transliterated_pin.py"))
    tp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tp)
    mod, _ = tp.load_reference(REF_JAVA, str(tmp_path / "kgj_transliterated.py"))
    rt = sys.modules["j2py_runtime"]
    rng = np.random.default_rng(31337)
    ncalls = 0
    for case in range(120):
        ncont = 1 if case % 2 else 6
        flags = dict(order_constraint=bool(rng.integers(2)), min_hits=int(rng.integers(2, 7)), min_weighted_hits=int(rng.integers(0, 4)),
                     max_gap=int(rng.choice([0, 5, 50, 200, 1000])))
        k = mod.KmerGutsJava()
        k.minHits, k.maxGap, k.minWeightedHits, k.orderConstraint = flags["min_hits"], flags["max_gap"], flags["min_weighted_hits"], flags["order_constraint"]
        otu, buf = rt.ArrayList(), io.StringIO()
        pw = rt.PrintWriter(buf)
        parts, text = [], []
        for c in range(ncont):
            pos, fI, avg, oI, wt = _random_container(rng, case % 3)
            h = np.zeros(len(pos), dtype=HITS)
            h["container"], h["pos"], h["fI"], h["avg"], h["oI"], h["wt"] = c, pos, fI, avg, oI, wt
            parts.append(h)
            lst = rt.ArrayList()
            for r in h:
                x = mod.Hit()
                x.from0InProt, x.fI, x.oI, x.avgOffFromEnd, x.functionWt = int(r["pos"]), int(r["fI"]), int(r["oI"]), int(r["avg"]), float(r["wt"])
                lst.add(x)
            pw.println(f"CONTAINER {c}")
            k.gatherHits(0, "+", 0, lst, rt.ArrayList(["F%d" % i for i in range(6)]), otu, pw)
        k.tabulateOtuDataForContig("s", 0, otu, pw)
        hits = np.concatenate(parts)
        for path in (0, 1):
            calls, o = fsm(path, flags, ncont, hits)
            lines, ci = [], 0
            for c in range(ncont):
                lines.append(f"CONTAINER {c}")
                while ci < len(calls) and calls[ci][0] == c:
                    _, s, e, cnt, f, w, _ = calls[ci]
                    lines.append(f"CALL\t{s}\t{e}\t{cnt}\t{f}\tF{f}\t{rt.java_format_f(float(np.frombuffer(w, np.float32)[0]), 6)}")
                    ci += 1
            lines.append("OTU-COUNTS\ts[0]" + "".join(f"\t{a}-{b}" for a, b in o))
            assert "\n".join(lines) + "\n" == buf.getvalue(), (case, path, flags)
        ncalls += buf.getvalue().count("CALL\t")
    assert ncalls > 150
