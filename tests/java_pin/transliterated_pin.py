#!/usr/bin/env python
"""tests/java_pin/transliterated_pin.py -- pins the CPU oracle to the reference's OWN SOURCE on a box without a JVM.

    python tests/java_pin/transliterated_pin.py <checkout of rsutormin/KmerGutsJava> <workdir> [--jobs N] [--record]

What pin_oracle.sh does with javac + java, this does with tests/java_pin/j2py.py: the unmodified
lib/src/kmergutsjava/KmerGutsJava.java is transliterated, statement by statement, into Python (the tool knows Java syntax and
Java's arithmetic, nothing about k-mers), and THAT text is executed:

  * KmerGutsJava.main(args) for the eight configs[0] runs (the reference's two E. coli fixtures against the derived table;
    protein mode and 6-frame mode; flag sets [], -d, -O, -m 3 -g 50 -M 2) -> <workdir>/java/<item>.txt
  * gatherHits + tabulateOtuDataForContig on the 21 hand-traced FSM vectors     -> <workdir>/java_kats.txt

then `pin_oracle.py compare` diffs them against the oracle's reports and the SHA-256 values under tests/golden/.  With
--record the outcome is written to tests/golden/java_transliteration_pin.json (asserted by tests/test_oracle_golden.py).
The full-size runs take tens of minutes of pure Python; tests/test_java_transliteration.py repeats the comparison on small
inputs inside the CPU test-suite whenever the reference checkout is present.
"""
import argparse
import hashlib
import importlib.util
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
JAVA_REL = os.path.join("lib", "src", "kmergutsjava", "KmerGutsJava.java")
RECORD = os.path.join(ROOT, "tests", "golden", "java_transliteration_pin.json")


def load_reference(java_path, module_path):
    """Transliterate the Java file, write the Python text next to the other work files, import it."""
    sys.path.insert(0, HERE)
    import j2py
    src = open(java_path, encoding="utf-8").read()
    text = j2py.transliterate(src)
    with open(module_path, "w") as f:
        f.write(text)
    spec = importlib.util.spec_from_file_location("kgj_transliterated", module_path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod, hashlib.sha256(src.encode()).hexdigest()


def run_main(mod, args):
    """KmerGutsJava.main(String[] args), KGJ:560"""
    mod.KmerGutsJava.main(list(args))


def run_kats(mod, kats_txt, out_txt):
    """What GoldenDump.kats does on a JVM: one KmerGutsJava instance per vector, gatherHits (KGJ:457) on the hits, then
    tabulateOtuDataForContig (KGJ:516)."""
    rt = sys.modules["j2py_runtime"]
    with open(kats_txt) as f, open(out_txt, "w", newline="") as out:
        pw = rt.PrintWriter(out)
        lines = f.read().splitlines()
        i = 0
        while i < len(lines):
            h = lines[i].split(" ")
            i += 1
            if h[0] != "KAT":
                continue
            k = mod.KmerGutsJava()
            k.minHits, k.maxGap, k.minWeightedHits, k.orderConstraint = int(h[2]), int(h[3]), int(h[4]), int(h[5]) != 0
            n, max_fi = int(h[6]), 0
            hits = rt.ArrayList()
            for _ in range(n):
                t = lines[i].strip().split(" ")
                i += 1
                x = mod.Hit()
                x.from0InProt, x.fI, x.oI = int(t[0]), int(t[1]), int(t[2])
                x.functionWt = rt.Float.intBitsToFloat(int(t[3]))
                x.avgOffFromEnd = int(t[4])
                max_fi = max(max_fi, x.fI)
                hits.add(x)
            functions = rt.ArrayList(["F%d" % j for j in range(max_fi + 1)])
            otu = rt.ArrayList()
            pw.println("KAT " + h[1])
            k.gatherHits(100000, "+", 0, hits, functions, otu, pw)
            k.tabulateOtuDataForContig("kat", 0, otu, pw)
        pw.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("reference")
    ap.add_argument("workdir")
    ap.add_argument("--jobs", type=int, default=min(8, os.cpu_count() or 1))
    ap.add_argument("--record", action="store_true")
    ap.add_argument("--one", help="(internal) run KmerGutsJava.main on this tab-separated argument line and exit")
    a = ap.parse_args()
    java = os.path.join(a.reference, JAVA_REL)
    w = os.path.abspath(a.workdir)
    os.makedirs(w, exist_ok=True)
    if a.one is not None:
        mod, _ = load_reference(java, os.path.join(w, f"kgj_transliterated_{os.getpid()}.py"))
        run_main(mod, a.one.split("\t"))
        os.remove(os.path.join(w, f"kgj_transliterated_{os.getpid()}.py"))
        return 0
    pin = os.path.join(HERE, "pin_oracle.py")
    if subprocess.run([sys.executable, pin, "prepare", w]).returncode != 0:
        return 1
    mod, java_sha = load_reference(java, os.path.join(w, "kgj_transliterated.py"))
    run_kats(mod, os.path.join(w, "kats.txt"), os.path.join(w, "java_kats.txt"))
    print("KATs done", flush=True)
    runs = [l.rstrip("\n") for l in open(os.path.join(w, "runs.txt")) if l.strip()]
    t0 = time.time()
    procs = []
    pending = list(runs)
    failed = 0
    while pending or procs:
        while pending and len(procs) < a.jobs:
            line = pending.pop(0)
            procs.append((line, subprocess.Popen([sys.executable, os.path.abspath(__file__), a.reference, w, "--one=" + line],
                                                 stdout=subprocess.DEVNULL)))
        time.sleep(1.0)
        for line, p in list(procs):
            if p.poll() is not None:
                procs.remove((line, p))
                item = os.path.basename(line.split("\t")[-1])
                print(f"  {item}: exit {p.returncode} after {time.time() - t0:.0f} s", flush=True)
                failed += p.returncode != 0
    if failed:
        print(f"{failed} run(s) of the transliterated main failed")
        return 1
    rc = subprocess.run([sys.executable, pin, "compare", w]).returncode
    if a.record:
        sys.path.insert(0, HERE)
        import pin_oracle
        items = {}
        for item, _, _ in pin_oracle.runs(w):
            items[item] = pin_oracle.sha(pin_oracle.strip(open(os.path.join(w, "java", item + ".txt"), errors="replace").read()).encode())
        rec = {"what": "SHA-256 of the reports KmerGutsJava.main wrote when the reference's unmodified Java source was executed through "
                       "tests/java_pin/j2py.py (mechanical transliteration; no JVM in the image), wall-clock lines dropped",
               "java_source_sha256": java_sha, "compare_exit_code": rc, "reports": items,
               "fsm_kats_sha256": hashlib.sha256(open(os.path.join(w, "java_kats.txt"), "rb").read()).hexdigest(),
               "generated_by": "python tests/java_pin/transliterated_pin.py /root/reference <workdir> --record"}
        json.dump(rec, open(RECORD, "w"), indent=1, sort_keys=True)
        print(f"wrote {RECORD}")
    return rc


if __name__ == "__main__":
    sys.exit(main())
