#!/usr/bin/env python
"""tests/java_pin/pin_oracle.py -- the Python half of tests/java_pin/pin_oracle.sh (pinning the oracle to the real Java).

    prepare <workdir>   builds the configs[0] fixture (KmerData/ from the committed E. coli .faa.gz), writes kats.txt for
                        GoldenDump, runs the ORACLE on the eight configs[0] runs and checks its reports against the SHA-256
                        values committed in tests/golden/c0_report_sha256.json (so the oracle that is compared with Java is
                        the oracle the test-suite pins everything else to);
    compare <workdir>   compares what GoldenDump wrote (workdir/java/*.txt, workdir/java_kats.txt) with the oracle's reports
                        and with tests/golden/fsm_kats.json: byte-identical after dropping the wall-clock lines; prints one
                        PASS / FAIL line per item and exits non-zero on any FAIL;
    hashes              (maintainers) regenerates tests/golden/c0_report_sha256.json from the oracle.

Needs no GPU and no JVM itself; only `compare` needs the files a JVM produced.
"""
import hashlib
import json
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
FAA = os.path.join(ROOT, "tests", "data", "Ecoli_K12_W3110.faa.gz")
FNA = os.path.join(ROOT, "tests", "data", "Ecoli_K12_W3110.fna.gz")
GOLDEN = os.path.join(ROOT, "tests", "golden", "c0_report_sha256.json")
KATS = os.path.join(ROOT, "tests", "golden", "fsm_kats.json")
FLAGSETS = [("default", []), ("debug", ["-d"]), ("order", ["-O"]), ("m3g50M2", ["-m", "3", "-g", "50", "-M", "2"])]
# wall-clock / environment lines of KmerGutsJava.run (KGJ:775, 794-796, 803-804, 819, 1019-1025, printInfoLine KGJ:891-898)
VOLATILE = ("Temp. directory:", "Preparation time:", "Lookup time:", "Grouping time:", "Processed:")


def strip(text: str) -> str:
    return "".join(line + "\n" for line in text.splitlines() if not line.startswith(VOLATILE))


def sha(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


DEBUG_PROTEINS, DEBUG_BASES = 400, 300_000   # the -d runs use a prefix: the after-hit / after-call dumps grow quadratically


def small_queries(workdir: str):
    """First DEBUG_PROTEINS proteins of the .faa.gz and the first DEBUG_BASES bases of the .fna.gz, as plain FASTA."""
    from tools import kg_synth as synth
    faa, fna = os.path.join(workdir, "debug_query.faa"), os.path.join(workdir, "debug_query.fna")
    ids, _, seqs = synth.read_fasta_simple(FAA)
    synth.write_fasta(faa, ids[:DEBUG_PROTEINS], seqs[:DEBUG_PROTEINS])
    ids, _, seqs = synth.read_fasta_simple(FNA)
    synth.write_fasta(fna, ids[:1], [seqs[0][:DEBUG_BASES]])
    return faa, fna


def runs(workdir: str):
    faa_s, fna_s = os.path.join(workdir, "debug_query.faa"), os.path.join(workdir, "debug_query.fna")
    for mode, q, qs in (("aa", FAA, faa_s), ("dna", FNA, fna_s)):
        for name, flags in FLAGSETS:
            yield f"{mode}_{name}", (["-a"] if mode == "aa" else []) + flags, (qs if "-d" in flags else q)


def oracle_reports(workdir: str) -> dict:
    """Fixture + the oracle's eight reports; returns {item: sha256 of the stripped text} (+ the two fixture files)."""
    from oracle import kgo
    from tools import kg_synth as synth
    kgo.build()
    data = os.path.join(workdir, "KmerData")
    synth.build_c0_fixture(FAA, data)
    out = {"kmer.table.mem_map": sha(open(os.path.join(data, "kmer.table.mem_map"), "rb").read()),
           "function.index": sha(open(os.path.join(data, "function.index"), "rb").read())}
    os.makedirs(os.path.join(workdir, "oracle"), exist_ok=True)
    os.makedirs(os.path.join(workdir, "java"), exist_ok=True)
    small_queries(workdir)
    with open(os.path.join(workdir, "runs.txt"), "w") as rf:   # the same runs, for GoldenDump (KmerGutsJava.main arguments)
        for item, flags, q in runs(workdir):
            path = os.path.join(workdir, "oracle", item + ".txt")
            kgo.run_cli(flags + ["-D", data, "-q", q, "-o", path])
            out[item] = sha(strip(open(path, errors="replace").read()).encode())
            rf.write("\t".join(flags + ["-D", data, "-q", os.path.abspath(q), "-o", os.path.join(workdir, "java", item + ".txt")]) + "\n")
    return out


def write_kats(path: str):
    with open(path, "w") as f:
        for k in json.load(open(KATS)):
            p = k["params"]
            f.write(f"KAT {k['name']} {p['min_hits']} {p['max_gap']} {p['min_weighted_hits']} {p['order_constraint']} {len(k['hits'])}\n")
            for pos, fi, oi, wt, avg in k["hits"]:
                bits = struct.unpack("<i", struct.pack("<f", wt))[0]
                f.write(f"{pos} {fi} {oi} {bits} {avg}\n")


def java_f(x: float) -> str:
    """String.format("%f") of a float as Java prints it (the oracle's restatement of it, through its C library)."""
    from oracle import kgo
    return kgo.java_format_f(x, 6)


def expected_kat_text() -> str:
    lines = []
    for k in json.load(open(KATS)):
        lines.append(f"KAT {k['name']}")
        for start, end, count, fi, w in k["calls"]:
            lines.append(f"CALL\t{start}\t{end}\t{count}\t{fi}\tF{fi}\t{java_f(struct.unpack('<f', struct.pack('<f', w))[0])}")
        lines.append("OTU-COUNTS\tkat[0]" + "".join(f"\t{c}-{o}" for c, o in k["otu"]))
    return "".join(x + "\n" for x in lines)


def main():
    cmd = sys.argv[1] if len(sys.argv) > 1 else ""
    if cmd == "hashes":
        import tempfile
        with tempfile.TemporaryDirectory() as d:
            h = oracle_reports(d)
        json.dump(h, open(GOLDEN, "w"), indent=1, sort_keys=True)
        print(f"wrote {GOLDEN}")
        return 0
    if cmd == "prepare":
        w = sys.argv[2]
        os.makedirs(w, exist_ok=True)
        got, want = oracle_reports(w), json.load(open(GOLDEN))
        bad = [k for k in want if got.get(k) != want[k]]
        write_kats(os.path.join(w, "kats.txt"))
        print(f"fixture and oracle reports in {w}; " + ("they match tests/golden/c0_report_sha256.json" if not bad else f"MISMATCH with the committed hashes: {bad}"))
        return 1 if bad else 0
    if cmd == "compare":
        w = sys.argv[2]
        want = json.load(open(GOLDEN))
        fails = 0
        for item, _, _ in runs(w):
            jp, op = os.path.join(w, "java", item + ".txt"), os.path.join(w, "oracle", item + ".txt")
            if not os.path.exists(jp):
                print(f"FAIL {item}: {jp} is missing")
                fails += 1
                continue
            j, o = strip(open(jp, errors="replace").read()), strip(open(op, errors="replace").read())
            ok = j == o and sha(j.encode()) == want[item]
            print(f"{'PASS' if ok else 'FAIL'} {item}: java {len(j.splitlines())} lines, sha256 {sha(j.encode())[:16]}; committed {want[item][:16]}")
            if not ok:
                fails += 1
                jl, ol = j.splitlines(), o.splitlines()
                for i in range(min(len(jl), len(ol))):
                    if jl[i] != ol[i]:
                        print(f"   first difference at line {i + 1}:\n     java  : {jl[i]!r}\n     oracle: {ol[i]!r}")
                        break
        jk = os.path.join(w, "java_kats.txt")
        if os.path.exists(jk):
            ok = open(jk).read() == expected_kat_text()
            print(f"{'PASS' if ok else 'FAIL'} fsm_kats: gatherHits on the 21 hand-traced vectors")
            fails += 0 if ok else 1
        else:
            print(f"FAIL fsm_kats: {jk} is missing")
            fails += 1
        print("oracle pinned to the Java: every report identical" if not fails else f"{fails} item(s) differ")
        return 1 if fails else 0
    print(__doc__)
    return 2


if __name__ == "__main__":
    sys.exit(main())
