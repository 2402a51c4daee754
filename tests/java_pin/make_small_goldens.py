#!/usr/bin/env python
"""tests/java_pin/make_small_goldens.py <checkout of rsutormin/KmerGutsJava> -- golden REPORT TEXTS written by the reference's own
source (executed through tests/java_pin/j2py.py) for small deterministic configs[0]-style runs, committed under
tests/golden/java_reports/ so that they travel to the GPU box, where tests/test_gpu_c0_ecoli.py diffs the product's reports
against them directly (no oracle in between).  tests/test_oracle_golden.py checks on CPU that the oracle writes the same texts.

Inputs are rebuilt from the committed E. coli fixtures by small_inputs() below (deterministic; the SHA-256 of every input
file is recorded in the manifest and re-checked wherever the goldens are used):
  aa : table = C0 rule on the first 300 proteins of the .faa.gz, query = the first 150 proteins
  dna: table = C0 rule on all proteins,                          query = bases 100 000 .. 160 000 of the genome
Flag sets: [], -O, -m 3 -g 50 -M 2 (the -d reports carry dumps the product omits on purpose; they are compared through the
oracle instead)."""
import gzip
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
OUT = os.path.join(ROOT, "tests", "golden", "java_reports")
FAA = os.path.join(ROOT, "tests", "data", "Ecoli_K12_W3110.faa.gz")
FNA = os.path.join(ROOT, "tests", "data", "Ecoli_K12_W3110.fna.gz")
FLAGSETS = [("default", []), ("order", ["-O"]), ("m3g50M2", ["-m", "3", "-g", "50", "-M", "2"])]


def sha_file(path):
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


def small_inputs(d):
    """-> {mode: (data dir, query path)} under directory d, and {relative name: sha256} of every file made"""
    from tools import kg_synth as synth
    ids, descr, seqs = synth.read_fasta_simple(FAA)
    plain = os.path.join(d, "first300.faa")
    synth.write_fasta(plain, ids[:300], seqs[:300], descr=descr[:300])
    sub = os.path.join(d, "first300.faa.gz")
    with open(plain, "rb") as f, gzip.GzipFile(sub, "wb", mtime=0) as g:
        g.write(f.read())
    data_aa, data_dna = os.path.join(d, "KmerData300"), os.path.join(d, "KmerDataAll")
    synth.build_c0_fixture(sub, data_aa)
    synth.build_c0_fixture(FAA, data_dna)
    q_aa, q_dna = os.path.join(d, "q150.faa"), os.path.join(d, "q60k.fna")
    synth.write_fasta(q_aa, ids[:150], seqs[:150], descr=descr[:150])
    gids, _, gseqs = synth.read_fasta_simple(FNA)
    synth.write_fasta(q_dna, ["slice_100000_160000"], [gseqs[0][100_000:160_000]])
    shas = {}
    for name in ("q150.faa", "q60k.fna", "KmerData300/kmer.table.mem_map", "KmerData300/function.index",
                 "KmerDataAll/kmer.table.mem_map", "KmerDataAll/function.index"):
        shas[name] = sha_file(os.path.join(d, name))
    return {"aa": (data_aa, q_aa), "dna": (data_dna, q_dna)}, shas


def strip(text):
    volatile = ("Temp. directory:", "Preparation time:", "Lookup time:", "Grouping time:", "Processed:", "Table load time:")
    return "".join(line + "\n" for line in text.splitlines() if not line.startswith(volatile))


def runs(inputs):
    for mode in ("aa", "dna"):
        data, q = inputs[mode]
        for name, flags in FLAGSETS:
            yield f"{mode}_{name}", (["-a"] if mode == "aa" else []) + flags + ["-D", data, "-q", q]


def main():
    import tempfile
    import transliterated_pin as tp
    ref = sys.argv[1]
    os.makedirs(OUT, exist_ok=True)
    with tempfile.TemporaryDirectory() as d:
        mod, java_sha = tp.load_reference(os.path.join(ref, tp.JAVA_REL), os.path.join(d, "kgj_transliterated.py"))
        inputs, shas = small_inputs(d)
        manifest = {"java_source_sha256": java_sha, "inputs": shas, "reports": {},
                    "generated_by": "python tests/java_pin/make_small_goldens.py /root/reference"}
        for item, args in runs(inputs):
            out = os.path.join(d, item + ".txt")
            mod.KmerGutsJava.main(args + ["-o", out])
            text = strip(open(out, errors="replace").read())
            open(os.path.join(OUT, item + ".txt"), "w", newline="").write(text)
            manifest["reports"][item] = hashlib.sha256(text.encode()).hexdigest()
            print(item, len(text.splitlines()), "lines", text.count("\nCALL\t"), "calls")
        json.dump(manifest, open(os.path.join(OUT, "manifest.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
