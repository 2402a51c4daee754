"""The pin between the CPU oracle and the real Java (tests/java_pin/): the oracle's reports for the configs[0] runs -- the
reference's own E. coli fixtures, protein mode and 6-frame mode, the four flag sets of SURVEY.md 8(d) -- must hash to the
values committed in tests/golden/c0_report_sha256.json.  Those are the values tests/java_pin/pin_oracle.sh compares the
UNMODIFIED KmerGutsJava's reports with on a box that has a JDK (none exists in this image), so "oracle == committed hashes"
here plus "Java == committed hashes" there pins every GPU-vs-oracle test of this suite to the Java."""
import importlib.util
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _pin():
    spec = importlib.util.spec_from_file_location("pin_oracle", os.path.join(ROOT, "tests", "java_pin", "pin_oracle.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def test_oracle_reports_match_committed_hashes(tmp_path):
    pin = _pin()
    got = pin.oracle_reports(str(tmp_path))
    want = json.load(open(pin.GOLDEN))
    assert set(got) == set(want) and len(want) == 10
    assert got == want, {k: (got[k][:12], want[k][:12]) for k in want if got[k] != want[k]}
    # the run list GoldenDump executes names exactly these reports, with KmerGutsJava.main's own flags
    runs = [l.rstrip("\n").split("\t") for l in open(tmp_path / "runs.txt")]
    assert len(runs) == 8 and all("-D" in r and "-q" in r and "-o" in r for r in runs)
    assert sum(1 for r in runs if "-a" in r) == 4 and sum(1 for r in runs if "-d" in r) == 2


def test_kat_text_for_the_java_driver(tmp_path):
    """kats.txt carries every hand-traced vector with its flags, weights as float bits; the text expected from Java's
    gatherHits is the CALL / OTU-COUNTS lines of the KAT file (the weighted column as Java's %f prints a float)."""
    pin = _pin()
    pin.write_kats(str(tmp_path / "kats.txt"))
    kats = json.load(open(pin.KATS))
    lines = open(tmp_path / "kats.txt").read().splitlines()
    assert sum(1 for l in lines if l.startswith("KAT ")) == len(kats) == 21
    assert len(lines) == len(kats) + sum(len(k["hits"]) for k in kats)
    exp = pin.expected_kat_text()
    assert exp.count("KAT ") == 21 and "CALL\t0\t47\t5\t7\tF7\t2.500000" in exp and "OTU-COUNTS\tkat[0]\t3-3\t2-6\t2-2\t1-5\t1-4" in exp
    src = open(os.path.join(ROOT, "tests", "java_pin", "GoldenDump.java")).read()
    for name in ("gatherHits", "tabulateOtuDataForContig", "minHits", "maxGap", "minWeightedHits", "orderConstraint", "KmerGutsJava.main"):
        assert name in src


def test_transliterated_java_reports_are_the_committed_hashes():
    """tests/golden/java_transliteration_pin.json records what the reference's unmodified Java source printed for the eight
    configs[0] runs when tests/java_pin/transliterated_pin.py executed it (mechanical Java -> Python transliteration; the image
    has no JVM): every report must be the one the oracle hashes to, i.e. the value every GPU-vs-oracle test is pinned to."""
    import hashlib
    pin = _pin()
    rec = json.load(open(os.path.join(ROOT, "tests", "golden", "java_transliteration_pin.json")))
    want = json.load(open(pin.GOLDEN))
    assert rec["compare_exit_code"] == 0
    assert set(rec["reports"]) == {k for k in want if k not in ("kmer.table.mem_map", "function.index")} and len(rec["reports"]) == 8
    assert all(rec["reports"][k] == want[k] for k in rec["reports"])
    assert rec["fsm_kats_sha256"] == hashlib.sha256(pin.expected_kat_text().encode()).hexdigest()
    java = os.path.join(os.environ.get("KG_REFERENCE", "/root/reference"), "lib", "src", "kmergutsjava", "KmerGutsJava.java")
    if os.path.exists(java):   # the record was made from exactly the source that is here
        assert hashlib.sha256(open(java, encoding="utf-8").read().encode()).hexdigest() == rec["java_source_sha256"]


def test_oracle_writes_the_java_written_golden_reports(tmp_path):
    """tests/golden/java_reports/*.txt were written by the reference's own source (tests/java_pin/make_small_goldens.py); the
    inputs are rebuilt here from the committed E. coli fixtures (their hashes are in the manifest) and the oracle must write
    the same texts.  The GPU suite diffs the product's reports against the same files."""
    from oracle import kgo
    spec = importlib.util.spec_from_file_location("make_small_goldens", os.path.join(ROOT, "tests", "java_pin", "make_small_goldens.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    kgo.build()
    manifest = json.load(open(os.path.join(g.OUT, "manifest.json")))
    inputs, shas = g.small_inputs(str(tmp_path))
    assert shas == manifest["inputs"]
    items = list(g.runs(inputs))
    assert [i for i, _ in items] == sorted(manifest["reports"], key=[i for i, _ in items].index) and len(items) == 6
    for item, args in items:
        out = str(tmp_path / (item + ".txt"))
        kgo.run_cli(args + ["-o", out])
        want = open(os.path.join(g.OUT, item + ".txt"), newline="").read()
        assert g.strip(open(out, errors="replace").read()) == want, item
        assert want.count("\nCALL\t") >= 50
