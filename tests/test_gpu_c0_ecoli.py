"""configs[0]: the reference's own CPU-runnable case.  The repo ships no kmer table, so the table is
derived from the E. coli protein FASTA the reference ships as test data (tools/kg_synth.build_c0_fixture); then the
reference's two fixtures are run in protein mode (.faa.gz) and 6-frame mode (.fna.gz) under the flag sets of
SURVEY.md section 8(d), through the command lines, and the text reports are compared byte for byte."""
import os
import subprocess

import numpy as np
import pytest

from tests.parity import assert_same
from tools import kg_synth as synth

pytestmark = pytest.mark.gpu
DATA = os.path.join(os.path.dirname(__file__), "data")
FAA = os.path.join(DATA, "Ecoli_K12_W3110.faa.gz")
FNA = os.path.join(DATA, "Ecoli_K12_W3110.fna.gz")


@pytest.fixture(scope="module")
def c0(tmp_path_factory):
    d = str(tmp_path_factory.mktemp("c0") / "KmerData")
    info = synth.build_c0_fixture(FAA, d)
    assert info["num_signatures"] > 300_000
    return d


@pytest.fixture(scope="module")
def kg():
    import kmergutsjava_b200 as kg
    return kg


FLAGSETS = [[], ["-d"], ["-O"], ["-m", "3", "-g", "50", "-M", "2"]]


def _strip(text, debug):
    """Drop what is wall-clock or deliberately not reproduced: timer lines everywhere; in -d output also the oracle's
    list dumps (after-hit / after-call) and its 'Kmers found' counter (divergences)."""
    out = []
    for line in text.splitlines():
        if line.startswith(("Temp. directory:", "Preparation time:", "Lookup time:", "Grouping time:", "Processed:", "Table load time:")):
            continue
        if debug and line.startswith(("after-hit:", "after-call:", "Kmers found:")):
            continue
        out.append(line)
    return out


@pytest.mark.parametrize("flags", FLAGSETS, ids=["default", "debug", "order", "m3g50M2"])
@pytest.mark.parametrize("mode", ["aa", "dna"])
def test_c0_reports_identical(kg, oracle, c0, tmp_path, mode, flags):
    query = FAA if mode == "aa" else FNA
    small = "-d" in flags   # the reference's -d output dumps the whole open run after every hit (quadratic): use a slice
    if small:
        ids, descr, seqs = synth.read_fasta_simple(query)
        query = str(tmp_path / ("q.faa" if mode == "aa" else "q.fna"))
        if mode == "aa":
            synth.write_fasta(query, ids[:400], seqs[:400])
        else:
            synth.write_fasta(query, ["slice1", "slice2"], [seqs[0][100000:160000], seqs[0][2000000:2030001]])
    base = (["-a"] if mode == "aa" else []) + flags + ["-D", c0, "-q", query]
    o_out, g_out = str(tmp_path / "oracle.txt"), str(tmp_path / "gpu.txt")
    oracle.run_cli(base + ["-o", o_out])
    r = subprocess.run([kg.CLI_PATH, *base, "-o", g_out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a = _strip(open(o_out).read(), "-d" in flags)
    b = _strip(open(g_out).read(), "-d" in flags)
    assert len(a) == len(b)
    ncall = sum(1 for x in a if x.startswith("CALL\t"))
    assert ncall > ((5000 if mode == "aa" else 3000) if not small else 30), ncall
    if small:
        assert sum(1 for x in a if x.startswith("HIT\t")) > 1000
    if a != b:
        bad = [(i, x, y) for i, (x, y) in enumerate(zip(a, b)) if x != y][:5]
        raise AssertionError(f"{len(bad)}+ differing lines, first: {bad}")
    if not small:
        # ... and, without the oracle in between: the GPU report hashes to the value committed under tests/golden/, which is
        # what the reference's own Java source wrote for this run when it was executed through tests/java_pin/ (a JVM via
        # pin_oracle.sh, or the mechanical transliteration recorded in tests/golden/java_transliteration_pin.json)
        import hashlib
        import json
        golden = os.path.join(os.path.dirname(__file__), "golden")
        name = {(): "default", ("-O",): "order"}.get(tuple(flags), "m3g50M2")
        want = json.load(open(os.path.join(golden, "c0_report_sha256.json")))[f"{mode}_{name}"]
        assert hashlib.sha256("".join(x + "\n" for x in b).encode()).hexdigest() == want
        rec = os.path.join(golden, "java_transliteration_pin.json")
        if os.path.exists(rec):
            assert json.load(open(rec))["reports"][f"{mode}_{name}"] == want


def test_reports_equal_the_java_written_goldens(kg, tmp_path):
    """No oracle in this one: tests/golden/java_reports/*.txt were written by the reference's own Java source (executed through
    tests/java_pin/j2py.py, see make_small_goldens.py); the inputs are rebuilt from the committed E. coli fixtures (hashes in
    the manifest) and the product's command line must write the same text, byte for byte."""
    import importlib.util
    import json
    spec = importlib.util.spec_from_file_location("make_small_goldens", os.path.join(os.path.dirname(__file__), "java_pin", "make_small_goldens.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    manifest = json.load(open(os.path.join(g.OUT, "manifest.json")))
    inputs, shas = g.small_inputs(str(tmp_path))
    assert shas == manifest["inputs"]
    n = 0
    for item, args in g.runs(inputs):
        out = str(tmp_path / (item + ".gpu.txt"))
        r = subprocess.run([kg.CLI_PATH, *args, "-o", out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        want = open(os.path.join(g.OUT, item + ".txt"), newline="").read()
        got = g.strip(open(out, errors="replace").read())
        if got != want:
            bad = [(i, x, y) for i, (x, y) in enumerate(zip(want.splitlines(), got.splitlines())) if x != y][:3]
            raise AssertionError(f"{item}: {len(want.splitlines())} vs {len(got.splitlines())} lines; first differences (java, gpu): {bad}")
        n += 1
    assert n == 6


def test_c0_library_level(kg, oracle, c0):
    ctx = kg.Context(0)
    table = ctx.load_table(c0)
    otab = oracle.Table(path=os.path.join(c0, "kmer.table.mem_map"))
    ti = table.info
    assert ti.num_unreachable == 0 and ti.tail_run == 0 and ti.num_slots == otab.num_sigs
    fa = kg.Fasta(FAA)
    assert fa.n == 13645
    res = ctx.run(table, kg.MODE_AA, fa.bytes, fa.offsets, kg.default_params(emit_hits=1))
    ref = oracle.run(otab, oracle.make_params(aa=True), fa.bytes, fa.offsets, oracle.STREAM_JOIN)
    assert ref.num_kmers == 4037833          # windows the reference enumerates from this file
    assert_same(res, ref, what="C0 aa")
    res.free()
    fn = kg.Fasta(FNA)
    res = ctx.run(table, kg.MODE_DNA, fn.bytes, fn.offsets, kg.default_params(emit_hits=1))
    ref = oracle.run(otab, oracle.make_params(aa=False), fn.bytes, fn.offsets, oracle.STREAM_JOIN)
    assert_same(res, ref, what="C0 dna")
    assert len(ref.calls) > 3000 and set(ref.calls["sf"]) == set(range(6))
    res.free()
    table.free()
    ctx.close()


def test_c0_gz_table_and_duplicate_ids(kg, oracle, c0, tmp_path):
    """kmer.table.mem_map.gz wins over the plain file (KGJ:749-753); repeated FASTA ids collapse as in the reference's
    LinkedHashMap bookkeeping (KGJ:772, 805-809): first position, last length, last record's hits."""
    import gzip
    import shutil
    d = tmp_path / "KmerDataGz"
    d.mkdir()
    with open(os.path.join(c0, "kmer.table.mem_map"), "rb") as f, gzip.open(d / "kmer.table.mem_map.gz", "wb", compresslevel=1) as g:
        shutil.copyfileobj(f, g)
    shutil.copy(os.path.join(c0, "function.index"), d / "function.index")
    ids, descr, seqs = synth.read_fasta_simple(FAA)
    pick = [0, 1, 2, 3, 1, 4, 0]
    q = str(tmp_path / "dup.faa")
    synth.write_fasta(q, [ids[i] if k != 4 else ids[1] for k, i in enumerate(pick)], [seqs[i] if k != 4 else seqs[5] for k, i in enumerate(pick)])
    o_out, g_out = str(tmp_path / "o.txt"), str(tmp_path / "g.txt")
    oracle.run_cli(["-a", "-D", str(d), "-q", q, "-o", o_out])
    r = subprocess.run([kg.CLI_PATH, "-a", "-D", str(d), "-q", q, "-o", g_out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    a, b = _strip(open(o_out).read(), False), _strip(open(g_out).read(), False)
    assert a == b
    assert sum(1 for x in a if x.startswith("PROTEIN-ID")) == 5


def test_c0_cli_table_cache(kg, oracle, c0, tmp_path):
    """-C: the first run builds the table from -D and writes the cache, the second run reads the cache (the reference
    table may be gone by then); both reports equal the oracle's."""
    import shutil
    d = tmp_path / "KmerData"
    shutil.copytree(c0, d)
    cache = str(tmp_path / "table.kgcache")
    ids, descr, seqs = synth.read_fasta_simple(FAA)
    q = str(tmp_path / "q.faa")
    synth.write_fasta(q, ids[:300], seqs[:300])
    o_out = str(tmp_path / "o.txt")
    oracle.run_cli(["-a", "-D", str(d), "-q", q, "-o", o_out])
    want = _strip(open(o_out).read(), False)
    for attempt in range(2):
        g_out = str(tmp_path / f"g{attempt}.txt")
        r = subprocess.run([kg.CLI_PATH, "-a", "-D", str(d), "-C", cache, "-q", q, "-o", g_out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert _strip(open(g_out).read(), False) == want
        assert os.path.getsize(cache) > 1 << 20
        if attempt == 0:
            os.remove(d / "kmer.table.mem_map")  # only the cache can serve the second run


@pytest.mark.parametrize("mode", ["aa", "dna"])
def test_c0_cli_streaming_and_dna_ranges(kg, oracle, c0, tmp_path, mode):
    """-B (extension): the query is read, run and reported in batches of whole records; with unique ids the report is the
    one-shot report (and the oracle's).  -c (extension): a DNA-RANGE line after every CALL of a 6-frame run."""
    base = ["-a", "-D", c0, "-q", FAA] if mode == "aa" else ["-D", c0, "-q", FNA]   # aa: ~4 batches of 1 MB; dna: one contig
    o_out = str(tmp_path / "o.txt")
    oracle.run_cli(base + ["-o", o_out])
    want = _strip(open(o_out).read(), False)
    outs = {}
    for name, extra in (("oneshot", []), ("stream", ["-B", "1"]), ("stream_c", ["-B", "1", "-c"])):
        g_out = str(tmp_path / f"{name}.txt")
        r = subprocess.run([kg.CLI_PATH] + base + extra + ["-o", g_out], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs[name] = _strip(open(g_out).read(), False)
        if name != "oneshot":
            assert "batches" in r.stdout
    assert outs["oneshot"] == want and outs["stream"] == want
    plain = [l for l in outs["stream_c"] if not l.startswith("DNA-RANGE")]
    assert plain == want
    ranges = [l for l in outs["stream_c"] if l.startswith("DNA-RANGE")]
    if mode == "aa":
        assert not ranges
    else:
        ncalls = sum(l.startswith("CALL") for l in want)
        assert len(ranges) == ncalls > 100
        for i, l in enumerate(outs["stream_c"]):
            if l.startswith("DNA-RANGE"):
                call = outs["stream_c"][i - 1].split("\t")
                _, b, e, sd = l.split("\t")
                assert call[0] == "CALL" and (int(e) - int(b) + 1) == 3 * (int(call[2]) - int(call[1]) + 1) and sd in "+-"


def test_report_rejects_function_index_outside_the_index(kg, oracle, c0, tmp_path):
    """functionArray.get(currentFI) throws in the reference (KGJ:403); a function.index that is too short for the
    table must not produce CALL lines with an empty name."""
    import shutil
    d = tmp_path / "KmerData"
    shutil.copytree(c0, d)
    lines = open(d / "function.index").read().splitlines()
    with open(d / "function.index", "w") as f:
        f.write("\n".join(lines[:5]) + "\n")
    ids, descr, seqs = synth.read_fasta_simple(FAA)
    q = str(tmp_path / "q.faa")
    synth.write_fasta(q, ids[:300], seqs[:300])
    r = subprocess.run([kg.CLI_PATH, "-a", "-D", str(d), "-q", q, "-o", str(tmp_path / "g.txt")], capture_output=True, text=True)
    assert r.returncode != 0 and "function index" in r.stderr and "KGJ:403" in r.stderr


def test_c0_cli_stale_table_cache(kg, oracle, c0, tmp_path):
    """-C with a cache that was built from ANOTHER kmer.table.mem_map than the one in -D: the cache records the
    source's size and mtime, the command line notices the mismatch, rebuilds from -D and reports what -D says."""
    import shutil
    d = tmp_path / "KmerData"
    shutil.copytree(c0, d)
    cache = str(tmp_path / "table.kgcache")
    ids, descr, seqs = synth.read_fasta_simple(FAA)
    q = str(tmp_path / "q.faa")
    synth.write_fasta(q, ids[:200], seqs[:200])
    g0 = str(tmp_path / "g0.txt")
    r = subprocess.run([kg.CLI_PATH, "-a", "-D", str(d), "-C", cache, "-q", q, "-o", g0], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # a different table in -D: every payload's function index swapped to 0 (same size, new mtime)
    img = bytearray(open(d / "kmer.table.mem_map", "rb").read())
    a = np.frombuffer(img, dtype=np.uint8)[24:].reshape(-1, 24)
    a[:, 16:20] = 0
    with open(d / "kmer.table.mem_map", "wb") as f:
        f.write(bytes(img))
    os.utime(d / "kmer.table.mem_map", ns=(1, 1))
    o_out, g1 = str(tmp_path / "o.txt"), str(tmp_path / "g1.txt")
    oracle.run_cli(["-a", "-D", str(d), "-q", q, "-o", o_out])
    r = subprocess.run([kg.CLI_PATH, "-a", "-D", str(d), "-C", cache, "-q", q, "-o", g1], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert "table cache not used" in r.stderr
    assert _strip(open(g1).read(), False) == _strip(open(o_out).read(), False)
    assert _strip(open(g1).read(), False) != _strip(open(g0).read(), False)
