"""The reference's OWN SOURCE, executed here without a JVM, against the CPU oracle.

tests/java_pin/j2py.py transliterates the unmodified lib/src/kmergutsjava/KmerGutsJava.java into Python statement by statement
(it knows Java syntax and Java's arithmetic, nothing about k-mers); these tests run that text -- main(), run(), readFasta,
lookup, gatherHits, processSetOfHits, the scalar helpers -- on small inputs and demand what the oracle gives, byte for byte.
The full-size runs of the same comparison (the eight configs[0] reports) are recorded in
tests/golden/java_transliteration_pin.json by tests/java_pin/transliterated_pin.py and asserted in test_oracle_golden.py.

Needs the reference checkout (KG_REFERENCE, default /root/reference): present where the CPU suite runs, absent on the GPU
boxes -- there the module is skipped.  No GPU, and nothing of the product is involved."""
import ctypes as C
import gzip
import importlib.util
import os
import sys

import numpy as np
import pytest

from tools import kg_synth as synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("KG_REFERENCE", "/root/reference")
JAVA = os.path.join(REF, "lib", "src", "kmergutsjava", "KmerGutsJava.java")
pytestmark = pytest.mark.skipif(not os.path.exists(JAVA), reason="the reference checkout is not on this box")

FAA = os.path.join(ROOT, "tests", "data", "Ecoli_K12_W3110.faa.gz")
FNA = os.path.join(ROOT, "tests", "data", "Ecoli_K12_W3110.fna.gz")
FLAGSETS = [[], ["-d"], ["-O"], ["-m", "3", "-g", "50", "-M", "2"]]


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.fixture(scope="module")
def pin():
    return _load("pin_oracle", os.path.join(ROOT, "tests", "java_pin", "pin_oracle.py"))


@pytest.fixture(scope="module")
def java(tmp_path_factory):
    """(module holding the transliterated classes, the runtime module)"""
    tp = _load("transliterated_pin", os.path.join(ROOT, "tests", "java_pin", "transliterated_pin.py"))
    mod, _ = tp.load_reference(JAVA, str(tmp_path_factory.mktemp("kgj") / "kgj_transliterated.py"))
    return mod, sys.modules["j2py_runtime"], tp


@pytest.fixture(scope="module")
def small_c0(tmp_path_factory):
    """configs[0] in small (SURVEY 8(d) C0 rule).  Protein mode: the table derived from the first 300 E. coli proteins, queried
    with the first 150.  6-frame mode: the table derived from ALL proteins (the .faa.gz is not in genome order, so a small
    table would leave a genome prefix almost without hits), queried with the first 60 kb of the genome."""
    d = tmp_path_factory.mktemp("c0small")
    ids, descr, seqs = synth.read_fasta_simple(FAA)
    sub = str(d / "first300.faa.gz")
    plain = str(d / "first300.faa")
    synth.write_fasta(plain, ids[:300], seqs[:300], descr=descr[:300])
    with open(plain, "rb") as f, gzip.open(sub, "wb") as g:
        g.write(f.read())
    data = str(d / "KmerData")
    synth.build_c0_fixture(sub, data)
    data_all = str(d / "KmerDataAll")
    synth.build_c0_fixture(FAA, data_all)
    faa = str(d / "q.faa")
    synth.write_fasta(faa, ids[:150], seqs[:150], descr=descr[:150])
    gids, _, gseqs = synth.read_fasta_simple(FNA)
    fna = str(d / "q.fna")
    synth.write_fasta(fna, gids[:1], [gseqs[0][:60_000]])
    return {"aa": (data, faa), "dna": (data_all, fna)}, faa, fna, d


def _both(oracle, pin, java, args, out_dir, tag):
    """The same command line through the oracle's CLI and through the transliterated KmerGutsJava.main -> stripped texts"""
    mod = java[0]
    o, j = os.path.join(out_dir, tag + ".oracle.txt"), os.path.join(out_dir, tag + ".java.txt")
    oracle.run_cli(list(args) + ["-o", o])
    mod.KmerGutsJava.main(list(args) + ["-o", j])
    return pin.strip(open(o, errors="replace").read()), pin.strip(open(j, errors="replace").read())


def test_scalar_helpers_and_tables(oracle, java):
    """KGJ:85-99, 111-175, 177-272, 274-318 against the oracle's C restatements, every byte value"""
    K = java[0].KmerGutsJava
    L = C.CDLL(oracle.lib()._name)   # a private handle: the argtypes set below must not leak into other tests' calls
    L.kgo_to_amino_acid_off.restype = L.kgo_compl.restype = L.kgo_dna_char.restype = C.c_int
    assert (K.K, K.CORE, K.MAX_ENCODED, K.MAX_HITS_PER_SEQ, K.OI_BUFSZ, K.VERSION) == (8, 20 ** 7, 20 ** 8, 40000, 5, 1)
    L.kgo_genetic_code.restype = C.c_char
    assert [c for c in K.GENETIC_CODE] == [L.kgo_genetic_code(i).decode() for i in range(64)]
    assert "".join(K.PROT_ALPHA) == synth.PROT_ALPHA
    for c in range(256):
        assert K.toAminoAcidOff(chr(c)) == L.kgo_to_amino_acid_off(c), c
        assert ord(K.compl(chr(c))) == L.kgo_compl(c), c
        assert K.dnaChar(chr(c)) == L.kgo_dna_char(c), c
    L.kgo_encoded_kmer.restype = C.c_int64
    L.kgo_encoded_kmer.argtypes = [C.c_char_p, C.c_size_t]
    rng = np.random.default_rng(5)
    for _ in range(300):
        codes = rng.integers(0, 22 if rng.random() < 0.3 else 20, 12).astype(np.uint8)
        for pos in range(5):
            assert K.encodedKmer([int(x) for x in codes], pos) == L.kgo_encoded_kmer(codes.tobytes(), pos)
    assert K.encodedKmer([19] * 8, 0) == 20 ** 8 - 1 and K.encodedKmer([1] + [0] * 7, 0) == 20 ** 7
    L.kgo_rev_comp.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    L.kgo_translate.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_char_p, C.c_char_p, C.c_size_t]
    alphabet = b"ACGTacgtuUNnRYKMSWBDHVxX-*"
    for n in list(range(0, 14)) + [50, 301]:
        seq = bytes(alphabet[int(i)] for i in rng.integers(0, len(alphabet), n)) if n < 301 else bytes(b"ACGT"[int(i)] for i in rng.integers(0, 4, n))
        rc = C.create_string_buffer(max(n, 1))
        L.kgo_rev_comp(seq, n, rc)
        assert "".join(K.revComp(list(seq.decode()))) == rc.raw[:n].decode()
        plen = n // 3 + 1                                    # KGJ:1061: the buffers are reused across the frames
        pseq, piseq = C.create_string_buffer(plen), C.create_string_buffer(plen)
        jp, ji = ["\0"] * plen, [0] * plen
        for frame in range(3):
            L.kgo_translate(seq, n, frame, pseq, piseq, plen)
            K.translate(list(seq.decode()), frame, jp, ji)
            assert [x for x in piseq.raw] == ji, (n, frame)
            assert [ord(c) if isinstance(c, str) else c for c in jp] == [x for x in pseq.raw], (n, frame)


def test_gather_hits_on_the_hand_traced_vectors(java, pin, tmp_path):
    """KGJ:457-514 + 385-455 + 516-524 on tests/golden/fsm_kats.json: the text Java prints == the text the KAT file implies"""
    mod, rt, tp = java
    pin.write_kats(str(tmp_path / "kats.txt"))
    tp.run_kats(mod, str(tmp_path / "kats.txt"), str(tmp_path / "java_kats.txt"))
    assert open(tmp_path / "java_kats.txt").read() == pin.expected_kat_text()


@pytest.mark.parametrize("flags", FLAGSETS, ids=lambda f: "_".join(f) or "default")
@pytest.mark.parametrize("mode", ["aa", "dna"])
def test_main_reports_equal_the_oracles(oracle, pin, java, small_c0, mode, flags):
    """KmerGutsJava.main (KGJ:560-654) -> run (742-820) -> readFasta, prepareQuery, the comparator sort, lookup (944-1034),
    processAASeq / processSeq: whole reports, wall-clock lines dropped, byte for byte -- incl. the -d dumps (HIT lines,
    after-hit / after-call lists, `Kmers found`), which the product's report omits but the oracle prints."""
    dirs, faa, fna, d = small_c0
    data, q = dirs[mode]
    args = (["-a"] if mode == "aa" else []) + flags + ["-D", data, "-q", q]
    o, j = _both(oracle, pin, java, args, str(d), f"{mode}_{'_'.join(flags)}")
    assert j == o
    assert o.count("\nCALL\t") > (30 if mode == "aa" else 15) and o.count("OTU-COUNTS") == (150 if mode == "aa" else 1)


def test_duplicate_ids_and_odd_fasta(oracle, pin, java, small_c0):
    """Q10 (LinkedHashMap: first position, last length, last container), Q11 (lines are appended untrimmed), Q2 (lower case is
    invalid), \\r\\n and blank lines: the same report from both"""
    dirs, faa, fna, d = small_c0
    data = dirs["aa"][0]
    ids, descr, seqs = synth.read_fasta_simple(faa)
    s = [x.decode() for x in seqs]
    text = (f">{ids[0]} first copy\n{s[0]}\n>{ids[1]}\r\n{s[1][:80]} \r\n{s[1][80:]}\r\n\n>{ids[0]} second copy, shorter\n{s[2]}\n"
            f"  >{ids[3]}\tx\n\n \n{s[3].lower()}\n{s[3]}\n>{ids[4]}\n{s[4][:50]}X{s[4][51:]}\n")
    q = str(d / "odd.faa")
    open(q, "w", newline="").write(text)
    for flags in ([], ["-d"]):
        o, j = _both(oracle, pin, java, ["-a"] + flags + ["-D", data, "-q", q], str(d), "odd" + "".join(flags))
        assert j == o
        assert o.count("PROTEIN-ID") == 4 and f"PROTEIN-ID\t{ids[0]}\t{len(s[2])}\n" in o


def _write_dir(d, img, nfun):
    os.makedirs(d, exist_ok=True)
    open(os.path.join(d, "kmer.table.mem_map"), "wb").write(img)
    synth.write_function_index(d, [f"FUNC_{i:03d}" for i in range(nfun)])
    return d


def test_lookup_on_malformed_tables(oracle, pin, java, tmp_path):
    """KGJ:959-1026 where it is most delicate: a key outside its probe chain (never found), a repeated key (the first copy
    wins), an occupied-but-unmatchable slot (20^8: extends chains), and a chain that runs off the end of the file
    (EOFException -> lookup aborted, earlier hits kept, `Error: null` in the -d report; KGJ:799-802, 1102-1103)."""
    prot = b"MKVLAAGIVGLCAHHHWYYRRDDEEFFGGHHIIKKLLMMNNPPQQ"
    wk = synth.window_keys(synth.aa_codes(prot))
    n = len(wk)
    img = bytearray(synth.build_table_image(wk, np.arange(n), np.arange(n) + 100, np.full(n, 5), np.ones(n, np.float32), num_slots=211))
    ent = np.frombuffer(img, dtype=synth.ENTRY_DTYPE, offset=24)
    occ = np.flatnonzero(ent["which"] <= synth.MAX_ENCODED)
    empty = np.flatnonzero(ent["which"] > synth.MAX_ENCODED)
    victim = occ[3]
    far = [e for e in empty if e > 0 and e + 1 < 211 and ent["which"][e - 1] > synth.MAX_ENCODED and ent["which"][e + 1] > synth.MAX_ENCODED
           and e != ent["which"][victim] % 211][0]
    ent[far] = ent[victim]
    ent["which"][victim] = synth.EMPTY_KEY
    dup = [s for s in occ if s != victim and s + 1 < 210 and ent["which"][s + 1] > synth.MAX_ENCODED and s + 1 != far][0]
    ent[dup + 1] = ent[dup]
    ent["fi"][dup + 1] = 99
    um = [e for e in empty if e not in (far, dup + 1) and e + 1 < 211][-1]
    ent["which"][um] = synth.MAX_ENCODED
    d1 = _write_dir(str(tmp_path / "malformed"), bytes(img), 100)
    q = str(tmp_path / "q.faa")
    synth.write_fasta(q, ["p1", "p2"], [prot, prot[5:] + prot[:9]])
    for flags in (["-m", "2"], ["-d", "-m", "2"]):
        o, j = _both(oracle, pin, java, ["-a"] + flags + ["-D", d1, "-q", q], str(tmp_path), "mal" + "".join(flags))
        assert j == o
        assert "CALL\t" in o and "\t99\t" not in o
    # the chain that runs off the end of the file
    prot2 = b"MKVLAAGIVGLCAHHHWYYRR"
    keys = synth.window_keys(synth.aa_codes(prot2))
    m = len(keys)
    good = synth.build_table_image(keys, np.zeros(m), np.zeros(m), np.ones(m), np.ones(m, np.float32), num_slots=101)
    ent2 = np.frombuffer(good, dtype=synth.ENTRY_DTYPE, offset=24).copy()
    ent2["which"][100] = 12345
    probe = (20 ** 8 - 1) - ((20 ** 8 - 1 - 100) % 101)
    codes, v = [], probe
    for _ in range(8):
        codes.append(v % 20)
        v //= 20
    seq2 = bytes(synth.PROT_ALPHA[c].encode()[0] for c in reversed(codes)) + b"A"
    d2 = _write_dir(str(tmp_path / "eof"), good[:24] + ent2.tobytes(), 2)
    q2 = str(tmp_path / "q2.faa")
    synth.write_fasta(q2, ["whole", "runs_off_the_end"], [prot2, seq2])
    for flags in (["-m", "2"], ["-d", "-m", "2"]):
        o, j = _both(oracle, pin, java, ["-a"] + flags + ["-D", d2, "-q", q2], str(tmp_path), "eof" + "".join(flags))
        assert j == o
    assert "Error: null" in o and "CALL\t" in o


def test_read_fasta_fuzz(oracle, java, tmp_path):
    """readFasta (KGJ:1132-1192) on adversarial text -- blank lines, lines of blanks, blanks before '>', a lone '>', \\r and
    \\r\\n, captions without a sequence, text before the first caption: same records or the same first error as the oracle's
    reader (which tests/test_host.py in turn holds the product's multi-threaded reader to)."""
    from tests.test_host import _oracle_fasta
    mod, rt, _ = java
    rng = np.random.default_rng(11)
    caps = [b">id%d desc\n", b"  >sp%d\tx y\n", b">dup\n", b">id%d\r\n", b">\t id%d  two  words \n"]
    seqs = [b"ACDEFGHIK\n", b"LMNP QRST \r\n", b"VWY\r", b"A\n", b"XX>notcaption\n", b"ACGT" * 30 + b"\n", b"\x0bAC\x0c\n"]
    blanks = [b" \n", b"\n", b"  \t \n", b">\n", b"\r\n", b"\x01\n"]
    n_err = n_ok = 0

    class Collect(rt.JObject):
        def __init__(self):
            self.rec = []

        def nextEntry(self, id_, seq, descr):
            self.rec.append((id_, seq, descr))

    for trial in range(150):
        parts = []
        for r in range(int(rng.integers(1, 20))):
            c = caps[int(rng.integers(0, len(caps)))]
            parts.append(c % (100 * trial + r) if b"%d" in c else c)
            while rng.random() < 0.3:
                parts.append(blanks[int(rng.integers(0, len(blanks)))])
            if rng.random() < 0.97:
                for _ in range(int(rng.integers(1, 5))):
                    parts.append(seqs[int(rng.integers(0, len(seqs)))])
                    if rng.random() < 0.2:
                        parts.append(blanks[int(rng.integers(0, 3))])
        body = b"".join(parts)
        if trial % 10 == 0:
            body = b"garbage before the first caption\n" + body
        if trial % 7 == 0:
            body = b"\n \nA\n" + body
        if trial % 9 == 0 and body.endswith(b"\n"):
            body = body[:-1]
        path = str(tmp_path / f"fz{trial}.fa")
        open(path, "wb").write(body)
        want = _oracle_fasta(oracle, path)
        cb = Collect()
        try:
            mod.KmerGutsJava.readFasta(rt.BufferedReader(rt.FileReader(rt.File(path))), cb)
            got = cb.rec
        except rt.IllegalStateException as e:
            got = e.getMessage()
        # ... and the PRODUCT's multi-threaded reader (kg_fasta_read, host part of the library: no GPU involved) against the
        # Java source directly, cut into tiny ranges so that its two parsing passes meet every kind of range boundary
        import kmergutsjava_b200 as kg
        for chunk in ("1", "53"):
            os.environ["KG_FASTA_CHUNK"] = chunk
            try:
                f = kg.Fasta(path)
                prod = (f.ids, bytes(f.bytes), [int(x) for x in f.offsets])
                f.free()
            except kg.KgError as e:
                prod = str(e)
            finally:
                os.environ.pop("KG_FASTA_CHUNK", None)
            if isinstance(got, str):
                assert isinstance(prod, str) and got in prod, (trial, chunk, prod, got)
            else:
                assert prod == ([g[0] for g in got], "".join(g[1] for g in got).encode("latin-1"),
                                [int(x) for x in np.cumsum([0] + [len(g[1]) for g in got])]), (trial, chunk, body)
        if isinstance(want, str):
            assert got == want, (trial, body)
            n_err += 1
        else:
            assert isinstance(got, list), (trial, got, body)
            assert [g[0] for g in got] == want[0], (trial, body)
            assert "".join(g[1] for g in got).encode("latin-1") == bytes(want[1]), (trial, body)
            assert np.array_equal(np.cumsum([0] + [len(g[1]) for g in got]), want[2]), (trial, body)
            n_ok += 1
    assert n_err > 15 and n_ok > 40


def test_function_index_reader(java, tmp_path):
    """loadIndexedArray, KGJ:345-373"""
    mod, rt, _ = java
    p = tmp_path / "function.index"
    p.write_text("0\thypothetical protein\n1\tDNA polymerase (EC 2.7.7.7)\n2\t\n")
    names = mod.KmerGutsJava.loadIndexedArray(rt.File(str(p)))
    assert [names.get(i) for i in range(names.size())] == ["hypothetical protein", "DNA polymerase (EC 2.7.7.7)", ""]
    p.write_text("0\ta\n2\tb\n")
    with pytest.raises(rt.IllegalStateException) as e:
        mod.KmerGutsJava.loadIndexedArray(rt.File(str(p)))
    assert e.value.getMessage() == "Your index must be dense and in order (see line 1)"


def test_the_tool_refuses_what_it_does_not_model():
    """j2py.py must fail loudly on Java it has no rule for, never guess"""
    j2py = _load("j2py", os.path.join(ROOT, "tests", "java_pin", "j2py.py"))
    for bad in ("class A { void f() { outer: for (;;) { break outer; } } }",
                "class A { void f() { try (java.io.Reader r = null) { } } }",
                "class A { <T> T f(T x) { return x; } }",
                "class A { void f(int x) { switch (x) { default: return; case 1: return; } } }",
                "class A { void f(int[] a, int i) { int y = a[i]++ + 1; } }"):
        with pytest.raises(j2py.ParseError):
            j2py.transliterate(bad)
    # and a few constructs it does model, checked against hand-computed Java results
    src = """class T {
      static int wrap() { int x = 2147483647; x += 1; return x; }
      static long mix(int a) { long v = 0; for (int i = 0; i < 5; i++) { if (i == 2) continue; v = v * 31 + (a << i); } return v; }
      static int div() { return (-7) / 2 * 10 + (-7) % 3; }
      static float fsum() { float s = 0; for (int i = 0; i < 10; i++) s += 0.1f; return s; }
      static int sw(char c) { int r = 0; switch (c) { case 'a': r += 1; case 'b': r += 10; break; case 'c': r += 100; default: r += 1000; } return r; }
      static int post() { int i = 3; int[] a = new int[8]; a[i++] = i; a[i--] = i + 10; return a[3] * 100 + a[4] + i; }
      static String cat(int n) { return "n=" + n + 1 + (n + 1) + 'c' + 1.5 + null + true; }
    }"""
    ns = {}
    sys.path.insert(0, os.path.join(ROOT, "tests", "java_pin"))
    exec(j2py.transliterate(src), ns)
    T = ns["T"]
    assert T.wrap() == -2147483648
    assert T.mix(3) == ((((3 << 0) * 31 + (3 << 1)) * 31 + (3 << 3)) * 31 + (3 << 4))
    assert T.div() == -31
    assert T.fsum() == float(np.float32(sum([np.float32(0.1)] * 10, np.float32(0))))
    assert [T.sw(c) for c in "abcd"] == [11, 10, 1100, 1000]
    assert T.post() == 4 * 100 + 13 + 3   # the index is evaluated before the right-hand side
    assert T.cat(4) == "n=415c1.5nulltrue"


def _java_fsm(java, flags, hits_rec):
    """gatherHits (KGJ:457) + tabulateOtuDataForContig (KGJ:516) of the transliterated source on one container -> the text"""
    import io
    mod, rt, _ = java
    k = mod.KmerGutsJava()
    k.minHits, k.maxGap = flags.get("min_hits", 5), flags.get("max_gap", 200)
    k.minWeightedHits, k.orderConstraint = flags.get("min_weighted_hits", 0), bool(flags.get("order_constraint", False))
    lst, max_fi = rt.ArrayList(), 0
    for h in hits_rec:
        x = mod.Hit()
        x.from0InProt, x.fI, x.oI, x.avgOffFromEnd, x.functionWt = int(h["pos"]), int(h["fI"]), int(h["oI"]), int(h["avg"]), float(h["wt"])
        max_fi = max(max_fi, x.fI)
        lst.add(x)
    buf = io.StringIO()
    pw = rt.PrintWriter(buf)
    k.gatherHits(0, "+", 0, lst, rt.ArrayList(["F%d" % i for i in range(max_fi + 1)]), otu := rt.ArrayList(), pw)
    k.tabulateOtuDataForContig("x", 0, otu, pw)
    return buf.getvalue()


def _oracle_fsm_text(oracle, flags, hits_rec):
    calls, otu = oracle.gather_hits(oracle.make_params(aa=True, **flags), hits_rec, max_calls=1 << 16)
    out = [f"CALL\t{int(c['start'])}\t{int(c['end'])}\t{int(c['count'])}\t{int(c['fI'])}\tF{int(c['fI'])}\t"
           f"{oracle.java_format_f(float(c['weighted']))}\n" for c in calls]
    k = int(otu["n"][0])
    out.append("OTU-COUNTS\tx[0]" + "".join(f"\t{int(otu['count'][0][j])}-{int(otu['oI'][0][j])}" for j in range(k)) + "\n")
    return "".join(out)


@pytest.mark.parametrize("seed", range(6))
def test_fsm_fuzz_java_source_vs_oracle(oracle, java, seed):
    """Random adversarial hit lists (runs of one function, pair switches, singletons, gaps around max_gap, ties in the OTU
    counts, order-constraint offsets, fp32 weights) through the SOURCE's gatherHits / processSetOfHits and through the oracle's:
    the printed CALL and OTU-COUNTS lines must be the same text (the generator is test_oracle_cross's)."""
    rng = np.random.default_rng(900 + seed)
    for case in range(30):
        n = int(rng.integers(0, 140))
        pos = np.sort(rng.choice(3000, size=n, replace=False)) if n else np.zeros(0, int)
        if case % 3 == 0 and n:
            pos = np.unique(np.cumsum(rng.choice([1, 1, 2, 5, 190, 200, 201, 230], size=n)))
            n = len(pos)
        fI = rng.integers(1, 4, size=n)
        for i in range(1, n):
            if rng.random() < 0.7:
                fI[i] = fI[i - 1]
        oI = rng.integers(0, 8, size=n)
        avg = (3000 - pos + rng.integers(-25, 25, size=n)) if n else np.zeros(0, int)
        wt = (rng.integers(1, 600, size=n) / 256.0).astype(np.float32)
        if case % 5 == 1:
            wt = rng.random(n).astype(np.float32)   # weights that are not dyadic: the fp32 sum and its %f rounding matter
        flags = dict(order_constraint=bool(rng.integers(2)), min_hits=int(rng.integers(2, 7)), min_weighted_hits=int(rng.integers(0, 4)),
                     max_gap=int(rng.choice([0, 5, 50, 200, 1000])))
        hits = np.zeros(n, dtype=oracle.HIT_DTYPE)
        hits["pos"], hits["fI"], hits["oI"], hits["avg"], hits["wt"] = pos, fI, oI, avg, wt
        assert _java_fsm(java, flags, hits) == _oracle_fsm_text(oracle, flags, hits), (seed, case, flags)


def test_open_run_cap_java_source_vs_oracle(oracle, java):
    """Q9 (KGJ:496-504): MAX_HITS_PER_SEQ - 2 = 39998 hits in an open run; later hits are not appended, yet the pair-switch test
    still runs on them -- 41 000 hits of one function, then a foreign pair, then a new run"""
    n = 41000
    hits = np.zeros(n + 12, dtype=oracle.HIT_DTYPE)
    hits["pos"] = np.arange(n + 12)
    hits["fI"][:n] = 7
    hits["fI"][n:n + 2] = 9
    hits["fI"][n + 2:] = 7
    hits["oI"] = np.arange(n + 12) % 3
    hits["avg"] = 50000 - np.arange(n + 12)
    hits["wt"] = np.float32(0.5)
    for flags in (dict(), dict(order_constraint=True)):
        want = _oracle_fsm_text(oracle, flags, hits)
        assert "\t39998\t7\t" in want
        assert _java_fsm(java, flags, hits) == want


SYN_FLAGS = [[], ["-O"], ["-m", "3", "-g", "50", "-M", "2"], ["-m", "2", "-g", "5"], ["-d", "-m", "4"]]


@pytest.fixture(scope="module")
def synthetic(tmp_path_factory):
    """A seeded synthetic universe (tools/kg_synth: 60 families, homologs with substitutions, decoys, X) as KmerData + FASTA,
    with the edge sequences of SURVEY 8(c): shorter than / exactly K and K+1 residues, lower case, X and *, DNA shorter than a
    codon, lower-case DNA, U for T, IUPAC codes."""
    d = tmp_path_factory.mktemp("syn")
    u = synth.Universe(n_families=60, seed=0x4B470001)
    keys, otu, avg, fi, wt = u.signatures()
    data = _write_dir(str(d / "KmerData"), synth.build_table_image(keys, otu, avg, fi, wt), int(fi.max()) + 1)
    prots = u.proteins(40, seed=5) + [b"A", b"ACDEFGH", b"ACDEFGHI", b"ACDEFGHIK", b"acdefghiklmnp", b"ACDEFGHIKXLMNPQRSTVWY*ACDEFGHIK"]
    prots.append(u.proteins(1, seed=6)[0].lower() + u.proteins(1, seed=6)[0])
    faa = str(d / "q.faa")
    synth.write_fasta(faa, [f"p{i}" for i in range(len(prots))], prots)
    g0, g1 = synth.genome(u, 6000, seed=7, index=0), synth.genome(u, 6000, seed=7, index=1)
    iupac = bytearray(synth.genome(u, 3001, seed=8))
    for i, c in zip(range(50, 3000, 97), b"NRYKMSWBDHVnrykmswbdhv-*" * 2):
        iupac[i] = c
    dnas = [g0, g1.lower(), g0.replace(b"T", b"U"), bytes(iupac), b"AC", b"ATG", b"ATGAAACCCGGGTTTACGTACGTAGCTAGCTAGCATCGATCGAT",
            synth.genome(u, 3002, seed=9)]
    fna = str(d / "q.fna")
    synth.write_fasta(fna, [f"c{i}" for i in range(len(dnas))], dnas)
    return data, faa, fna, d


@pytest.mark.parametrize("flags", SYN_FLAGS, ids=lambda f: "_".join(f) or "default")
@pytest.mark.parametrize("mode", ["aa", "dna"])
def test_synthetic_universe_reports(oracle, pin, java, synthetic, mode, flags):
    data, faa, fna, d = synthetic
    args = (["-a"] if mode == "aa" else []) + flags + ["-D", data, "-q", faa if mode == "aa" else fna]
    o, j = _both(oracle, pin, java, args, str(d), f"syn_{mode}_{'_'.join(flags)}")
    assert j == o
    assert o.count("\nCALL\t") > 5
    if mode == "dna":
        assert all(f"\t{s}\t{f}\n" in o for s in "+-" for f in range(3))


def test_runtime_follows_the_java_api():
    """Spot checks of tests/java_pin/j2py_runtime.py against behaviour the Java API documents (values below are what a JVM
    prints / throws; none of this needs the reference checkout, but the module is skipped as a whole without it)."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "java_pin"))
    import j2py_runtime as rt
    f32 = rt._f32
    # Formatter %f: float widened to double, shortest repr, HALF_UP  (C's printf gives 0.007812 for the first one)
    assert rt.String.format_("%f", 0.0078125) == "0.007813"
    assert rt.String.format_("%f", f32(1.0000001)) == "1.000000" and rt.String.format_("%f", f32(0.1) * 1) == "0.100000"
    assert rt.String.format_("%f", 2.5) == "2.500000" and rt.String.format_("%f", 0) == "0.000000"
    assert rt.String.format_("%1.3f", f32(0.0005)) == "0.001" and rt.String.format_("%1.3f", 1.0005) == "1.001"   # 1.0005 is 1.000499999999999989... in binary: HALF_UP works on the DECIMAL repr
    assert rt.String.format_("%1.3f", -0.25) == "-0.250"
    assert rt.String.format_("CALL\t%d\t%s\t%c|%5d|%-4s|%%", 7, None, "x", 42, "ab") == "CALL\t7\tnull\tx|   42|ab  |%"
    # int / long / float arithmetic
    assert rt._i32(0x7FFFFFFF + 1) == -0x80000000 and rt._i64(2 ** 63) == -2 ** 63 and rt._i8(200) == -56
    assert rt._idiv(-7, 2) == -3 and rt._idiv(7, -2) == -3 and rt._rem(-7, 3) == -1 and rt._rem(7, -3) == 1
    assert rt._cast_int(3.99) == 3 and rt._cast_int(-3.99) == -3 and rt._cast_int(1e20) == 0x7FFFFFFF and rt._cast_int(float("nan")) == 0
    assert rt._cast_long(0xFF) << 56 == 0xFF << 56 and rt._i64(0xFF << 56) == -(1 << 56)
    with pytest.raises(rt.ArithmeticException):
        rt._rem(5, 0)
    assert rt.Float.intBitsToFloat(0x3F000000) == 0.5 and rt.Float.intBitsToFloat(-1) != rt.Float.intBitsToFloat(-1)   # NaN
    assert rt.Math.abs_(-0x80000000) == -0x80000000 and rt.Math.abs_(-5) == 5
    # String
    assert rt._s_trim("\x00\t a b \r\n\x1f") == "a b" and rt._s_trim(" x") == " x"   # only <= U+0020 goes
    assert rt._cat("n=", 1.0) == "n=1.0" and rt._cat("x", 1e7) == "x1.0E7" and rt._cat(None, True) == "nulltrue"
    with pytest.raises(rt.StringIndexOutOfBoundsException):
        rt._s_substring("abc", 0, -1)
    assert rt._s_hashCode("hello") == 99162322 and rt._s_hashCode("polygenelubricants") == -0x80000000
    st = rt.StringTokenizer(" a\tb  c ", " \t")
    assert [st.nextToken() for _ in range(3)] == ["a", "b", "c"] and not st.hasMoreTokens()
    for bad in ("", " 1", "1 ", "1.0", "2147483648", None):
        with pytest.raises(rt.NumberFormatException):
            rt.Integer.parseInt(bad)
    assert rt.Integer.parseInt("-2147483648") == -0x80000000 and rt.Integer.parseInt("+7") == 7
    # collections
    a = rt.ArrayList()
    a.add("x")
    for i in (-1, 1):
        with pytest.raises(rt.IndexOutOfBoundsException):
            a.get(i)
    m = rt.LinkedHashMap()
    for k, v in (("b", 1), ("a", 2), ("b", 3)):
        m.put(k, v)
    assert [k for k in m.keySet()] == ["b", "a"] and m.get("b") == 3 and m.get("zz") is None   # re-insertion keeps the place
    lst = rt.ArrayList([3, 1, 2, 1])

    class ByParity(rt.JObject):
        def compare(self, x, y):
            return rt.Integer.compare(x % 2, y % 2)
    rt.Collections.sort(lst, ByParity())
    assert list(lst) == [2, 3, 1, 1]   # stable
    # BufferedReader.readLine: \n, \r and \r\n end a line; no empty line after a final terminator
    br = rt.BufferedReader(rt.InputStreamReader(rt.InputStream(b"a\nb\r\nc\rd\n\n e ")))
    lines = []
    while (line := br.readLine()) is not None:
        lines.append(line)
    assert lines == ["a", "b", "c", "d", "", " e "]
    s = rt.InputStream(bytes([1, 255]))
    assert [s.read(), s.read(), s.read()] == [1, 255, -1] and rt.InputStream(b"abcdef").skip(10) == 6


def test_external_merge_drops_query_kmers_q4(java, synthetic, tmp_path):
    """SURVEY.md Q4 (KGJ:683-715, 832-853): when more than inputSizeLimit query k-mers are in flight, the source spills sorted
    runs to files and merges them -- and mergeTwoQueryKmerFiles copies only ONE leftover record after its main loop (KGJ:705-709),
    so the tail of the longer run is lost and the lookup sees fewer k-mers.  That is why parity is defined on the in-RAM path
    (<= 20 M k-mers per run, the default limit; sequences are independent, so larger inputs are batches of that).  Executed here
    through the transliterated source: same input, in RAM and with a small limit."""
    import re
    mod, rt, _ = java
    data, faa, fna, d = synthetic

    def run(limit, name):
        k = mod.KmerGutsJava()
        k.aa, k.debug = True, True
        if limit:
            k.inputSizeLimit, k.tempDirPath = limit, str(tmp_path / f"spill{limit}")
        out = str(tmp_path / name)
        pw = rt.PrintWriter(rt.FileWriter(rt.File(out)))
        k.run(rt.File(data), rt.File(faa), pw, False)
        pw.close()
        text = open(out).read()
        m = re.search(r"Kmers found: (\d+) \(pos-count=(\d+)\)", text)
        return int(m.group(1)), int(m.group(2)), text

    found_ram, pos_ram, _ = run(0, "ram.txt")
    assert pos_ram > 500
    lost = []
    for limit in (4000, 1500, 400):
        found, pos, text = run(limit, f"ext{limit}.txt")
        assert not os.listdir(str(tmp_path / f"spill{limit}"))      # the temporary files are deleted (KGJ:874-887)
        assert pos <= pos_ram and found <= found_ram
        lost.append(pos_ram - pos)
    assert max(lost) > 0, "the external merge kept every k-mer: Q4 would not be a quirk"


def test_fenced_quirks_q3_q8_as_the_source_behaves(java, synthetic, monkeypatch):
    """The two behaviours the product deliberately does NOT reproduce, shown on the source:
    Q8  minHits = 1: processSetOfHits indexes hits[n-2] of a one-hit run (KGJ:442) -> IndexOutOfBounds; the ABI requires min_hits >= 2.
    Q3  -t / -l can never succeed: `case 't'` has no break and falls into `case 'l'`, which has none either and falls into
        `default`, which throws (KGJ:605-610); after ANY flag error main prints the usage text and carries on into
        new File(null) (KGJ:616-647).  The product's command line accepts and ignores both and exits 2 on a flag error."""
    mod, rt, _ = java
    data, faa, fna, d = synthetic
    k = mod.KmerGutsJava()
    k.minHits = 1
    h = mod.Hit()
    h.from0InProt, h.fI, h.oI, h.functionWt, h.avgOffFromEnd = 10, 3, 1, 1.0, 100
    h2 = mod.Hit()
    h2.from0InProt, h2.fI, h2.oI, h2.functionWt, h2.avgOffFromEnd = 500, 3, 1, 1.0, 100
    import io
    with pytest.raises(rt.IndexOutOfBoundsException):
        k.gatherHits(0, "+", 0, rt.ArrayList([h, h2]), rt.ArrayList(["F%d" % i for i in range(5)]), rt.ArrayList(), rt.PrintWriter(io.StringIO()))
    # -l 5: the value parses, then `default` throws.  -t 5: falls into case 'l', which polls the NEXT flag as its number.
    for flag, message in (("-l", "Error: Unknown parameter: -l"), ("-t", 'Error: For input string: "-D"')):
        buf = io.StringIO()
        monkeypatch.setattr(rt.System, "out", rt.PrintStream(buf))
        with pytest.raises(rt.NullPointerException):
            mod.KmerGutsJava.main([flag, "5", "-D", data, "-q", faa])
        out = buf.getvalue()
        assert message in out and "Usage: kmer_guts [options] -D DataDir" in out
