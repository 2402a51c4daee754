"""Host-side mirror of run()'s CPU parts (FASTA, function.index, %f) -- no GPU needed.  The reference's behaviours
under test are cited from KmerGutsJava.java (KGJ)."""
import gzip
import os

import numpy as np
import pytest

import kmergutsjava_b200 as kg


def _write(tmp_path, name, data: bytes):
    p = tmp_path / name
    with (gzip.open(p, "wb") if name.endswith(".gz") else open(p, "wb")) as f:
        f.write(data)
    return str(p)


def test_fasta_basic_and_gz(tmp_path):
    body = b">id1 some description here\nACDE\nFGHI\n\n>id2\tx\nKLMN\n"
    for name in ("a.fa", "a.fa.gz"):
        f = kg.Fasta(_write(tmp_path, name, body))
        assert f.ids == ["id1", "id2"]
        assert bytes(f.bytes) == b"ACDEFGHIKLMN" and list(f.offsets) == [0, 8, 12]
        f.free()


def test_fasta_lines_are_not_trimmed(tmp_path):
    # KGJ:1176 appends the raw line: inner and trailing blanks become (invalid) residues; \r\n is a line end
    f = kg.Fasta(_write(tmp_path, "b.fa", b">x\r\nAC DE \r\n  \r\nFG\r\n>y\nAA\n"))
    assert f.ids == ["x", "y"]
    assert bytes(f.bytes[: int(f.offsets[1])]) == b"AC DE   FG"
    f.free()


def test_fasta_short_lines_before_caption_are_skipped(tmp_path):
    # KGJ:1145: only lines whose trimmed length is > 1 are examined while looking for a caption
    f = kg.Fasta(_write(tmp_path, "c.fa", b"\n \nA\n>s1\nACD\n"))
    assert f.ids == ["s1"] and bytes(f.bytes) == b"ACD"
    f.free()


def test_fasta_errors(tmp_path):
    with pytest.raises(kg.KgError, match="Wrong caption line: ACDEF"):     # KGJ:1158
        kg.Fasta(_write(tmp_path, "d.fa", b"ACDEF\n>s\nAA\n"))
    with pytest.raises(kg.KgError, match="No sequence for caption: s1"):   # KGJ:1170
        kg.Fasta(_write(tmp_path, "e.fa", b">s1\n>s2\nAA\n"))
    with pytest.raises(kg.KgError, match="No sequence for caption: s2"):
        kg.Fasta(_write(tmp_path, "f.fa", b">s1\nAA\n>s2\n\n"))
    with pytest.raises(kg.KgError):
        kg.Fasta(str(tmp_path / "missing.fa"))


def test_fasta_matches_oracle_reader_on_ecoli(oracle):
    import ctypes as C
    path = os.path.join(os.path.dirname(__file__), "data", "Ecoli_K12_W3110.faa.gz")
    f = kg.Fasta(path)
    assert f.n == 13645 and int(f.offsets[-1]) == 4147102
    L = oracle.lib()
    L.kgo_fasta_read.restype = C.c_void_p
    L.kgo_fasta_read.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]

    class KF(C.Structure):
        _fields_ = [("n", C.c_size_t), ("id", C.POINTER(C.c_char_p)), ("seq", C.POINTER(C.c_uint8)), ("off", C.POINTER(C.c_uint64))]
    err = C.create_string_buffer(256)
    h = L.kgo_fasta_read(path.encode(), err, 256)
    o = C.cast(h, C.POINTER(KF)).contents
    assert o.n == f.n
    assert [o.id[i].decode() for i in range(0, f.n, 997)] == f.ids[::997]
    assert np.array_equal(np.ctypeslib.as_array(o.off, (f.n + 1,)), f.offsets)
    assert np.array_equal(np.ctypeslib.as_array(o.seq, (int(f.offsets[-1]),)), f.bytes)
    f.free()


def _oracle_fasta(oracle, path):
    """(ids, bytes, offsets) from the oracle's reader, or its error message."""
    import ctypes as C
    L = oracle.lib()
    L.kgo_fasta_read.restype = C.c_void_p
    L.kgo_fasta_read.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]

    class KF(C.Structure):
        _fields_ = [("n", C.c_size_t), ("id", C.POINTER(C.c_char_p)), ("seq", C.POINTER(C.c_uint8)), ("off", C.POINTER(C.c_uint64))]
    err = C.create_string_buffer(512)
    h = L.kgo_fasta_read(path.encode(), err, 512)
    if not h:
        return err.value.decode()
    o = C.cast(h, C.POINTER(KF)).contents
    off = np.ctypeslib.as_array(o.off, (o.n + 1,)).copy()
    seq = np.ctypeslib.as_array(o.seq, (max(int(off[-1]), 1),))[: int(off[-1])].copy()
    return [o.id[i].decode("latin1") for i in range(o.n)], seq, off


def test_fasta_parallel_reader_fuzz(oracle, tmp_path, monkeypatch):
    """The reader cuts the text at caption lines and parses the ranges on several threads (KG_FASTA_CHUNK forces many tiny
    ranges here).  Whatever the text looks like -- blank lines, lines of blanks, leading blanks before '>', lone '>', \r and
    \r\n line ends, captions without a sequence -- it must return exactly what the oracle's sequential reader returns,
    including which error comes first."""
    rng = np.random.default_rng(11)
    caps = [b">id%d desc\n", b"  >sp%d\tx y\n", b">dup\n", b">id%d\r\n"]
    seqs = [b"ACDEFGHIK\n", b"LMNP QRST \r\n", b"VWY\r", b"A\n", b"XX>notcaption\n", b"ACGT" * 30 + b"\n"]
    blanks = [b" \n", b"\n", b"  \t \n", b">\n", b"\r\n"]
    n_err = n_ok = 0
    for trial in range(120):
        parts = []
        for r in range(int(rng.integers(1, 25))):
            c = caps[int(rng.integers(0, len(caps)))]
            parts.append(c % (100 * trial + r) if b"%d" in c else c)
            while rng.random() < 0.3:
                parts.append(blanks[int(rng.integers(0, len(blanks)))])
            if rng.random() < 0.97:  # now and then a caption without a sequence: an error, and it must be the FIRST one reported
                for _ in range(int(rng.integers(1, 5))):
                    parts.append(seqs[int(rng.integers(0, len(seqs)))])
                    if rng.random() < 0.2:
                        parts.append(blanks[int(rng.integers(0, 3))])
        body = b"".join(parts)
        if trial % 10 == 0:
            body = b"garbage before the first caption\n" + body
        if trial % 7 == 0:
            body = b"\n \nA\n" + body
        path = _write(tmp_path, f"fz{trial}.fa", body)
        want = _oracle_fasta(oracle, path)
        for chunk in ("1", "37", "100000000"):
            monkeypatch.setenv("KG_FASTA_CHUNK", chunk)
            if isinstance(want, str):
                with pytest.raises(kg.KgError) as e:
                    kg.Fasta(path)
                assert want in str(e.value), (trial, chunk, body)
                n_err += 1
            else:
                f = kg.Fasta(path)
                assert f.ids == want[0], (trial, chunk, body)
                assert np.array_equal(f.offsets, want[2]) and bytes(f.bytes) == bytes(want[1]), (trial, chunk, body)
                f.free()
                n_ok += 1
    assert n_err > 20 and n_ok > 60


def test_function_index(tmp_path):
    import ctypes as C
    L = kg.lib()
    p = _write(tmp_path, "function.index", b"0\thypothetical protein\n1\tDNA polymerase (EC 2.7.7.7)\n2\t\n")
    h = C.c_void_p()
    assert L.kg_functions_load(str(tmp_path).encode(), C.byref(h)) == 0
    assert L.kg_functions_count(h) == 3
    assert L.kg_functions_name(h, 1) == b"DNA polymerase (EC 2.7.7.7)" and L.kg_functions_name(h, 2) == b""
    L.kg_functions_free(h)
    bad = _write(tmp_path, "bad.index", b"0\ta\n2\tb\n")
    assert L.kg_functions_read(bad.encode(), C.byref(h)) == -5
    assert b"dense and in order (see line 1)" in L.kg_last_error()     # KGJ:361-364


def test_java_format_matches_oracle(oracle):
    rng = np.random.default_rng(3)
    vals = [0.0, 1 / 128, 3 / 128, 2.5, 0.1, 1.0000001, 123456.0, 0.9999999403953552, 1e-7, 5e-7, 4.9999997e-7]
    vals += list(rng.integers(0, 1 << 16, 300) / 256.0) + list(rng.random(300) * 100) + list(rng.integers(0, 4096, 200) / 128.0)
    for v in vals:
        v = float(np.float32(v))
        for prec in (6, 3):
            assert kg.java_format_f(v, prec) == oracle.java_format_f(v, prec), (v, prec)
    assert kg.java_format_f(1 / 128) == "0.007813"


def test_pack_aa_layout_and_codes():
    """kg_pack_aa (include/kmerguts.h): toAminoAcidOff codes (KGJ:111-175), 8 per 5 bytes, ceil((len+1)/8) groups per sequence
    (at least one padding code 31 after the last residue); one thread and many threads give the same bytes."""
    import kmergutsjava_b200 as kg
    from oracle import kgo
    rng = np.random.default_rng(7)
    alpha = np.frombuffer(b"ACDEFGHIKLMNPQRSTVWYXacdx*-\x00 BZUO", np.uint8)
    seqs = [bytes(rng.choice(alpha, int(L))) for L in list(rng.integers(0, 70, 1500)) + [0, 1, 7, 8, 9, 15, 16, 17, 4000]]
    sb, off = kgo.concat(seqs)
    packed, goff = kg.pack_aa(sb, off, threads=1)
    packed_mt, goff_mt = kg.pack_aa(sb, off, threads=5)
    assert np.array_equal(packed, packed_mt) and np.array_equal(goff, goff_mt)
    assert [int(goff[i + 1] - goff[i]) for i in range(len(seqs))] == [(len(s) + 1 + 7) // 8 for s in seqs]
    assert all(kg.lib().kg_pack_aa_groups(L) == (L + 1 + 7) // 8 for L in (0, 1, 7, 8, 1000))
    code = {c: i for i, c in enumerate(b"ACDEFGHIKLMNPQRSTVWY")}
    for i in (0, 3, 700, len(seqs) - 1, len(seqs) - 2, len(seqs) - 9):
        s = seqs[i]
        raw = packed[5 * int(goff[i]):5 * int(goff[i + 1])]
        bits = int.from_bytes(raw.tobytes(), "little")
        got = [(bits >> (5 * k)) & 31 for k in range(8 * int(goff[i + 1] - goff[i]))]
        assert got[:len(s)] == [code.get(c, 20) for c in s]
        assert got[len(s):] == [31] * (len(got) - len(s)) and len(got) > len(s)


def test_fasta_stream_equals_whole_file(tmp_path, monkeypatch):
    """kg_fasta_stream_*: the batches, concatenated, are exactly what kg_fasta_read gives for the whole file -- for any batch
    size, line ending, blank lines, records longer than a batch, plain and .gz."""
    import gzip
    rng = np.random.default_rng(5)
    recs = []
    for i in range(400):
        n = int(rng.integers(1, 900)) if i % 37 else 20000
        seq = "".join(rng.choice(list("ACDEFGHIKLMNPQRSTVWYXacgt "), n))
        w = int(rng.integers(20, 90))
        lines = [seq[j:j + w] for j in range(0, n, w)]
        if i % 11 == 0:
            lines.insert(1, "")            # a blank line inside a record is kept out of the sequence but ends nothing
        recs.append((f">seq{i} description {i}", lines))
    for eol in ("\n", "\r\n", "\r"):
        text = "".join(cap + eol + eol.join(lines) + eol for cap, lines in recs)
        plain = tmp_path / "q.fa"
        plain.write_bytes(text.encode())
        gz = tmp_path / "q.fa.gz"
        with gzip.open(gz, "wb") as f:
            f.write(text.encode())
        whole = kg.Fasta(str(plain))
        for path in (plain, gz):
            for batch in (1, 700, 5000, 64 << 10, 1 << 30):
                ids, chunks, lens, nb = [], [], [], 0
                for b in kg.fasta_batches(str(path), batch):
                    ids += b.ids
                    chunks.append(b.bytes)
                    lens += list(np.diff(b.offsets.astype(np.int64)))
                    assert b.n > 0
                    nb += 1
                    b.free()
                assert ids == whole.ids, (eol, batch)
                assert lens == list(np.diff(whole.offsets.astype(np.int64)))
                assert np.array_equal(np.concatenate(chunks), whole.bytes)
                if batch == 1:
                    assert nb == len(recs)       # every record alone
                if batch == 1 << 30:
                    assert nb == 1
        whole.free()
    # a reader error (caption without a sequence) surfaces from the batch that holds it
    bad = tmp_path / "bad.fa"
    bad.write_text(">a\nACGT\n>b\n>c\nAC\n")
    with pytest.raises(kg.KgError) as e:
        for b in kg.fasta_batches(str(bad), 1):
            b.free()
    assert "No sequence for caption: b" in str(e.value)


def test_call_dna_range(oracle):
    """kg_call_dna_range: the nucleotides it names, read on the strand it names, translate to the residues of the call."""
    rng = np.random.default_rng(9)
    L = 3001
    contig = "".join(rng.choice(list("ACGT"), L))
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rc = "".join(comp[c] for c in reversed(contig))
    code = "KNKNTTTTRSRSIIMIQHQHPPPPRRRRLLLLEDEDAAAAGGGGVVVV*Y*YSSSS*CWCLFLF"   # KGJ:88-93
    idx = {"A": 0, "C": 1, "G": 2, "T": 3}

    def translate(s, off):
        return "".join(code[idx[s[i]] * 16 + idx[s[i + 1]] * 4 + idx[s[i + 2]]] for i in range(off, len(s) - 2, 3))

    for sf in range(6):
        strand_seq = contig if sf < 3 else rc
        prot = translate(strand_seq, sf % 3)
        for _ in range(50):
            a = int(rng.integers(0, len(prot) - 8))
            b = int(rng.integers(a + 7, len(prot)))
            call = {"seq": 0, "sf": sf, "start": a, "end": b, "count": 5, "fI": 0, "weighted": 1.0, "hits_before": 0}
            begin, end, sd = kg.call_dna_range(call, L)
            assert sd == ("+" if sf < 3 else "-") and 0 <= begin <= end < L and (end - begin + 1) == 3 * (b - a + 1)
            piece = contig[begin:end + 1]
            if sd == "-":
                piece = "".join(comp[c] for c in reversed(piece))
            assert translate(piece, 0) == prot[a:b + 1]
    with pytest.raises(kg.KgError):
        kg.call_dna_range({"seq": 0, "sf": 0, "start": 0, "end": L, "count": 5, "fI": 0, "weighted": 1.0, "hits_before": 0}, L)


def test_pack_dna_layout_and_codes():
    """kg_pack_dna: 2 bits per nucleotide in dnaChar's codes (KGJ:294-318), four per byte, low bits first, every contig on a byte
    boundary; every other character in the exception list; any number of threads gives the same bytes."""
    rng = np.random.default_rng(3)
    alphabet = np.frombuffer(b"ACGTacgtUuNnRYKMSWBDHV*- X", dtype=np.uint8)
    seqs = [b"", b"A", b"ACG", b"ACGT", b"ACGTN", b"nnnn"] + [bytes(rng.choice(alphabet, int(rng.integers(0, 3000)), p=None)) for _ in range(60)]
    seqs.append(bytes(rng.choice(alphabet[:4], 1 << 18)))
    sb = np.frombuffer(b"".join(seqs), dtype=np.uint8)
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in seqs])
    code = np.full(256, 4, dtype=np.uint8)
    for ch, c in zip(b"aAcCgGtTuU", [0, 0, 1, 1, 2, 2, 3, 3, 3, 3]):
        code[ch] = c
    want_exc = np.nonzero(code[sb] == 4)[0].astype(np.uint64)
    ref = None
    for threads in (1, 3, 16):
        packed, boff, exc = kg.pack_dna(sb, off, threads=threads)
        assert np.array_equal(exc, want_exc)
        assert list(np.diff(boff.astype(np.int64))) == [(len(x) + 3) // 4 for x in seqs]
        if ref is None:
            ref = packed.copy()
            for s, x in enumerate(seqs):
                b = packed[int(boff[s]):int(boff[s + 1])]
                dec = np.stack([(b >> (2 * k)) & 3 for k in range(4)], axis=1).reshape(-1)[:len(x)]
                c = code[np.frombuffer(x, dtype=np.uint8)]
                assert np.array_equal(dec[c != 4], c[c != 4])
                assert not dec[c == 4].any()          # an exception's two bits are 0
        else:
            assert np.array_equal(packed, ref)
