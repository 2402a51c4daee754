"""The CPU oracle against the hand-traced known-answer vectors (tests/golden/fsm_kats.json) and the scalar KATs of
SURVEY.md section 8(c).  The reference has no golden vectors for this path ; these are hand-traced pins (the reference's own source reproduces them through
tests/java_pin/j2py.py: tests/test_java_transliteration.py)."""
import json
import os

import numpy as np
import pytest

from oracle import kg_oracle_py as pyo

GOLD = os.path.join(os.path.dirname(__file__), "golden")
KATS = json.load(open(os.path.join(GOLD, "fsm_kats.json")))


def _params(kgo, p):
    return kgo.make_params(aa=True, order_constraint=p["order_constraint"], min_hits=p["min_hits"],
                           min_weighted_hits=p["min_weighted_hits"], max_gap=p["max_gap"])


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_fsm_kat_c_oracle(oracle, kat):
    hits = np.zeros(len(kat["hits"]), dtype=oracle.HIT_DTYPE)
    for i, (pos, fI, oI, wt, avg) in enumerate(kat["hits"]):
        hits[i] = (0, 0, pos, oI, avg, fI, wt)
    calls, otu = oracle.gather_hits(_params(oracle, kat["params"]), hits)
    got = [[int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]), float(c["weighted"])] for c in calls]
    assert got == [[a, b, c, d, float(np.float32(w))] for a, b, c, d, w in kat["calls"]]
    n = int(otu["n"][0])
    assert [[int(otu["count"][0][j]), int(otu["oI"][0][j])] for j in range(n)] == kat["otu"]


JAVA_VECTORS = json.load(open(os.path.join(GOLD, "java_fsm_vectors.json")))["vectors"]


@pytest.mark.parametrize("vec", JAVA_VECTORS, ids=[v["name"] for v in JAVA_VECTORS])
def test_fsm_vectors_printed_by_the_java_source(oracle, vec):
    """tests/golden/java_fsm_vectors.json: 64 adversarial hit lists whose CALL / OTU-COUNTS lines were printed by the
    reference's own source (tests/java_pin/make_fsm_vectors.py); the weighted score is compared as the text %f gives."""
    hits = np.zeros(len(vec["hits"]), dtype=oracle.HIT_DTYPE)
    for i, (pos, fI, oI, wt, avg) in enumerate(vec["hits"]):
        hits[i] = (0, 0, pos, oI, avg, fI, wt)
    calls, otu = oracle.gather_hits(_params(oracle, vec["params"]), hits)
    got = [[int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]), oracle.java_format_f(float(c["weighted"]))] for c in calls]
    assert got == vec["calls"]
    n = int(otu["n"][0])
    assert [[int(otu["count"][0][j]), int(otu["oI"][0][j])] for j in range(n)] == vec["otu"]


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_fsm_kat_py_oracle(kat):
    p = kat["params"]
    fsm = pyo.Fsm(pyo.Params(aa=True, order_constraint=bool(p["order_constraint"]), min_hits=p["min_hits"],
                             min_weighted_hits=p["min_weighted_hits"], max_gap=p["max_gap"]))
    fsm.gather(0, [pyo.Hit(oI, pos, avg, fI, np.float32(wt)) for pos, fI, oI, wt, avg in kat["hits"]])
    assert [[s, e, c, f, float(w)] for _, s, e, c, f, w in fsm.calls] == \
           [[a, b, c, d, float(np.float32(w))] for a, b, c, d, w in kat["calls"]]
    assert [list(x) for x in fsm.otu] == kat["otu"]


def test_encoding_kats(oracle):
    L = oracle.lib()
    code = lambda s: bytes(pyo.to_amino_acid_off(c) for c in s)
    assert L.kgo_encoded_kmer(code("AAAAAAAA"), 0) == 0
    assert L.kgo_encoded_kmer(code("AAAAAAAC"), 0) == 1
    assert L.kgo_encoded_kmer(code("CAAAAAAA"), 0) == 20 ** 7 == 1280000000
    assert L.kgo_encoded_kmer(code("YYYYYYYY"), 0) == 20 ** 8 - 1 == 25599999999
    for bad in "XU*xa- \n":
        assert L.kgo_encoded_kmer(code("AAAA" + bad + "AAA"), 0) == -1
        assert pyo.encoded_kmer(list(code("AAAA" + bad + "AAA")), 0) == -1
    for c in range(256):
        assert L.kgo_to_amino_acid_off(c) == pyo.to_amino_acid_off(chr(c))
        assert L.kgo_dna_char(c) == pyo.dna_char(chr(c))
        assert L.kgo_compl(c) == ord(pyo.rev_comp(chr(c)))
    assert [L.kgo_to_amino_acid_off(ord(c)) for c in "ACDEFGHIKLMNPQRSTVWY"] == list(range(20))
    assert L.kgo_compl(ord("s")) == ord("S")  # KGJ:218-219


def test_codon_kats(oracle):
    L = oracle.lib()
    idx = lambda cod: pyo.dna_char(cod[0]) * 16 + pyo.dna_char(cod[1]) * 4 + pyo.dna_char(cod[2])
    assert idx("ATG") == 14 and L.kgo_genetic_code(14) == b"M" and pyo.to_amino_acid_off("M") == 10
    for stop in ("TAA", "TAG", "TGA"):
        assert L.kgo_genetic_code(idx(stop)) == b"*"
    assert "".join(L.kgo_genetic_code(i).decode() for i in range(64)) == pyo.GENETIC_CODE


@pytest.mark.parametrize("seq", ["", "A", "AT", "ATG", "ATGN", "ATGAAATAGNNNCCCGGGTTTACGT", "acgtuACGTURYKMSWBDHVN" * 3])
def test_translate_matches_python(oracle, seq):
    L = oracle.lib()
    n = len(seq) // 3 + 1
    ps, pi = (np.zeros(n, np.uint8), np.zeros(n, np.uint8))
    pps, ppi = ["\0"] * n, [0] * n
    rc = np.zeros(len(seq), np.uint8)
    L.kgo_rev_comp(seq.encode(), len(seq), rc.ctypes.data)
    assert rc.tobytes().decode() == pyo.rev_comp(seq)
    for s in (seq, pyo.rev_comp(seq)):
        for off in range(3):
            L.kgo_translate(s.encode(), len(s), off, ps.ctypes.data, pi.ctypes.data, n)
            pyo.translate(s, off, pps, ppi)
            assert list(pi) == ppi and [chr(c) for c in ps] == pps


def test_java_format_f(oracle):
    f = oracle.java_format_f
    assert f(2.5) == "2.500000"
    assert f(float(np.float32(0.1))) == "0.100000"
    assert f(float(sum([np.float32(0.1)] * 10, np.float32(0)))) == "1.000000"
    assert f(1 / 128) == "0.007813"          # HALF_UP on the tie 0.0078125 (C printf gives 0.007812)
    assert f(3 / 128) == "0.023438"
    assert f(0.0) == "0.000000"
    assert f(123456.0) == "123456.000000"
    assert f(16777216.0) == "16777216.000000"
    assert f(0.9999999403953552) == "1.000000"   # largest float below 1
    assert f(1.5, 3) == "1.500" and f(0.0625, 3) == "0.063" and f(0.9996, 3) == "1.000"
    assert f(-2.5) == "-2.500000"
