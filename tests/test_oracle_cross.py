"""Double-entry check of the CPU oracle: C stream-join == C direct-probe == independent pure-Python restatement, on
seeded synthetic universes in both modes and under the flag combinations SURVEY.md section 8(d) lists."""
import numpy as np
import pytest

from oracle import kg_oracle_py as pyo
from tools import kg_synth as synth

FLAGS = [dict(), dict(order_constraint=True), dict(min_hits=3, max_gap=50, min_weighted_hits=2), dict(min_hits=2, max_gap=5)]


@pytest.fixture(scope="module")
def small_universe():
    u = synth.Universe(n_families=60, seed=0x4B470001)
    keys, otu, avg, fi, wt = u.signatures()
    img = synth.build_table_image(keys, otu, avg, fi, wt)
    return u, img


def _as_lists(res):
    hits = [(int(h["seq"]), int(h["sf"]), int(h["pos"]), int(h["oI"]), int(h["avg"]), int(h["fI"]), float(h["wt"]))
            for h in res.hits]
    calls = [(int(c["seq"]), int(c["sf"]), int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]),
              float(c["weighted"])) for c in res.calls]
    otus = [[(int(o["count"][j]), int(o["oI"][j])) for j in range(int(o["n"]))] for o in res.otus]
    return hits, calls, otus


@pytest.mark.parametrize("flags", FLAGS)
def test_aa_mode_three_way(oracle, small_universe, flags):
    u, img = small_universe
    seqs = u.proteins(40, seed=5) + [b"", b"A", b"ACDEFGHI", b"ACDEFGHIK", b"acdefghiklmnp", b"ACDEFGHIKXLMNPQRSTVWY"]
    t = oracle.Table(data=img)
    p = oracle.make_params(aa=True, **flags)
    sb, off = oracle.concat(seqs)
    a = oracle.run(t, p, sb, off, oracle.STREAM_JOIN)
    b = oracle.run(t, p, sb, off, oracle.DIRECT_PROBE)
    assert a.lookup_error == 0 and b.lookup_error == 0
    la, lb = _as_lists(a), _as_lists(b)
    assert la == lb
    assert len(la[0]) > 100 and len(la[1]) > 5          # the case is not vacuous
    pt = pyo.Table.from_bytes(img)
    pp = pyo.Params(aa=True, order_constraint=flags.get("order_constraint", False), min_hits=flags.get("min_hits", 5),
                    min_weighted_hits=flags.get("min_weighted_hits", 0), max_gap=flags.get("max_gap", 200))
    h, c, o = pyo.run(pt, pp, [s.decode() for s in seqs])
    assert [(x[0], x[1], x[2], x[3], x[4], x[5], float(x[6])) for x in h] == la[0]
    assert [(x[0], x[1], x[2], x[3], x[4], x[5], float(x[6])) for x in c] == la[1]
    assert o == la[2]


@pytest.mark.parametrize("flags", FLAGS[:3])
def test_dna_mode_three_way(oracle, small_universe, flags):
    u, img = small_universe
    seqs = [synth.genome(u, 6000, seed=7, index=i) for i in range(2)] + [b"", b"AC", b"ATG", b"ATGAAACCCGGGTTTACGTACGTAGCTAGCTAGCATCGATCGAT",
                                                                        synth.genome(u, 3001, seed=8), synth.genome(u, 3002, seed=9)]
    t = oracle.Table(data=img)
    p = oracle.make_params(aa=False, **flags)
    sb, off = oracle.concat(seqs)
    a = oracle.run(t, p, sb, off, oracle.STREAM_JOIN)
    b = oracle.run(t, p, sb, off, oracle.DIRECT_PROBE)
    la, lb = _as_lists(a), _as_lists(b)
    assert la == lb
    assert len(la[0]) > 100 and len(la[1]) > 3
    assert {h[1] for h in la[0]} == set(range(6))       # hits on all six frames
    pt = pyo.Table.from_bytes(img)
    pp = pyo.Params(aa=False, order_constraint=flags.get("order_constraint", False), min_hits=flags.get("min_hits", 5),
                    min_weighted_hits=flags.get("min_weighted_hits", 0), max_gap=flags.get("max_gap", 200))
    h, c, o = pyo.run(pt, pp, [s.decode() for s in seqs])
    assert [(x[0], x[1], x[2], x[3], x[4], x[5], float(x[6])) for x in h] == la[0]
    assert [(x[0], x[1], x[2], x[3], x[4], x[5], float(x[6])) for x in c] == la[1]
    assert o == la[2]


def test_aa_mode_drops_last_window(oracle):
    """Q1 (KGJ:912, 1055): in aa mode the window starting at len-8 is never looked up; in DNA mode it is."""
    prot = b"MKVLAAGIVGLCAHHHW"
    codes = synth.aa_codes(prot)
    keys = synth.window_keys(codes)
    n = len(keys)
    img = synth.build_table_image(keys, np.arange(n), np.full(n, 9), np.full(n, 3), np.ones(n, np.float32))
    t = oracle.Table(data=img)
    sb, off = oracle.concat([prot])
    r = oracle.run(t, oracle.make_params(aa=True, min_hits=2), sb, off)
    assert list(r.hits["pos"]) == list(range(n - 1))
    assert r.num_kmers == n - 1
    dna = "".join(synth._CODONS[chr(c)][0] for c in prot).encode()
    sb, off = oracle.concat([dna])
    r = oracle.run(t, oracle.make_params(aa=False, min_hits=2), sb, off)
    f0 = r.hits[r.hits["sf"] == 0]
    assert list(f0["pos"]) == list(range(n))


def test_no_wraparound_and_eof(oracle):
    """Q5 (KGJ:995, 1102-1103, 799-802): a chain that reaches the end of the file aborts the lookup; earlier hits stay."""
    prot = b"MKVLAAGIVGLCAHHHWYYRR"
    keys = synth.window_keys(synth.aa_codes(prot))
    n = len(keys)
    good = synth.build_table_image(keys, np.zeros(n), np.zeros(n), np.ones(n), np.ones(n, np.float32), num_slots=101)
    ent = np.frombuffer(good, dtype=synth.ENTRY_DTYPE, offset=24).copy()
    ent["which"][100] = 12345                 # occupy the last slot with a foreign key: chains through it run off the end
    probe = (20 ** 8 - 1) - ((20 ** 8 - 1 - 100) % 101)   # a valid key whose home is slot 100
    assert probe % 101 == 100 and probe < 20 ** 8
    img = good[:24] + ent.tobytes()
    t = oracle.Table(data=img)
    codes = []
    v = probe
    for _ in range(8):
        codes.append(v % 20)
        v //= 20
    seq2 = bytes(synth.PROT_ALPHA[c].encode()[0] for c in reversed(codes)) + b"A"
    sb, off = oracle.concat([prot, seq2])
    for variant in (oracle.STREAM_JOIN, oracle.DIRECT_PROBE):
        r = oracle.run(t, oracle.make_params(aa=True, min_hits=2), sb, off, variant)
        assert r.lookup_error == 1
        assert len(r.hits) == n - 1 and set(r.hits["seq"]) == {0}


def test_duplicate_positions_of_one_kmer(oracle):
    """The same 8-mer at several positions / sequences: every position gets the hit (KGJ:1004-1015)."""
    rep = b"ACDEFGHIK" * 6
    keys = synth.window_keys(synth.aa_codes(rep))[:9]
    img = synth.build_table_image(keys, np.arange(9), np.arange(9) + 50, np.full(9, 4), np.ones(9, np.float32))
    t = oracle.Table(data=img)
    sb, off = oracle.concat([rep, rep[3:]])
    a = oracle.run(t, oracle.make_params(aa=True), sb, off, oracle.STREAM_JOIN)
    b = oracle.run(t, oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
    assert _as_lists(a) == _as_lists(b)
    assert len(a.hits) == (len(rep) - 8) + (len(rep) - 3 - 8)
    assert len(a.calls) == 2 and int(a.calls["count"][0]) == len(rep) - 8


@pytest.mark.parametrize("seed", range(8))
def test_fsm_fuzz_c_vs_python(oracle, seed):
    """The two independently written FSMs (C, pure Python) on random adversarial hit lists, all flag combinations."""
    rng = np.random.default_rng(500 + seed)
    for case in range(25):
        n = int(rng.integers(0, 120))
        pos = np.sort(rng.choice(3000, size=n, replace=False)) if n else np.zeros(0, int)
        if case % 3 == 0 and n:      # clustered
            pos = np.unique(np.cumsum(rng.choice([1, 1, 2, 5, 190, 200, 201, 230], size=n)))
            n = len(pos)
        fI = rng.integers(1, 4, size=n)
        for i in range(1, n):
            if rng.random() < 0.7:
                fI[i] = fI[i - 1]
        oI = rng.integers(0, 8, size=n)
        avg = (3000 - pos + rng.integers(-25, 25, size=n)) if n else np.zeros(0, int)
        wt = (rng.integers(1, 600, size=n) / 256.0).astype(np.float32)
        flags = dict(order_constraint=bool(rng.integers(2)), min_hits=int(rng.integers(2, 7)), min_weighted_hits=int(rng.integers(0, 4)),
                     max_gap=int(rng.choice([0, 5, 50, 200, 1000])))
        hits = np.zeros(n, dtype=oracle.HIT_DTYPE)
        hits["pos"], hits["fI"], hits["oI"], hits["avg"], hits["wt"] = pos, fI, oI, avg, wt
        calls, otu = oracle.gather_hits(oracle.make_params(aa=True, **flags), hits)
        fsm = pyo.Fsm(pyo.Params(aa=True, **flags))
        fsm.gather(0, [pyo.Hit(int(o), int(p), int(a), int(f), np.float32(w)) for p, f, o, a, w in zip(pos, fI, oI, avg, wt)])
        got = [(int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]), float(c["weighted"])) for c in calls]
        assert got == [(s, e, c, f, float(w)) for _, s, e, c, f, w in fsm.calls], (seed, case, flags)
        k = int(otu["n"][0])
        assert [[int(otu["count"][0][j]), int(otu["oI"][0][j])] for j in range(k)] == [list(x) for x in fsm.otu], (seed, case, flags)
