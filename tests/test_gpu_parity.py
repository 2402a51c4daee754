"""GPU parity tests (run with -m gpu on a B200): the CUDA path through the C ABI against the CPU oracle, bit-exact for
hit positions, function/OTU ids, calls and (because the fp32 sum keeps the reference's order) weighted scores."""
import json
import os

import ctypes as C

import numpy as np
import pytest

from tests.parity import assert_same
from tools import kg_synth as synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FLAGS = [dict(), dict(order_constraint=1), dict(min_hits=3, max_gap=50, min_weighted_hits=2), dict(min_hits=2, max_gap=5)]


@pytest.fixture(scope="module")
def kg():
    import kmergutsjava_b200 as kg
    return kg


@pytest.fixture(scope="module")
def ctx(kg):
    c = kg.Context(0)
    yield c
    c.close()


@pytest.fixture(scope="module")
def universe():
    u = synth.Universe(n_families=300, seed=0x4B470003)
    keys, otu, avg, fi, wt = u.signatures()
    return u, synth.build_table_image(keys, otu, avg, fi, wt), len(keys)


def test_table_build(kg, ctx, universe):
    u, img, nsig = universe
    t = ctx.table_from_image(img)
    ti = t.info
    assert ti.num_signatures == nsig and ti.num_unreachable == 0 and ti.tail_run == 0
    assert ti.entry_size == 24 and ti.num_slots == int(np.frombuffer(img[:8], "<i8")[0])
    assert ti.flagged_buckets < ti.num_buckets * 0.2
    t.free()


@pytest.mark.parametrize("flags", FLAGS)
def test_aa_parity(kg, ctx, oracle, universe, flags):
    u, img, _ = universe
    seqs = u.proteins(500, seed=21) + [b"", b"A", b"ACDEFGHI", b"ACDEFGHIK", b"acdefghiklmnp", b"ACDEFGHIKXLMNPQRSTVWY", b""]
    sb, off = oracle.concat(seqs)
    t = ctx.table_from_image(img)
    res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1, **flags))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True, **flags), sb, off, oracle.STREAM_JOIN)
    assert len(ref.calls) > 20
    assert_same(res, ref, what=f"aa {flags}")
    res.free()
    t.free()


@pytest.mark.parametrize("flags", FLAGS[:3])
def test_dna_parity(kg, ctx, oracle, universe, flags):
    u, img, _ = universe
    seqs = [synth.genome(u, 30000, seed=31, index=i) for i in range(3)] + [
        b"", b"AC", b"ATG", b"ATGAAACCCGGGTTTACGTACGTAGCTAGCTAGCATCGATCGAT", synth.genome(u, 3001, seed=32),
        synth.genome(u, 3002, seed=33), synth.genome(u, 4097 * 3, seed=34), b"acgtnACGTNryk" * 40]
    sb, off = oracle.concat(seqs)
    t = ctx.table_from_image(img)
    res = ctx.run(t, kg.MODE_DNA, sb, off, kg.default_params(emit_hits=1, **flags))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=False, **flags), sb, off, oracle.STREAM_JOIN)
    assert len(ref.calls) > 10 and set(ref.hits["sf"]) == set(range(6))
    assert_same(res, ref, what=f"dna {flags}")
    res.free()
    t.free()


@pytest.mark.parametrize("probe,stages", [("cascade", "2"), ("fused", "2"), ("fused", "1"), ("fused", "0"), ("halves", "1")],
                         ids=["cascade", "fused", "fused-one-filter", "fused-no-filter", "two-pass-halves"])
def test_probe_variants(kg, ctx, oracle, universe, monkeypatch, probe, stages):
    """The probe stage exists as a three-kernel cascade (two prefilters that take turns in L2, the default) and as one
    fused kernel (one prefilter, also what the hash-sharded mode's k_answer shares its code with): both must give the
    oracle's hits, in aa mode and in 6-frame mode, with dense tiles (a protein of the table's own windows) included."""
    u, img, _ = universe
    if probe == "halves":   # k_probe_half: one pass per half of the key space, each with its own filter
        monkeypatch.setenv("KG_FILTER_HALVES", "1")
    else:
        monkeypatch.setenv("KG_PROBE", probe)
    if stages == "0":
        monkeypatch.setenv("KG_FILTER_BITS", "0")
    else:
        monkeypatch.setenv("KG_FILTER_STAGES", stages)
    t = ctx.table_from_image(img)
    alpha = np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8)
    fams = [alpha[u.consensus(f)].tobytes() for f in range(40)]  # unmutated consensus: tiles where a third of the windows hit
    seqs = u.proteins(700, seed=77) + fams + [b"", b"ACDEFGHIK", b"A" * 3000]
    sb, off = oracle.concat(seqs)
    res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
    assert len(ref.hits) > 5000
    assert_same(res, ref, what=f"aa {probe} {stages}")
    st = res.stats
    assert (st.ms_filter > 0) == (probe in ("cascade", "halves")), (st.ms_filter, st.ms_refilter, st.ms_lines)
    res.free()
    sb, off = oracle.concat([synth.genome(u, 40000, seed=41, index=i) for i in range(2)])
    res = ctx.run(t, kg.MODE_DNA, sb, off, kg.default_params(emit_hits=1))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=False), sb, off, oracle.DIRECT_PROBE)
    assert_same(res, ref, what=f"dna {probe} {stages}")
    res.free()
    t.free()


@pytest.mark.parametrize("flags", FLAGS[:3])
def test_packed_aa_input_and_compact_otus(kg, ctx, oracle, universe, monkeypatch, flags):
    """kg_run_packed_aa (5-bit residue codes, 8 per 5 bytes, unpacked on the device) must give exactly what kg_run gives on
    the original characters -- hits, calls and OTU counts, edge sequences included -- and the compact OTU form that both
    calls bring home expands to the oracle's records.  Small slices so that the call runs through several pipeline slices."""
    u, img, _ = universe
    monkeypatch.setenv("KG_SLICE_MB", "1")
    seqs = u.proteins(900, seed=55) + [b"", b"A", b"ACDEFGH", b"ACDEFGHI", b"ACDEFGHIK", b"acdefghiklmnp", b"ACDEFGHIKXLMNPQRSTVWY\x00\x00",
                                       b"", b"WWWWWWWWWWWWWWWW", b"ACDEFGHIKLMNPQRS", b"ACDEFGHIKLMNPQRST"] + u.proteins(300, seed=56)
    sb, off = oracle.concat(seqs)
    packed, goff = kg.pack_aa(sb, off, threads=3)
    t = ctx.table_from_image(img)
    p = kg.default_params(emit_hits=1, **flags)
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True, **flags), sb, off, oracle.DIRECT_PROBE)
    a = ctx.run(t, kg.MODE_AA, sb, off, p)
    b = ctx.run_packed_aa(t, packed, goff, p)
    assert_same(a, ref, what=f"kg_run {flags}")
    assert_same(b, ref, what=f"kg_run_packed_aa {flags}")
    for r in (a, b):
        n_per, ent = r.otus_compact
        assert n_per.dtype == np.uint8 and len(n_per) == len(seqs)
        assert np.array_equal(n_per, ref.otus["n"].astype(np.uint8))
        flat_c = np.concatenate([ref.otus["count"][i][:n] for i, n in enumerate(ref.otus["n"])] + [np.zeros(0, np.int32)])
        flat_o = np.concatenate([ref.otus["oI"][i][:n] for i, n in enumerate(ref.otus["n"])] + [np.zeros(0, np.int32)])
        assert np.array_equal(ent["count"], flat_c) and np.array_equal(ent["oI"], flat_o)
    assert a.stats.num_kmers == b.stats.num_kmers == ref.num_kmers
    # a device-resident result (kg_batch_run) serves the compact form too
    bt = ctx.upload(kg.MODE_AA, sb, off)
    c = ctx.run_batch(t, bt, p)
    n_per, ent = c.otus_compact
    assert np.array_equal(n_per, ref.otus["n"].astype(np.uint8)) and len(ent) == int(ref.otus["n"].sum())
    for x in (a, b, c):
        x.free()
    bt.free()
    t.free()


def test_packed_first_call_with_overflowing_slice(kg, oracle, universe, monkeypatch):
    """Regression (r02): on a FRESH context the hit buffers of slice s+1 are sized from slice s; a slice with more hits
    overflows them and the pass is repeated -- the repeat must not patch the padded residue stream a second time (it dropped
    the last allowed window of every protein of that slice).  Sparse proteins first, dense ones last, small slices."""
    u, img, _ = universe
    monkeypatch.setenv("KG_SLICE_MB", "1")
    alpha = np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8)
    rng = np.random.default_rng(3)
    sparse = [bytes(rng.choice(alpha, 300)) for _ in range(9000)]          # ~no hits: tiny hit buffers
    dense = [alpha[u.consensus(f % 300)].tobytes() for f in range(9000)]   # a third of the windows hit
    sb, off = oracle.concat(sparse + dense)
    packed, goff = kg.pack_aa(sb, off, threads=2)
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE, threads=4)
    for which in ("packed", "raw"):
        c = kg.Context(0)   # fresh: no buffer sizes remembered
        t = c.table_from_image(img)
        p = kg.default_params(emit_hits=1)
        res = c.run_packed_aa(t, packed, goff, p) if which == "packed" else c.run(t, kg.MODE_AA, sb, off, p)
        assert_same(res, ref, what=f"first call, {which}")
        res.free()
        t.free()
        c.close()


KATS = json.load(open(os.path.join(GOLD, "fsm_kats.json")))


JAVA_VECTORS = json.load(open(os.path.join(GOLD, "java_fsm_vectors.json")))["vectors"]


@pytest.mark.parametrize("fsm", ["seq", "seg"])
def test_fsm_vectors_printed_by_the_java_source_on_gpu(kg, ctx, fsm, monkeypatch):
    """No oracle in this one: the 64 vectors of tests/golden/java_fsm_vectors.json carry the CALL / OTU-COUNTS lines the
    reference's own Java source printed for them (tests/java_pin/make_fsm_vectors.py).  Each vector becomes a random protein
    whose windows at the vector's positions are the only table entries; both FSM paths must give Java's numbers (the weighted
    score as the text %f prints)."""
    monkeypatch.setenv("KG_FSM", fsm)
    rng = np.random.default_rng(4321)
    ncalls = 0
    for vec in JAVA_VECTORS:
        L = max(h[0] for h in vec["hits"]) + 40
        prot = bytes(rng.choice(np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8), L))
        wk = synth.window_keys(synth.aa_codes(prot))
        pos = [h[0] for h in vec["hits"]]
        if len(np.unique(wk[pos])) != len(pos):
            continue   # a repeated 8-mer among the chosen windows (never seen at these lengths): skip rather than alias two hits
        img = synth.build_table_image(wk[pos], [h[2] for h in vec["hits"]], [h[4] for h in vec["hits"]],
                                      [h[1] for h in vec["hits"]], np.array([h[3] for h in vec["hits"]], np.float32))
        t = ctx.table_from_image(img)
        res = ctx.run(t, kg.MODE_AA, np.frombuffer(prot, np.uint8), np.array([0, L], np.uint64), kg.default_params(emit_hits=1, **vec["params"]))
        hit_pos = set(int(x) for x in res.hits["pos"])
        assert set(pos) <= hit_pos, vec["name"]
        if hit_pos == set(pos):   # (a window elsewhere in the random protein may equal a table key: then the vector is not this run)
            got = [[int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]), kg.java_format_f(float(c["weighted"]), 6)] for c in res.calls]
            assert got == vec["calls"], vec["name"]
            o = res.otus[0]
            assert [[int(o["count"][j]), int(o["oI"][j])] for j in range(int(o["n"]))] == vec["otu"], vec["name"]
            ncalls += len(got)
        res.free()
        t.free()
    assert ncalls > 120


@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_fsm_kats_on_gpu(kg, ctx, kat):
    """The hand-traced FSM vectors, driven through the whole GPU path: a random protein whose windows at the KAT's
    positions are the only table entries, carrying the KAT's (fI, oI, wt, avg)."""
    rng = np.random.default_rng(1234)
    L = max(h[0] for h in kat["hits"]) + 40
    prot = bytes(rng.choice(np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8), L))
    wk = synth.window_keys(synth.aa_codes(prot))
    assert len(np.unique(wk)) == len(wk)
    pos = [h[0] for h in kat["hits"]]
    img = synth.build_table_image(wk[pos], [h[2] for h in kat["hits"]], [h[4] for h in kat["hits"]],
                                  [h[1] for h in kat["hits"]], np.array([h[3] for h in kat["hits"]], np.float32))
    t = ctx.table_from_image(img)
    p = kat["params"]
    sb = np.frombuffer(prot, np.uint8)
    off = np.array([0, L], np.uint64)
    res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1, **p))
    assert list(res.hits["pos"]) == sorted(pos)
    got = [[int(c["start"]), int(c["end"]), int(c["count"]), int(c["fI"]), float(c["weighted"])] for c in res.calls]
    assert got == [[a, b, c, d, float(np.float32(w))] for a, b, c, d, w in kat["calls"]]
    o = res.otus[0]
    assert [[int(o["count"][j]), int(o["oI"][j])] for j in range(int(o["n"]))] == kat["otu"]
    res.free()
    t.free()


def test_hit_cap_40000(kg, ctx, oracle):
    """Q9 (KGJ:496-504): an open run never holds more than 39998 hits; later hits of the run are dropped."""
    rng = np.random.default_rng(99)
    L = 41000
    prot = bytes(rng.choice(np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8), L))
    wk = synth.window_keys(synth.aa_codes(prot))
    keys, first = np.unique(wk, return_index=True)
    n = len(keys)
    img = synth.build_table_image(keys, np.full(n, 3), np.full(n, 1), np.full(n, 7), np.full(n, 0.5, np.float32))
    t = ctx.table_from_image(img)
    sb, off = oracle.concat([prot])
    res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
    assert int(ref.calls["count"][0]) == 39998
    assert_same(res, ref, what="cap")
    res.free()
    t.free()


def test_malformed_table_reachability(kg, ctx, oracle):
    """A key stored where the reference's no-wrap probe chain cannot reach it is never returned; a repeated key
    returns the first copy on the chain; unmatchable keys (>= 20^8 or < 0) only extend chains."""
    prot = b"MKVLAAGIVGLCAHHHWYYRRDDEEFFGGHHIIKKLLMMNNPPQQ"
    wk = synth.window_keys(synth.aa_codes(prot))
    n = len(wk)
    img = bytearray(synth.build_table_image(wk, np.arange(n), np.arange(n) + 100, np.full(n, 5), np.ones(n, np.float32), num_slots=211))
    ent = np.frombuffer(img, dtype=synth.ENTRY_DTYPE, offset=24)
    occ = np.flatnonzero(ent["which"] <= synth.MAX_ENCODED)
    empty = np.flatnonzero(ent["which"] > synth.MAX_ENCODED)
    # 1. move one key to an isolated empty slot far from its home: unreachable
    victim = occ[3]
    far = [e for e in empty if e > 0 and e + 1 < 211 and ent["which"][e - 1] > synth.MAX_ENCODED and ent["which"][e + 1] > synth.MAX_ENCODED
           and e != ent["which"][victim] % 211][0]
    ent[far] = ent[victim]
    ent["which"][victim] = synth.EMPTY_KEY
    # 2. duplicate another key right after itself with a different payload: the first copy must win
    dup = [s for s in occ if s != victim and s + 1 < 210 and ent["which"][s + 1] > synth.MAX_ENCODED and s + 1 != far][0]
    ent[dup + 1] = ent[dup]
    ent["fi"][dup + 1] = 99
    # 3. an unmatchable key
    um = [e for e in empty if e not in (far, dup + 1) and e + 1 < 211][-1]
    ent["which"][um] = synth.MAX_ENCODED  # == 20^8: occupied (not > MAX_ENCODED) but no 8-mer encodes to it
    t = ctx.table_from_image(bytes(img))
    sb, off = oracle.concat([prot])
    res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1, min_hits=2))
    for variant in (oracle.STREAM_JOIN, oracle.DIRECT_PROBE):
        ref = oracle.run(oracle.Table(data=bytes(img)), oracle.make_params(aa=True, min_hits=2), sb, off, variant)
        assert len(ref.hits) == n - 1 - 1 and 99 not in ref.hits["fI"]
        assert_same(res, ref, what="malformed")
    assert t.info.num_unreachable >= 2
    res.free()
    t.free()


def test_errors(kg, ctx, universe):
    u, img, _ = universe
    t = ctx.table_from_image(img)
    sb = np.frombuffer(b"ACDEFGHIKLMNP", np.uint8)
    off = np.array([0, 13], np.uint64)
    with pytest.raises(kg.KgError) as e:
        ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(min_hits=1))
    assert e.value.code == -1
    bad = bytearray(img)
    bad[8:16] = (32).to_bytes(8, "little")
    with pytest.raises(kg.KgError) as e:
        ctx.table_from_image(bytes(bad))
    assert e.value.code == -5
    with pytest.raises(kg.KgError):
        ctx.load_table("/nonexistent/dir")
    t.free()


def test_batch_rerun_is_idempotent(kg, ctx, oracle, universe):
    u, img, _ = universe
    seqs = u.proteins(200, seed=41)
    sb, off = oracle.concat(seqs)
    t = ctx.table_from_image(img)
    b = ctx.upload(kg.MODE_AA, sb, off)
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
    for _ in range(3):
        res = ctx.run_batch(t, b, kg.default_params(emit_hits=1))
        assert_same(res, ref, what="rerun")
        res.free()
    b.free()
    t.free()


@pytest.mark.parametrize("mode", ["aa", "dna"])
def test_sliced_pipeline_matches_oracle(kg, ctx, oracle, universe, mode, monkeypatch):
    """kg_run cuts a batch into slices that overlap H2D / compute / D2H; force many tiny slices (the smallest allowed,
    64 KiB) and check that records, sequence indices and ordering survive the stitching."""
    u, img, _ = universe
    monkeypatch.setenv("KG_SLICE_MB", "0")     # clamps to the 64 KiB minimum
    t = ctx.table_from_image(img)
    if mode == "aa":
        seqs = u.proteins(1500, seed=51) + [b""] + u.proteins(10, seed=52)
        m = kg.MODE_AA
    else:
        seqs = [synth.genome(u, 50000 + 7 * i, seed=53, index=i) for i in range(8)] + [b"", b"ACG"]
        m = kg.MODE_DNA
    sb, off = oracle.concat(seqs)
    assert int(off[-1]) > 5 * 65536
    res = ctx.run(t, m, sb, off, kg.default_params(emit_hits=1))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=(mode == "aa")), sb, off, oracle.DIRECT_PROBE)
    assert_same(res, ref, what=f"sliced {mode}")
    assert res.stats.num_launches > 30
    res.free()
    t.free()


@pytest.mark.parametrize("fsm", ["seg", "seq"])
@pytest.mark.parametrize("kat", KATS, ids=[k["name"] for k in KATS])
def test_fsm_kats_both_paths(kg, ctx, kat, fsm, monkeypatch):
    """Both run-FSM implementations (thread per sequence / thread per gap-delimited segment + OTU replay) against the
    hand-traced vectors."""
    monkeypatch.setenv("KG_FSM", fsm)
    test_fsm_kats_on_gpu(kg, ctx, kat)


@pytest.mark.parametrize("fsm", ["seg", "seq"])
@pytest.mark.parametrize("flags", FLAGS)
def test_parity_both_fsm_paths(kg, ctx, oracle, universe, flags, fsm, monkeypatch):
    monkeypatch.setenv("KG_FSM", fsm)
    u, img, _ = universe
    t = ctx.table_from_image(img)
    aa = u.proteins(300, seed=61) + [b"", b"ACDEFGHIK"]
    sb, off = oracle.concat(aa)
    res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1, **flags))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True, **flags), sb, off, oracle.DIRECT_PROBE)
    assert_same(res, ref, what=f"aa {fsm} {flags}")
    res.free()
    dna = [synth.genome(u, 40000, seed=62, index=i) for i in range(3)] + [b"", b"ACGT"]
    sb, off = oracle.concat(dna)
    res = ctx.run(t, kg.MODE_DNA, sb, off, kg.default_params(emit_hits=1, **flags))
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=False, **flags), sb, off, oracle.DIRECT_PROBE)
    assert_same(res, ref, what=f"dna {fsm} {flags}")
    res.free()
    t.free()


def test_otu_replay_many_otus(kg, ctx, oracle, monkeypatch):
    """OTU top-5 buffer with more than five OTU indices in alternating and repeated order (overwrite + bubbling), through
    the run-length replay of the segment path and through the per-sequence FSM."""
    rng = np.random.default_rng(5)
    L = 3000
    prot = bytes(rng.choice(np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8), L))
    wk = synth.window_keys(synth.aa_codes(prot))
    keys, first = np.unique(wk, return_index=True)
    pos = np.sort(first)
    n = len(pos)
    otu = rng.integers(0, 9, size=n)
    otu[::7] = otu[1::7][: len(otu[::7])]          # some immediate repeats
    fI = np.where((pos // 400) % 2 == 0, 7, 9)      # function switches every 400 residues
    img = synth.build_table_image(wk[pos], otu, np.full(n, 3), fI, np.full(n, 0.25, np.float32))
    t = ctx.table_from_image(img)
    sb, off = oracle.concat([prot, prot[100:2500], prot[::-1]])
    ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
    assert int(ref.otus["n"][0]) == 5 and len(ref.calls) >= 8
    for fsm in ("seg", "seq"):
        monkeypatch.setenv("KG_FSM", fsm)
        res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))
        assert_same(res, ref, what=f"otu {fsm}")
        res.free()
    t.free()


@pytest.mark.parametrize("seed", range(6))
def test_fsm_fuzz(kg, ctx, oracle, seed, monkeypatch):
    """Adversarial hit streams for the run FSM: sticky-but-switching function indices, OTU churn over more than five
    indices, clustered positions with gaps around max_gap, odd weights and offsets -- several proteins per case, every flag
    set, both FSM implementations, against the oracle."""
    rng = np.random.default_rng(1000 + seed)
    alpha = np.frombuffer(synth.PROT_ALPHA.encode(), np.uint8)
    prots, keys, otu, avg, fI, wt = [], [], [], [], [], []
    used = set()
    for _ in range(12):
        L = int(rng.integers(200, 6000))
        prot = bytes(rng.choice(alpha, L))
        wk = synth.window_keys(synth.aa_codes(prot))
        # clustered positions: bursts of hits separated by gaps of 150..260 (around the default max_gap of 200)
        pos, p = [], int(rng.integers(0, 50))
        while p < len(wk) - 1:
            for _ in range(int(rng.integers(1, 40))):
                if p >= len(wk) - 1:
                    break
                pos.append(p)
                p += int(rng.choice([1, 1, 1, 2, 3, 7, 30]))
            p += int(rng.integers(150, 260))
        f = int(rng.integers(1, 5))
        for q in pos:
            k = int(wk[q])
            if k in used:
                continue
            used.add(k)
            if rng.random() < 0.25:
                f = int(rng.integers(1, 5))
            keys.append(k)
            fI.append(f if rng.random() > 0.1 else int(rng.integers(1, 5)))
            otu.append(int(rng.integers(0, 9)) if rng.random() < 0.5 else 3)
            avg.append(int(L - q + rng.integers(-30, 30)))
            wt.append(float(rng.integers(1, 700)) / 256.0)
        prots.append(prot)
    img = synth.build_table_image(np.array(keys), otu, avg, fI, np.array(wt, np.float32))
    t = ctx.table_from_image(img)
    sb, off = oracle.concat(prots)
    for flags in FLAGS + [dict(min_hits=2, max_gap=0), dict(min_hits=7, max_gap=1000, min_weighted_hits=9)]:
        ref = oracle.run(oracle.Table(data=img), oracle.make_params(aa=True, **flags), sb, off, oracle.DIRECT_PROBE)
        for fsm in ("seq", "seg"):
            monkeypatch.setenv("KG_FSM", fsm)
            res = ctx.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1, **flags))
            assert_same(res, ref, what=f"fuzz seed {seed} {fsm} {flags}")
            res.free()
    t.free()


def test_table_cache_roundtrip(kg, ctx, oracle, universe, tmp_path):
    """kg_table_save / kg_table_load_cached: the cached GPU layout answers exactly like the table it was saved from;
    foreign, truncated or corrupted files are refused (KG_EFORMAT / KG_EIO), never half-loaded."""
    u, img, nsig = universe
    t = ctx.table_from_image(img)
    path = str(tmp_path / "table.kgcache")
    t.save(path)
    c = ctx.load_table_cached(path)
    for f in ("num_slots", "entry_size", "version", "num_signatures", "num_unreachable", "tail_run", "num_buckets", "flagged_buckets", "device_bytes"):
        assert getattr(c.info, f) == getattr(t.info, f), f
    sb, off = oracle.concat(u.proteins(400, seed=3))
    p = kg.default_params(emit_hits=1)
    a, b = ctx.run(t, kg.MODE_AA, sb, off, p), ctx.run(c, kg.MODE_AA, sb, off, p)
    assert len(a.hits) > 1000
    for name in ("hits", "calls", "otus"):
        assert getattr(a, name).tobytes() == getattr(b, name).tobytes(), name
    a.free()
    b.free()
    # a shard keeps its identity through the cache
    s = ctx.table_from_image_sharded(img, 1, 3)
    s.save(path + ".shard")
    s2 = ctx.load_table_cached(path + ".shard")
    with pytest.raises(kg.KgError, match="shard 1 of 3"):
        ctx.run(s2, kg.MODE_AA, sb, off, p)
    raw = open(path, "rb").read()
    for bad in (raw[:len(raw) - 128], b"XXXXXXXX" + raw[8:], raw[:40] + bytes([raw[40] ^ 1]) + raw[41:], raw + b"\0" * 8, b""):
        with open(path + ".bad", "wb") as f:
            f.write(bad)
        with pytest.raises(kg.KgError) as e:
            ctx.load_table_cached(path + ".bad")
        assert e.value.code in (-5, -4)
    # bit rot inside the body (not the header): the body checksum refuses it (a damaged cache must not yield calls)
    mid = len(raw) // 2
    with open(path + ".rot", "wb") as f:
        f.write(raw[:mid] + bytes([raw[mid] ^ 0x10]) + raw[mid + 1:])
    with pytest.raises(kg.KgError, match="checksum") as e:
        ctx.load_table_cached(path + ".rot")
    assert e.value.code == -5
    with pytest.raises(kg.KgError):
        ctx.load_table_cached(path + ".missing")
    for x in (t, c, s, s2):
        x.free()


def test_two_contexts_share_a_table(kg, ctx, oracle, universe):
    """One context per host thread (include/kmerguts.h, threading): a second context attaches to the table and both run
    different batches at the same time; each result equals the oracle's."""
    import threading
    u, img, _ = universe
    t = ctx.table_from_image(img)
    ctx2 = kg.Context(0)
    t.attach(ctx2)
    otable = oracle.Table(data=img)
    jobs = []
    for k, c in enumerate((ctx, ctx2)):
        sb, off = oracle.concat(u.proteins(3000, seed=50 + k))
        jobs.append((c, sb, off))
    out = [[None] * 4, [None] * 4]

    def lane(k):
        c, sb, off = jobs[k]
        for rep in range(4):
            out[k][rep] = c.run(t, kg.MODE_AA, sb, off, kg.default_params(emit_hits=1))

    th = [threading.Thread(target=lane, args=(k,)) for k in range(2)]
    for x in th:
        x.start()
    for x in th:
        x.join()
    for k in range(2):
        _, sb, off = jobs[k]
        ref = oracle.run(otable, oracle.make_params(aa=True), sb, off, oracle.DIRECT_PROBE)
        for rep in range(4):
            assert_same(out[k][rep], ref, what=f"context {k}, run {rep}")
            out[k][rep].free()
    ctx2.close()
    t.free()


def test_run_batch_many_equals_single_runs(kg, ctx, oracle):
    """kg_batch_run_many (two batches in flight): every result equals the one kg_batch_run gives for that batch."""
    u = synth.Universe(n_families=300, seed=0x4B470011)
    keys, otu, avg, fi, wt = u.signatures()
    table = ctx.table_from_image(synth.build_table_image(keys, otu, avg, fi, wt))
    params = kg.default_params(emit_hits=1)
    batches, singles = [], []
    for k, (mode, n) in enumerate([(kg.MODE_AA, 700), (kg.MODE_AA, 5), (kg.MODE_DNA, 3), (kg.MODE_AA, 1200), (kg.MODE_AA, 0)]):
        seqs = u.proteins(n, seed=50 + k) if mode == kg.MODE_AA else [synth.genome(u, 30000, seed=60 + k, index=i) for i in range(n)]
        sb, off = oracle.concat(seqs)
        batches.append(ctx.upload(mode, sb, off))
    order = [0, 1, 2, 3, 4, 0, 3]          # a batch may appear more than once
    for i in order:
        singles.append(ctx.run_batch(table, batches[i], params))
    many = ctx.run_batch_many(table, [batches[i] for i in order], params)
    assert len(many) == len(order)
    piped = list(ctx.run_batches(table, [batches[i] for i in order], params))   # submit / collect, two in flight
    for a, b in zip(piped, singles):
        for name in ("hits", "calls", "otus"):
            assert getattr(a, name).tobytes() == getattr(b, name).tobytes(), name
    for r in piped:
        r.free()
    with pytest.raises(kg.KgError):
        kg._check(kg.lib().kg_batch_collect(ctx._h, C.byref(C.c_void_p())))   # nothing in flight
    for a, b in zip(many, singles):
        for name in ("hits", "calls", "otus"):
            assert getattr(a, name).tobytes() == getattr(b, name).tobytes(), name
        assert a.stats.num_kmers == b.stats.num_kmers
    assert sum(len(r.calls) for r in many) > 50
    for r in many + singles:
        r.free()
    for b in batches:
        b.free()
    table.free()


def test_packed_dna_equals_raw(kg, ctx, oracle):
    """kg_run_packed_dna (2-bit nucleotides + exception list over the host link) gives the records of kg_run on the characters:
    genomes with N / IUPAC / lowercase / U, empty and tiny contigs, several slices."""
    u = synth.Universe(n_families=300, seed=0x4B470012)
    keys, otu, avg, fi, wt = u.signatures()
    table = ctx.table_from_image(synth.build_table_image(keys, otu, avg, fi, wt))
    rng = np.random.default_rng(12)
    contigs = []
    for i in range(7):
        g = bytearray(synth.genome(u, 40000 + 1000 * i, seed=80 + i, index=i))
        for p in rng.integers(0, len(g), 40):
            g[int(p)] = int(rng.choice(np.frombuffer(b"NRYnacgtU", dtype=np.uint8)))
        contigs.append(bytes(g))
    contigs[3:3] = [b"", b"AC", b"ACGTACGTA", b"NNNNNNNNNNNNNNNNNNNNNNNNNNNNNNNN"]
    sb, off = oracle.concat(contigs)
    packed, boff, exc = kg.pack_dna(sb, off, threads=4)
    assert len(exc) > 100 and packed.nbytes < sb.nbytes / 3.9
    for flags in (dict(), dict(min_hits=3, max_gap=50)):
        params = kg.default_params(emit_hits=1, **flags)
        raw = ctx.run(table, kg.MODE_DNA, sb, off, params)
        os.environ["KG_SLICE_MB"] = "1"          # cut the call into several slices: the exception list is cut with them
        try:
            pk = ctx.run_packed_dna(table, packed, off, boff, exc, params)
        finally:
            del os.environ["KG_SLICE_MB"]
        for name in ("hits", "calls", "otus"):
            assert getattr(pk, name).tobytes() == getattr(raw, name).tobytes(), (name, flags)
        assert len(raw.calls) > 20
        ref = oracle.run(oracle.Table(data=synth.build_table_image(keys, otu, avg, fi, wt)), oracle.make_params(aa=False, **flags), sb, off, oracle.DIRECT_PROBE)
        assert_same(pk, ref, what=f"packed dna {flags}")
        raw.free()
        pk.free()
    table.free()
