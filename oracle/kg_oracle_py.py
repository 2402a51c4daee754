"""oracle/kg_oracle_py.py -- second, independently written restatement of KmerGutsJava's hot path.

TEST INFRASTRUCTURE ONLY.  The reference ships no golden vectors for this path and cannot be run in this image (no JVM;
its source IS executed through tests/java_pin/j2py.py, see oracle/kg_oracle.h).  This file exists so that the C oracle (oracle/kg_oracle.c) is checked by something that
was written separately from it, in a different style (pure Python objects, dictionary probe chains), directly from
lib/src/kmergutsjava/KmerGutsJava.java ("KGJ").  Pure-Python loops: small cases only.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

K = 8                                  # KGJ:85
MAX_ENCODED = 20 ** 8                  # KGJ:87
MAX_HITS_PER_SEQ = 40000               # KGJ:98
OI_BUFSZ = 5                           # KGJ:99
PROT_ALPHA = "ACDEFGHIKLMNPQRSTVWY"    # KGJ:94-96
GENETIC_CODE = ("KNKNTTTTRSRSIIMI" "QHQHPPPPRRRRLLLL" "EDEDAAAAGGGGVVVV" "*Y*YSSSS*CWCLFLF")  # KGJ:88-93

_COMPL = dict(zip("aAcCgGtuTUmMrRwWsSyYkKbBdDhHvVnN",   # KGJ:177-260 ('s' -> 'S' is the reference's own typo)
                  "tTgGcCaaAAkKyYwWSSrRmMvVhHdDbBnN"))


def to_amino_acid_off(c: str) -> int:      # KGJ:111-175
    i = PROT_ALPHA.find(c) if len(c) == 1 else -1
    return i if i >= 0 else 20


def dna_char(c: str) -> int:               # KGJ:294-318
    return {"a": 0, "A": 0, "c": 1, "C": 1, "g": 2, "G": 2, "t": 3, "u": 3, "T": 3, "U": 3}.get(c, 4)


def rev_comp(s: str) -> str:               # KGJ:263-272
    return "".join(_COMPL.get(c, c) for c in reversed(s))


def encoded_kmer(codes: Sequence[int], pos: int) -> int:   # KGJ:274-292
    v = 0
    for i in range(K):
        a = codes[pos + i]
        if a >= 20:
            return -1
        v = v * 20 + a
    return v


def translate(seq: str, off: int, pseq: List[str], piseq: List[int]) -> None:   # KGJ:320-343 (buffers reused)
    p = 0
    i = off
    while i <= len(seq) - 3:
        c1, c2, c3 = dna_char(seq[i]), dna_char(seq[i + 1]), dna_char(seq[i + 2])
        i += 3
        if c1 < 4 and c2 < 4 and c3 < 4:
            aa = GENETIC_CODE[c1 * 16 + c2 * 4 + c3]
            pseq[p] = aa
            piseq[p] = to_amino_acid_off(aa)
        else:
            pseq[p] = "x"
            piseq[p] = 20
        p += 1
    if p < len(pseq):
        pseq[p] = "\0"
        piseq[p] = 21


@dataclass
class Params:                               # KGJ:102-107
    aa: bool = False
    order_constraint: bool = False
    min_hits: int = 5
    min_weighted_hits: int = 0
    max_gap: int = 200
    debug: bool = False


@dataclass
class Hit:                                  # KGJ:1213-1219
    oI: int
    pos: int
    avg: int
    fI: int
    wt: np.float32


@dataclass
class Table:
    """kmer.table.mem_map image: header (KGJ:933-935) + 24-byte LE entries (KGJ:995-999)."""
    num_sigs: int
    entry_size: int
    version: int
    entries: np.ndarray     # structured: which, otu, avg, fi, wt

    ENTRY = np.dtype([("which", "<i8"), ("otu", "<i4"), ("avg", "<i4"), ("fi", "<i4"), ("wt", "<f4")])

    @classmethod
    def from_bytes(cls, b: bytes) -> "Table":
        num_sigs, entry_size, version = struct.unpack_from("<qqq", b, 0)
        n = (len(b) - 24) // 24
        return cls(num_sigs, entry_size, version, np.frombuffer(b, dtype=cls.ENTRY, count=n, offset=24))

    def probe(self, v: int) -> Optional[int]:
        """Slot of v following the chain h, h+1, ... with NO wrap (KGJ:959-1026); None = miss (or ran off the end)."""
        s = v % self.num_sigs
        while s < len(self.entries):
            w = int(self.entries["which"][s])
            if w > MAX_ENCODED:
                return None
            if w == v:
                return s
            s += 1
        return None


def enumerate_kmers(p: Params, seq: str) -> List[Tuple[int, List[Tuple[int, int]]]]:
    """prepareQuery + addKmers (KGJ:1051-1074, 900-922): [(strand_frame, [(pos, value), ...])]."""
    out = []
    if p.aa:
        codes = [to_amino_acid_off(c) for c in seq]
        out.append((0, [(i, encoded_kmer(codes, i)) for i in range(0, len(codes) - K)]))   # NB: '<', KGJ:912
    else:
        n = len(seq) // 3 + 1
        pseq, piseq = ["\0"] * n, [0] * n
        for strand, s in ((0, seq), (1, None)):
            if s is None:
                s = rev_comp(seq)
            for frame in range(3):
                translate(s, frame, pseq, piseq)
                out.append((3 * strand + frame, [(i, encoded_kmer(piseq, i)) for i in range(0, n - K)]))
    return [(sf, [(i, v) for i, v in lst if v >= 0]) for sf, lst in out]


@dataclass
class Fsm:
    """gatherHits (KGJ:457-514) + processSetOfHits (KGJ:385-455) for one sequence (OTU buffer shared by frames)."""
    p: Params
    otu: List[List[int]] = field(default_factory=list)      # [[count, oI], ...]
    calls: List[Tuple[int, int, int, int, int, np.float32]] = field(default_factory=list)  # (sf,start,end,cnt,fI,w)

    def process(self, sf: int, hits: List[Hit], cur: int) -> int:
        cnt, w, last = 0, np.float32(0), 0
        for i, h in enumerate(hits):
            if h.fI == cur:
                last, cnt, w = i, cnt + 1, np.float32(w + h.wt)
        if cnt >= self.p.min_hits and w >= np.float32(self.p.min_weighted_hits):
            self.calls.append((sf, hits[0].pos, hits[last].pos + K - 1, cnt, cur, w))
            for h in hits[: last + 1]:
                if h.fI != cur:
                    continue
                j = next((j for j, e in enumerate(self.otu) if e[1] == h.oI), len(self.otu))
                if j == len(self.otu):
                    if len(self.otu) == OI_BUFSZ:
                        j -= 1
                        self.otu[j] = [1, h.oI]
                    else:
                        self.otu.append([1, h.oI])
                else:
                    self.otu[j][0] += 1
                while j > 0 and self.otu[j - 1][0] <= self.otu[j][0]:
                    self.otu[j - 1], self.otu[j] = self.otu[j], self.otu[j - 1]
                    j -= 1
        if len(hits) >= 2 and hits[-2].fI != cur and hits[-2].fI == hits[-1].fI:
            cur = hits[-1].fI
            hits[:] = hits[-2:]
        else:
            hits.clear()
        return cur

    def gather(self, sf: int, all_hits: List[Hit]) -> None:
        p = self.p
        all_hits = sorted(all_hits, key=lambda h: h.pos)
        hits: List[Hit] = []
        cur = 0
        for ph in all_hits:
            if hits and hits[-1].pos + p.max_gap < ph.pos:
                if len(hits) >= p.min_hits:
                    cur = self.process(sf, hits, cur)
                else:
                    hits.clear()
            if not hits:
                cur = ph.fI
            if (not p.order_constraint) or not hits or (
                    ph.fI == hits[-1].fI and abs((ph.pos - hits[-1].pos) - (hits[-1].avg - ph.avg)) <= 20):
                if len(hits) < MAX_HITS_PER_SEQ - 2:
                    hits.append(ph)
                if len(hits) > 1 and cur != ph.fI and hits[-2].fI == hits[-1].fI:
                    cur = self.process(sf, hits, cur)
        if len(hits) >= p.min_hits:
            self.process(sf, hits, cur)


def run(table: Table, p: Params, seqs: Sequence[str]):
    """Returns (hits, calls, otus): hits = [(seq, sf, pos, oI, avg, fI, wt)], calls = [(seq, sf, start, end, cnt,
    fI, w)], otus = [[(count, oI), ...] per sequence] -- all in the order the reference prints them."""
    hits_out, calls_out, otus_out = [], [], []
    for si, seq in enumerate(seqs):
        fsm = Fsm(p)
        for sf, kmers in enumerate_kmers(p, seq):
            hs = []
            for pos, v in kmers:
                s = table.probe(v)
                if s is not None:
                    e = table.entries[s]
                    hs.append(Hit(int(e["otu"]), pos, int(e["avg"]), int(e["fi"]), np.float32(e["wt"])))
            hits_out += [(si, sf, h.pos, h.oI, h.avg, h.fI, h.wt) for h in hs]
            fsm.gather(sf, hs)
        calls_out += [(si,) + c for c in fsm.calls]
        otus_out.append([(c, o) for c, o in fsm.otu])
    return hits_out, calls_out, otus_out
