"""tools/kg_synth.py -- synthetic inputs in the REFERENCE's on-disk formats, plus the counter-based synthetic universe.

Neutral tooling (no oracle, no product code): writers for kmer.table.mem_map[.gz] (KGJ:924-942, 995-999),
function.index (KGJ:345-373) and FASTA (KGJ:1132-1192), the C0 fixture derived from the E. coli FASTA that the
reference ships as test data, and the family/protein/genome generators of SURVEY.md section 8(d).

The synthetic universe is defined with integer-only counter hashing (splitmix64 finaliser) so that the CUDA generators
in tools/benchlib/kg_bench.cu produce the very same bytes for the large configurations.
"""
from __future__ import annotations

import gzip
import os
import struct
from dataclasses import dataclass
from typing import Iterable, List, Sequence, Tuple

import numpy as np

K = 8
MAX_ENCODED = 20 ** 8
EMPTY_KEY = MAX_ENCODED + 1              # "whichKmer > MAX_ENCODED" marks an empty slot (KGJ:1000)
PROT_ALPHA = "ACDEFGHIKLMNPQRSTVWY"
ENTRY_DTYPE = np.dtype([("which", "<i8"), ("otu", "<i4"), ("avg", "<i4"), ("fi", "<i4"), ("wt", "<f4")])
assert ENTRY_DTYPE.itemsize == 24

U64 = np.uint64
_M1, _M2, _GOLD, _MB = U64(0xBF58476D1CE4E5B9), U64(0x94D049BB133111EB), U64(0x9E3779B97F4A7C15), U64(0xD6E8FEB86659FD93)


def mix64(x):
    """splitmix64 finaliser on uint64 arrays (wrapping arithmetic)."""
    with np.errstate(over="ignore"):
        z = np.asarray(x, dtype=U64) + _GOLD
        z = (z ^ (z >> U64(30))) * _M1
        z = (z ^ (z >> U64(27))) * _M2
        return z ^ (z >> U64(31))


def hash3(seed, a, b):
    """h(seed, a, b) = mix64(mix64(seed + a) ^ (b * MB)); the same formula is in tools/benchlib/kg_bench.cu."""
    with np.errstate(over="ignore"):
        return mix64(mix64(U64(seed) + np.asarray(a, dtype=U64)) ^ (np.asarray(b, dtype=U64) * _MB))


# ----------------------------------------------------------------------------------------------------------------
# reference-format writers
# ----------------------------------------------------------------------------------------------------------------
def is_prime(n: int) -> bool:
    if n < 2:
        return False
    if n % 2 == 0:
        return n == 2
    i = 3
    while i * i <= n:
        if n % i == 0:
            return False
        i += 2
    return True


def next_prime(n: int) -> int:
    while not is_prime(n):
        n += 1
    return n


def build_table_image(keys, otu, avg, fi, wt, num_slots: int | None = None, load: float = 0.5,
                      entry_size: int = 24, version: int = 1) -> bytes:
    """kmer.table.mem_map bytes: 3 x LE int64 header (numSigs = SLOT count, entrySize, version) then numSigs 24-byte
    LE entries; keys placed by linear probing from key % numSigs WITHOUT wrap-around (the reference never wraps,
    KGJ:959-1026); the final slot is left empty so that no probe chain can run off the end."""
    keys = np.asarray(keys, dtype=np.int64)
    n = len(keys)
    assert len(np.unique(keys)) == n, "duplicate keys"
    assert n == 0 or (keys.min() >= 0 and keys.max() < MAX_ENCODED)
    if num_slots is None:
        num_slots = next_prime(max(int(n / load) + 1, 11))
    while True:
        home = keys % num_slots
        order = np.lexsort((keys, home))
        h = home[order]
        r = np.arange(n, dtype=np.int64)
        slot = r + np.maximum.accumulate(h - r) if n else h      # first free slot >= home, keys taken in home order
        if n == 0 or slot.max() < num_slots - 1:
            break
        num_slots = next_prime(num_slots + 1)
    ent = np.zeros(num_slots, dtype=ENTRY_DTYPE)
    ent["which"] = EMPTY_KEY
    ent["which"][slot] = keys[order]
    ent["otu"][slot] = np.asarray(otu, dtype=np.int32)[order]
    ent["avg"][slot] = np.asarray(avg, dtype=np.int32)[order]
    ent["fi"][slot] = np.asarray(fi, dtype=np.int32)[order]
    ent["wt"][slot] = np.asarray(wt, dtype=np.float32)[order]
    return struct.pack("<qqq", num_slots, entry_size, version) + ent.tobytes()


def write_table(data_dir: str, keys, otu, avg, fi, wt, gz: bool = False, **kw) -> str:
    os.makedirs(data_dir, exist_ok=True)
    img = build_table_image(keys, otu, avg, fi, wt, **kw)
    path = os.path.join(data_dir, "kmer.table.mem_map" + (".gz" if gz else ""))
    with (gzip.open(path, "wb", compresslevel=1) if gz else open(path, "wb")) as f:
        f.write(img)
    return path


def write_function_index(data_dir: str, names: Sequence[str], gz: bool = False) -> str:
    os.makedirs(data_dir, exist_ok=True)
    path = os.path.join(data_dir, "function.index" + (".gz" if gz else ""))
    body = "".join(f"{i}\t{nm}\n" for i, nm in enumerate(names)).encode()
    with (gzip.open(path, "wb") if gz else open(path, "wb")) as f:
        f.write(body)
    return path


def write_fasta(path: str, ids: Sequence[str], seqs: Sequence[bytes], width: int = 70, descr: Sequence[str] | None = None):
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "wb") as f:
        for i, (name, s) in enumerate(zip(ids, seqs)):
            d = f" {descr[i]}" if descr else ""
            f.write(f">{name}{d}\n".encode())
            for a in range(0, len(s), width):
                f.write(s[a:a + width] + b"\n")


def read_fasta_simple(path: str) -> Tuple[List[str], List[str], List[bytes]]:
    """Well-formed FASTA only (ids, descriptions, sequences).  The faithful reader lives in the host library."""
    op = gzip.open if path.endswith(".gz") else open
    ids, descr, seqs, cur = [], [], [], []
    with op(path, "rb") as f:
        for line in f:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if ids:
                    seqs.append(b"".join(cur))
                parts = line[1:].decode().split(None, 1)
                ids.append(parts[0])
                descr.append(parts[1] if len(parts) > 1 else "")
                cur = []
            elif ids:
                cur.append(line)
    if ids:
        seqs.append(b"".join(cur))
    return ids, descr, seqs


# ----------------------------------------------------------------------------------------------------------------
# encoding helpers (numpy, vectorised) -- used to DERIVE tables from sequences, not to check anything
# ----------------------------------------------------------------------------------------------------------------
_AA_LUT = np.full(256, 20, dtype=np.uint8)
for _i, _c in enumerate(PROT_ALPHA):
    _AA_LUT[ord(_c)] = _i


def aa_codes(seq: bytes) -> np.ndarray:
    return _AA_LUT[np.frombuffer(seq, dtype=np.uint8)]


def window_keys(codes: np.ndarray) -> np.ndarray:
    """Base-20 value of every 8-residue window (first residue most significant); -1 where a code >= 20 occurs."""
    n = len(codes) - K + 1
    if n <= 0:
        return np.zeros(0, dtype=np.int64)
    v = np.zeros(n, dtype=np.int64)
    bad = np.zeros(n, dtype=bool)
    for i in range(K):
        c = codes[i:i + n].astype(np.int64)
        v = v * 20 + c
        bad |= c >= 20
    v[bad] = -1
    return v


def weight_of_key(keys) -> np.ndarray:
    """0.5 + b/256 with b = 8 hashed bits: exactly representable in fp32, and rich in %f rounding ties."""
    b = (mix64(np.asarray(keys, dtype=np.int64).astype(U64)) >> U64(17)) & U64(0xFF)
    return (0.5 + b.astype(np.float64) / 256.0).astype(np.float32)


# ----------------------------------------------------------------------------------------------------------------
# C0: table derived from the E. coli protein FASTA (the repo ships no kmer table)
# ----------------------------------------------------------------------------------------------------------------
def build_c0_fixture(faa_gz: str, data_dir: str, stride: int = 4, gz_table: bool = False) -> dict:
    ids, descr, seqs = read_fasta_simple(faa_gz)
    fun_names: List[str] = []
    fun_index = {}
    keys, otu, avg, fi = [], [], [], []
    for d, s in zip(descr, seqs):
        name = d.split(" [")[0]
        if name not in fun_index:
            fun_index[name] = len(fun_names)
            fun_names.append(name)
        f = fun_index[name]
        codes = aa_codes(s)
        wk = window_keys(codes)
        pos = np.arange(0, len(wk), stride)
        pos = pos[wk[pos] >= 0]
        keys.append(wk[pos])
        avg.append(len(s) - pos)
        fi.append(np.full(len(pos), f, dtype=np.int32))
    keys = np.concatenate(keys)
    avg = np.concatenate(avg).astype(np.int32)
    fi = np.concatenate(fi)
    _, first = np.unique(keys, return_index=True)      # first occurrence wins
    first.sort()
    keys, avg, fi = keys[first], avg[first], fi[first]
    otu = (fi % 97).astype(np.int32)
    wt = weight_of_key(keys)
    path = write_table(data_dir, keys, otu, avg, fi, wt, gz=gz_table)
    write_function_index(data_dir, fun_names)
    return {"table": path, "num_signatures": int(len(keys)), "num_functions": len(fun_names)}


# ----------------------------------------------------------------------------------------------------------------
# synthetic universe; integer-only so CUDA reproduces it bit for bit
# ----------------------------------------------------------------------------------------------------------------
# E. coli residue composition (per mille-ish, section 8(d)) in PROT_ALPHA order A C D E F G H I K L M N P Q R S T V W Y
_COMP = {"L": 10.7, "A": 9.5, "G": 7.4, "V": 7.1, "I": 6.0, "S": 5.8, "E": 5.8, "R": 5.5, "T": 5.4, "D": 5.1,
         "Q": 4.4, "P": 4.4, "K": 4.4, "N": 3.9, "F": 3.9, "Y": 2.8, "M": 2.8, "H": 2.3, "W": 1.5, "C": 1.2}


def residue_cdf16() -> np.ndarray:
    """20 cumulative thresholds on a 16-bit draw: residue = number of thresholds <= draw."""
    p = np.array([_COMP[c] for c in PROT_ALPHA], dtype=np.float64)
    cdf = np.floor(np.cumsum(p / p.sum()) * 65536.0 + 0.5).astype(np.int64)
    cdf[-1] = 65536
    return cdf.astype(np.uint32)


def length_quantiles(nq: int = 4096, median: float = 267.0, sigma: float = 0.60, lo: int = 30, hi: int = 5000) -> np.ndarray:
    """Quantile table of round(LogNormal(ln median, sigma)) clipped to [lo, hi]; computed once on the host and
    shipped to the device as data, so no transcendental function has to agree between CPU and GPU."""
    from statistics import NormalDist
    nd = NormalDist()
    q = (np.arange(nq) + 0.5) / nq
    z = np.array([nd.inv_cdf(float(x)) for x in q])
    return np.clip(np.rint(np.exp(np.log(median) + sigma * z)), lo, hi).astype(np.uint32)


@dataclass
class Universe:
    """F consensus families; everything is a pure function of (seed, family, position)."""
    n_families: int
    seed: int = 0x4B470000
    sig_keep_per_1024: int = 341          # ~1/3 of consensus windows are signatures
    n_functions: int = 50000
    n_otus: int = 1000

    def __post_init__(self):
        self.cdf = residue_cdf16()
        self.lenq = length_quantiles()

    # -- families --
    def family_len(self, f) -> np.ndarray:
        h = hash3(self.seed, f, 0xFFFFFFFF)
        return self.lenq[(h & U64(4095)).astype(np.int64)].astype(np.int64)

    def residue_code(self, f, i) -> np.ndarray:
        draw = (hash3(self.seed ^ 0x11, f, i) & U64(0xFFFF)).astype(np.uint32)
        return np.searchsorted(self.cdf, draw, side="right").astype(np.uint8)

    def random_code(self, seed, a, b) -> np.ndarray:
        draw = (hash3(seed, a, b) & U64(0xFFFF)).astype(np.uint32)
        return np.searchsorted(self.cdf, draw, side="right").astype(np.uint8)

    def consensus(self, f: int) -> np.ndarray:
        n = int(self.family_len(np.array([f]))[0])
        return self.residue_code(np.full(n, f), np.arange(n))

    # -- signature table --
    def signatures(self, max_sigs: int | None = None):
        """All selected consensus windows, first occurrence (family-major order) wins; optionally truncated."""
        F = self.n_families
        fam = np.arange(F, dtype=np.int64)
        lens = self.family_len(fam)
        nwin = np.maximum(lens - K + 1, 0)
        f_of = np.repeat(fam, lens)
        start = np.concatenate([[0], np.cumsum(lens)])
        i_of = np.arange(int(start[-1]), dtype=np.int64) - np.repeat(start[:-1], lens)
        codes = self.residue_code(f_of, i_of)
        keys_all = []
        # window keys per family without crossing family boundaries
        wk = window_keys(codes)                                     # length total-7; mask windows crossing ends
        wi = i_of[: len(wk)]
        wf = f_of[: len(wk)]
        ok = wi + K <= lens[wf]
        sel = (hash3(self.seed ^ 0x22, wf, wi) & U64(1023)).astype(np.int64) < self.sig_keep_per_1024
        m = ok & sel & (wk >= 0)
        keys, wf, wi = wk[m], wf[m], wi[m]
        _, first = np.unique(keys, return_index=True)
        first.sort()
        if max_sigs is not None:
            first = first[:max_sigs]
        keys, wf, wi = keys[first], wf[first], wi[first]
        fi = (wf % self.n_functions).astype(np.int32)
        otu = (wf % self.n_otus).astype(np.int32)
        avg = (lens[wf] - wi).astype(np.int32)
        wt = weight_of_key(keys)
        return keys, otu, avg, fi, wt

    # -- query proteins (C1 / C3 generator) --
    def proteins(self, n: int, seed: int = 1, subst_per_65536: int = 6554, decoy_per_256: int = 51,
                 x_per_2_20: int = 105, first: int = 0) -> List[bytes]:
        alpha = np.frombuffer(PROT_ALPHA.encode(), dtype=np.uint8)
        out = []
        for j in range(first, first + n):
            h = int(hash3(seed, j, 0xFFFFFFFF))
            if (h & 0xFF) < decoy_per_256:
                L = int(self.lenq[(h >> 8) & 4095])
                codes = self.random_code(seed ^ 0x33, np.full(L, j), np.arange(L))
            else:
                f = (h >> 8) % self.n_families
                cons = self.consensus(f)
                L = len(cons)
                hs = hash3(seed ^ 0x44, np.full(L, j), np.arange(L))
                sub = (hs & U64(0xFFFF)).astype(np.int64) < subst_per_65536
                rnd = self.random_code(seed ^ 0x55, np.full(L, j), np.arange(L))
                codes = np.where(sub, rnd, cons)
            hx = hash3(seed ^ 0x66, np.full(L, j), np.arange(L))
            isx = ((hx >> U64(20)) & U64(0xFFFFF)).astype(np.int64) < x_per_2_20
            s = alpha[codes]
            s[isx] = ord("X")
            out.append(s.tobytes())
        return out


# back-translation for the genome generator: one codon table per residue (uniform synonymous choice)
_GENETIC_CODE = ("KNKNTTTTRSRSIIMI" "QHQHPPPPRRRRLLLL" "EDEDAAAAGGGGVVVV" "*Y*YSSSS*CWCLFLF")
_CODONS = {a: [] for a in PROT_ALPHA}
for _idx, _aa in enumerate(_GENETIC_CODE):
    if _aa in _CODONS:
        _CODONS[_aa].append("ACGT"[_idx >> 4] + "ACGT"[(_idx >> 2) & 3] + "ACGT"[_idx & 3])
_COMPL = bytes.maketrans(b"ACGTN", b"TGCAN")


def genome(u: Universe, length: int, seed: int = 2, index: int = 0, n_per_100k: int = 1) -> bytes:
    """Genes = family consensus proteins back-translated with uniform synonymous codons on a random strand, separated
    by exponential(120 bp) uniform-ACGT spacers; ~1e-5 N.  (numpy RNG: only the SMALL test genomes come from here;
    the 5 Mbp bench genomes are made by the CUDA generator with its own counter-based definition.)"""
    rng = np.random.default_rng([seed, index])
    parts, n = [], 0
    while n < length:
        gap = int(rng.exponential(120)) + 1
        parts.append(bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), gap)))
        n += gap
        f = int(rng.integers(u.n_families))
        cons = u.consensus(f)
        gene = "".join(_CODONS[PROT_ALPHA[c]][int(rng.integers(len(_CODONS[PROT_ALPHA[c]])))] for c in cons).encode()
        if rng.random() < 0.5:
            gene = gene.translate(_COMPL)[::-1]
        parts.append(gene)
        n += len(gene)
    g = bytearray(b"".join(parts)[:length])
    for p in rng.integers(0, length, size=max(1, length * n_per_100k // 100000)):
        g[int(p)] = ord("N")
    return bytes(g)
