"""oracle/kgo.py -- ctypes loader for the C oracle (oracle/kg_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs, never by kmergutsjava_b200/.  Parity pin: see oracle/kg_oracle.h (the reference's own source executed through tests/java_pin/j2py.py).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "build", "libkg_oracle.so")
CLI_PATH = os.path.join(HERE, "build", "kmer_guts_oracle")

STREAM_JOIN, DIRECT_PROBE = 0, 1

HIT_DTYPE = np.dtype([("seq", "<i4"), ("sf", "<i4"), ("pos", "<i4"), ("oI", "<i4"), ("avg", "<i4"), ("fI", "<i4"),
                      ("wt", "<f4")])
CALL_DTYPE = np.dtype([("seq", "<i4"), ("sf", "<i4"), ("start", "<i4"), ("end", "<i4"), ("count", "<i4"),
                       ("fI", "<i4"), ("weighted", "<f4"), ("hits_before", "<i4")])
OTU_DTYPE = np.dtype([("n", "<i4"), ("count", "<i4", (5,)), ("oI", "<i4", (5,))])


class Params(C.Structure):
    _fields_ = [("aa", C.c_int32), ("order_constraint", C.c_int32), ("min_hits", C.c_int32),
                ("min_weighted_hits", C.c_int32), ("max_gap", C.c_int32), ("debug", C.c_int32)]


def make_params(aa=False, order_constraint=False, min_hits=5, min_weighted_hits=0, max_gap=200, debug=False):
    return Params(int(aa), int(order_constraint), int(min_hits), int(min_weighted_hits), int(max_gap), int(debug))


def build(force: bool = False) -> None:
    """Compile the C oracle (gcc via oracle/Makefile).  Building the checker is not using it."""
    if force or not (os.path.exists(LIB_PATH) and os.path.exists(CLI_PATH)) or any(
            os.path.getmtime(os.path.join(HERE, f)) > os.path.getmtime(LIB_PATH)
            for f in ("kg_oracle.c", "kg_oracle.h", "kg_oracle_main.c")):
        subprocess.run(["make", "-C", HERE], check=True, stdout=subprocess.DEVNULL)


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.kgo_table_open.restype = C.c_void_p
        L.kgo_table_open.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t]
        L.kgo_table_from_memory.restype = C.c_void_p
        L.kgo_table_from_memory.argtypes = [C.c_void_p, C.c_size_t]
        L.kgo_table_free.argtypes = [C.c_void_p]
        for f in ("kgo_table_num_sigs", "kgo_table_entry_size", "kgo_table_version"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.kgo_run.restype = C.c_void_p
        L.kgo_run.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        L.kgo_run_parallel.restype = C.c_void_p
        L.kgo_run_parallel.argtypes = [C.c_void_p, C.POINTER(Params), C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
        L.kgo_table_borrow.restype = C.c_void_p
        L.kgo_table_borrow.argtypes = [C.c_void_p, C.c_size_t]
        L.kgo_result_free.argtypes = [C.c_void_p]
        for f in ("kgo_result_num_hits", "kgo_result_num_calls", "kgo_result_num_otus"):
            getattr(L, f).restype = C.c_size_t
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("kgo_result_hits", "kgo_result_calls", "kgo_result_otus"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("kgo_result_num_kmers", "kgo_result_kmers_found"):
            getattr(L, f).restype = C.c_int64
            getattr(L, f).argtypes = [C.c_void_p]
        L.kgo_result_lookup_error.restype = C.c_int
        L.kgo_result_lookup_error.argtypes = [C.c_void_p]
        L.kgo_gather_hits.restype = C.c_size_t
        L.kgo_gather_hits.argtypes = [C.POINTER(Params), C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t]
        L.kgo_java_format_f.argtypes = [C.c_float, C.c_int, C.c_char_p, C.c_size_t]
        L.kgo_translate.argtypes = [C.c_char_p, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_size_t]
        L.kgo_rev_comp.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
        L.kgo_encoded_kmer.restype = C.c_int64
        L.kgo_encoded_kmer.argtypes = [C.c_char_p, C.c_size_t]
        L.kgo_genetic_code.restype = C.c_char
        _lib = L
    return _lib


def _copy(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Table:
    def __init__(self, path: Optional[str] = None, data: Optional[bytes] = None, borrow: Optional[np.ndarray] = None):
        L = lib()
        if borrow is not None:      # numpy uint8 image kept alive by this object (no copy: 10 GB tables)
            self._keep = borrow
            self._h = L.kgo_table_borrow(borrow.ctypes.data, borrow.nbytes)
        elif path is not None:
            err = C.create_string_buffer(512)
            self._h = L.kgo_table_open(path.encode(), err, 512)
            if not self._h:
                raise IOError(err.value.decode())
        else:
            self._h = L.kgo_table_from_memory(data, len(data))
        self.num_sigs = L.kgo_table_num_sigs(self._h)
        self.entry_size = L.kgo_table_entry_size(self._h)
        self.version = L.kgo_table_version(self._h)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().kgo_table_free(self._h)
            self._h = None


class Result:
    def __init__(self, h):
        L = lib()
        self.hits = _copy(L.kgo_result_hits(h), L.kgo_result_num_hits(h), HIT_DTYPE)
        self.calls = _copy(L.kgo_result_calls(h), L.kgo_result_num_calls(h), CALL_DTYPE)
        self.otus = _copy(L.kgo_result_otus(h), L.kgo_result_num_otus(h), OTU_DTYPE)
        self.num_kmers = L.kgo_result_num_kmers(h)
        self.kmers_found = L.kgo_result_kmers_found(h)
        self.lookup_error = L.kgo_result_lookup_error(h)
        L.kgo_result_free(h)


def concat(seqs: Sequence[bytes]):
    off = np.zeros(len(seqs) + 1, dtype=np.uint64)
    if len(seqs):
        off[1:] = np.cumsum([len(s) for s in seqs], dtype=np.uint64)
    return np.frombuffer(b"".join(seqs), dtype=np.uint8), off


def run(table: Table, params: Params, seq_bytes: np.ndarray, offsets: np.ndarray, variant: int = STREAM_JOIN,
        threads: int = 1) -> Result:
    seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    if threads > 1:
        h = lib().kgo_run_parallel(table._h, C.byref(params), seq_bytes.ctypes.data, offsets.ctypes.data,
                                   len(offsets) - 1, variant, threads)
    else:
        h = lib().kgo_run(table._h, C.byref(params), seq_bytes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, variant)
    return Result(h)


def gather_hits(params: Params, hits: np.ndarray, otu: Optional[np.ndarray] = None, max_calls: int = 4096):
    """FSM alone on one container's hits (HIT_DTYPE).  Returns (calls, otu)."""
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
    otu = np.zeros(1, dtype=OTU_DTYPE) if otu is None else otu.copy()
    calls = np.zeros(max_calls, dtype=CALL_DTYPE)
    n = lib().kgo_gather_hits(C.byref(params), hits.ctypes.data, len(hits), otu.ctypes.data, calls.ctypes.data,
                              max_calls)
    assert n <= max_calls
    return calls[:n], otu


def java_format_f(v: float, prec: int = 6) -> str:
    out = C.create_string_buffer(128)
    lib().kgo_java_format_f(C.c_float(v), prec, out, 128)
    return out.value.decode()


def run_cli(args: Sequence[str], check: bool = True) -> subprocess.CompletedProcess:
    if not os.path.exists(CLI_PATH):
        build()
    return subprocess.run([CLI_PATH, *args], check=check, capture_output=True, text=True)
