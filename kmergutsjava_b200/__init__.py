"""kmergutsjava_b200 -- ctypes binding of libkmerguts_b200.so (include/kmerguts.h, kmerguts_host.h, kmerguts_shard.h).

The product is the CUDA library; this module only loads it and wraps handles in small Python classes so that the
tests and bench.py read like the reference's own driver (KmerGutsJava.run: load table, feed sequences, print calls).
There is no CPU fallback: if the shared library is missing, or no B200 is visible, every entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("KG_LIB") or os.path.join(HERE, "libkmerguts_b200.so")  # KG_LIB: kernel-variant experiments
CLI_PATH = os.path.join(HERE, "bin", "kmer_guts_b200")

MODE_DNA, MODE_AA = 0, 1

CALL_DTYPE = np.dtype([("seq", "<u4"), ("sf", "<i4"), ("start", "<i4"), ("end", "<i4"), ("count", "<i4"), ("fI", "<i4"),
                       ("weighted", "<f4"), ("hits_before", "<i4")])
OTU_DTYPE = np.dtype([("n", "<i4"), ("count", "<i4", (5,)), ("oI", "<i4", (5,))])
OTU_ENTRY_DTYPE = np.dtype([("count", "<i4"), ("oI", "<i4")])
HIT_DTYPE = np.dtype([("seq", "<u4"), ("sf", "<i4"), ("pos", "<i4"), ("oI", "<i4"), ("avg", "<i4"), ("fI", "<i4"),
                      ("wt", "<f4")])

# every symbol the three headers declare (tests/test_abi.py checks the library exports exactly these)
EXPORTS = [
    "kg_init", "kg_shutdown", "kg_last_error", "kg_version",
    "kg_table_load", "kg_table_load_file", "kg_table_from_image", "kg_table_from_device_entries", "kg_table_get_info",
    "kg_table_save", "kg_table_load_cached", "kg_table_load_cached_checked", "kg_table_attach", "kg_table_free", "kg_params_default", "kg_run", "kg_run_packed_aa", "kg_pack_aa", "kg_pack_aa_groups", "kg_pack_dna", "kg_run_packed_dna", "kg_result_otus_compact", "kg_batch_upload", "kg_batch_from_device", "kg_batch_free",
    "kg_batch_run", "kg_batch_run_many", "kg_batch_submit", "kg_batch_collect", "kg_result_fetch", "kg_result_stats", "kg_result_calls", "kg_result_otus", "kg_result_hits",
    "kg_result_free",
    "kg_fasta_read", "kg_fasta_count", "kg_fasta_id", "kg_fasta_bytes", "kg_fasta_offsets", "kg_fasta_free",
    "kg_functions_load", "kg_functions_read", "kg_functions_count", "kg_functions_name", "kg_functions_free",
    "kg_format_java_f", "kg_report_write", "kg_main",
    "kg_fasta_stream_open", "kg_fasta_stream_next", "kg_fasta_stream_close",
    "kg_report_open", "kg_report_add", "kg_report_close", "kg_call_dna_range",
    "kg_shard_owner", "kg_comm_unique_id", "kg_comm_init", "kg_comm_init_local", "kg_comm_free", "kg_comm_last_stats",
    "kg_table_load_sharded", "kg_table_from_image_sharded", "kg_table_from_device_entries_sharded",
    "kg_batch_run_sharded", "kg_batch_run_sharded_local",
]


class KgError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"kmerguts error {code}: {msg}")
        self.code = code


class Params(C.Structure):
    _fields_ = [("min_hits", C.c_int32), ("min_weighted_hits", C.c_int32), ("max_gap", C.c_int32),
                ("order_constraint", C.c_int32), ("emit_hits", C.c_int32)]


class TableInfo(C.Structure):
    _fields_ = [(n, C.c_int64) for n in ("num_slots", "entry_size", "version", "num_signatures", "num_unreachable",
                                         "tail_run", "num_buckets", "flagged_buckets", "device_bytes")]


class RunStats(C.Structure):
    _fields_ = [("num_sequences", C.c_uint64), ("num_positions", C.c_uint64), ("num_kmers", C.c_uint64),
                ("num_hits", C.c_uint64), ("num_calls", C.c_uint64), ("num_launches", C.c_uint32),
                ("ms_h2d", C.c_float), ("ms_device", C.c_float), ("ms_d2h", C.c_float),
                ("ms_prepare", C.c_float), ("ms_probe", C.c_float), ("ms_group", C.c_float),
                ("ms_filter", C.c_float), ("ms_refilter", C.c_float), ("ms_lines", C.c_float),
                ("num_survivors1", C.c_uint64), ("num_survivors2", C.c_uint64)]


class ShardStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("keys_sent", "keys_remote", "keys_received", "replies_sent", "replies_received",
                                          "bytes_sent")] + \
               [(n, C.c_float) for n in ("ms_route", "ms_keys", "ms_answer", "ms_replies", "ms_merge", "ms_total")] + \
               [("chunks", C.c_int32)]


_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """The shared library, or an exception: the product path must fail loudly when the CUDA extension is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KgError(-2, f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'` "
                          f"or `make -C kmergutsjava_b200/csrc`); there is no CPU fallback")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64, sz = C.c_void_p, C.c_int, C.c_uint64, C.c_size_t
    pp = C.POINTER(vp)
    sig = {
        "kg_init": (i32, [i32, pp]), "kg_shutdown": (None, [vp]), "kg_last_error": (C.c_char_p, []),
        "kg_version": (C.c_char_p, []),
        "kg_table_load": (i32, [vp, C.c_char_p, pp]), "kg_table_load_file": (i32, [vp, C.c_char_p, pp]),
        "kg_table_from_image": (i32, [vp, vp, sz, pp]), "kg_table_from_device_entries": (i32, [vp, vp, vp, sz, pp]),
        "kg_table_get_info": (i32, [vp, C.POINTER(TableInfo)]), "kg_table_free": (None, [vp]),
        "kg_table_attach": (i32, [vp, vp]),
        "kg_table_save": (i32, [vp, vp, C.c_char_p]), "kg_table_load_cached": (i32, [vp, C.c_char_p, pp]),
        "kg_table_load_cached_checked": (i32, [vp, C.c_char_p, C.c_char_p, pp]),
        "kg_params_default": (None, [C.POINTER(Params)]),
        "kg_run": (i32, [vp, vp, i32, vp, vp, sz, C.POINTER(Params), pp]),
        "kg_run_packed_aa": (i32, [vp, vp, vp, vp, sz, C.POINTER(Params), pp]),
        "kg_pack_aa": (i32, [vp, vp, sz, vp, vp, i32]), "kg_pack_aa_groups": (u64, [u64]),
        "kg_pack_dna": (i32, [vp, vp, sz, vp, vp, vp, C.POINTER(sz), i32]),
        "kg_run_packed_dna": (i32, [vp, vp, vp, vp, vp, vp, sz, sz, C.POINTER(Params), pp]),
        "kg_result_otus_compact": (i32, [vp, pp, pp, C.POINTER(sz), C.POINTER(sz)]),
        "kg_batch_upload": (i32, [vp, i32, vp, vp, sz, pp]),
        "kg_batch_from_device": (i32, [vp, i32, vp, vp, sz, u64, pp]), "kg_batch_free": (None, [vp]),
        "kg_batch_run": (i32, [vp, vp, vp, C.POINTER(Params), pp]), "kg_result_fetch": (i32, [vp]),
        "kg_batch_run_many": (i32, [vp, vp, pp, sz, C.POINTER(Params), pp]),
        "kg_batch_submit": (i32, [vp, vp, vp, C.POINTER(Params)]), "kg_batch_collect": (i32, [vp, pp]),
        "kg_result_stats": (i32, [vp, C.POINTER(RunStats)]),
        "kg_result_calls": (i32, [vp, pp, C.POINTER(sz)]), "kg_result_otus": (i32, [vp, pp, C.POINTER(sz)]),
        "kg_result_hits": (i32, [vp, pp, C.POINTER(sz)]), "kg_result_free": (None, [vp]),
        "kg_fasta_read": (i32, [C.c_char_p, pp]), "kg_fasta_count": (sz, [vp]), "kg_fasta_id": (C.c_char_p, [vp, sz]),
        "kg_fasta_bytes": (vp, [vp]), "kg_fasta_offsets": (vp, [vp]), "kg_fasta_free": (None, [vp]),
        "kg_functions_load": (i32, [C.c_char_p, pp]), "kg_functions_read": (i32, [C.c_char_p, pp]),
        "kg_functions_count": (sz, [vp]), "kg_functions_name": (C.c_char_p, [vp, sz]), "kg_functions_free": (None, [vp]),
        "kg_format_java_f": (i32, [C.c_float, i32, C.c_char_p, sz]),
        "kg_report_write": (i32, [C.c_char_p, i32, i32, vp, vp, vp, vp]),
        "kg_main": (i32, [i32, C.POINTER(C.c_char_p)]),
        "kg_fasta_stream_open": (i32, [C.c_char_p, sz, pp]), "kg_fasta_stream_next": (i32, [vp, pp]),
        "kg_fasta_stream_close": (None, [vp]),
        "kg_report_open": (i32, [C.c_char_p, pp]), "kg_report_add": (i32, [vp, i32, i32, vp, vp, vp, vp]),
        "kg_report_close": (i32, [vp]),
        "kg_call_dna_range": (i32, [vp, u64, C.POINTER(u64), C.POINTER(u64), C.c_char_p]),
        "kg_shard_owner": (i32, [u64, i32]), "kg_comm_unique_id": (i32, [vp]),
        "kg_comm_init": (i32, [vp, i32, i32, vp, pp]), "kg_comm_init_local": (i32, [pp, i32, pp]),
        "kg_comm_free": (None, [vp]), "kg_comm_last_stats": (i32, [vp, C.POINTER(ShardStats)]),
        "kg_table_load_sharded": (i32, [vp, C.c_char_p, i32, i32, pp]),
        "kg_table_from_image_sharded": (i32, [vp, vp, sz, i32, i32, pp]),
        "kg_table_from_device_entries_sharded": (i32, [vp, vp, vp, sz, i32, i32, pp]),
        "kg_batch_run_sharded": (i32, [vp, vp, vp, C.POINTER(Params), pp]),
        "kg_batch_run_sharded_local": (i32, [pp, pp, pp, i32, C.POINTER(Params), pp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def _check(rc: int):
    if rc != 0:
        raise KgError(rc, lib().kg_last_error().decode(errors="replace"))


def pack_aa(seq_bytes: np.ndarray, offsets: np.ndarray, threads: int = 1):
    """kg_pack_aa: (packed uint8[5 * groups], group_offsets uint64[n+1]) -- toAminoAcidOff codes, 8 per 5 bytes."""
    seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    goff = np.zeros(n + 1, dtype=np.uint64)
    _check(lib().kg_pack_aa(seq_bytes.ctypes.data, offsets.ctypes.data, n, None, goff.ctypes.data, 1))
    packed = np.empty(5 * int(goff[-1]), dtype=np.uint8)
    _check(lib().kg_pack_aa(seq_bytes.ctypes.data, offsets.ctypes.data, n, packed.ctypes.data, goff.ctypes.data, threads))
    return packed, goff


def pack_dna(seq_bytes: np.ndarray, offsets: np.ndarray, threads: int = 1):
    """kg_pack_dna: (packed 2-bit nucleotides, byte offsets (n + 1), sorted positions of the non-ACGTU characters)."""
    seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
    offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
    n = len(offsets) - 1
    boff = np.zeros(n + 1, dtype=np.uint64)
    ne = C.c_size_t(0)
    _check(lib().kg_pack_dna(seq_bytes.ctypes.data, offsets.ctypes.data, n, None, boff.ctypes.data, None, C.byref(ne), threads))
    packed = np.zeros(int(boff[-1]) + 64, dtype=np.uint8)
    exc = np.zeros(max(int(ne.value), 1), dtype=np.uint64)
    _check(lib().kg_pack_dna(seq_bytes.ctypes.data, offsets.ctypes.data, n, packed.ctypes.data, boff.ctypes.data, exc.ctypes.data,
                             C.byref(ne), threads))
    return packed, boff, exc[:int(ne.value)]


def default_params(**kw) -> Params:
    p = Params()
    lib().kg_params_default(C.byref(p))
    for k, v in kw.items():
        setattr(p, k, int(v))
    return p


def java_format_f(v: float, prec: int = 6) -> str:
    out = C.create_string_buffer(128)
    _check(lib().kg_format_java_f(C.c_float(v), prec, out, 128))
    return out.value.decode()


class Context:
    """One GPU (kg_context)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _check(lib().kg_init(device, C.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            lib().kg_shutdown(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # -- tables --
    def load_table(self, data_dir: str) -> "Table":
        h = C.c_void_p()
        _check(lib().kg_table_load(self._h, data_dir.encode(), C.byref(h)))
        return Table(self, h)

    def load_table_file(self, path: str) -> "Table":
        """One reference-format file (kmer.table.mem_map, or .gz by its suffix, KGJ:927)."""
        h = C.c_void_p()
        _check(lib().kg_table_load_file(self._h, path.encode(), C.byref(h)))
        return Table(self, h)

    def load_table_cached(self, path: str, data_dir: Optional[str] = None) -> "Table":
        """data_dir given: only if the cache was built from the kmer.table.mem_map[.gz] that directory holds now."""
        h = C.c_void_p()
        if data_dir is None:
            _check(lib().kg_table_load_cached(self._h, path.encode(), C.byref(h)))
        else:
            _check(lib().kg_table_load_cached_checked(self._h, path.encode(), data_dir.encode(), C.byref(h)))
        return Table(self, h)

    def table_from_image(self, image: bytes) -> "Table":
        h = C.c_void_p()
        buf = np.frombuffer(image, dtype=np.uint8)
        _check(lib().kg_table_from_image(self._h, buf.ctypes.data, len(buf), C.byref(h)))
        return Table(self, h)

    def table_from_device_entries(self, d_keys: int, d_payload: int, n: int) -> "Table":
        h = C.c_void_p()
        _check(lib().kg_table_from_device_entries(self._h, d_keys, d_payload, n, C.byref(h)))
        return Table(self, h)

    # -- the path --
    # ---- hash-sharded table (include/kmerguts_shard.h) ----
    def load_table_sharded(self, data_dir: str, rank: int, nranks: int) -> "Table":
        h = C.c_void_p()
        _check(lib().kg_table_load_sharded(self._h, data_dir.encode(), rank, nranks, C.byref(h)))
        return Table(self, h)

    def table_from_image_sharded(self, image: bytes, rank: int, nranks: int) -> "Table":
        h = C.c_void_p()
        buf = np.frombuffer(image, dtype=np.uint8)
        _check(lib().kg_table_from_image_sharded(self._h, buf.ctypes.data, buf.size, rank, nranks, C.byref(h)))
        return Table(self, h)

    def table_from_device_entries_sharded(self, d_keys: int, d_payload: int, n: int, rank: int, nranks: int) -> "Table":
        h = C.c_void_p()
        _check(lib().kg_table_from_device_entries_sharded(self._h, d_keys, d_payload, n, rank, nranks, C.byref(h)))
        return Table(self, h)

    def run(self, table: "Table", mode: int, seq_bytes: np.ndarray, offsets: np.ndarray, params: Params) -> "Result":
        """Host buffers in, host results out (kg_run): the end-to-end call."""
        seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h = C.c_void_p()
        _check(lib().kg_run(self._h, table._h, mode, seq_bytes.ctypes.data, offsets.ctypes.data, len(offsets) - 1,
                            C.byref(params), C.byref(h)))
        return Result(h)

    def run_ptr(self, table: "Table", mode: int, seq_ptr: int, off_ptr: int, n: int, params: Params) -> "Result":
        h = C.c_void_p()
        _check(lib().kg_run(self._h, table._h, mode, seq_ptr, off_ptr, n, C.byref(params), C.byref(h)))
        return Result(h)

    def run_packed_aa(self, table: "Table", packed: np.ndarray, group_offsets: np.ndarray, params: Params) -> "Result":
        """kg_run for proteins in the 5-bit packed form of pack_aa()."""
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        group_offsets = np.ascontiguousarray(group_offsets, dtype=np.uint64)
        return self.run_packed_aa_ptr(table, packed.ctypes.data, group_offsets.ctypes.data, len(group_offsets) - 1, params)

    def run_packed_dna(self, table: "Table", packed: np.ndarray, offsets: np.ndarray, byte_offsets: np.ndarray, exceptions: np.ndarray,
                       params: Params) -> "Result":
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        return self.run_packed_dna_ptr(table, packed.ctypes.data, offsets.ctypes.data, byte_offsets.ctypes.data,
                                       exceptions.ctypes.data if len(exceptions) else None, len(exceptions), len(offsets) - 1, params,
                                       _keep=(packed, offsets, byte_offsets, exceptions))

    def run_packed_dna_ptr(self, table: "Table", packed_ptr, off_ptr, boff_ptr, exc_ptr, n_exc: int, n: int, params: Params, _keep=None) -> "Result":
        h = C.c_void_p()
        _check(lib().kg_run_packed_dna(self._h, table._h, packed_ptr, off_ptr, boff_ptr, exc_ptr, n_exc, n, C.byref(params), C.byref(h)))
        return Result(h)

    def run_packed_aa_ptr(self, table: "Table", packed_ptr: int, goff_ptr: int, n: int, params: Params) -> "Result":
        h = C.c_void_p()
        _check(lib().kg_run_packed_aa(self._h, table._h, packed_ptr, goff_ptr, n, C.byref(params), C.byref(h)))
        return Result(h)

    def upload(self, mode: int, seq_bytes: np.ndarray, offsets: np.ndarray) -> "Batch":
        seq_bytes = np.ascontiguousarray(seq_bytes, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        h = C.c_void_p()
        _check(lib().kg_batch_upload(self._h, mode, seq_bytes.ctypes.data, offsets.ctypes.data, len(offsets) - 1, C.byref(h)))
        return Batch(self, h)

    def batch_from_device(self, mode: int, d_seq: int, d_off: int, n: int, total: int) -> "Batch":
        h = C.c_void_p()
        _check(lib().kg_batch_from_device(self._h, mode, d_seq, d_off, n, total, C.byref(h)))
        return Batch(self, h)

    def run_batch_many(self, table: "Table", batches: Sequence["Batch"], params: Params) -> list["Result"]:
        """kg_batch_run_many: resident batches, two in flight (the FSM of batch i overlaps the probe of batch i+1)."""
        n = len(batches)
        arr = (C.c_void_p * n)(*[b._h for b in batches])
        out = (C.c_void_p * n)()
        _check(lib().kg_batch_run_many(self._h, table._h, arr, n, C.byref(params), out))
        return [Result(C.c_void_p(out[i])) for i in range(n)]

    def run_batches(self, table: "Table", batches, params: Params):
        """kg_batch_submit / kg_batch_collect: yields the results in order with two batches in flight; the caller frees each
        result before asking for the next, so the buffers recycle."""
        it = iter(batches)
        pending = 0
        for b in it:
            _check(lib().kg_batch_submit(self._h, table._h, b._h, C.byref(params)))
            pending += 1
            if pending == 2:
                h = C.c_void_p()
                _check(lib().kg_batch_collect(self._h, C.byref(h)))
                pending -= 1
                yield Result(h)
        while pending:
            h = C.c_void_p()
            _check(lib().kg_batch_collect(self._h, C.byref(h)))
            pending -= 1
            yield Result(h)

    def run_batch(self, table: "Table", batch: "Batch", params: Params) -> "Result":
        h = C.c_void_p()
        _check(lib().kg_batch_run(self._h, table._h, batch._h, C.byref(params), C.byref(h)))
        return Result(h)


def shard_owner(key: int, nranks: int) -> int:
    return int(lib().kg_shard_owner(int(key), int(nranks)))


class Comm:
    """One rank of a sharded-table communicator.  `Comm.unique_id()` on rank 0, pass the 128 bytes to every rank, then
    `Comm(ctx, rank, nranks, id)` on all of them (collective).  `Comm.local(ctxs)` puts all ranks in this process."""

    def __init__(self, ctx: Context, rank: int = 0, nranks: int = 1, uid: Optional[bytes] = None, _h=None):
        self.ctx, self.rank, self.nranks = ctx, rank, nranks
        if _h is not None:
            self._h = _h
            return
        h = C.c_void_p()
        buf = C.create_string_buffer(uid, 128) if uid is not None else None
        _check(lib().kg_comm_init(ctx._h, rank, nranks, C.cast(buf, C.c_void_p) if buf is not None else None, C.byref(h)))
        self._h = h

    @staticmethod
    def unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().kg_comm_unique_id(C.cast(buf, C.c_void_p)))
        return buf.raw

    @staticmethod
    def local(ctxs: Sequence[Context]) -> list["Comm"]:
        n = len(ctxs)
        arr = (C.c_void_p * n)(*[c._h for c in ctxs])
        out = (C.c_void_p * n)()
        _check(lib().kg_comm_init_local(arr, n, out))
        return [Comm(ctxs[r], r, n, _h=C.c_void_p(out[r])) for r in range(n)]

    def run(self, table: "Table", batch: "Batch", params: Params) -> "Result":
        h = C.c_void_p()
        _check(lib().kg_batch_run_sharded(self._h, table._h, batch._h, C.byref(params), C.byref(h)))
        return Result(h)

    @property
    def stats(self) -> ShardStats:
        s = ShardStats()
        _check(lib().kg_comm_last_stats(self._h, C.byref(s)))
        return s

    def free(self):
        if self._h:
            lib().kg_comm_free(self._h)
            self._h = None


def run_sharded_local(comms: Sequence[Comm], tables: Sequence["Table"], batches: Sequence["Batch"], params: Params) -> list["Result"]:
    n = len(comms)
    ca = (C.c_void_p * n)(*[c._h for c in comms])
    ta = (C.c_void_p * n)(*[t._h for t in tables])
    ba = (C.c_void_p * n)(*[b._h for b in batches])
    out = (C.c_void_p * n)()
    _check(lib().kg_batch_run_sharded_local(ca, ta, ba, n, C.byref(params), out))
    return [Result(C.c_void_p(out[r])) for r in range(n)]


class Table:
    def __init__(self, ctx: Context, h):
        self.ctx, self._h = ctx, h

    @property
    def info(self) -> TableInfo:
        ti = TableInfo()
        _check(lib().kg_table_get_info(self._h, C.byref(ti)))
        return ti

    def attach(self, ctx: "Context"):
        """Use this table from another context of the same device (one context per host thread)."""
        _check(lib().kg_table_attach(ctx._h, self._h))

    def save(self, path: str):
        _check(lib().kg_table_save(self.ctx._h, self._h, path.encode()))

    def free(self):
        if self._h:
            lib().kg_table_free(self._h)
            self._h = C.c_void_p()


class Batch:
    def __init__(self, ctx: Context, h):
        self.ctx, self._h = ctx, h

    def free(self):
        if self._h:
            lib().kg_batch_free(self._h)
            self._h = C.c_void_p()


def _view(ptr, n, dtype) -> np.ndarray:
    if not ptr or n == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_char * (n * dtype.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n).copy()


class Result:
    def __init__(self, h):
        self._h = h

    @property
    def stats(self) -> RunStats:
        s = RunStats()
        _check(lib().kg_result_stats(self._h, C.byref(s)))
        return s

    def _arr(self, fn, dtype):
        p, n = C.c_void_p(), C.c_size_t()
        _check(fn(self._h, C.byref(p), C.byref(n)))
        return _view(p.value, n.value, dtype)

    @property
    def calls(self) -> np.ndarray:
        return self._arr(lib().kg_result_calls, CALL_DTYPE)

    @property
    def otus(self) -> np.ndarray:
        return self._arr(lib().kg_result_otus, OTU_DTYPE)

    @property
    def hits(self) -> np.ndarray:
        return self._arr(lib().kg_result_hits, HIT_DTYPE)

    @property
    def otus_compact(self):
        """(n_per_seq uint8[n], entries [(count, oI)]): the OTU counts without the unused slots (kg_result_otus_compact)."""
        pn, pe, ns, ne = C.c_void_p(), C.c_void_p(), C.c_size_t(), C.c_size_t()
        _check(lib().kg_result_otus_compact(self._h, C.byref(pn), C.byref(pe), C.byref(ns), C.byref(ne)))
        return _view(pn.value, ns.value, np.dtype("u1")), _view(pe.value, ne.value, OTU_ENTRY_DTYPE)

    def fetch(self):
        _check(lib().kg_result_fetch(self._h))

    def free(self):
        if self._h:
            lib().kg_result_free(self._h)
            self._h = C.c_void_p()


def call_dna_range(call, contig_len: int):
    """(begin, end, strand) of a CALL of a 6-frame run on the contig: 0-based, inclusive (kg_call_dna_range)."""
    rec = np.zeros(1, dtype=CALL_DTYPE)
    for f in CALL_DTYPE.names:
        rec[f] = call[f]
    b, e, sd = C.c_uint64(), C.c_uint64(), C.create_string_buffer(2)
    _check(lib().kg_call_dna_range(rec.ctypes.data, contig_len, C.byref(b), C.byref(e), sd))
    return int(b.value), int(e.value), sd.raw[:1].decode()


def fasta_batches(path: str, batch_bytes: int):
    """kg_fasta_stream_*: the file as a sequence of Fasta objects of about batch_bytes of text each (whole records)."""
    h = C.c_void_p()
    _check(lib().kg_fasta_stream_open(path.encode(), batch_bytes, C.byref(h)))
    try:
        while True:
            b = C.c_void_p()
            _check(lib().kg_fasta_stream_next(h, C.byref(b)))
            if not b:
                break
            yield Fasta(None, _h=b)
    finally:
        lib().kg_fasta_stream_close(h)


class Fasta:
    """readFasta (KGJ:1132-1192) through the host library."""

    def __init__(self, path, _h=None):
        self._h = _h if _h is not None else C.c_void_p()
        if _h is None:
            _check(lib().kg_fasta_read(path.encode(), C.byref(self._h)))
        L = lib()
        self.n = L.kg_fasta_count(self._h)
        self.ids = [L.kg_fasta_id(self._h, i).decode(errors="replace") for i in range(self.n)]
        off_p = L.kg_fasta_offsets(self._h)
        self.offsets = np.frombuffer((C.c_uint64 * (self.n + 1)).from_address(off_p), dtype=np.uint64).copy()
        total = int(self.offsets[-1])
        b_p = L.kg_fasta_bytes(self._h)
        self.bytes = np.frombuffer((C.c_uint8 * total).from_address(b_p), dtype=np.uint8).copy() if total else np.zeros(0, np.uint8)

    def free(self):
        if self._h:
            lib().kg_fasta_free(self._h)
            self._h = C.c_void_p()
