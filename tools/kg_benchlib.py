"""tools/kg_benchlib.py -- ctypes binding of tools/benchlib/libkmerguts_bench.so (tools/benchlib/kmerguts_bench.h): CUDA
generators of the synthetic universe of tools/kg_synth.py, the device-side writer of the reference's table format, the naive
cross-check scan and the random-sector roofline microbenchmark.  Bench / test tooling only: the product library
(kmergutsjava_b200/libkmerguts_b200.so) carries none of it."""
import ctypes as C
import os

import numpy as np

import kmergutsjava_b200 as kg
from kmergutsjava_b200 import Context, Table, HIT_DTYPE

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "benchlib", "libkmerguts_bench.so")


class UniverseStruct(C.Structure):
    _fields_ = [("n_families", C.c_uint64), ("seed", C.c_uint64), ("sig_keep_per_1024", C.c_uint32),
                ("n_functions", C.c_uint32), ("n_otus", C.c_uint32), ("cdf16", C.c_uint32 * 20),
                ("lenq", C.c_uint32 * 4096)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    kg.lib()  # the handles come from the product library; load it first so that the dependency resolves to the same copy
    if not os.path.exists(LIB_PATH):
        raise kg.KgError(-2, f"{LIB_PATH} is not built (make -C tools/benchlib)")
    L = C.CDLL(LIB_PATH)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    pp = C.POINTER(vp)
    sig = {
        "kg_synth_signatures": (i32, [vp, C.POINTER(UniverseStruct), u64, pp, pp, C.POINTER(u64)]),
        "kg_synth_signatures_sharded": (i32, [vp, C.POINTER(UniverseStruct), u64, i32, i32, pp, pp, C.POINTER(u64)]),
        "kg_synth_proteins": (i32, [vp, C.POINTER(UniverseStruct), u64, u64, u64, pp, pp, C.POINTER(u64)]),
        "kg_synth_genomes": (i32, [vp, C.POINTER(UniverseStruct), u64, u64, u64, pp, pp, C.POINTER(u64)]),
        "kg_synth_genomes_range": (i32, [vp, C.POINTER(UniverseStruct), u64, u64, u64, u64, pp, pp, C.POINTER(u64)]),
        "kg_synth_reference_image": (i32, [vp, vp, vp, u64, u64, C.POINTER(u64), pp, C.POINTER(C.c_double)]),
        "kg_synth_naive_scan_aa": (i32, [vp, vp, vp, vp, u64, u64, vp]),
        "kg_synth_hits_checksum": (i32, [vp, vp, u64, vp, C.POINTER(u64)]),
        "kg_device_free": (None, [vp]), "kg_device_to_host": (i32, [vp, vp, vp, u64]),
        "kg_probe_roofline": (i32, [vp, u64, u64, i32, i32, C.POINTER(C.c_double)]),
        "kg_probe_roofline_table": (i32, [vp, vp, u64, i32, i32, C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


_check = kg._check


def probe_roofline(ctx: Context, nbytes: int, n_loads: int, tpb: int = 256, inflight: int = 4) -> float:
    out = C.c_double()
    _check(lib().kg_probe_roofline(ctx._h, nbytes, n_loads, tpb, inflight, C.byref(out)))
    return out.value


def probe_roofline_table(ctx: Context, table: Table, n_loads: int, tpb: int = 256, inflight: int = 4) -> float:
    out = C.c_double()
    _check(lib().kg_probe_roofline_table(ctx._h, table._h, n_loads, tpb, inflight, C.byref(out)))
    return out.value


def to_host(ctx: Context, d_ptr: int, nbytes: int) -> np.ndarray:
    out = np.empty(nbytes, dtype=np.uint8)
    _check(lib().kg_device_to_host(ctx._h, out.ctypes.data, d_ptr, nbytes))
    return out


def make_universe(u) -> UniverseStruct:
    """tools.kg_synth.Universe -> the C struct the CUDA generators take."""
    s = UniverseStruct()
    s.n_families, s.seed, s.sig_keep_per_1024 = u.n_families, u.seed, u.sig_keep_per_1024
    s.n_functions, s.n_otus = u.n_functions, u.n_otus
    for i, v in enumerate(u.cdf):
        s.cdf16[i] = int(v)
    for i, v in enumerate(u.lenq):
        s.lenq[i] = int(v)
    return s


def synth_signatures(ctx: Context, u, max_sigs: int = 0):
    dk, dp, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    us = make_universe(u)
    _check(lib().kg_synth_signatures(ctx._h, C.byref(us), max_sigs, C.byref(dk), C.byref(dp), C.byref(n)))
    return dk.value, dp.value, n.value


def synth_signatures_sharded(ctx: Context, u, rank: int, nranks: int):
    """The signatures of the universe that `rank` owns (device arrays); the union over ranks = synth_signatures(u, 0)."""
    dk, dp, n = C.c_void_p(), C.c_void_p(), C.c_uint64()
    us = make_universe(u)
    _check(lib().kg_synth_signatures_sharded(ctx._h, C.byref(us), 0, rank, nranks, C.byref(dk), C.byref(dp), C.byref(n)))
    return dk.value, dp.value, n.value


def synth_proteins(ctx: Context, u, first: int, n: int, seed: int):
    ds, do, total = C.c_void_p(), C.c_void_p(), C.c_uint64()
    us = make_universe(u)
    _check(lib().kg_synth_proteins(ctx._h, C.byref(us), first, n, seed, C.byref(ds), C.byref(do), C.byref(total)))
    return ds.value, do.value, total.value


def synth_genomes(ctx: Context, u, n_genomes: int, length: int, seed: int):
    ds, do, total = C.c_void_p(), C.c_void_p(), C.c_uint64()
    us = make_universe(u)
    _check(lib().kg_synth_genomes(ctx._h, C.byref(us), n_genomes, length, seed, C.byref(ds), C.byref(do), C.byref(total)))
    return ds.value, do.value, total.value


def synth_genomes_range(ctx: Context, u, first: int, n_genomes: int, length: int, seed: int):
    """Genomes first .. first+n-1 of the job synth_genomes(n_total) generates, byte for byte."""
    ds, do, total = C.c_void_p(), C.c_void_p(), C.c_uint64()
    us = make_universe(u)
    _check(lib().kg_synth_genomes_range(ctx._h, C.byref(us), first, n_genomes, length, seed, C.byref(ds), C.byref(do), C.byref(total)))
    return ds.value, do.value, total.value


def to_host_at(ctx: Context, d_ptr: int, offset: int, nbytes: int) -> np.ndarray:
    return to_host(ctx, d_ptr + offset, nbytes)


def synth_reference_image(ctx: Context, d_keys: int, d_payload: int, n: int, min_slots: int) -> np.ndarray:
    """kmer.table.mem_map image written on the device (reference format, no wrap-around), copied to the host."""
    ns, dimg, disp = C.c_uint64(), C.c_void_p(), C.c_double()
    _check(lib().kg_synth_reference_image(ctx._h, d_keys, d_payload, n, min_slots, C.byref(ns), C.byref(dimg), C.byref(disp)))
    synth_reference_image.mean_displacement = disp.value
    try:
        return to_host(ctx, dimg.value, 24 + 24 * ns.value)
    finally:
        device_free(dimg.value)


def naive_scan_aa(ctx: Context, table: Table, d_seq: int, d_off: int, n: int, total: int):
    """(valid windows, hits, checksum) by the naive one-thread-per-position kernel."""
    out = (C.c_uint64 * 3)()
    _check(lib().kg_synth_naive_scan_aa(ctx._h, table._h, d_seq, d_off, n, total, out))
    return int(out[0]), int(out[1]), int(out[2])


def hits_checksum(ctx: Context, hits: np.ndarray, d_off: int) -> int:
    hits = np.ascontiguousarray(hits, dtype=HIT_DTYPE)
    out = C.c_uint64()
    _check(lib().kg_synth_hits_checksum(ctx._h, hits.ctypes.data, len(hits), d_off, C.byref(out)))
    return int(out.value)


def device_free(ptr):
    if ptr:
        lib().kg_device_free(ptr)
