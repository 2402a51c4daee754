"""CPU-only checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol the headers
under include/ declare, and refuses to run (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import kmergutsjava_b200 as kg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in ("kmerguts.h", "kmerguts_host.h", "kmerguts_shard.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b(kg_[a-z0-9_]+)\s*\(", src))
    return names


def test_library_exports_every_declared_symbol():
    L = kg.lib()
    declared = _declared()
    assert declared == set(kg.EXPORTS), declared ^ set(kg.EXPORTS)
    out = subprocess.run(["nm", "-D", "--defined-only", kg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert declared <= exported, declared - exported
    for name in declared:
        assert getattr(L, name) is not None


def test_no_torch_types_in_signatures():
    for h in ("kmerguts.h", "kmerguts_host.h", "kmerguts_shard.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        assert "torch" not in src and "at::" not in src and "std::" not in src


def test_product_library_carries_only_the_path():
    """Generators, the naive cross-check scan and the roofline microbenchmark live in tools/benchlib/libkmerguts_bench.so;
    the product library exports nothing of them."""
    out = subprocess.run(["nm", "-D", "--defined-only", kg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l and l.split()[-1].startswith("kg_")}
    assert not [n for n in exported if "synth" in n or "roofline" in n or "naive" in n], exported
    assert exported == _declared(), exported ^ _declared()
    from tools import kg_benchlib as bl
    bl.lib()  # the bench library loads (no GPU needed) and resolves against the product library


def test_struct_layouts_match_numpy_views():
    assert kg.CALL_DTYPE.itemsize == 32 and kg.OTU_DTYPE.itemsize == 44 and kg.HIT_DTYPE.itemsize == 28
    assert C.sizeof(kg.Params) == 20 and C.sizeof(kg.TableInfo) == 72


def test_shard_owner_partitions_keys():
    """kg_shard_owner (include/kmerguts_shard.h): every key has exactly one owner, the split is even, and it is not
    correlated with the bucket hash (a python restatement of the two murmur mixes)."""
    M = (1 << 64) - 1

    def mix(k):
        k ^= k >> 33
        k = (k * 0xff51afd7ed558ccd) & M
        k ^= k >> 33
        k = (k * 0xc4ceb9fe1a85ec53) & M
        return k ^ (k >> 33)

    rng = np.random.default_rng(5)
    keys = [int(k) for k in rng.integers(0, 20 ** 8, 4000)] + [0, 1, 20 ** 8 - 1]
    for R in (1, 2, 3, 8, 16):
        own = [kg.shard_owner(k, R) for k in keys]
        assert own == [((mix(k ^ 0x5851F42D4C957F2D) >> 32) * R) >> 32 for k in keys]
        cnt = np.bincount(own, minlength=R)
        assert cnt.min() > 0.8 * len(keys) / R and cnt.max() < 1.2 * len(keys) / R
    assert C.sizeof(kg.ShardStats) == 6 * 8 + 6 * 4 + 8  # + chunks, padded to 8


def test_defaults_match_reference():
    p = kg.default_params()   # KGJ:102-107
    assert (p.min_hits, p.min_weighted_hits, p.max_gap, p.order_constraint, p.emit_hits) == (5, 0, 200, 0, 0)


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(kg.KgError) as e:
        kg.Context(0)
    assert e.value.code == -2 and "no CPU fallback" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py may use oracle/."""
    pkg = os.path.join(ROOT, "kmergutsjava_b200")
    for dp, _, files in os.walk(pkg):
        if os.sep + "build" in dp:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".java", "Makefile")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in src.lower() or f == "__init__.py" and "oracle" not in src.replace("# oracle", ""), (dp, f)
    # tools/ is neutral tooling: nothing there imports, runs or links the oracle either
    for dp, _, files in os.walk(os.path.join(ROOT, "tools")):
        for f in files:
            if f.endswith((".py", ".sh", ".cu")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"(from|import)\s+oracle|oracle/build|kgo\.", src), (dp, f)
    so = subprocess.run(["ldd", kg.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in so
