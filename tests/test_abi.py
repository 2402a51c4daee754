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
    for h in ("kmerguts.h", "kmerguts_host.h", "kmerguts_synth.h"):
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
    for h in ("kmerguts.h", "kmerguts_host.h", "kmerguts_synth.h"):
        src = open(os.path.join(ROOT, "include", h)).read()
        assert "torch" not in src and "at::" not in src and "std::" not in src


def test_struct_layouts_match_numpy_views():
    assert kg.CALL_DTYPE.itemsize == 32 and kg.OTU_DTYPE.itemsize == 44 and kg.HIT_DTYPE.itemsize == 28
    assert C.sizeof(kg.Params) == 20 and C.sizeof(kg.TableInfo) == 72


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
