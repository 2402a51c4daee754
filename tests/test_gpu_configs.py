"""The per-config scripts under tests/configs/ (BASELINE.json configs[2..4]) at a reduced size: each one generates its
workload on the device, runs it through the C ABI, checks it against the CPU oracle (or its size-independent properties)
and prints one JSON line.  The full-size runs are the ones recorded under profiles/."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(script, *args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "configs", script), *args], cwd=ROOT, capture_output=True, text=True,
                       timeout=600)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    return json.loads(line)


def test_config2_six_frame_contigs():
    d = _run("config2_dna.py", "--genomes", "4", "--length", "400000", "--families", "20000", "--sigs", "2000000", "--steps", "2",
             "--parity-genomes", "4")
    assert "on 4 of 4 genomes" in d["parity"] and d["parity"].startswith("bit-exact") and d["calls"] > 100 and d["e2e"]["ms_per_step"] > 0


def test_config3_orfs_in_batches():
    d = _run("config3_orfs.py", "--orfs", "300000", "--batch", "100000", "--families", "20000", "--sigs", "2000000", "--parity", "3000")
    assert d["batches_per_rank"] == 3 and d["parity"].startswith("bit-exact") and d["calls"] > 1000


def test_config4_sharded_single_rank():
    d = _run("config4_sharded.py", "--families", "20000", "--proteins", "50000", "--steps", "2", "--sample", "500")
    assert "parity_hits" in d and "parity_calls" in d and d["lookups_per_s"] > 0


def test_config4_sharded_virtual_ranks():
    d = _run("config4_sharded.py", "--families", "20000", "--proteins", "20000", "--steps", "1", "--warmup", "1", "--local-ranks", "3")
    assert d["rank0_phase_ms"]["total"] > 0
