"""bench.py's reference arm (`--impl reference`) runs on the host only, so its JSON contract can be checked here:
one line, the keys the driver reads, `impl: reference`, an e2e object with zero copy bytes."""
import json
import os
import subprocess
import sys

from tests.conftest import ROOT


def test_reference_arm_prints_one_contract_line():
    env = dict(os.environ, HGT_BENCH_CPU_SCALE="0.002", OMP_NUM_THREADS="4")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "edges/s" and d["higher_is_better"] is True
    assert d["vs_baseline"] is None and d["value"] > 0 and d["config"]["workload"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert abs(d["e2e"]["value"] - d["value"]) < 1e-6 * d["value"]


def test_edge_algorithmic_bytes_formula():
    sys.path.insert(0, ROOT)
    import bench
    # E*(2*d*4 + 4) + N_dst*(2*d*4 + 4): ~2.24 KB per edge at the ogbn-mag-shaped config
    b = bench.edge_algorithmic_bytes(21_111_007, 1_939_743, 256)
    assert b == 21_111_007 * (2 * 256 * 4 + 4) + 1_939_743 * (2 * 256 * 4 + 4)
    assert 2.2e3 < b / 21_111_007 < 2.3e3
