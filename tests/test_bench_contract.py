"""bench.py's output contract, checked on CPU through the reference arm (`--impl reference` times the
reference's own functors on the host cores and needs no GPU) and through static inspection of the
own-arm code path."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _oracle_available() -> bool:
    sys.path.insert(0, ROOT)
    from oracle import ref

    return ref.available()


def test_reference_arm_prints_one_json_line_with_the_contract_keys():
    if not _oracle_available():
        pytest.skip("oracle/_ref not built")
    for extra, metric, unit in (
            (["--cpu-stencil-n", "200", "--cpu-stencil-iters", "2"], "stencil_points_per_second",
             "points/s"),
            (["--workload", "black_scholes", "--cpu-sample", "200000"],
             "black_scholes_elements_per_second", "options/s")):
        res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                              "--steps", "3", "--warmup", "1"] + extra,
                             capture_output=True, text=True, timeout=300, cwd=ROOT)
        assert res.returncode == 0, res.stderr[-2000:]
        lines = [ln for ln in res.stdout.splitlines() if ln.strip()]
        assert len(lines) == 1, lines
        r = json.loads(lines[0])
        for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
                    "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
                    "cpu_baseline", "e2e"):
            assert key in r, key
        assert r["impl"] == "reference" and r["metric"] == metric and r["unit"] == unit
        assert r["steps"] == 3 and r["warmup"] == 1          # --steps / --warmup are honoured
        assert r["higher_is_better"] is True and r["vs_baseline"] is None
        assert r["value"] > 0 and r["cpu_baseline"]["kind"] == "reference"
        assert r["cpu_baseline"]["value"] == r["value"] == r["e2e"]["value"]
        assert r["e2e"]["h2d_bytes_per_step"] == 0 and r["e2e"]["d2h_bytes_per_step"] == 0
        assert "workload" in r["config"] and "model" not in r["config"]
        assert "sample" in r["config"]["workload"]           # the sampled size is stated there


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference",
                          "--gpus", "2", "--steps", "1", "--warmup", "1"],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_own_arm_emits_every_contract_key():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"',
                '"higher_is_better"', '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"',
                '"roofline"', '"cpu_baseline"', '"e2e"', '"gpu_launches"', '"clocks"',
                '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"bound"', '"achieved"', '"peak"',
                '"frac"', '"traffic"'):
        assert key in src, key
