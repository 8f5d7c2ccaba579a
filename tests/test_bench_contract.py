"""CPU: the bench contract's reference arm runs without a GPU and prints one well-formed JSON line."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "voxels/s" and d["value"] > 0
    # "reference" = the reference's own modules from oracle/_ref (build container / GPU box snapshot), "port" = oracle only
    have_ref = os.path.isfile(os.path.join(ROOT, "oracle", "_ref", "models", "vit_3d_2d_pretrain.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if have_ref else "port") and d["cpu_baseline"]["cores"] >= 1
    assert "batch 64" in d["cpu_baseline"]["sample"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]


def test_default_workload_is_documented():
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.DEFAULT_CONFIG in bench.CONFIGS
    assert set(bench.CONFIGS) == {"cfg2", "cfg3", "cfg4", "cfg5"}
    for cfg in bench.CONFIGS.values():
        x, y = bench.synthetic_batch(cfg, 2, seed=1)
        assert x.shape[0] == 2 and y.shape[0] == 2


def test_nonzero_rank_reference_arm_is_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and "{" not in r.stdout
