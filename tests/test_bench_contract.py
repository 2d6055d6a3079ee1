"""bench.py contract pieces that run without a GPU: the FLOP model equals BASELINE.md's numbers and the CPU reference
sample (what `--impl reference` / `cpu_baseline` time) runs and reports sane values."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_flop_model_matches_baseline_md():
    sys.path.insert(0, ROOT)
    import bench

    f_step, f_lin, f_attn, f_lora = bench.flux_flops()
    assert abs(f_step / 1e12 - 172.4) < 0.1 and abs(f_lin / 1e12 - 59.50) < 0.05
    assert abs(f_attn / 1e12 - 14.87) < 0.02 and abs(f_lora / 1e12 - 0.448) < 0.002
    assert abs(bench.flux_flops(4, 64)[0] / 1e12 - 705.8) < 0.3  # rank-sweep row of BASELINE.md section 3
    w = bench.wan_flops()  # F_lin 33.44, F_attn 33.92 -> 187.1 TFLOP
    assert abs(w[0] / 1e12 - 187.1) < 0.1 and abs(w[1] / 1e12 - 33.44) < 0.01 and abs(w[2] / 1e12 - 33.92) < 0.01
    assert 1 <= bench.host_cores() <= (os.cpu_count() or 1)


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, OMP_NUM_THREADS="8")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype",
              "data", "config", "impl", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["cpu_baseline"]["kind"] == "port" and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and "workload" in d["config"]
