"""bench.py's reference arm runs without a GPU and prints the contract's JSON line (metric, unit, config, cpu_baseline, e2e)."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def test_reference_arm_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "c1", "--steps", "1",
                          "--warmup", "0", "--particles", "200000"], capture_output=True, text=True, timeout=600,
                         env={"OMP_NUM_THREADS": "1", "PATH": "/usr/bin:/bin:/usr/local/bin", "CUDA_VISIBLE_DEVICES": ""})
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "Gparticles/s splatted" and line["unit"] == "Gparticles/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": line["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]


def test_non_zero_ranks_of_the_reference_arm_do_nothing():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2"], capture_output=True,
                         text=True, timeout=120, env={"RANK": "1", "WORLD_SIZE": "2", "PATH": "/usr/bin:/bin:/usr/local/bin"})
    assert out.returncode == 0 and out.stdout.strip() == ""
