"""bench.py contract pieces that can be checked without a GPU: both arms describe the same workload with the same `config` dict, and the
reference (CPU) arm runs end to end and prints the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _args(**kw):
    import argparse
    base = dict(batch=32, ctx=4096, steps=20, warmup=5, parity_steps=3, config="q4k", kv=None, gpus=1)
    base.update(kw)
    return argparse.Namespace(**base)


def test_both_arms_print_the_same_config_and_window():
    import bench
    a = _args()
    first, last = bench.timed_window(a)
    assert (first, last) == (4096 + 3 + 1 + 5 + 1, 4096 + 3 + 1 + 5 + 20)          # ctx of the 20 timed steps, incl. the decoded token
    c1 = bench.workload_config(a, 1)
    assert c1 == bench.workload_config(_args(), 1) and f"ctx {first}->{last}" in c1["workload"]
    assert bench.workload_config(_args(gpus=8), 8)["parallelism"] == "tp8"
    assert "fp8 paged KV" in bench.workload_config(_args(config="gptq_fp8kv"), 1)["workload"]


def test_reference_arm_runs_on_the_host_and_prints_the_contract_keys():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "tokens/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and "nothing scaled" in line["cpu_baseline"]["sample"]
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    import bench
    assert line["config"] == bench.workload_config(_args(steps=1, warmup=0), 1)
    # one complete decode step was run: the claimed step time is real wall time, not an extrapolation
    assert line["steps"] == 1 and line["ms_per_step"] > 100
