"""bench.py contract pieces that run without a GPU: the reference arm and the CPU-port baseline."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "1", "--ref-samples", "24"], capture_output=True, text=True, timeout=600)
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1, out.stderr[-500:]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "saa_linearize_assemble_throughput"
    assert d["unit"] == "samples*steps/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["config"]["workload"].startswith("drone SAA linearize+assemble")


def test_cpu_port_baseline_dict():
    sys.path.insert(0, ROOT)
    import bench
    d = bench.cpu_port_baseline(M_s=2000, budget_s=0.5)
    assert d["kind"] == "port" and d["cores"] >= 1 and d["value"] > 0 and "sample" in d
