"""bench.py contract checks that need no GPU: the reference arm runs the compiled reference (oracle/_ref) on the host cores
and prints one JSON line with the agreed keys; our arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_harness")


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_harness not built (run __graft_entry__.build())")
def test_reference_arm_json_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "phase-space cell-updates/s per RK stage" and d["unit"] == "cell-updates/s"
    assert d["higher_is_better"] is True and d["value"] > 1e5 and d["dtype"] == "f64"
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample of config 3" in cb["sample"]
    assert d["gpu_launches"] == 0 and d["steps"] == 1 and d["warmup"] == 1


def test_our_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1"], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                       text=True, timeout=600)
    assert r.returncode != 0 and "no CPU fallback" in (r.stderr + r.stdout)


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_harness not built (run __graft_entry__.build())")
def test_reference_arm_under_torchrun_prints_once():
    """launched as the driver launches N > 1: rank 0 alone runs and prints the line, the other ranks exit 0 without work"""
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29541", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "0"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["scaling"] == "strong" and "sample of config 5" in d["config"]["workload"]


@pytest.mark.skipif(not os.path.exists(REF), reason="oracle/_ref/ref_harness not built (run __graft_entry__.build())")
def test_reference_arm_on_an_amr_workload():
    """--workload c4 (BASELINE.json configs[3], 3 levels): the reference arm runs that very configuration (same_config)"""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c4", "--steps", "2", "--warmup", "1"],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["value"] > 1e5 and "3 levels" in d["config"]["workload"]
    assert d["cpu_baseline"]["same_config"] is True and d["cpu_baseline"]["kind"] == "reference"


def test_committed_kernel_counts_and_traffic_describe_the_shipped_kernel():
    """The fp64 roof of the bench line uses profiles/fused_sass_counts.json and `traffic` uses profiles/fused_traffic.json: both must
    belong to the kernel version bench.py names, and the instruction counts must be those of the object that is actually built."""
    import re
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import sass_count
    version = re.search(r'FUSED_KERNEL_VERSION = "(\w+)"', open(os.path.join(ROOT, "bench.py")).read()).group(1)
    counts = json.load(open(os.path.join(ROOT, "profiles", "fused_sass_counts.json")))
    traffic = json.load(open(os.path.join(ROOT, "profiles", "fused_traffic.json")))
    assert counts["kernel_version"] == version and traffic["kernel_version"] == version
    assert set(traffic["stages"]) >= {"0", "3", "5"}
    obj = os.path.join(ROOT, "veritas_b200", "build", "vrt_fused.cu.o")
    if not os.path.exists(obj):
        pytest.skip("object not built")
    for name, body in sass_count.functions(obj):
        m = re.search(r"k_fused_stageILi(\d)ELi(\d)ELi128E", name)
        if m:
            r = sass_count.analyse(body, int(m.group(2)))
            c = counts["stages"][f"S{m.group(1)}"]
            assert abs(r["fp64_per_column"] - c["fp64_per_column"]) < 0.01 and abs(r["instr_per_column"] - c["instr_per_column"]) < 0.01, name
