"""bench.py contract checks that need no GPU: the reference arm runs here (it times the reference binary,
or the oracle port when oracle/_ref is absent) and prints the JSON line the driver parses."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-500:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["metric"] == "input GB/s (scan mode)" and line["dtype"] == "u8" and line["vs_baseline"] is None
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"] and 0.001 < line["value"] < 10


def test_non_zero_ranks_of_the_reference_arm_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1"],
                       capture_output=True, text=True, timeout=120, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_others_plan_names_existing_workloads_and_labels_every_kernel():
    """The `others` block runs after the headline inside try/except: a wrong workload name would only show up as an
    {"error": ...} entry on the GPU box.  Check the plan and the kernel labels here."""
    sys.path.insert(0, ROOT)
    import bench
    import trre_b200 as T
    for world in (1, 2, 8):
        for label, wl, per_gpu, coll in bench.others_plan(world):
            assert wl in bench.WORKLOADS and wl in bench.DESCR and per_gpu > 0, (world, label, wl)
    for name, spec in bench.WORKLOADS.items():
        info = dict(T.compile(spec[0], spec[1]).info, mode=spec[8] if len(spec) > 8 else None)
        assert isinstance(bench.kernel_label(T, info), str) and bench.kernel_label(T, info), name
    assert bench.kernel_label(T, dict(T.compile("cat:dog", 0).info, mode=None)) == "k_lit"
    assert "k_lines<dfs, -m>" in bench.kernel_label(T, dict(T.compile("cat:dog", 0).info, mode=(1, 0)))
