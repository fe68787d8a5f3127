"""bench.py's reference arm (the CPU path the driver times next to the GPU arm) runs here without a GPU: the JSON line
must carry the contract's keys.  cfg1 (BASELINE.json configs[0], the reference's own CPU-runnable case) keeps it short."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_json_contract():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "voxels/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["ms_per_step"] > 0 and line["n_gpus"] == 1 and line["steps"] == 1
    assert line["e2e"] == {"value": line["value"], "unit": "voxels/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "oracle" in cb["sample"]
    assert line["config"]["workload"] == "cfg1" and line["config"]["shape"] == [1, 1, 128, 128, 128]
    assert line["gpu_launches"] == 0 and line["vs_baseline"] is None and line["dtype"] == "f32"
    # both arms print the same config dict (the driver compares them)
    sys.path.insert(0, ROOT)
    import argparse

    import bench
    ours = bench.config_dict(argparse.Namespace(workload="cfg1", sw_batch=8), bench.WORKLOADS["cfg1"])
    assert ours == line["config"] and line["config"]["windows"] == 8


def test_reference_arm_other_ranks_exit_quietly():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "cfg1",
                          "--gpus", "2", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=600,
                         cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
