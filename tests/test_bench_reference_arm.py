"""bench.py --impl reference (the CPU arm the driver times beside ours): prints one JSON line with
the contract's keys and maps the oracle and the scan generator only - none of the product's
shared objects (libformgpu.so / libformhost.so) may be in its address space.  CPU only."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line_and_mapped_libraries():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--sensor", "vlp-16",
                          "--steps", "2", "--warmup", "3", "--preroll", "0"], cwd=ROOT, capture_output=True,
                         text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "scans/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["e2e"]["value"] == line["value"]
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["d2h_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    rp = line["reference_code_pipeline"]  # FORM's own register_scan next to the host logic over the restatement
    if rp is not None:                    # (None only where oracle/_ref is not built)
        assert "unavailable" not in rp, rp
        assert rp["keypoints_identical"] is True and rp["max_pose_difference"] < 1e-9
    mapped = line["native_libs_mapped"]
    assert any(p.endswith("liboracle.so") for p in mapped)
    assert not any("libformgpu" in p or "libformhost" in p or "_core" in p for p in mapped), mapped


def test_other_ranks_of_the_reference_arm_exit_without_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip() == ""
