"""The driver-facing contract of bench.py that can be checked without a GPU: exactly one JSON line on stdout whatever the
libraries print, the reference arm's "unavailable" line (also when launched with a rank environment), and the keys the record
must carry -- checked against the committed 1-GPU record, which is what a GPU run of the same script wrote."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, capture_output=True, text=True, timeout=300, env=e, cwd=ROOT)


def test_reference_arm_prints_one_unavailable_line_from_rank_zero_only():
    r = _run(["--impl", "reference", "--gpus", "1"])
    assert r.returncode == 0
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and "unavailable" in d and "\n" not in d["unavailable"]
    r1 = _run(["--impl", "reference", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert r1.returncode == 0 and r1.stdout.strip() == ""


def test_without_a_gpu_the_own_arm_says_so_in_one_json_line():
    import torch
    if torch.cuda.is_available():
        return                                   # on a GPU box this is the real bench: covered by the driver, not by a unit test
    r = _run([], env={"NCCL_DEBUG": "VERSION"})
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["value"] is None and "unavailable" in d


def test_committed_record_carries_every_contract_key():
    d = json.load(open(os.path.join(ROOT, "profiles", "bench_1gpu.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data",
              "config", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["steps"] >= 1 and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(d["e2e"])
    assert d["e2e"]["h2d_bytes_per_step"] > 0 and d["e2e"]["d2h_bytes_per_step"] > 0 and d["e2e"]["value"] < d["value"]
    assert {"sm_mhz", "sm_max_mhz", "reasons"} <= set(d["clocks"]) and not any("slowdown" in r for r in d["clocks"]["reasons"])
    assert d["gpu_launches"] == d["steps"] and d["verified"] is True
    assert d["config"]["timed_region_s"] >= 1.0 and d["config"]["msg_bytes"] > 126 << 20        # inputs larger than L2
    assert abs(d["ms_per_step"] * d["steps"] / 1e3 - d["config"]["timed_region_s"]) < 0.05
    assert {"cudaMemcpyAsync_d2d_gbps", "host_posted_gbps", "host_staged_gbps"} <= set(d["baselines"])
