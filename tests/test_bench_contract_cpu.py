"""bench.py contract checks that run without a GPU: the reference arm (oracle port on host cores) prints ONE JSON line
with the keys the driver reads, and the committed round profile carries the same contract for the GPU arm."""
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
BASE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
             "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches"}


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--cpu-seconds", "2", "--events", "20000", "--batch", "2"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert BASE_KEYS <= set(j) and j["impl"] == "reference" and j["higher_is_better"] is True
    assert j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["value"] == j["value"] > 0
    assert j["e2e"]["h2d_bytes_per_step"] == 0 and j["e2e"]["d2h_bytes_per_step"] == 0
    assert j["gpu_launches"] == 0 and "workload" in j["config"]


def test_committed_gpu_bench_line_has_the_contract_keys():
    j = json.loads((ROOT / "profiles" / "r02_bench_n1.json").read_text())
    assert BASE_KEYS | {"clocks", "roofline", "cpu_baseline"} <= set(j)
    r = j["roofline"]
    assert {"bound", "achieved", "peak", "unit", "frac", "traffic"} <= set(r) and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9
    assert j["e2e"]["h2d_bytes_per_step"] > 0 and j["e2e"]["d2h_bytes_per_step"] > 0 and j["gpu_launches"] > 0
    assert not set(j["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    # the dominant kernel is the real maximum of the per-op times, and every kernel >= 10 % of the step has an entry
    per_op = r["per_op_ms"]
    assert r["op"] == max(per_op, key=per_op.get) == r["kernels"][0]["op"]
    tot = sum(per_op.values())
    assert {e["op"] for e in r["kernels"]} >= {k for k, v in per_op.items() if v >= 0.10 * tot and k in ("l1_build", "l1_conv_b_pool_voxel")}
    for e in r["kernels"]:
        assert abs(e["frac"] - e["achieved"] / e["peak"]) < 1e-9 and e["traffic"] is not None
    # the rest of BASELINE.json's story is in the default line
    assert j["sustained"]["seconds"] >= 1.9 and j["clustered"]["value"] > 0
    ifr = j["interframe_latency_ms"]
    assert len(ifr["batch8"]["steps"]) == 10 and ifr["batch8"]["steps"][0]["events"] == 0 and len(ifr["batch1"]["steps"]) == 10
    st = j["streaming"]
    assert st["chunk_us"] == 1000 and st["stream_seconds"] >= 1.99 and st["latency_ms"]["p99"] > 0 and not st["overflow"]
