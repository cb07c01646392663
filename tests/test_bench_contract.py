"""The bench line contract, checked on the committed round profiles (profiles/r01_bench_*.json are verbatim `bench.py`
output from the B200): every key the driver reads is present and self-consistent.  No GPU needed."""
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _line(name):
    with open(os.path.join(ROOT, "profiles", name)) as f:
        return json.loads(f.read())


@pytest.mark.parametrize("name,n", [("r01_bench_n1.json", 1), ("r01_bench_n2.json", 2), ("r01_bench_n8.json", 8)])
def test_bench_line_has_the_contract_keys(name, n):
    d = _line(name)
    with open(os.path.join(ROOT, "BASELINE.json")) as f:
        base = json.load(f)
    assert d["metric"].split(" at ")[0] in base["metric"] and d["unit"] == "proofs/s"
    assert d["n_gpus"] == n and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["warmup"] >= 3 and d["steps"] >= 1 and d["data"] == "synthetic" and "workload" in d["config"]
    assert abs(d["value"] - n * d["steps"] * d["config"]["batch_per_gpu"] / (d["ms_per_step"] * d["steps"] / 1000.0)) < 1e-6 * d["value"]
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["value"] != d["value"]
    assert d["gpu_launches"] > 0
    c = d["clocks"]
    assert c["sm_mhz"] > 0.9 * c["sm_max_mhz"] and not set(c["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and r["traffic"] > 0 and r["launches"] > 0
    if n == 1:
        cb = d["cpu_baseline"]
        assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] > 0 and cb["proof_bytes_equal_gpu"] is True
        assert d["e2e_from_pass_uris"]["same_public_signals_as_e2e"] is True and d["verify"]["all_valid"] is True


def test_reference_arm_line():
    d = _line("r01_bench_reference_arm.json")
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["value"] == d["value"] and d["cpu_baseline"]["kind"] in ("port", "reference")
