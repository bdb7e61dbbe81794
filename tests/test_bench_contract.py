"""The JSON line bench.py prints is a contract with the driver.  These tests check the committed artefacts of the
last GPU run (profiles/) for every key the contract names, so a refactor of bench.py that drops one shows up on the
CPU box; bench.py itself is parsed for syntax and for the flags the driver passes."""
import ast
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def last_json_line(path):
    with open(path) as f:
        lines = [ln for ln in f.read().strip().splitlines() if ln.startswith("{")]
    return json.loads(lines[-1])


def check_common(d):
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e"):
        assert key in d, key
    assert d["metric"] == "ed25519_msm_points_per_sec" and d["unit"] == "points/s"
    assert d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["data"] == "synthetic" and "workload" in d["config"]
    assert "model" not in d["config"]
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in d["e2e"], key


def test_own_arm_artefact_has_every_contract_key():
    d = last_json_line(os.path.join(ROOT, "profiles", "r02_bench_1gpu.json"))
    check_common(d)
    assert d["n_gpus"] == 1 and d["warmup"] >= 3 and d["value"] > 0
    assert d["gpu_launches"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 160 * (1 << 20) and d["e2e"]["d2h_bytes_per_step"] == 128
    assert 0 < d["e2e"]["value"] < d["value"]  # host buffers + copies can only be slower than device-resident
    for key in ("sm_mhz", "sm_max_mhz", "reasons"):
        assert key in d["clocks"], key
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
    r = d["roofline"]
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in r, key
    assert abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-9 and 0 < r["frac"] < 1
    c = d["cpu_baseline"]
    for key in ("value", "unit", "cores", "kind", "sample"):
        assert key in c, key
    assert c["kind"] in ("reference", "port") and c["cores"] >= 1
    assert "measured" in r["peak_source"] or "bpk_measure_int_peak" in r["peak_source"]  # the peak comes from the same run
    assert r["traffic"] and r["ncu"]["source_hash"]  # parsed from a committed ncu summary of the same kernel sources
    assert 0.95 * r["peak"] < r["peak_all_T_per_s"]["imad_wide_carry"] < 1.05 * r["peak"]
    s = d["secondary"]
    assert s["metric"] == "range_proof_verifies_per_sec" and s["decisions_correct"] is True
    assert 0 < s["roofline"]["frac"] < 1
    assert s["one_by_one"]["same_decisions"] is True and s["algorithm"]["grouped"] >= 2
    # pageable and affine rows of the end-to-end call, every result equal to the device-resident one
    assert d["e2e"]["result_matches_device_path"] is True
    for row in list(d["e2e"]["pageable"].values()) + [d["e2e"]["affine_extension"]]:
        assert row["result_matches_device_path"] is True and row["value"] > 0
    # strong scaling: ONE global MSM per row, checked against the scalar identity
    for row in d["secondary_scaling"]["rows"]:
        assert row["result_check"] is True and len(row["result_xy"]) == 128


def test_reference_arm_artefact():
    d = last_json_line(os.path.join(ROOT, "profiles", "r02_bench_reference_arm.json"))
    check_common(d)
    assert d["impl"] == "reference"
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["e2e"]["value"] == d["value"]
    assert d["cpu_baseline"]["kind"] == "reference" and d["cpu_baseline"]["value"] == d["value"]


def test_scaling_artefacts_are_weak_scaling_lines():
    single = last_json_line(os.path.join(ROOT, "profiles", "r02_bench_1gpu.json"))
    one = single["value"]
    xy = {r["log2_n"]: r["result_xy"] for r in single["secondary_scaling"]["rows"]}
    for n in (2, 4, 8):
        d = last_json_line(os.path.join(ROOT, "profiles", f"r02_bench_{n}gpu.json"))
        check_common(d)
        assert d["n_gpus"] == n
        assert 0.85 * n * one < d["value"] < 1.1 * n * one  # whole-job aggregate, near-linear
        for r in d["secondary_scaling"]["rows"]:  # the same global MSM on every N: identical bytes, checked result
            assert r["result_check"] is True and r["result_xy"] == xy[r["log2_n"]]
            assert r["n_gpus"] == n


def test_bench_py_accepts_the_driver_flags():
    src = open(os.path.join(ROOT, "bench.py")).read()
    ast.parse(src)
    for flag in ("--gpus", "--steps", "--warmup", "--impl"):
        assert f'"{flag}"' in src, flag
