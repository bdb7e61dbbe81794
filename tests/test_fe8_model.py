"""CPU: the lane-level model the octet-form device arithmetic (csrc/fe8.cuh) was written from — every register-width
mask and every bound its comments claim is asserted on random and edge inputs against Python integers."""
import importlib.util
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fe8_lane_model_selftest():
    spec = importlib.util.spec_from_file_location("fe8_model", os.path.join(ROOT, "tools", "fe8_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.selftest(rounds=1500, seed=7)
