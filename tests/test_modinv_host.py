"""CPU: csrc/modinv.cuh (Bernstein-Yang divsteps inversion, the source the device compiles) built for the host with
g++ and pinned against Python big integers mod p and mod l; plus the limb-level model it was written from."""
import importlib.util
import os
import random
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**255 - 19
L = 2**252 + 27742317777372353535851937790883648493


def test_modinv_model_selftest():
    spec = importlib.util.spec_from_file_location("modinv_model", os.path.join(ROOT, "tools", "modinv_model.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    assert m.selftest(rounds=150, seed=11) <= 21


def test_modinv_host_build_matches_python(tmp_path):
    exe = str(tmp_path / "modinv_host")
    subprocess.run(["g++", "-O2", "-std=c++17", "-Wall", os.path.join(ROOT, "tests", "c_abi", "modinv_host.cpp"), "-o", exe],
                   check=True, capture_output=True, text=True)
    rng = random.Random(5)
    cases = []
    for m in (P, L):
        edge = [0, 1, 2, 3, m - 1, m - 2, (m + 1) // 2, m, m + 1, 2 * m, 2**255, 2**256 - 1, 2**128, 2**64 - 1, 2**62, 2**62 - 1,
                2**124, 19, 38]
        cases += [(m, x) for x in edge] + [(m, rng.getrandbits(256)) for _ in range(600)]
    text = "".join(f"{m:064x} {x:064x}\n" for m, x in cases)
    out = subprocess.run([exe], input=text, capture_output=True, text=True, check=True).stdout.split()
    assert len(out) == len(cases)
    for (m, x), got in zip(cases, out):
        want = pow(x % m, -1, m) if x % m else 0
        assert int(got, 16) == want, (hex(m), hex(x))
