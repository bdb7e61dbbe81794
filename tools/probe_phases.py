"""Phase timers (library CUDA events: front / accumulate / tail / total) of the device-resident MSM per size.

  python tools/probe_phases.py 10 14 17 20        # one JSON line per log2 n
  python tools/probe_phases.py --once 14          # a single warm MSM of that size (for an ncu launch list)
"""
import ctypes as C
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp  # noqa: E402


def prof(lib, kind):
    ms, cnt = C.c_float(0), C.c_int(0)
    lib.bpk_profile_read(kind, C.byref(ms), C.byref(cnt))
    return round(ms.value, 4)


def main():
    lib = cbp.load()
    args = sys.argv[1:]
    once = "--once" in args
    normalize = "--no-normalize" not in args
    sizes = [int(a) for a in args if not a.startswith("--")] or [10, 12, 14, 16, 17, 18, 20]
    for lg in sizes:
        n = 1 << lg if lg <= 30 else lg  # values above 30 are taken as n itself (147 = one 64-bit verification)
        pts, _ = cbp.synth_points(n, seed=0xC3 + lg)
        sc = cbp.synth_scalars(n, seed=0x5CA1A000 + lg, bits=252)
        msm = cbp.Msm(n)
        for _ in range(1 if once else 3):
            msm(sc, pts, normalize=normalize)
        torch.cuda.synchronize()
        if once:
            continue
        lib.bpk_profile_reset()
        lib.bpk_profile_enable(1)
        reps = 10
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = lib.bpk_kernel_launches()
        e0.record()
        for _ in range(reps):
            msm(sc, pts, normalize=normalize)
        e1.record()
        torch.cuda.synchronize()
        lib.bpk_profile_enable(0)
        launches = (lib.bpk_kernel_launches() - l0) // reps
        # the profile events themselves cost a few microseconds each; "ms" is the honest per-call time
        print(json.dumps({"log2_n": lg, "window_bits": msm.window_bits, "ms": round(e0.elapsed_time(e1) / reps, 4),
                          "front_ms": prof(lib, 5), "accumulate_ms": prof(lib, 0), "tail_ms": prof(lib, 6),
                          "total_ms": prof(lib, 1), "launches": int(launches), "normalize": normalize}), flush=True)


if __name__ == "__main__":
    main()
