"""Quick device-timed probe of the MSM at several sizes (CUDA events on the launching stream)."""
import json
import sys

import torch

sys.path.insert(0, ".")
import cudabulletproof_b200 as cbp  # noqa: E402


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [10, 14, 16, 18, 20, 22]
    for log_n in sizes:
        n = 1 << log_n
        pts, _ = cbp.synth_points(n, seed=1)
        sc = cbp.synth_scalars(n, seed=2, bits=252)
        msm = cbp.Msm(n)
        for _ in range(3):
            msm(sc, pts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            msm(sc, pts)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        print(json.dumps({"log_n": log_n, "window_bits": msm.window_bits, "ms": round(ms, 4),
                          "Mpoints_per_s": round(n / ms / 1e3, 2)}), flush=True)


if __name__ == "__main__":
    main()
