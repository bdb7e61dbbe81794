"""device MSM time against the window width c: python tools/probe_c.py [log_n ...] (auto choice first)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
sizes = [int(a) for a in sys.argv[1:]] or [12, 13, 14, 15, 16, 17, 18, 19, 20, 21, 22]
n = 1 << max(sizes)
pts, _ = cbp.synth_points(n, seed=1)
sc = cbp.synth_scalars(n, seed=2, bits=253)
for lg in sizes:
    m = 1 << lg
    row = []
    auto = cbp.Msm(m).window_bits
    for c in [0] + [c for c in range(min(9, max(4, auto - 3)), 18)]:
        msm = cbp.Msm(m, window_bits=c) if c else cbp.Msm(m)
        for _ in range(3):
            msm(sc[:m], pts[:m])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if lg <= 20 else 4
        e0.record()
        for _ in range(reps):
            msm(sc[:m], pts[:m])
        e1.record(); torch.cuda.synchronize()
        row.append(f"{'auto ' + str(msm.window_bits) if not c else c}:{e0.elapsed_time(e1)/reps:.3f}")
    print(f"2^{lg}: " + "  ".join(row), flush=True)
