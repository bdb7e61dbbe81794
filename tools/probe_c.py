"""device MSM time against the window width c for mid-size inputs"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
n = 1 << 24
pts, _ = cbp.synth_points(n, seed=1)
sc = cbp.synth_scalars(n, seed=2, bits=253)
for lg in (19, 20, 21, 22, 24):
    m = 1 << lg
    row = []
    for c in (0, 15, 16, 17, 18):
        msm = cbp.Msm(m, window_bits=c) if c else cbp.Msm(m)
        for _ in range(3):
            msm(sc[:m], pts[:m])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            msm(sc[:m], pts[:m])
        e1.record(); torch.cuda.synchronize()
        row.append(f"c={msm.window_bits if not c else c}:{e0.elapsed_time(e1)/10:.3f}")
    print(f"2^{lg}: " + "  ".join(row), flush=True)
