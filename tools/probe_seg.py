"""MSM time against (window width, segment length)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
sizes = [int(a) for a in sys.argv[1:]] or [14, 16, 17, 18]
for lg in sizes:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2, bits=252)
    for c in (11, 12, 13, 14, 15):
        row = []
        for sh in (-1, 3, 4, 5, 6):
            lib.bpk_debug_set_option(8, sh)
            msm = cbp.Msm(n, window_bits=c)
            ref = None
            for _ in range(3):
                r = msm(sc, pts)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                msm(sc, pts)
            e1.record(); torch.cuda.synchronize()
            row.append(f"s{sh}:{e0.elapsed_time(e1) / 10:.3f}")
        lib.bpk_debug_set_option(8, -1)
        print(f"2^{lg} c={c}: " + "  ".join(row), flush=True)
