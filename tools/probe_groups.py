"""MSM time against the window grouping of the pipeline (BPK_OPT_MSM_GROUPS, hex digits top group first)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
sizes = [int(a) for a in sys.argv[1:]] or [17, 18, 20]
for lg in sizes:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2, bits=252)
    msm = cbp.Msm(n)
    W = (256 + msm.window_bits - 1) // msm.window_bits
    cands = [0, 0x8422, 0x844, 0x88, 0x862, 0xA42, 0xA6, 0xC4, 0x6442, 0x664] if W == 16 else [0, 0x9432, 0x963, 0x99, 0x6642, 0x666, 0xA44, 0xC6]
    row = []
    for g in cands:
        lib.bpk_debug_set_option(4, g)
        for _ in range(3):
            msm(sc, pts)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            msm(sc, pts)
        e1.record(); torch.cuda.synchronize()
        row.append(f"{g:x}:{e0.elapsed_time(e1) / 10:.3f}")
    lib.bpk_debug_set_option(4, 0)
    print(f"2^{lg} (c={msm.window_bits}, W={W}): " + "  ".join(row), flush=True)
