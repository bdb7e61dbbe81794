"""BASELINE config 2: device-resident MSM sweep 2^10 .. 2^22 (+ 2^24), one JSON line per size"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
top = 24
pts, _ = cbp.synth_points(1 << top, seed=0xC3)
sc = cbp.synth_scalars(1 << top, seed=0x5CA1A000, bits=253)
for lg in list(range(10, 23)) + [24]:
    m = 1 << lg
    msm = cbp.Msm(m)
    for _ in range(3):
        msm(sc[:m], pts[:m])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if lg <= 20 else 5
    e0.record()
    for _ in range(reps):
        msm(sc[:m], pts[:m])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    W = (256 + msm.window_bits - 1) // msm.window_bits
    print(json.dumps({"log2_n": lg, "window_bits": msm.window_bits, "ms": round(ms, 4), "points_per_s": round(m / ms * 1e3),
                      "frac_int_roofline_whole_msm": round(m * W * 504 / (ms * 1e-3) / 9.0e12, 3)}), flush=True)
