import sys, torch
sys.path.insert(0, ".")
import cudabulletproof_b200 as cbp
n = 1 << 17
pts, _ = cbp.synth_points(n, seed=1)
sc = cbp.synth_scalars(n, seed=2, bits=252)
for c in (13, 15):
    msm = cbp.Msm(n, window_bits=c)
    msm(sc, pts); msm(sc, pts)
    torch.cuda.synchronize()
