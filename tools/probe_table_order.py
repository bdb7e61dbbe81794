import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
def prof(kind):
    ms, cnt = C.c_float(0), C.c_int(0)
    lib.bpk_profile_read(kind, C.byref(ms), C.byref(cnt))
    return round(ms.value, 4)
for lg in [18, 20, 22]:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2, bits=252)
    msm = cbp.Msm(n)
    for v in (0, 1, 2):
        lib.bpk_debug_set_option(12, v)
        for _ in range(3):
            msm(sc, pts)
        torch.cuda.synchronize()
        lib.bpk_profile_reset(); lib.bpk_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            msm(sc, pts)
        e1.record(); torch.cuda.synchronize()
        lib.bpk_profile_enable(0)
        print(json.dumps({"log2_n": lg, "variant": v, "ms": round(e0.elapsed_time(e1) / 10, 4), "front": prof(5), "acc": prof(0), "tail": prof(6), "pre": prof(4)}), flush=True)
lib.bpk_debug_set_option(12, 0)
