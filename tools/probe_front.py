"""MSM time with the fused (one cooperative launch) and the separate front-end kernels (BPK_OPT_MSM_FUSED_FRONT)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
def prof(kind):
    ms, cnt = C.c_float(0), C.c_int(0)
    lib.bpk_profile_read(kind, C.byref(ms), C.byref(cnt))
    return round(ms.value, 4)
for lg in [int(a) for a in sys.argv[1:]] or [14, 16, 17, 18, 19, 20, 22]:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2, bits=252)
    msm = cbp.Msm(n)
    row = {"log2_n": lg}
    ref = None
    for fused in (0, 1):
        lib.bpk_debug_set_option(14, fused)
        for _ in range(3):
            out = msm(sc, pts)
        torch.cuda.synchronize()
        r = out.cpu().numpy().tobytes()
        ref = ref or r
        assert r == ref
        lib.bpk_profile_reset(); lib.bpk_profile_enable(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            msm(sc, pts)
        e1.record(); torch.cuda.synchronize()
        lib.bpk_profile_enable(0)
        row["fused" if fused else "separate"] = {"ms": round(e0.elapsed_time(e1) / 10, 4), "front": prof(5)}
    print(json.dumps(row), flush=True)
lib.bpk_debug_set_option(14, 1)
