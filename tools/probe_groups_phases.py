"""Phase timers (front / accumulate / tail) of the 2^20 MSM for several window groupings.

  python tools/probe_groups_phases.py [log2 n] [acc-stream modes, e.g. 0,1] [groups as hex, e.g. 0,8422,844]
"""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
lg = int(sys.argv[1]) if len(sys.argv) > 1 else 20
n = 1 << lg
pts, _ = cbp.synth_points(n, seed=1)
sc = cbp.synth_scalars(n, seed=2, bits=252)
msm = cbp.Msm(n)
def prof(kind):
    ms, cnt = C.c_float(0), C.c_int(0)
    lib.bpk_profile_read(kind, C.byref(ms), C.byref(cnt))
    return round(ms.value, 4)
acc_modes = [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else [-1]
glist = [int(a, 16) for a in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 0x8422, 0x844, 0x88, 0xC4, 0xF1, 0x6442, 0x4444, 0x22222222]
for acc_mode, g in [(a, g) for g in glist for a in acc_modes]:
    lib.bpk_debug_set_option(4, g)
    lib.bpk_debug_set_option(9, acc_mode)  # BPK_OPT_MSM_ACC_STREAMS
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
    print(json.dumps({"groups": hex(g), "acc_streams": acc_mode, "ms": round(e0.elapsed_time(e1) / 10, 4), "front": prof(5), "acc": prof(0), "tail": prof(6)}), flush=True)
lib.bpk_debug_set_option(4, 0)
lib.bpk_debug_set_option(9, -1)
