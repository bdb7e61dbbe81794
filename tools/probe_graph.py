"""Mid-size device MSM as plain launches and as the cached CUDA graph of the same launch DAG (BPK_OPT_MSM_GRAPH)."""
import ctypes as C, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
for lg in [int(a) for a in sys.argv[1:]] or [14, 15, 16, 17, 18]:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2, bits=252)
    msm = cbp.Msm(n)
    row = {"log2_n": lg}
    ref = None
    for graph in (0, 1, 0, 1):
        lib.bpk_debug_set_option(15, graph)
        for _ in range(3):
            out = msm(sc, pts)
        torch.cuda.synchronize()
        r = out.cpu().numpy().tobytes()
        ref = ref or r
        assert r == ref
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            msm(sc, pts)
        e1.record(); torch.cuda.synchronize()
        row.setdefault("graph" if graph else "plain", []).append(round(e0.elapsed_time(e1) / 20, 4))
    print(json.dumps(row), flush=True)
lib.bpk_debug_set_option(15, 1)
