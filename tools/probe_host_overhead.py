"""Host enqueue time of one MSM call against its device time (is the GPU starved by launches?)."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
for lg in [int(a) for a in sys.argv[1:]] or [10, 14, 16, 17, 18, 20]:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2, bits=252)
    msm = cbp.Msm(n)
    for _ in range(3):
        msm(sc, pts)
    torch.cuda.synchronize()
    host = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        t0 = time.perf_counter()
        msm(sc, pts)
        host.append(time.perf_counter() - t0)
        torch.cuda.synchronize()  # one call at a time: the enqueue of the next cannot hide behind this one
    e1.record(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(20):
        msm(sc, pts)
    back_to_back_host = (time.perf_counter() - t0) / 20
    torch.cuda.synchronize()
    print(f"2^{lg}: host enqueue {sum(host) / len(host) * 1e3:.3f} ms per call (isolated), {back_to_back_host * 1e3:.3f} ms back to back", flush=True)
