"""Times the host-pointer MSM drop-in (pinned buffers) for several chunk sizes, and device MSMs of chunk size."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp

n = 1 << 20
pts, _ = cbp.synth_points(n, seed=1)
sc = cbp.synth_scalars(n, seed=2, bits=253)
h_p = torch.empty((n, 128), dtype=torch.uint8).pin_memory(); h_p.copy_(pts)
h_s = torch.empty((n, 32), dtype=torch.uint8).pin_memory(); h_s.copy_(sc)
hp = h_p.numpy().view(np.uint64).reshape(n, 16)
hs = h_s.numpy().view(np.uint64).reshape(n, 4)
for lg in (17, 18):
    os.environ["CBP_HOST_CHUNK_LOG2"] = str(lg)
    for _ in range(3):
        cbp.cuda_point_vector_multi_scalar_mul(hs, hp)
    t0 = time.perf_counter()
    for _ in range(10):
        r = cbp.cuda_point_vector_multi_scalar_mul(hs, hp)
    dt = (time.perf_counter() - t0) / 10
    print(f"chunk 2^{lg}: {dt*1e3:.3f} ms  {n/dt/1e6:.1f} M points/s", flush=True)
for lg in (16, 17, 18, 19):
    m = 1 << lg
    msm = cbp.Msm(m)
    for _ in range(3):
        msm(sc[:m], pts[:m])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        msm(sc[:m], pts[:m])
    e1.record(); torch.cuda.synchronize()
    print(f"device MSM 2^{lg}: {e0.elapsed_time(e1)/20:.3f} ms (c={msm.window_bits})", flush=True)
# raw copy time
d = torch.empty_like(pts)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10):
    d.copy_(h_p, non_blocking=True)
torch.cuda.synchronize()
print(f"H2D 128 MB: {(time.perf_counter()-t0)/10*1e3:.3f} ms")
