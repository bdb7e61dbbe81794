"""End-to-end time of the host-pointer MSM drop-in from pinned, pageable and registered-once host buffers."""
import ctypes as C, os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
trace = False
for a in sys.argv[1:]:
    if a.startswith("--taper="):
        lib.bpk_debug_set_option(10, int(a.split("=")[1]))  # BPK_OPT_HOST_TAPER_LOG2
    if a == "--trace":
        trace = True
    if a.startswith("--variant="):
        lib.bpk_debug_set_option(12, int(a.split("=")[1]))  # BPK_OPT_DEBUG_VARIANT
    if a.startswith("--chunk="):
        lib.bpk_debug_set_option(2, int(a.split("=")[1]))  # BPK_OPT_HOST_CHUNK_LOG2
for lg in [int(a) for a in args] or [16, 18, 19, 20, 22]:
    n = 1 << lg
    pts, _ = cbp.synth_points(n, seed=5)
    sc = cbp.synth_scalars(n, seed=6, bits=252)
    ref = cbp.Msm(n)(sc, pts).cpu().numpy().tobytes()
    h_sc = torch.empty((n, 32), dtype=torch.uint8).pin_memory(); h_sc.copy_(sc)
    h_pts = torch.empty((n, 128), dtype=torch.uint8).pin_memory(); h_pts.copy_(pts)
    p_sc, p_pts = h_sc.numpy().copy(), h_pts.numpy().copy()
    out = np.zeros(16, dtype=np.uint64)
    row = {"log2_n": lg}
    for label, (a, b, reg) in {"pinned": (h_sc.data_ptr(), h_pts.data_ptr(), 0), "malloc": (p_sc.ctypes.data, p_pts.ctypes.data, 0),
                               "malloc_registered": (p_sc.ctypes.data, p_pts.ctypes.data, 1)}.items():
        lib.bpk_debug_set_option(6, reg)
        fv, pv = cbp.FieldVector(a, n), cbp.PointVector(b, n)
        for _ in range(2):
            lib.cuda_point_vector_multi_scalar_mul(out.ctypes.data_as(C.c_void_p), C.byref(fv), C.byref(pv))
        t0 = time.perf_counter()
        for _ in range(5):
            lib.cuda_point_vector_multi_scalar_mul(out.ctypes.data_as(C.c_void_p), C.byref(fv), C.byref(pv))
        row[label + "_ms"] = round((time.perf_counter() - t0) / 5 * 1e3, 3)
        if trace and label == "pinned":
            lib.bpk_debug_set_option(11, 1)
            lib.cuda_point_vector_multi_scalar_mul(out.ctypes.data_as(C.c_void_p), C.byref(fv), C.byref(pv))
            lib.bpk_debug_set_option(11, 0)
        assert out.tobytes() == ref, label
    lib.bpk_debug_set_option(6, 0)
    lib.bpk_host_release()
    print(row, flush=True)
