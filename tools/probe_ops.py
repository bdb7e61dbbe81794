"""One pass of the HBM-side operations (inner product, batch inversion, point pack) at 2^22 / 2^20 elements: timing, or an
ncu target."""
import ctypes as C, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
cnt = 1 << 22
a = cbp.synth_scalars(cnt, seed=0xF1E1D, bits=253)
b = cbp.synth_scalars(cnt, seed=0xF1E1E, bits=253)
o = torch.empty_like(a)
pts, _ = cbp.synth_points(1 << 20, seed=0xC0DEC)
enc = torch.empty((1 << 20, 32), dtype=torch.uint8, device="cuda")
wsb = C.c_size_t(0)
lib.bpk_sc_inner_product_workspace_bytes(cnt, C.byref(wsb))
ws = torch.empty(max(1, wsb.value), dtype=torch.uint8, device="cuda")
ip = torch.empty(32, dtype=torch.uint8, device="cuda")
iwb = C.c_size_t(0)
lib.bpk_fe_batch_invert_workspace_bytes(cnt, C.byref(iwb))
iws = torch.empty(max(1, iwb.value), dtype=torch.uint8, device="cuda")
ops = {
    "sc_inner_product": (lambda: lib.bpk_sc_inner_product_device(ip.data_ptr(), a.data_ptr(), b.data_ptr(), cnt, ws.data_ptr(), ws.numel(), None), 64 * cnt),
    "fe_batch_invert": (lambda: lib.bpk_fe_batch_invert_device(o.data_ptr(), a.data_ptr(), cnt, iws.data_ptr(), iws.numel(), None), 64 * cnt),
    "point_pack": (lambda: lib.bpk_point_pack_device(enc.data_ptr(), pts.data_ptr(), 1 << 20, None), 160 << 20),
}
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
variants = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else [0]
for variant, (name, (fn, nbytes)) in [(v, it) for v in variants for it in ops.items()]:
    lib.bpk_debug_set_option(12, variant)  # BPK_OPT_DEBUG_VARIANT
    if len(variants) > 1:
        name = f"v{variant} {name}"
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{name}: {ms * 1e3:.1f} us, {nbytes / ms / 1e6:.0f} GB/s algorithmic", flush=True)
