"""Batch verification of 2^14 distinct 64-bit proofs (16-bit generator tables): timing, or an ncu target."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
m = 1 << 14
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
wb = int(sys.argv[2]) if len(sys.argv) > 2 else 16
gp, _ = cbp.synth_points(130, seed=0xB0070002)
gens = cbp.Generators(gp[:64], gp[64:128], gp[128], gp[129], window_bits=wb)
rng = np.random.default_rng(0xC5)
vals = rng.integers(0, 2**63, size=m, dtype=np.uint64)
gam = rng.integers(0, 2**63, size=(m, 4), dtype=np.uint64); gam[:, 3] &= np.uint64((1 << 59) - 1)
proofs = cbp.range_prove_batch(gens, vals, gam, np.arange(m, dtype=np.uint64))
ver = cbp.RangeVerifier(gens, m)
acc = ver(proofs); torch.cuda.synchronize()
assert bool(acc.all())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    ver(proofs)
e1.record(); torch.cuda.synchronize()
print(f"verify 2^14 ({wb}-bit tables): {e0.elapsed_time(e1) / reps:.3f} ms per batch, {m * reps / e0.elapsed_time(e1) / 1e3:.3f} M verifies/s", flush=True)
