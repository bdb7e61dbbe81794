"""prover throughput probe: proves `m` 64-bit proofs on the device (16-bit generator tables)"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
m = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
gpts, _ = cbp.synth_points(130, seed=0xB0070002)
gens = cbp.Generators(gpts[:64], gpts[64:128], gpts[128], gpts[129], window_bits=16)
rng = np.random.default_rng(1)
vals = rng.integers(0, 2**63, size=m, dtype=np.uint64)
gam = rng.integers(0, 2**63, size=(m, 4), dtype=np.uint64); gam[:, 3] &= np.uint64((1 << 59) - 1)
seeds = np.arange(m, dtype=np.uint64)
cbp.range_prove_batch(gens, vals, gam, seeds)
torch.cuda.synchronize(); t0 = time.perf_counter()
p = cbp.range_prove_batch(gens, vals, gam, seeds)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"{m} proofs in {dt*1e3:.1f} ms = {m/dt:.0f} proofs/s")
ver = cbp.RangeVerifier(gens, m)
print("all verify:", bool(ver(p).all().item()))
