"""Batch verification of 2^14 distinct 64-bit proofs of which 1 % are tampered (BASELINE config 5), per group size of the
grouped verification (BPK_OPT_VERIFY_GROUP; 0 = one by one): ms per batch, decisions checked."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp
lib = cbp.load()
m = 1 << 14
gp, _ = cbp.synth_points(130, seed=0xB0070002)
gens = cbp.Generators(gp[:64], gp[64:128], gp[128], gp[129], window_bits=16)
rng = np.random.default_rng(0xC5)
vals = rng.integers(0, 2**63, size=m, dtype=np.uint64)
gam = rng.integers(0, 2**63, size=(m, 4), dtype=np.uint64); gam[:, 3] &= np.uint64((1 << 59) - 1)
proofs = cbp.range_prove_batch(gens, vals, gam, np.arange(m, dtype=np.uint64))
h = proofs.cpu().numpy().copy()
bad = rng.choice(m, size=m // 100, replace=False)
for i in bad:
    h[i, rng.integers(0, h.shape[1])] ^= np.uint8(1 << rng.integers(0, 8))
expect = np.ones(m, dtype=bool); expect[bad] = False
tampered = torch.from_numpy(h).cuda()
ver = cbp.RangeVerifier(gens, m)
for group in [int(a) for a in sys.argv[1:]] or [0, 2, 4, 8, 16]:
    lib.bpk_debug_set_option(13, group)
    for name, recs, exp in (("honest", proofs, np.ones(m, dtype=bool)), ("1% tampered", tampered, expect)):
        acc = ver(recs); torch.cuda.synchronize()
        ok = bool((acc.cpu().numpy().astype(bool) == exp).all())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ver(recs)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"group {group:2d}  {name:12s} {ms:.3f} ms per 2^14 proofs, {m / ms / 1e3:.3f} M verifies/s, decisions correct: {ok}", flush=True)
lib.bpk_debug_set_option(13, -1)
