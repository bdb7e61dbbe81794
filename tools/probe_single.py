"""One warm pass of the latency-bound entry points, for an ncu launch list or wall-clock timing:
   single 64-bit range proof through cuda_range_proof_verify (host pointers), IPA prove n = 4096."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudabulletproof_b200 as cbp

what = sys.argv[1] if len(sys.argv) > 1 else "all"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
if what in ("verify", "all"):
    for nb in (16, 64):
        gp, _ = cbp.synth_points(2 * nb + 2, seed=0xB0070002 + nb)
        gens = cbp.Generators(gp[:nb], gp[nb:2 * nb], gp[2 * nb], gp[2 * nb + 1])
        gam = np.array([[7, 0, 0, 0]], dtype=np.uint64)
        rec = cbp.range_prove_batch(gens, [42], gam, [1])
        t0 = time.perf_counter()
        for _ in range(reps):
            rec = cbp.range_prove_batch(gens, [42], gam, [1])
        prove_ms = (time.perf_counter() - t0) / reps * 1e3
        hG = gp[:nb].cpu().numpy().view(np.uint64).reshape(nb, 16)
        hH = gp[nb:2 * nb].cpu().numpy().view(np.uint64).reshape(nb, 16)
        hg, hh = gp[2 * nb].cpu().numpy().view(np.uint64), gp[2 * nb + 1].cpu().numpy().view(np.uint64)
        proof, hV, keep = cbp.record_to_range_proof(rec[0].cpu().numpy(), nb)
        ok = cbp.cuda_range_proof_verify(proof, hV, nb, hG, hH, hg, hh)
        t0 = time.perf_counter()
        for _ in range(reps):
            ok = cbp.cuda_range_proof_verify(proof, hV, nb, hG, hH, hg, hh) and ok
        print(f"range proof {nb} bit: prove {prove_ms:.3f} ms, verify (host API) {(time.perf_counter() - t0) / reps * 1e3:.3f} ms, accepted {ok}", flush=True)
if what in ("ipa", "all"):
    n = 4096
    ig, _ = cbp.synth_points(2 * n + 1, seed=0xA66E0040)
    ia = cbp.synth_scalars(n, seed=0xA66E0041, bits=252)
    ib = cbp.synth_scalars(n, seed=0xA66E0042, bits=252)
    cbp.ipa_prove(ig[:n], ig[n:2 * n], ig[2 * n], ia, ib)
    t0 = time.perf_counter()
    for _ in range(reps):
        cbp.ipa_prove(ig[:n], ig[n:2 * n], ig[2 * n], ia, ib)
    print(f"ipa_prove n={n}: {(time.perf_counter() - t0) / reps * 1e3:.3f} ms", flush=True)
