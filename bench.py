#!/usr/bin/env python
"""bench.py — the benchmark contract of this repo.

  python bench.py --gpus N --steps K --warmup W          # this repo's CUDA path (default)
  python bench.py --impl reference --gpus N ...           # the reference's own CPU path (oracle/_ref)
  torchrun --nproc-per-node N bench.py --gpus N ...       # N > 1: one rank per GPU, NCCL

Headline metric (BASELINE.json): Ed25519 MSM points/s at 2^20 points per GPU.  A step is one
multi-scalar multiplication over synthetic on-curve points and 252-bit scalars; with N ranks the
MSM has N * 2^20 pairs partitioned by point range, each rank reduces its range and the N partial
points are combined through an NCCL all-gather plus a 128-byte point-sum kernel (weak scaling).
The second BASELINE metric, 64-bit range-proof verifies/s, is measured in the same run and
reported under "secondary" (or as the headline with --workload verify).

The JSON line carries: value (inputs resident in HBM), e2e (through the reference-facing host-pointer
C ABI with pinned HOST buffers, copies inside the timed region), roofline (dominant kernel,
CUDA-event timed inside the timed region), cpu_baseline (the unmodified reference timed on this
box's host cores), clocks, gpu_launches.
"""
import argparse
import glob
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_N_DEFAULT = 20
VERIFY_PROOFS_DEFAULT = 1 << 14
INT_PEAK_PLANNING_TIMAD = 18.6  # SURVEY.md section 8d planning figure (148 SM x 4 x 16 lanes x 1.965 GHz, 32-bit IMAD)
INT_PEAK_FALLBACK_TIMAD = 9.0   # round-1 measurement (profiles/r01_microbench_int_pipe.jsonl); used only if the in-run
                                # microbenchmark fails, and then labelled "fallback"
KERNEL_SOURCES = {"msm": ("msm.cu", "ge25519.cuh", "fe25519.cuh", "fe8.cuh"),
                  "verify": ("rangeproof.cu", "rangeproof.cuh", "ge25519.cuh", "fe25519.cuh", "fe8.cuh", "sc25519.cuh")}
P25519 = 2**255 - 19
L25519 = 2**252 + 27742317777372353535851937790883648493


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def measure_int_peak(lib, target_ms=20.0):
    """The integer roofline denominator, measured in THIS run on this GPU (csrc/intpeak.cu): lane operations per
    second of IMAD.WIDE.U32 carry chains (fe_mul's inner pattern), next to the other multiply flavours so that
    the record shows why IMAD.WIDE is the right denominator here."""
    names = ["imad_wide_carry", "imad_wide_indep", "imad_lo", "imad_hi", "dfma_fp64", "imad_wide_carry_beside_equal_dfma"]
    rates = (C.c_double * len(names))()
    try:
        rc = lib.bpk_measure_int_peak(target_ms, rates, len(names))
    except Exception:  # noqa: BLE001
        rc = -1
    if rc != 0 or not rates[0] > 0:
        return {"T_per_s": INT_PEAK_FALLBACK_TIMAD, "source": "fallback (round-1 measurement)", "all": None}
    allr = {n: rates[i] / 1e12 for i, n in enumerate(names)}
    rec = {"T_per_s": allr["imad_wide_carry"], "all_T_per_s": allr, "target_ms_each": target_ms,
           "source": "bpk_measure_int_peak in this run (IMAD.WIDE.U32.X carry chains, 8 CTAs x 256 threads per SM)"}
    for d in (os.path.join(ROOT, "profiles"), os.path.join(ROOT, "gpurun_out")):
        try:
            if os.path.isdir(d):
                with open(os.path.join(d, "INT_PEAK.json"), "w") as f:
                    json.dump(dict(rec, when=time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())), f, indent=1)
        except OSError:
            pass
    return rec


def kernel_source_hash(kind):
    """sha256 over the CUDA sources that define the profiled kernel (`kind`: an ncu summary name, msm_* or a verifier
    kernel): ties a committed ncu summary to the code it was taken from (the GPU box has no .git, so a commit hash
    cannot be checked there)."""
    import hashlib
    h = hashlib.sha256()
    for name in KERNEL_SOURCES["msm" if kind.startswith("msm") else "verify"]:
        with open(os.path.join(ROOT, "cudabulletproof_b200", "csrc", name), "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def ncu_summary(kind):
    """profiles/ncu_<kind>.json (written by tools/ncu_extract.py from a committed ncu raw CSV): DRAM traffic and
    pipe utilisation of the dominant kernel.  Returned only when it was taken from the kernel sources this run was
    built from; otherwise the caller prints null and says why."""
    path = os.path.join(ROOT, "profiles", f"ncu_{kind}.json")
    try:
        with open(path) as f:
            d = json.load(f)
    except (OSError, ValueError):
        return None, f"no profiles/ncu_{kind}.json"
    if d.get("source_hash") != kernel_source_hash(kind):
        return None, (f"profiles/ncu_{kind}.json was taken at source hash {d.get('source_hash')} (git {d.get('git')}), "
                      f"this build is {kernel_source_hash(kind)}: not reported")
    return d, None


def dot_mod_l(sc_h, ks_h):
    """sum_i s_i k_i mod l, vectorised ((n,4) uint64 little-endian scalars, (n,) uint64 k): 32x32-bit partial products
    are exact in uint64, their halves are summed separately so nothing overflows"""
    import numpy as np
    s32 = np.ascontiguousarray(sc_h, dtype=np.uint64).view(np.uint32).reshape(-1, 8).astype(np.uint64)
    k32 = np.ascontiguousarray(ks_h, dtype=np.uint64).view(np.uint32).reshape(-1, 2).astype(np.uint64)
    mask, total = np.uint64(0xFFFFFFFF), 0
    for a in range(8):
        for b in range(2):
            prod = s32[:, a] * k32[:, b]
            total += (int((prod & mask).sum(dtype=np.uint64)) + (int((prod >> np.uint64(32)).sum(dtype=np.uint64)) << 32)) << (32 * (a + b))
    return total % L25519


def base_multiple_xy_hex(k):
    """hex of the 64 bytes x || y (little-endian) of k*B on Ed25519, in plain Python integers (independent of the
    library and of oracle/): the check value for an MSM over points P_i = k_i B"""
    d = -121665 * pow(121666, P25519 - 2, P25519) % P25519

    def add(p, q):
        (x1, y1), (x2, y2) = p, q
        t = d * x1 * x2 * y1 * y2 % P25519
        x3 = (x1 * y2 + x2 * y1) * pow(1 + t, P25519 - 2, P25519) % P25519
        y3 = (y1 * y2 + x1 * x2) * pow(1 - t, P25519 - 2, P25519) % P25519
        return x3, y3

    by = 4 * pow(5, P25519 - 2, P25519) % P25519
    bx = 0x216936d3cd6e53fec0a4e231fdd6dc5c692cc7609525a7b2c9562d608f25d51a
    acc, cur = (0, 1), (bx, by)
    k %= L25519
    while k:
        if k & 1:
            acc = add(acc, cur)
        cur = add(cur, cur)
        k >>= 1
    return (acc[0].to_bytes(32, "little") + acc[1].to_bytes(32, "little")).hex()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for nm, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the UNMODIFIED reference CPU MSM (oracle/_ref/libref_verbatim.so,
# point_vector_multi_scalar_mul, bulletproof_vectors.cu:189-224) on all host cores.
# ------------------------------------------------------------------------------------------------------
def cpu_reference_msm(points_per_thread, threads=None, repeats=1):
    import numpy as np
    from oracle import binding as ob
    lib = ob.load_verbatim()
    kind = "reference"
    if lib is None:  # never built here: fall back to this repo's CPU restatement
        lib, kind = ob.load_oracle(), "port"
    threads = threads or os.cpu_count() or 1
    rng = np.random.default_rng(0x5CA1A000 + LOG_N_DEFAULT)
    n = points_per_thread
    sc = rng.integers(0, 2**63, size=(threads, n, 4), dtype=np.uint64)
    sc[:, :, 3] &= np.uint64((1 << 60) - 1)  # 252-bit scalars, like the GPU arm
    pts = rng.integers(0, 2**63, size=(threads, n, 16), dtype=np.uint64)  # the reference never checks curve membership
    pts[:, :, 8:12] = 0
    pts[:, :, 8] = 1  # Z = 1
    outs = np.zeros((threads, 16), dtype=np.uint64)

    def work(t):
        fv, pv = ob.field_vector(sc[t]), ob.point_vector(pts[t])
        lib.point_vector_multi_scalar_mul(ob.ptr(outs[t]), C.byref(fv), C.byref(pv))

    best = None
    for _ in range(repeats):
        ths = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return {"value": threads * n / best, "unit": "points/s", "cores": threads, "kind": kind,
            "sample": f"{threads} threads x {n} points each through point_vector_multi_scalar_mul "
                      f"(reference bulletproof_vectors.cu:189-224), wall {best:.2f} s"}


def cpu_reference_verify(count=8, threads=None):
    """the reference's CPU range_proof_verify (bulletproof_range_proof.cu:1717) on its own 64-bit proofs"""
    import numpy as np
    from oracle import binding as ob
    lib = ob.load_verbatim()
    if lib is None:
        return None
    threads = min(threads or os.cpu_count() or 1, 16)
    n = 64
    rng = np.random.default_rng(7)
    G = rng.integers(0, 2**63, size=(n, 16), dtype=np.uint64)
    H = rng.integers(0, 2**63, size=(n, 16), dtype=np.uint64)
    gh = rng.integers(0, 2**63, size=(2, 16), dtype=np.uint64)
    for a in (G, H, gh):
        a[:, 8:12] = 0
        a[:, 8] = 1
    Gv, Hv = ob.point_vector(G), ob.point_vector(H)
    devnull = os.open(os.devnull, os.O_WRONLY)
    saved = os.dup(1)
    sys.stdout.flush()
    os.dup2(devnull, 1)  # the reference verifier prints ~100 lines per call
    try:
        lib.refv_seed_rng(1)
        proof = ob.RangeProof()
        v, gam = ob.int_to_fe(42), ob.int_to_fe(12345)
        lib.generate_range_proof(C.byref(proof), ob.ptr(v), ob.ptr(gam), n, C.byref(Gv), C.byref(Hv), ob.ptr(gh[0]), ob.ptr(gh[1]))
        V = np.frombuffer(bytes(proof.V), dtype=np.uint64).copy()

        def work():
            for _ in range(count):
                lib.range_proof_verify(C.byref(proof), ob.ptr(V), n, C.byref(Gv), C.byref(Hv), ob.ptr(gh[0]), ob.ptr(gh[1]))

        ths = [threading.Thread(target=work) for _ in range(threads)]
        t0 = time.perf_counter()
        for th in ths:
            th.start()
        for th in ths:
            th.join()
        dt = time.perf_counter() - t0
    finally:
        sys.stdout.flush()
        os.dup2(saved, 1)
        os.close(devnull)
        os.close(saved)
    return {"value": threads * count / dt, "unit": "verifies/s", "cores": threads, "kind": "reference",
            "sample": f"{threads} threads x {count} range_proof_verify calls on one 64-bit proof "
                      f"(reference bulletproof_range_proof.cu:1717), wall {dt:.2f} s"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_thread = args.cpu_sample
    vals = []
    for _ in range(max(1, args.warmup)):
        cpu_reference_msm(max(64, per_thread // 8))
    t0 = time.perf_counter()
    res = None
    for _ in range(args.steps):
        res = cpu_reference_msm(per_thread)
        vals.append(res["value"])
    wall = time.perf_counter() - t0
    value = sum(vals) / len(vals)
    res["value"] = value
    line = {"impl": "reference", "metric": "ed25519_msm_points_per_sec", "value": value, "unit": "points/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": wall / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": f"Ed25519 MSM 2^{args.log_n} random scalars/points per GPU (BASELINE.json configs[2])",
                       "note": "reference CPU path, bounded sample per step"},
            "cpu_baseline": res,
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# this repo's arm
# ------------------------------------------------------------------------------------------------------
def bind_to_gpu_numa(local_rank):
    """Pin this process to the CPUs next to its GPU (NVML's affinity mask) BEFORE any pinned host buffer is allocated:
    with the default first-touch policy the buffers then live in that GPU's NUMA node, so N ranks uploading at once draw
    on every node's memory bandwidth instead of the one the launcher happened to start them on.  Returns what was done."""
    info = {"bound": False}
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        pr = torch.cuda.get_device_properties(local_rank)
        bus = "%08x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), pr.pci_bus_id, pr.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        nodes = {}
        for path in glob.glob("/sys/devices/system/node/node*/cpulist"):
            node = int(path.split("node")[-1].split("/")[0])
            got = set()
            for part in open(path).read().strip().split(","):
                if part:
                    lo, _, hi = part.partition("-")
                    got |= set(range(int(lo), int(hi or lo) + 1))
            nodes[node] = got
        info.update({"gpu_pci": bus, "gpu_cpus": len(cpus), "numa_nodes": len(nodes),
                     "gpu_nodes": sorted(n for n, c in nodes.items() if c & cpus)})
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            info["bound"] = True
    except Exception as exc:  # noqa: BLE001 - best effort: no NVML, no permission, one node
        info["error"] = repr(exc)[:120]
    return info


def run_cuda(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cudabulletproof_b200 as cbp

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    numa = bind_to_gpu_numa(local_rank) if (world > 1 and not args.no_numa_bind) else {"bound": False}
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"  # NCCL prints its version banner on stdout; stdout carries ONE JSON line
        dist.init_process_group("nccl", device_id=dev)
    lib = cbp.load()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def prof_read(kind):
        ms, cnt = C.c_float(0), C.c_int(0)
        lib.bpk_profile_read(kind, C.byref(ms), C.byref(cnt))
        return ms.value, cnt.value

    hbm_peak, peak_src = measured_peaks()
    int_peak = measure_int_peak(lib)  # ~0.1 s, before any timed region
    INT_PEAK_TIMAD = int_peak["T_per_s"]

    # ---------------- MSM ----------------
    def bench_msm():
        n = 1 << args.log_n
        pts, _ = cbp.synth_points(n, seed=0xC3 + args.log_n + 1000 * rank, device=dev)
        sc = cbp.synth_scalars(n, seed=0x5CA1A000 + args.log_n + 1000 * rank, bits=252, device=dev)
        from cudabulletproof_b200.multi import ShardedMsm
        smsm = ShardedMsm(n, world, dev)  # this rank's point range; NCCL all-gather of 128 B partials + point sum
        msm = smsm.msm

        def step():
            return smsm(sc, pts)

        for _ in range(args.warmup):
            step()
        barrier()
        lib.bpk_profile_reset()
        lib.bpk_profile_enable(1)
        launches0 = lib.bpk_kernel_launches()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            res = step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        clocks = sampler.stop() if rank == 0 else None
        lib.bpk_profile_enable(0)
        launches = lib.bpk_kernel_launches() - launches0 + (args.steps if world > 1 else 0)
        acc_ms, acc_n = prof_read(0)
        pre_ms, _ = prof_read(4)
        phases = {"front_ms": prof_read(5)[0], "accumulate_ms": acc_ms, "tail_ms": prof_read(6)[0],
                  "total_ms": prof_read(1)[0], "table_build_ms_overlapped": pre_ms}
        result_hex = bytes(res.cpu().numpy().tobytes()[:64]).hex()
        W = (256 + msm.window_bits - 1) // msm.window_bits
        imad_acc = n * W * 504.0  # 7 fe_mul x 72 IMAD per mixed addition, N*W additions
        ncu_acc, ncu_why = ncu_summary(f"msm_accumulate_2_{args.log_n}")
        achieved = imad_acc / (acc_ms * 1e-3) / 1e12
        roofline = {"bound": "int", "kernel": "msm_accumulate_kernel", "achieved": achieved,
                    "peak": INT_PEAK_TIMAD, "unit": "TIMAD/s", "frac": achieved / INT_PEAK_TIMAD,
                    "peak_source": int_peak["source"], "peak_all_T_per_s": int_peak.get("all_T_per_s"),
                    "frac_of_planning_peak": achieved / INT_PEAK_PLANNING_TIMAD,
                    "planning_peak_note": f"{INT_PEAK_PLANNING_TIMAD} T/s = SURVEY.md section 8d's 32-bit IMAD figure; the "
                                          "32x32->64 IMAD.WIDE this path needs issues at half that rate (measured above)",
                    "traffic": ncu_acc["dram_bytes"] if ncu_acc else None,
                    "traffic_note": (ncu_acc["note"] if ncu_acc else ncu_why),
                    "ncu": ({k: ncu_acc[k] for k in ("file", "git", "source_hash", "launches", "fmaheavy_pct_per_launch",
                                                       "duration_us_per_launch") if k in ncu_acc} if ncu_acc else None),
                    "algorithmic_gather_bytes": n * W * 100.0,
                    "launch_ms": acc_ms, "launches_timed": acc_n,
                    "launch_note": "span of the window-group accumulation launches per MSM: every group accumulates on its own "
                                   "stream, so the span runs from the first launch to the end of the last one (CUDA events on the "
                                   "launching stream, which waits for every group); the groups' bucket reductions share the span",
                    "algorithmic_imad_per_launch": imad_acc,
                    # the W * 2^c additions (9M = 648) of the bucket reductions that run inside the same span
                    "frac_with_reductions": (imad_acc + W * (1 << msm.window_bits) * 648.0) / (acc_ms * 1e-3) / 1e12 / INT_PEAK_TIMAD,
                    "whole_call_frac": imad_acc / (ms / args.steps * 1e-3) / 1e12 / INT_PEAK_TIMAD}
        roofline_hbm = {"bound": "hbm", "kernel": "msm_precompute_kernel", "achieved": n * 224.0 / (pre_ms * 1e-3) / 1e9,
                        "peak": hbm_peak, "unit": "GB/s", "frac": n * 224.0 / (pre_ms * 1e-3) / 1e9 / hbm_peak,
                        "traffic": None, "peak_source": f"MEASURED_PEAKS.json ({peak_src})", "launch_ms": pre_ms}
        # ---- e2e: the reference-facing host-pointer call with pinned HOST buffers ----
        h_sc = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
        h_pts = torch.empty((n, 128), dtype=torch.uint8).pin_memory()
        h_sc.copy_(sc)
        h_pts.copy_(pts)
        torch.cuda.synchronize()
        h_out = np.zeros(16, dtype=np.uint64)
        fv = cbp.FieldVector(h_sc.data_ptr(), n)
        pv = cbp.PointVector(h_pts.data_ptr(), n)
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(2):
            lib.cuda_point_vector_multi_scalar_mul(h_out.ctypes.data_as(C.c_void_p), C.byref(fv), C.byref(pv))
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            lib.cuda_point_vector_multi_scalar_mul(h_out.ctypes.data_as(C.c_void_p), C.byref(fv), C.byref(pv))
        torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        e2e_hex = h_out.tobytes()[:64].hex()
        # the same call from PAGEABLE buffers (plain malloc, what the reference's own callers hand over:
        # bulletproof_vectors.cu:18,136), then with the library's opt-in registration cache
        pageable = {}
        if world == 1:
            p_sc, p_pts = np.empty((n, 32), dtype=np.uint8), np.empty((n, 128), dtype=np.uint8)
            p_sc[:] = h_sc.numpy()
            p_pts[:] = h_pts.numpy()
            fvp, pvp = cbp.FieldVector(p_sc.ctypes.data, n), cbp.PointVector(p_pts.ctypes.data, n)
            for label, reg in (("malloc", 0), ("malloc_registered_once", 1)):
                lib.bpk_debug_set_option(6, reg)  # BPK_OPT_HOST_REGISTER
                h_out[:] = 0
                for _ in range(2):
                    lib.cuda_point_vector_multi_scalar_mul(h_out.ctypes.data_as(C.c_void_p), C.byref(fvp), C.byref(pvp))
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    lib.cuda_point_vector_multi_scalar_mul(h_out.ctypes.data_as(C.c_void_p), C.byref(fvp), C.byref(pvp))
                torch.cuda.synchronize()
                pms = (time.perf_counter() - t0) * 1e3 / e2e_steps
                pageable[label] = {"ms_per_step": pms, "value": n / (pms * 1e-3), "unit": "points/s",
                                   "result_matches_device_path": h_out.tobytes()[:64].hex() == result_hex}
            lib.bpk_debug_set_option(6, 0)
            lib.bpk_host_release()
        # the documented extension for callers that can hand over affine points (x || y, 64 B): 96 B per pair on the wire
        affine = None
        if world == 1:
            h_xy = torch.empty((n, 64), dtype=torch.uint8).pin_memory()
            h_xy.copy_(pts[:, :64])  # the synthetic points are normalised (Z = 1)
            h_out[:] = 0
            for _ in range(2):
                lib.bpk_msm_host_affine(h_out.ctypes.data_as(C.c_void_p), h_sc.data_ptr(), h_xy.data_ptr(), n)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                lib.bpk_msm_host_affine(h_out.ctypes.data_as(C.c_void_p), h_sc.data_ptr(), h_xy.data_ptr(), n)
            torch.cuda.synchronize()
            ams = (time.perf_counter() - t0) * 1e3 / e2e_steps
            affine = {"ms_per_step": ams, "value": n / (ams * 1e-3), "unit": "points/s", "h2d_bytes_per_step": n * 96,
                      "api": "bpk_msm_host_affine (extension: 64-byte affine points, pinned host buffers)",
                      "result_matches_device_path": h_out.tobytes()[:64].hex() == result_hex}
        return {"ms": ms, "n": n, "clocks": clocks, "launches": int(launches), "roofline": roofline,
                "roofline_hbm": roofline_hbm, "window_bits": msm.window_bits, "phases": phases,
                "e2e": {"value": world * n * e2e_steps / (e2e_ms * 1e-3), "unit": "points/s",
                        "h2d_bytes_per_step": n * 160, "d2h_bytes_per_step": 128,
                        "api": "cuda_point_vector_multi_scalar_mul (host pointers, pinned)", "ms_per_step": e2e_ms / e2e_steps,
                        "result_matches_device_path": (e2e_hex == result_hex) if world == 1 else None,
                        "pageable": pageable or None, "affine_extension": affine},
                "result_xy": result_hex}

    # ---------------- strong scaling: ONE global MSM cut into point-range shards ------
    def bench_strong(log_n):
        """The same 2^log_n global inputs for every world size; rank r owns shard_range(n, r, world); the partial points
        meet through an all_gather of 128 B + the point-sum kernel.  result_xy must be identical for N = 1, 2, 4, 8 and
        equal (sum s_i k_i mod l) B, checked in plain Python outside the timed region."""
        from cudabulletproof_b200.multi import all_gather_bytes, shard_range
        from cudabulletproof_b200.host import Msm, point_sum
        n = 1 << log_n
        pts, ks = cbp.synth_points(n, seed=0x57A0 + log_n, device=dev)  # every rank derives the same global arrays
        sc = cbp.synth_scalars(n, seed=0x57B0 + log_n, bits=252, device=dev)
        lo, hi = shard_range(n, rank, world)
        my_pts, my_sc = pts[lo:hi].contiguous(), sc[lo:hi].contiguous()
        msm = Msm(hi - lo, device=dev)
        partial = torch.zeros(128, dtype=torch.uint8, device=dev)
        gathered = torch.zeros((world, 128), dtype=torch.uint8, device=dev)

        def step():
            if world == 1:
                return msm(my_sc, my_pts, normalize=True)
            msm(my_sc, my_pts, normalize=False, out=partial)
            all_gather_bytes(partial, world, out=gathered)
            return point_sum(gathered, normalize=True)

        for _ in range(max(3, args.warmup)):
            res = step()
        barrier()
        ssteps = max(5, min(args.steps, 20))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(ssteps):
            res = step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1)) / ssteps
        got = bytes(res.cpu().numpy().tobytes()[:64]).hex()
        row = {"log2_n": log_n, "ms_per_step": ms, "value": n / (ms * 1e-3), "unit": "points/s", "n_gpus": world,
               "shard_points": hi - lo, "window_bits": msm.window_bits, "result_xy": got}
        if rank == 0:
            want = base_multiple_xy_hex(dot_mod_l(sc.cpu().numpy().view(np.uint64).reshape(n, 4),
                                                  ks.cpu().numpy().astype(np.uint64)))
            row["result_check"] = got == want
            row["check"] = "(sum s_i k_i mod l) * B in plain Python integers, outside the timed region"
        del pts, sc, ks, my_pts, my_sc, msm
        torch.cuda.empty_cache()
        return row

    # ---------------- range-proof batch verification ----------------
    def bench_verify():
        nbits, m = 64, args.proofs
        distinct = min(m, args.distinct_proofs)
        gpts, _ = cbp.synth_points(2 * nbits + 2, seed=0xB0070002, device=dev)
        gens = cbp.Generators(gpts[:nbits], gpts[nbits:2 * nbits], gpts[2 * nbits], gpts[2 * nbits + 1], device=dev,
                              window_bits=args.fixed_window_bits)
        nwin = 256 // args.fixed_window_bits
        rng = np.random.default_rng(0xC5 + rank)
        vals = rng.integers(0, 2**63, size=distinct, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=distinct, dtype=np.uint64)
        gam = rng.integers(0, 2**63, size=(distinct, 4), dtype=np.uint64)
        gam[:, 3] &= np.uint64((1 << 59) - 1)
        seeds = np.arange(distinct, dtype=np.uint64) + np.uint64(100000 * rank)
        base = cbp.range_prove_batch(gens, vals, gam, seeds)  # also warms the prover up
        torch.cuda.synchronize()
        prove_ms = float("inf")
        for _ in range(3):  # best of three: the call allocates its 0.5 GB workspace, which the caching allocator
            pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)  # may or may not have
            pe0.record()
            base = cbp.range_prove_batch(gens, vals, gam, seeds)
            pe1.record()
            torch.cuda.synchronize()
            prove_ms = min(prove_ms, pe0.elapsed_time(pe1))
        reps = (m + distinct - 1) // distinct
        proofs = base.repeat(reps, 1)[:m].contiguous()
        bad = rng.choice(m, size=max(1, m // 100), replace=False)  # 1 % tampered: one bit flipped
        hb = proofs[bad].cpu().numpy()
        for i in range(len(bad)):
            hb[i, rng.integers(0, hb.shape[1])] ^= np.uint8(1 << rng.integers(0, 8))
        proofs[torch.from_numpy(bad).to(dev)] = torch.from_numpy(hb).to(dev)
        ver = cbp.RangeVerifier(gens, m)
        masks = torch.zeros((world, m), dtype=torch.uint8, device=dev)
        # grouped verification (csrc/rangeproof.cu): K proofs share one combined identity; 0 / 1 = one by one
        K = args.verify_group if args.verify_group >= 0 else (12 if m >= 256 else 0)
        lib.bpk_debug_set_option(13, K)  # BPK_OPT_VERIFY_GROUP
        groups_hit = len(set(int(b) // K for b in bad)) if K >= 2 else 0
        second_pass = min(m, groups_hit * K) if K >= 2 else 0  # proofs verified again one by one (upper bound)

        def step():
            acc = ver(proofs)
            if world > 1:
                dist.all_gather_into_tensor(masks.view(-1), acc)
            return acc

        for _ in range(args.warmup):
            step()
        barrier()
        lib.bpk_profile_reset()
        lib.bpk_profile_enable(1)
        launches0 = lib.bpk_kernel_launches()
        vsteps = max(3, min(args.steps, 10))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(vsteps):
            acc = step()
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        lib.bpk_profile_enable(0)
        launches = lib.bpk_kernel_launches() - launches0
        k_ms, k_n = prof_read(2)
        chunks = max(1, k_n // vsteps)
        rejected = int((acc == 0).sum().item())
        ok = rejected == len(bad) and bool((acc[torch.from_numpy(bad).to(dev)] == 0).all().item())
        # the same batch one by one (two identities per proof, the round-1 algorithm), for comparison
        one_by_one = None
        if K >= 2:
            lib.bpk_debug_set_option(13, 0)
            for _ in range(2):
                acc1 = ver(proofs)
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for _ in range(3):
                acc1 = ver(proofs)
            o1.record()
            torch.cuda.synchronize()
            one_by_one = {"value": m * 3 / (o0.elapsed_time(o1) * 1e-3), "unit": "verifies/s per GPU",
                          "ms_per_step": o0.elapsed_time(o1) / 3, "same_decisions": bool((acc1 == acc).all().item())}
            lib.bpk_debug_set_option(13, K)
        # algorithmic IMAD (SURVEY.md §8d units).  Whole proof: 131 bases x nwin windows mixed additions (504)
        # + 17 points x 51 signed 5-bit windows x 8M (576) + 2 x (255 doublings (464) + 51 additions (648)) Horner.
        # The timed kernel (verify_fixed_kernel, nwin lanes per proof) does the first term plus two
        # log2(nwin)-level shuffle trees of 9M additions.
        imad_single = (131 * nwin * 504 + 17 * 51 * 576 + 2 * (255 * 464 + 51 * 648)) * 1.0
        if K >= 2:
            # grouped: per proof its own points (A carries a 128-bit weight: 26 windows) + 1/K of one fixed-base sum over
            # 130 bases and one Horner chain, + the second pass (members of failed groups, each as a group of one)
            var_proof = (16 * 51 + 26) * 576.0
            shared = 130 * nwin * 504 + 255 * 464 + 51 * 648.0
            imad_proof = var_proof + shared / K + (var_proof + shared) * second_pass / m
            imad_kernel, kernel_name, ncu_kind = var_proof, "vg_winsum_kernel", f"vg_winsum_{m}"
        else:
            imad_proof = imad_single
            imad_kernel = 131 * nwin * 504 + 2 * (nwin.bit_length() - 1) * nwin * 648.0
            kernel_name, ncu_kind = "verify_fixed_kernel", f"verify_fixed_{args.fixed_window_bits}_{m}"
        per_launch = imad_kernel * (m / chunks)
        ncu_v, ncu_v_why = ncu_summary(ncu_kind)
        roofline = {"bound": "int", "kernel": kernel_name, "achieved": per_launch / (k_ms * 1e-3) / 1e12,
                    "peak": INT_PEAK_TIMAD, "unit": "TIMAD/s", "frac": per_launch / (k_ms * 1e-3) / 1e12 / INT_PEAK_TIMAD,
                    "traffic": ncu_v["dram_bytes"] if ncu_v else None,
                    "traffic_note": ncu_v["note"] if ncu_v else ncu_v_why,
                    "peak_source": int_peak["source"],
                    "launch_ms": k_ms, "launches_timed": k_n,
                    "algorithmic_imad_per_launch": per_launch, "algorithmic_imad_per_proof_total": imad_proof,
                    "algorithmic_imad_per_proof_one_by_one": imad_single,
                    "whole_batch_frac": imad_proof * m / (ms / vsteps * 1e-3) / 1e12 / INT_PEAK_TIMAD}
        # e2e: proof records in pinned host memory -> device -> accept mask back on the host
        h_proofs = torch.empty_like(proofs, device="cpu").pin_memory()
        h_proofs.copy_(proofs)
        h_acc = torch.empty(m, dtype=torch.uint8).pin_memory()
        d_stage = torch.empty_like(proofs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(vsteps):
            d_stage.copy_(h_proofs, non_blocking=True)
            a = ver(d_stage)
            h_acc.copy_(a, non_blocking=True)
            torch.cuda.synchronize()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        lib.bpk_debug_set_option(13, -1)
        return {"metric": "range_proof_verifies_per_sec", "value": world * m * vsteps / (ms * 1e-3), "unit": "verifies/s",
                "ms_per_step": ms / vsteps, "steps": vsteps, "proofs_per_gpu": m, "distinct_proofs": distinct,
                "tampered": len(bad), "decisions_correct": ok, "gpu_launches": int(launches), "roofline": roofline,
                "algorithm": ({"grouped": K, "second_pass_proofs": second_pass,
                               "note": f"groups of {K} proofs share one combined identity (128-bit weights from SHA-256 over "
                                       "all records of the group); members of failed groups are verified again on their own; "
                                       "accept bits per proof"} if K >= 2 else {"grouped": 0, "note": "two identities per proof"}),
                "one_by_one": one_by_one,
                "prover": {"proofs_per_s": distinct / (prove_ms * 1e-3), "ms": prove_ms, "proofs": distinct,
                           "note": "bpk_range_prove_batch_device, 64-bit proofs, byte-identical to the CPU oracle's "
                                   "(host-to-device copy of values / blinding factors included)"},
                "e2e": {"value": world * m * vsteps / (e2e_ms * 1e-3), "unit": "verifies/s",
                        "h2d_bytes_per_step": int(proofs.numel()), "d2h_bytes_per_step": m,
                        "api": "bpk_range_verify_batch_device on records staged from pinned host memory"},
                "config": {"workload": f"batch verification of 2^{m.bit_length() - 1} 64-bit range proofs per GPU, 1% tampered "
                                       "(BASELINE.json configs[4]); records prover-generated on device, "
                                       f"{distinct} distinct proofs; generator tables with {args.fixed_window_bits}-bit windows"},
                "fixed_window_bits": args.fixed_window_bits}

    # ---------------- the HBM-bound rows of SURVEY.md §8d (rank 0, one pass each, inputs larger than L2) ----------------
    def bench_other_ops():
        cnt = 1 << 22  # 128 MiB per array of field elements
        a = cbp.synth_scalars(cnt, seed=0xF1E1D, bits=253, device=dev)
        b = cbp.synth_scalars(cnt, seed=0xF1E1E, bits=253, device=dev)
        o = torch.empty_like(a)
        pts, _ = cbp.synth_points(1 << 20, seed=0xC0DEC, device=dev)
        enc = torch.empty((1 << 20, 32), dtype=torch.uint8, device=dev)
        back = torch.empty_like(pts)
        okm = torch.empty(1 << 20, dtype=torch.uint8, device=dev)
        wsb = C.c_size_t(0)
        lib.bpk_sc_inner_product_workspace_bytes(cnt, C.byref(wsb))
        ws = torch.empty(max(1, wsb.value), dtype=torch.uint8, device=dev)
        ip = torch.empty(32, dtype=torch.uint8, device=dev)
        st = torch.cuda.current_stream().cuda_stream

        def timed(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps

        rows = {}

        def hbm_row(name, fn, nbytes, units, imad_per_unit):
            ms = timed(fn)
            rows[name] = {"ms": ms, "GB/s": nbytes / (ms * 1e-3) / 1e9, "frac_hbm": nbytes / (ms * 1e-3) / 1e9 / hbm_peak,
                          "units_per_s": units / (ms * 1e-3),
                          "frac_int": units * imad_per_unit / (ms * 1e-3) / 1e12 / INT_PEAK_TIMAD, "units": units}

        hbm_row("fe_batch_add", lambda: lib.bpk_fe_batch_device(0, o.data_ptr(), a.data_ptr(), b.data_ptr(), cnt, st), 96 * cnt, cnt, 0)
        hbm_row("fe_batch_mul", lambda: lib.bpk_fe_batch_device(2, o.data_ptr(), a.data_ptr(), b.data_ptr(), cnt, st), 96 * cnt, cnt, 72)
        hbm_row("fe_batch_square", lambda: lib.bpk_fe_batch_device(3, o.data_ptr(), a.data_ptr(), a.data_ptr(), cnt, st), 64 * cnt, cnt, 44)
        iwb = C.c_size_t(0)
        lib.bpk_fe_batch_invert_workspace_bytes(cnt, C.byref(iwb))
        iws = torch.empty(max(1, iwb.value), dtype=torch.uint8, device=dev)
        hbm_row("fe_batch_invert", lambda: lib.bpk_fe_batch_invert_device(o.data_ptr(), a.data_ptr(), cnt, iws.data_ptr(), iws.numel(), st),
                128 * cnt, cnt, 3 * 72)
        hbm_row("fe_batch_invert_single_kernel", lambda: lib.bpk_fe_batch_invert_device(o.data_ptr(), a.data_ptr(), cnt, None, 0, st),
                128 * cnt, cnt, 3 * 72)
        hbm_row("sc_inner_product", lambda: lib.bpk_sc_inner_product_device(ip.data_ptr(), a.data_ptr(), b.data_ptr(), cnt,
                                                                            ws.data_ptr(), ws.numel(), st), 64 * cnt, cnt, 64)
        hbm_row("point_pack", lambda: lib.bpk_point_pack_device(enc.data_ptr(), pts.data_ptr(), 1 << 20, st), 224 * (1 << 20), 1 << 20,
                5 * 72)
        hbm_row("point_unpack", lambda: lib.bpk_point_unpack_device(back.data_ptr(), okm.data_ptr(), enc.data_ptr(), 1 << 20, st),
                160 * (1 << 20), 1 << 20, 253 * 44 + 25 * 72)
        # BASELINE config 4: one inner-product argument over n = 4096 generators (12 rounds) proved on the device
        n_ipa = 4096
        ig, _ = cbp.synth_points(2 * n_ipa + 1, seed=0xA66E0040, device=dev)
        ia = cbp.synth_scalars(n_ipa, seed=0xA66E0041, bits=252, device=dev)
        ib = cbp.synth_scalars(n_ipa, seed=0xA66E0042, bits=252, device=dev)
        cbp.ipa_prove(ig[:n_ipa], ig[n_ipa:2 * n_ipa], ig[2 * n_ipa], ia, ib)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            cbp.ipa_prove(ig[:n_ipa], ig[n_ipa:2 * n_ipa], ig[2 * n_ipa], ia, ib)
        ipa_ms = (time.perf_counter() - t0) / 3 * 1e3
        rows["ipa_prove_n4096"] = {"ms": ipa_ms, "rounds": 12,
                                   "note": "bpk_ipa_prove_device: per round 2 inner products, 2 Pippenger MSMs of 2n'+1 points, "
                                           "challenge hash + inversion, a/b and G/H folds; wall time incl. workspace allocation"}
        # BASELINE configs 0-1: ONE range proof, proved on the device and verified through the reference-facing
        # host-pointer drop-in cuda_range_proof_verify (latency, not throughput)
        lat = {}
        for nb in (16, 64):
            gp, _ = cbp.synth_points(2 * nb + 2, seed=0xB0070002 + nb, device=dev)
            gsmall = cbp.Generators(gp[:nb], gp[nb:2 * nb], gp[2 * nb], gp[2 * nb + 1], device=dev)
            one_gam = np.array([[7, 0, 0, 0]], dtype=np.uint64)
            cbp.range_prove_batch(gsmall, [42], one_gam, [1])
            t0 = time.perf_counter()
            recd = cbp.range_prove_batch(gsmall, [42], one_gam, [1])
            prove_ms1 = (time.perf_counter() - t0) * 1e3
            hG = gp[:nb].cpu().numpy().view(np.uint64).reshape(nb, 16)
            hH = gp[nb:2 * nb].cpu().numpy().view(np.uint64).reshape(nb, 16)
            hg = gp[2 * nb].cpu().numpy().view(np.uint64)
            hh = gp[2 * nb + 1].cpu().numpy().view(np.uint64)
            proof, hV, keep = cbp.record_to_range_proof(recd[0].cpu().numpy(), nb)
            okv = cbp.cuda_range_proof_verify(proof, hV, nb, hG, hH, hg, hh)  # builds and caches the generator tables
            t0 = time.perf_counter()
            for _ in range(5):
                okv = cbp.cuda_range_proof_verify(proof, hV, nb, hG, hH, hg, hh) and okv
            lat[f"range_proof_{nb}bit"] = {"prove_ms": prove_ms1, "verify_host_api_ms": (time.perf_counter() - t0) / 5 * 1e3,
                                          "accepted": bool(okv)}
        rows["single_proof_latency"] = lat
        return {"note": "one kernel pass each on arrays larger than L2; bytes are algorithmic (SURVEY.md section 8d), "
                        f"HBM peak {hbm_peak:.0f} GB/s ({peak_src}), integer peak {INT_PEAK_TIMAD} T IMAD.WIDE/s", "rows": rows}

    msm_res = bench_msm() if args.workload in ("msm", "both") else None
    strong = None
    if args.workload in ("msm", "both") and not args.no_strong:
        strong = [bench_strong(lg) for lg in args.strong_log_n]
    ver_res = bench_verify() if args.workload in ("verify", "both") else None
    other = None
    if rank == 0 and world == 1 and args.workload == "both":
        try:
            other = bench_other_ops()
        except Exception as ex:  # never lose the headline numbers over the side table
            other = {"error": repr(ex)}

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            try:
                cpu = cpu_reference_msm(args.cpu_sample)
                if ver_res is not None:
                    ver_res["cpu_baseline"] = cpu_reference_verify()
            except Exception as ex:  # the baseline is a report, never a reason to lose the GPU numbers
                cpu = {"error": repr(ex)}
        if msm_res is not None:
            n = msm_res["n"]
            line = {"metric": "ed25519_msm_points_per_sec", "value": world * n * args.steps / (msm_res["ms"] * 1e-3),
                    "unit": "points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": msm_res["ms"] / args.steps, "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "u32 limbs (IMAD.WIDE.U32), exact integer arithmetic",
                    "data": "synthetic",
                    "config": {"workload": f"Ed25519 MSM 2^{args.log_n} random 252-bit scalars x on-curve points per GPU "
                                           "(BASELINE.json configs[2]), point-range sharded across ranks",
                               "window_bits": msm_res["window_bits"], "points_per_gpu": n,
                               "l2": "inputs (160 B/pair = 168 MB at 2^20) larger than the 126 MB L2; no explicit flush",
                               "parallelism": f"point-range x{world}" if world > 1 else "single GPU",
                               "combine": "NCCL all_gather of 128 B partial points + point-sum kernel" if world > 1 else None,
                               "host_numa": numa if world > 1 else None},
                    "e2e": msm_res["e2e"], "gpu_launches": msm_res["launches"], "clocks": msm_res["clocks"],
                    "roofline": msm_res["roofline"], "roofline_hbm": msm_res["roofline_hbm"], "phases": msm_res["phases"],
                    "cpu_baseline": cpu,
                    "result_xy": msm_res["result_xy"]}
            if strong is not None:
                line["secondary_scaling"] = {
                    "scaling": "strong", "rows": strong,
                    "note": "ONE global MSM per row, the same inputs for every N, cut into contiguous point-range "
                            "shards (rank r owns [r n/N, (r+1) n/N)); result_xy is the same string for every N and "
                            "equals the scalar identity (result_check).  Strong-scaling efficiency = "
                            "value(N) / (N * value(1)) over the driver's N = 1, 2, 4, 8 runs."}
            if ver_res is not None:
                line["secondary"] = ver_res
            if other is not None:
                line["other_ops"] = other
        else:
            line = dict(ver_res)
            line.update({"n_gpus": world, "warmup": args.warmup, "higher_is_better": True, "scaling": "weak",
                         "vs_baseline": None, "dtype": "u32 limbs (IMAD.WIDE.U32), exact integer arithmetic",
                         "data": "synthetic", "cpu_baseline": ver_res.get("cpu_baseline")})
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--workload", default="both", choices=["msm", "verify", "both"])
    ap.add_argument("--log-n", type=int, default=LOG_N_DEFAULT)
    ap.add_argument("--proofs", type=int, default=VERIFY_PROOFS_DEFAULT)
    ap.add_argument("--distinct-proofs", type=int, default=VERIFY_PROOFS_DEFAULT,
                    help="distinct prover-generated records per GPU (default: all distinct, so that table reads "
                         "are as uncorrelated as in production)")
    ap.add_argument("--fixed-window-bits", type=int, default=16, choices=[8, 16],
                    help="window width of the generator tables (8: 51 MB in L2; 16: 6.5 GB in HBM, half the additions)")
    ap.add_argument("--cpu-sample", type=int, default=4096, help="points per host thread for the CPU baseline")
    ap.add_argument("--verify-group", type=int, default=-1,
                    help="proofs per combined identity in batch verification (-1: library default, 0: one by one)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-numa-bind", action="store_true", help="N > 1: leave the ranks where the launcher put them")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling rows (secondary_scaling)")
    ap.add_argument("--strong-log-n", type=int, nargs="*", default=[20, 22],
                    help="global sizes of the strong-scaling MSMs (one MSM cut into N point-range shards)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args)
        return
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun when called directly with --gpus N
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", str(29500 + os.getpid() % 500)] + [os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_cuda(args)


if __name__ == "__main__":
    main()
