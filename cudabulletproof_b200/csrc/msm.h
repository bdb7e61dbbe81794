// msm.h — host-side plan/launcher for the Pippenger MSM (msm.cu).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace cbp {

struct MsmPlan {
    size_t n;           // number of (scalar, point) pairs
    int c;              // window width in bits (signed digits)
    int W;              // number of windows = ceil(256 / c)
    uint32_t B;         // buckets per window = 2^(c-1)
    uint32_t nbuckets;  // W * B
    uint32_t cap;       // entries per fixed bucket slot of the first digit pass (0: every window takes both passes)
    int w_exact;        // windows below this one are slotted, the others placed exactly by the second pass
    int seg_shift;      // log2 of the accumulation segment length
    size_t max_segs;    // upper bound on accumulation segments (buckets + entries / segment length)
    size_t off_ends;    // run ends of the buckets (fused front end)
    size_t off_table, off_counts, off_offsets, off_cursors, off_tiles, off_segoff, off_desc, off_order, off_bins, off_heavy,
        off_entries, off_toprank, off_buckets, off_segsums;
    size_t off_redX[2], off_redY[2], off_winX, off_winY, off_state;
    size_t workspace_bytes;
    // n <= the small-n threshold and no explicit window width: Straus in three launches (msm.cu, section 8)
    bool small;
    size_t off_stables, off_sdigits, off_ssums;
};

int msm_pick_window(size_t n);
void msm_make_plan(MsmPlan* p, size_t n, int c /* 0 = auto */);
// points_ready (optional): event after which d_points may be read (lets the caller overlap the point
// upload with the scalar-only front end: digit recoding, histogram, sort).
// kit_index < kMsmKits selects the set of internal side streams: MSMs that are in flight at the same time
// on one device (the chunked host path) must use different sets and different workspaces.
constexpr int kMsmKits = 3;
// flags, for inputs that arrive in chunks (all chunks share one workspace and one plan, processed in stream
// order): kMsmCarryIn — the bucket sums in the workspace already hold earlier chunks, add this one into them;
// kMsmNoTail — stop once the buckets are updated (no reduction, no result).  The last chunk passes
// kMsmCarryIn alone and produces the result for the whole input.
constexpr int kMsmCarryIn = 1, kMsmNoTail = 2;
// kMsmFrontOnly — only the scalar side (digit recoding, bucket sort, segments): needs no points and leaves its arrays
// in the front part of the workspace; kMsmBackOnly — everything else, on a front part prepared by an earlier
// kMsmFrontOnly call with the same plan.  d_front_workspace (optional, msm_front_bytes(plan) bytes) holds that front
// part instead of the head of d_workspace, so that the sorted scalars of several chunks can wait for their points
// while the chunks share one set of buckets (the chunked host path sorts every chunk while the points upload).
constexpr int kMsmFrontOnly = 4, kMsmBackOnly = 8;
// kMsmAffineXY — d_points holds affine points, 64 bytes each (x || y as fe25519 containers, Z = 1 implied) instead of
// the reference's 128-byte extended ge25519: 96 instead of 160 bytes per pair on the wire (bpk_msm_*_affine)
constexpr int kMsmAffineXY = 16;
inline size_t msm_front_bytes(const MsmPlan& p) { return p.off_buckets; }
int msm_run(const MsmPlan& p, const void* d_scalars, const void* d_points, void* d_result, void* d_workspace,
            int normalize, cudaStream_t stream, int* launches, cudaEvent_t points_ready, int kit_index = 0,
            int flags = 0, void* d_front_workspace = nullptr);

// bpk_msm_device's way in: mid-size calls (2^13 < n < 2^19, latency-bound: 36-70 dependent launches on up to ten streams)
// replay a CUDA graph of msm_run's launch DAG, captured on first use and cached per device by (buffers, n, window width,
// normalize); everything else — and every call while the profiler timers are on — goes straight to msm_run.
int msm_run_cached(const MsmPlan& p, const void* d_scalars, const void* d_points, void* d_result, void* d_workspace,
                   int normalize, cudaStream_t stream, int* launches);

}  // namespace cbp
