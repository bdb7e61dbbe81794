// msm.cu — Pippenger multi-scalar multiplication over ge25519 for sm_100a.
//
// Replaces cuda_point_vector_multi_scalar_mul and kernels K1-K3 of the reference
// (cuda_bulletproof_kernels.cu:26-207: one 256-step double-and-add per thread + a racy tree) and
// is bit-exact (as canonical encodings) with the CPU MSM point_vector_multi_scalar_mul
// (bulletproof_vectors.cu:189-224) as restated by oracle/ref_corrected.c.
//
// Scalar convention (cuda_bulletproof_kernels.cu:33-37): a scalar is an fe25519 container; the
// integer used is k = canonical(fe25519_tobytes(s)) = s mod p, all 255 bits, no reduction mod l —
// so the result is exact for every point of the curve, including points with a torsion component.
//
// Pipeline (no host synchronisation; from 2^15 points the windows go top-down in groups whose reductions and
// Horner steps run on side streams under the next group's accumulation):
//   1. msm_precompute   AoS ge25519 -> 96-byte affine (y+x, y-x, 2dxy) table   [HBM streaming, side stream]
//   2. msm_digits<0>    signed c-bit digit recoding + per-(window,bucket) counts; from 2^19 points the same pass
//                       places every window but the top one into fixed bucket slots and keeps the top window's ranks
//   3. scan             exclusive prefix sum over the counters of the compactly placed buckets
//   4. msm_digits<1>    exact placement: the top window by rank (or every window: small inputs, slot overflow)
//   4b. seg_*           runs cut into segments, segments ordered by (window group, length)
//   5. msm_accumulate   one thread per segment: gathers its run, 7M mixed additions  [IMAD-bound, ~2/3 of the time]
//   6. msm_reduce2d_*   sum_b b*B_b per window: row / column sums, per-bit subset sums, bit Horner (c >= 9);
//      msm_reduce_*     work-efficient running-sum levels (narrow windows; cross-check)
//   7. msm_horner       Horner over the windows (chain state kept between groups), normalise to Z = 1
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include <cooperative_groups.h>
#include "fe8.cuh"
#include "msm.h"
#include "common.h"

namespace cbp {

static constexpr int kReduceM = 8;  // buckets per thread per reduction level

// ---- scalar handling ------------------------------------------------------------------------
__device__ __forceinline__ void load_scalar_canon(uint32_t (&k)[8], const void* p) {
    fe t;
    fe_load(t, p);
    fe_canon(t);
#pragma unroll
    for (int i = 0; i < 8; i++) k[i] = t.v[i];
}
__device__ __forceinline__ uint32_t raw_digit(const uint32_t (&k)[8], int w, int c) {
    int bit = w * c;
    if (bit >= 256) return 0;
    int word = bit >> 5, sh = bit & 31;
    uint64_t v = k[word];
    if (word + 1 < 8) v |= (uint64_t)k[word + 1] << 32;
    return (uint32_t)(v >> sh) & ((1u << c) - 1u);
}

// ---- 1. points -> affine precomputed table ------------------------------------------------------
// Each thread converts kPreChunk consecutive points.  Points with Z == 1 (what the reference's CPU
// code produces everywhere, since it normalises after every operation) take the fast path; others
// share one inversion per chunk (Montgomery's trick).
static constexpr int kPreChunk = 4;
__global__ void __launch_bounds__(128) msm_precompute_kernel(const uint8_t* __restrict__ points, size_t n,
                                                             uint8_t* __restrict__ table) {
    size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * kPreChunk;
    if (base >= n) return;
    int cnt = (int)((n - base) < (size_t)kPreChunk ? (n - base) : kPreChunk);
    fe Z[kPreChunk], pre[kPreChunk];
    bool all_one = true;
    fe one;
    fe_set1(one);
    for (int i = 0; i < cnt; i++) {
        fe_load(Z[i], points + (base + i) * 128 + 64);
        all_one = all_one && fe_equal(Z[i], one);
    }
    fe inv;
    if (!all_one) {
        fe acc;
        fe_set1(acc);
        for (int i = 0; i < cnt; i++) {
            pre[i] = acc;
            fe_mul(acc, acc, Z[i]);
        }
        fe_invert(inv, acc);
    }
    for (int i = cnt - 1; i >= 0; i--) {
        fe x, y;
        fe_load(x, points + (base + i) * 128);
        fe_load(y, points + (base + i) * 128 + 32);
        if (!all_one) {
            fe zi;
            fe_mul(zi, inv, pre[i]);
            fe_mul(inv, inv, Z[i]);
            fe_mul(x, x, zi);
            fe_mul(y, y, zi);
        }
        ge_niels q;
        ge_to_niels_affine(q, x, y);
        ge_niels_store(table + (base + i) * 96, q);
    }
}

// the same from affine input (x || y, 64 bytes per point): one point per thread, no inversion
__global__ void __launch_bounds__(128) msm_precompute_affine_kernel(const uint8_t* __restrict__ xy, size_t n,
                                                                    uint8_t* __restrict__ table) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe x, y;
    fe_load(x, xy + i * 64);
    fe_load(y, xy + i * 64 + 32);
    ge_niels q;
    ge_to_niels_affine(q, x, y);
    ge_niels_store(table + i * 96, q);
}

// ---- 2./4. digit recoding: histogram and scatter ------------------------------------------------
// Measured and rejected: a two-level counting sort with shared-memory atomics only (keys first partitioned by
// (window, high 5 bucket bits) with one global atomic per (CTA, partition), then one CTA sorts each partition by
// the low 10 bits).  Count + partition take 0.12 ms, but one CTA per partition is hostage to the distribution:
// the top window of 253-bit scalars fills 4 of its 32 partitions with 8x the mean, and the sort kernel takes
// 0.39 ms (front end 0.66 instead of 0.49 ms).  Slicing heavy partitions needs a second reserve/scan round and
// would save < 0.1 ms over the two atomic passes below.
// signed digits d_w in [-(2^(c-1)-1), 2^(c-1)], sum d_w 2^(cw) = k; bucket index |d|-1.
// Two passes over the scalars (digits are recomputed, never stored):
//  PASS 0  counts every (window, bucket) pair.  With cap > 0, windows below w_exact are at the same time
//          scattered into FIXED slots of `cap` entries per bucket (entries[id * cap + pos]; counts[] is the
//          cursor), which spares them the second atomic pass: the digits of every window but the top one are
//          uniform for any scalars that are not adversarial, so a slot of mean + 8 sigma never fills up.  A
//          bucket that does outgrow its slot raises *overflow (its count stays exact).
//          The top window (w_exact = W - 1; scalars below the group order populate 1/16 of its buckets at 16x
//          the mean, so it gets no slots) is only counted, but the value the atomic returns — the entry's rank
//          inside its bucket — is kept per point (toprank: bucket | sign << 31, rank).
//  PASS 1  exact placement into compact runs.  With slots: the top window, at offsets[bucket] + rank, without
//          atomics or a second look at the scalar (0.04 -> 0.01 ms at 2^20: its 4096 populated buckets made the
//          atomics contend) — and, only when *overflow is set, the slotted windows again through the scanned
//          cursors (the slots of pass 0 are then ignored).  Without slots (cap = 0): every window.
template <int PASS>
__global__ void __launch_bounds__(256) msm_digits_kernel(const uint8_t* __restrict__ scalars, size_t n, int c, int W,
                                                         uint32_t B, uint32_t cap, int w_exact,
                                                         uint32_t* __restrict__ counters,
                                                         uint32_t* __restrict__ entries,
                                                         uint32_t* __restrict__ overflow,
                                                         const uint32_t* __restrict__ offsets,
                                                         uint2* __restrict__ toprank) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int w_to = W;  // windows [0, w_to) go through the loop below
    if (PASS == 1 && toprank) {
        const uint2 r = toprank[i];
        if (r.x != 0xffffffffu) entries[offsets[r.x & 0x7fffffffu] + r.y] = ((uint32_t)i << 1) | (r.x >> 31);
        if (!*overflow) return;
        w_to = w_exact;
    }
    uint2 top = make_uint2(0xffffffffu, 0u);
    uint32_t k[8];
    load_scalar_canon(k, scalars + i * 32);
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    // windows in batches of 4: the 4 atomics of a batch are issued back to back (independent), only then
    // are their return values consumed — the scatter pass is bound by atomic round-trip latency otherwise
    for (int w0 = 0; w0 < W; w0 += 4) {
        uint32_t id[4], val[4];
        bool live[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int w = w0 + j;
            live[j] = false;
            if (w < W) {
                uint32_t d = raw_digit(k, w, c) + carry;
                uint32_t neg = 0;
                carry = 0;
                if (d > half) {
                    d = (1u << c) - d;
                    neg = 1;
                    carry = 1;
                }
                if (d != 0 && w < w_to) {
                    live[j] = true;
                    id[j] = (uint32_t)w * B + (d - 1);
                    val[j] = ((uint32_t)i << 1) | neg;
                }
            }
        }
        if (PASS == 1) {
            uint32_t pos[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (live[j]) pos[j] = atomicAdd(&counters[id[j]], 1u);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (live[j]) entries[pos[j]] = val[j];
        } else if (toprank) {  // slots: the atomics' return values place (slotted windows) or rank (top window)
            uint32_t pos[4];
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (live[j]) pos[j] = atomicAdd(&counters[id[j]], 1u);
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (live[j]) {
                    if (w0 + j >= w_exact) top = make_uint2(id[j] | (val[j] << 31), pos[j]);
                    else if (pos[j] < cap) entries[id[j] * cap + pos[j]] = val[j];
                    else *overflow = 1u;
                }
        } else {
#pragma unroll
            for (int j = 0; j < 4; j++)
                if (live[j]) atomicAdd(&counters[id[j]], 1u);
        }
    }
    if (PASS == 0 && toprank) toprank[i] = top;
}

static constexpr int kMaxGroups = 8;
struct GroupMap {
    int ngroups;
    int log2B;                               // buckets per window = 2^log2B
    int w_hi[kMaxGroups], w_lo[kMaxGroups];  // processed top window first
    uint8_t group_of_window[64];
    uint8_t seg_shift_of_window[64];         // log2 segment length; smaller for the small last groups so that
                                             // their accumulation still fills the machine
};

// ---- 3. exclusive scan over `total` counters (three small kernels) -------------------------------
static constexpr int kScanThreads = 256, kScanPer = 16, kScanTile = kScanThreads * kScanPer;
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* smem, uint32_t& block_total) {
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) smem[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint32_t s = lane < (blockDim.x >> 5) ? smem[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, s, o);
            if (lane >= o) s += y;
        }
        smem[32 + lane] = s;
    }
    __syncthreads();
    uint32_t warp_off = wid ? smem[32 + wid - 1] : 0;
    block_total = smem[32 + (blockDim.x >> 5) - 1];
    return warp_off + x - v;
}
// SEG = true scans the per-bucket SEGMENT counts ceil(count / kSegLen) (>= 1) instead of the counts
__device__ __forceinline__ uint32_t seg_count(uint32_t c, int seg_shift) {
    return c ? (c + (1u << seg_shift) - 1u) >> seg_shift : 1u;
}
// SEG = false scans the entry counts of the buckets that are placed compactly: all of them, or — when the
// first digit pass slotted the windows below w_exact and no slot overflowed — the ids from `slotted` on.
template <bool SEG>
__device__ __forceinline__ uint32_t scan_input(uint32_t c, uint32_t id, const GroupMap& gm, uint32_t slotted) {
    return SEG ? seg_count(c, gm.seg_shift_of_window[id >> gm.log2B]) : (id >= slotted ? c : 0u);
}
template <bool SEG>
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint32_t total,
                                                                      GroupMap gm, uint32_t slotted_ids,
                                                                      const uint32_t* __restrict__ overflow,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t smem[64];
    const uint32_t slotted = (!SEG && slotted_ids && !*overflow) ? slotted_ids : 0u;
    uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer, s = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
        if (base + j < total) s += scan_input<SEG>(in[base + j], base + j, gm, slotted);
    uint32_t bt;
    block_exclusive_scan(s, smem, bt);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = bt;
}
__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t* tile_sums, uint32_t ntiles) {
    // ntiles <= 1024 (W*B <= 2^22 counters)
    __shared__ uint32_t smem[64];
    uint32_t v = threadIdx.x < ntiles ? tile_sums[threadIdx.x] : 0, bt;
    uint32_t ex = block_exclusive_scan(v, smem, bt);
    if (threadIdx.x < ntiles) tile_sums[threadIdx.x] = ex;
}
template <bool SEG>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t total,
                                                                  GroupMap gm, uint32_t slotted_ids, uint32_t cap,
                                                                  uint32_t ranked_from,
                                                                  const uint32_t* __restrict__ overflow,
                                                                  const uint32_t* __restrict__ tile_sums,
                                                                  uint32_t* __restrict__ offsets,
                                                                  uint32_t* __restrict__ cursors) {
    __shared__ uint32_t smem[64];
    const uint32_t slotted = (!SEG && slotted_ids && !*overflow) ? slotted_ids : 0u;
    uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    uint32_t v[kScanPer], s = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) {
        v[j] = base + j < total ? scan_input<SEG>(in[base + j], base + j, gm, slotted) : 0;
        s += v[j];
    }
    uint32_t bt;
    // the compact runs start behind the slot area
    uint32_t ex = block_exclusive_scan(s, smem, bt) + tile_sums[blockIdx.x] + slotted * cap;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) {
        const uint32_t id = base + j;
        if (!SEG && id < slotted) {  // already placed by the first digit pass: run = [id * cap, id * cap + count)
            offsets[id] = id * cap;
            cursors[id] = id * cap + in[id];
        } else {
            if (id <= total) offsets[id] = ex;  // offsets[total] = end of the compact area (sentinel)
            // the placing pass advances the cursors: afterwards cursors[id] is the END of bucket id's run
            // (buckets from ranked_from on are placed by rank, without cursors: the end is written here)
            if (cursors && id < total) cursors[id] = ex + (id >= ranked_from ? v[j] : 0u);
        }
        ex += v[j];
    }
}

// ---- warp-level point helpers ---------------------------------------------------------------------------
__device__ __forceinline__ void ge_shfl_down(ge_p3& out, const ge_p3& in, int delta) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        out.X.v[j] = __shfl_down_sync(0xffffffffu, in.X.v[j], delta);
        out.Y.v[j] = __shfl_down_sync(0xffffffffu, in.Y.v[j], delta);
        out.Z.v[j] = __shfl_down_sync(0xffffffffu, in.Z.v[j], delta);
        out.T.v[j] = __shfl_down_sync(0xffffffffu, in.T.v[j], delta);
    }
}
__device__ __forceinline__ void ge_warp_sum(ge_p3& v) {  // result in lane 0
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_p3 other;
        ge_shfl_down(other, v, o);
        ge_add(v, v, other);
    }
}

// ---- 4b. segments: split long runs, order by length -----------------------------------------------------
// A bucket's run of entries is cut into segments of at most kSegLen entries; one thread accumulates one
// segment.  This bounds the serial work per thread (adversarial inputs: all scalars equal; ordinary
// inputs: the top window of 252/253-bit scalars has few distinct digits, and the carry bucket of a
// partially filled top window receives half of all points) and, because segments are then ordered by
// length (counting sort, longest first, grouped by window group), the 32 segments of a warp run in
// lockstep instead of waiting for the longest of 32 Poisson-distributed runs.
// segment length = max(64, 2 * mean run length), a power of two (msm_make_plan)
static constexpr int kSegBinsPerGroup = 65;  // length classes ceil(64 len / seglen), longest first
static constexpr int kSegBins = 1024;  // >= kMaxGroups * kSegBinsPerGroup, = threads of the bin scan


__global__ void __launch_bounds__(256) seg_build_kernel(const uint32_t* __restrict__ counts,
                                                        const uint32_t* __restrict__ offsets,
                                                        const uint32_t* __restrict__ segoff, uint32_t nbuckets,
                                                        uint32_t B, GroupMap gm, uint2* __restrict__ desc,
                                                        uint32_t* __restrict__ heavy, uint32_t* __restrict__ heavy_cnt) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= nbuckets) return;
    const int seg_shift = gm.seg_shift_of_window[id >> gm.log2B];
    uint32_t c = counts[id], ns = seg_count(c, seg_shift), so = segoff[id], start = offsets[id];
    for (uint32_t q = 0; q < ns; q++) desc[so + q] = make_uint2(start + (q << seg_shift), id);
    if (ns > 1) {
        int g = gm.group_of_window[id >> gm.log2B];
        uint32_t idx = atomicAdd(&heavy_cnt[g], 1u);
        heavy[(uint32_t)gm.w_lo[g] * B + idx] = id;
    }
}
// `ends` = the cursors after the placing pass (end of each bucket's run; runs need not be adjacent)
__device__ __forceinline__ uint32_t seg_length(uint2 d, const uint32_t* __restrict__ ends, const GroupMap& gm) {
    uint32_t rest = ends[d.y] - d.x, cap = 1u << gm.seg_shift_of_window[d.y >> gm.log2B];
    return rest < cap ? rest : cap;
}
__device__ __forceinline__ uint32_t seg_bin(uint2 d, const uint32_t* __restrict__ offsets, const GroupMap& gm) {
    uint32_t len = seg_length(d, offsets, gm), w = d.y >> gm.log2B;
    const int seg_shift = gm.seg_shift_of_window[w];
    uint32_t cls = (len * 64u + (1u << seg_shift) - 1u) >> seg_shift;  // 0..64
    return (uint32_t)gm.group_of_window[w] * kSegBinsPerGroup + (64u - cls);  // longest first within a group
}
__global__ void __launch_bounds__(256) seg_hist_kernel(const uint2* __restrict__ desc,
                                                       const uint32_t* __restrict__ offsets,
                                                       const uint32_t* __restrict__ nsegs_p, GroupMap gm,
                                                       uint32_t* __restrict__ hist) {
    __shared__ uint32_t sh[kSegBins];
    for (int i = threadIdx.x; i < kSegBins; i += blockDim.x) sh[i] = 0;
    __syncthreads();
    uint32_t nsegs = *nsegs_p;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nsegs; i += gridDim.x * blockDim.x)
        atomicAdd(&sh[seg_bin(desc[i], offsets, gm)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < kSegBins; i += blockDim.x)
        if (sh[i]) atomicAdd(&hist[i], sh[i]);
}
// exclusive scan of the bins in place; binstart (a copy) keeps the group boundaries for the accumulate launches
__global__ void __launch_bounds__(kSegBins) seg_bin_scan_kernel(uint32_t* __restrict__ hist,
                                                                uint32_t* __restrict__ binstart) {
    __shared__ uint32_t smem[64];
    uint32_t v = hist[threadIdx.x], bt;
    uint32_t ex = block_exclusive_scan(v, smem, bt);
    hist[threadIdx.x] = ex;
    binstart[threadIdx.x] = ex;
    if (threadIdx.x == 0) binstart[kSegBins] = bt;
}
__global__ void __launch_bounds__(256) seg_scatter_kernel(const uint2* __restrict__ desc,
                                                          const uint32_t* __restrict__ offsets,
                                                          const uint32_t* __restrict__ nsegs_p, GroupMap gm,
                                                          uint32_t* __restrict__ cursors,
                                                          uint32_t* __restrict__ order) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nsegs = *nsegs_p;
    bool live = i < nsegs;
    uint32_t bin = live ? seg_bin(desc[i], offsets, gm) : 0xffffffffu;
    uint32_t active = __ballot_sync(0xffffffffu, live);
    if (!live) return;
    uint32_t peers = __match_any_sync(active, bin);  // warp-aggregated atomics
    int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (lane == leader) base = atomicAdd(&cursors[bin], __popc(peers));
    base = __shfl_sync(peers, base, leader);
    order[base + __popc(peers & ((1u << lane) - 1u))] = i;
}

// ---- 4c. everything between the first digit pass and the accumulation in ONE cooperative launch ----------------
// scan_tile_sums / scan_tiles / scan_apply (twice: entry offsets, segment offsets), seg_build, seg_hist, seg_bin_scan and
// seg_scatter are eleven launches of 3-20 us each, and a 2^20-point MSM spent 0.18 ms of its 0.39 ms front end in them
// and in the gaps between them (a 2^16-point one 0.08 of 0.10 ms).  Here they are the phases of one kernel whose CTAs
// are all resident (cooperative launch) and meet at four grid barriers:
//   A  per tile of 4096 buckets: sum of the compactly placed entry counts and of the segment counts
//   B  (CTA 0) exclusive scan of both rows of tile sums
//   C  per tile: offsets, cursors and run ends of every bucket, its segment offset, its segment descriptors, the list
//      of split buckets, and the histogram of (window group, length class) of its segments
//   D  (CTA 0) exclusive scan of the histogram
//   E  segments scattered into their bins (warp-aggregated atomics)
// The run ends — what the placing pass would leave in the cursors — are offsets + counts, so the placing pass
// (msm_digits_kernel<1>) runs AFTER this kernel and nothing here waits for it.  The separate kernels above remain as
// the fallback for devices without cooperative launch and as a cross-check (BPK_OPT_MSM_FUSED_FRONT = 0).
struct FrontTail {
    const uint32_t* counts;
    const uint32_t* overflow;
    uint32_t* tilesA;
    uint32_t* tilesS;
    uint32_t* offsets;
    uint32_t* cursors;
    uint32_t* ends;
    uint32_t* segoff;
    uint2* desc;
    uint32_t* heavy;
    uint32_t* heavy_cnt;
    uint32_t* hist;
    uint32_t* binstart;
    uint32_t* order;
    uint32_t total, B, slotted_ids, cap, ranked_from;
};
__device__ __forceinline__ uint32_t seg_bin_of(uint32_t len, uint32_t w, const GroupMap& gm) {
    const int seg_shift = gm.seg_shift_of_window[w];
    uint32_t cls = (len * 64u + (1u << seg_shift) - 1u) >> seg_shift;  // 0..64
    return (uint32_t)gm.group_of_window[w] * kSegBinsPerGroup + (64u - cls);
}
// exclusive scan of a[0..n) in place by one CTA (n <= 4 * blockDim.x * ... : looped), total returned
__device__ __forceinline__ uint32_t cta_scan_inplace(uint32_t* a, uint32_t n, uint32_t* smem) {
    uint32_t carry = 0;
    for (uint32_t base = 0; base < n; base += blockDim.x * 4) {
        const uint32_t i0 = base + threadIdx.x * 4;
        uint32_t v[4], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            v[j] = i0 + j < n ? a[i0 + j] : 0;
            sum += v[j];
        }
        uint32_t bt;
        uint32_t ex = block_exclusive_scan(sum, smem, bt) + carry;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i0 + j < n) a[i0 + j] = ex;
            ex += v[j];
        }
        carry += bt;
        __syncthreads();
    }
    return carry;
}
static constexpr int kFtPer = 4, kFtTile = kScanThreads * kFtPer;  // buckets per thread / per tile of the fused kernel (16: 61 us at 2^20)
__global__ void __launch_bounds__(kScanThreads) msm_front_tail_kernel(FrontTail f, GroupMap gm) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ uint32_t smemA[64], smemS[64];
    __shared__ uint32_t sh_hist[kSegBins];
    const uint32_t total = f.total;
    const uint32_t slotted = (f.slotted_ids && !*f.overflow) ? f.slotted_ids : 0u;
    const uint32_t ntiles = (total + 1 + kFtTile - 1) / kFtTile;
    // ---- A
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t base = tile * kFtTile + threadIdx.x * kFtPer;
        uint32_t sA = 0, sS = 0;
#pragma unroll
        for (int j = 0; j < kFtPer; j++) {
            const uint32_t id = base + j;
            if (id < total) {
                const uint32_t c = f.counts[id];
                sA += id >= slotted ? c : 0u;
                sS += seg_count(c, gm.seg_shift_of_window[id >> gm.log2B]);
            }
        }
        uint32_t btA, btS;
        block_exclusive_scan(sA, smemA, btA);
        block_exclusive_scan(sS, smemS, btS);
        if (threadIdx.x == 0) {
            f.tilesA[tile] = btA;
            f.tilesS[tile] = btS;
        }
        __syncthreads();
    }
    grid.sync();
    // ---- B
    if (blockIdx.x == 0) {
        cta_scan_inplace(f.tilesA, ntiles, smemA);
        cta_scan_inplace(f.tilesS, ntiles, smemS);
    }
    grid.sync();
    // ---- C
    for (int i = threadIdx.x; i < kSegBins; i += blockDim.x) sh_hist[i] = 0;
    __syncthreads();
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const uint32_t base = tile * kFtTile + threadIdx.x * kFtPer;
        uint32_t cnt[kFtPer], sA = 0, sS = 0;
#pragma unroll
        for (int j = 0; j < kFtPer; j++) {
            const uint32_t id = base + j;
            cnt[j] = id < total ? f.counts[id] : 0u;
            if (id < total) {
                sA += id >= slotted ? cnt[j] : 0u;
                sS += seg_count(cnt[j], gm.seg_shift_of_window[id >> gm.log2B]);
            }
        }
        uint32_t btA, btS;
        uint32_t exA = block_exclusive_scan(sA, smemA, btA) + f.tilesA[tile] + slotted * f.cap;  // compact runs start behind the slots
        uint32_t exS = block_exclusive_scan(sS, smemS, btS) + f.tilesS[tile];
#pragma unroll
        for (int j = 0; j < kFtPer; j++) {
            const uint32_t id = base + j;
            if (id == total) {  // sentinels: end of the compact area, number of segments
                f.offsets[id] = exA;
                f.segoff[id] = exS;
            }
            if (id >= total) continue;
            const uint32_t c = cnt[j];
            uint32_t start;
            if (id < slotted) {  // placed by the first digit pass: run = [id * cap, id * cap + count)
                start = id * f.cap;
                f.cursors[id] = start + c;
            } else {
                start = exA;
                // the placing pass advances the cursors of the buckets it places through atomics (from the start of the run);
                // buckets from ranked_from on are placed by rank: their cursor is the end already
                f.cursors[id] = exA + (id >= f.ranked_from ? c : 0u);
                exA += c;
            }
            f.offsets[id] = start;
            f.ends[id] = start + c;
            // segments of this bucket
            const uint32_t w = id >> gm.log2B;
            const int seg_shift = gm.seg_shift_of_window[w];
            const uint32_t ns = seg_count(c, seg_shift), seglen = 1u << seg_shift;
            f.segoff[id] = exS;
            for (uint32_t q = 0; q < ns; q++) {
                f.desc[exS + q] = make_uint2(start + (q << seg_shift), id);
                const uint32_t rest = c - (q << seg_shift);
                atomicAdd(&sh_hist[seg_bin_of(rest < seglen ? rest : seglen, w, gm)], 1u);
            }
            if (ns > 1) {
                const int g = gm.group_of_window[w];
                const uint32_t idx = atomicAdd(&f.heavy_cnt[g], 1u);
                f.heavy[(uint32_t)gm.w_lo[g] * f.B + idx] = id;
            }
            exS += ns;
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kSegBins; i += blockDim.x)
        if (sh_hist[i]) atomicAdd(&f.hist[i], sh_hist[i]);
    grid.sync();
    // ---- D
    if (blockIdx.x == 0) {
        const uint32_t nsegs = cta_scan_inplace(f.hist, kSegBins, smemA);
        for (int i = threadIdx.x; i < kSegBins; i += blockDim.x) f.binstart[i] = f.hist[i];
        if (threadIdx.x == 0) f.binstart[kSegBins] = nsegs;
    }
    grid.sync();
    // ---- E
    const uint32_t nsegs = f.segoff[total];
    const uint32_t rounded = (nsegs + 31u) & ~31u;  // whole warps take part in the ballots
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < rounded; i += gridDim.x * blockDim.x) {
        const bool live = i < nsegs;
        const uint32_t bin = live ? seg_bin(f.desc[i], f.ends, gm) : 0xffffffffu;
        const uint32_t active = __ballot_sync(0xffffffffu, live);
        if (live) {
            const uint32_t peers = __match_any_sync(active, bin);
            const int leader = __ffs(peers) - 1, lane = threadIdx.x & 31;
            uint32_t b0 = 0;
            if (lane == leader) b0 = atomicAdd(&f.hist[bin], __popc(peers));
            b0 = __shfl_sync(peers, b0, leader);
            f.order[b0 + __popc(peers & ((1u << lane) - 1u))] = i;
        }
    }
}

// ---- 5. bucket accumulation ---------------------------------------------------------------------
// Measured and rejected: processing the points in 2-16 tiles whose tables stay L2-resident through all
// windows (tiles add into the same buckets with the carry-in path): 2.31 / 2.44 / 2.78 / 3.45 ms for
// 1 / 2 / 4 / 8 tiles at 2^20 — the gathers are not what limits this kernel, the per-tile sort is pure cost.
// One thread per segment of group g (order[] range from binstart).  7M mixed additions from the
// 96-byte affine table, next operand prefetched while the current addition runs.
__global__ void __launch_bounds__(128, 4)
    msm_accumulate_kernel(const uint8_t* __restrict__ table, const uint32_t* __restrict__ entries,
                          const uint2* __restrict__ desc, const uint32_t* __restrict__ offsets,
                          const uint32_t* __restrict__ segoff, GroupMap gm, const uint32_t* __restrict__ order,
                          const uint32_t* __restrict__ binstart, int group, int carry,
                          uint8_t* __restrict__ bucket_sums, uint8_t* __restrict__ seg_sums) {
    uint32_t lo = binstart[group * kSegBinsPerGroup], hi = binstart[(group + 1) * kSegBinsPerGroup];
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= hi - lo) return;
    uint32_t seg = order[lo + t];
    uint2 d = desc[seg];
    uint32_t e = d.x, end = d.x + seg_length(d, offsets, gm), id = d.y;
    const bool single = segoff[id + 1] - segoff[id] == 1;
    ge_p3 acc;
    // carry: the bucket already holds the sum over earlier chunks of the input (chunked host path)
    if (carry && single) ge_load(acc, bucket_sums + (size_t)id * 128);
    else ge_p3_0(acc);
    if (e < end) {
        uint32_t ent = __ldg(entries + e);
        ge_niels q;
        ge_niels_load(q, table + (size_t)(ent >> 1) * 96);
        for (;;) {
            uint32_t neg = ent & 1u;
            ge_niels cur = q;
            ++e;
            if (e < end) {
                ent = __ldg(entries + e);
                ge_niels_load(q, table + (size_t)(ent >> 1) * 96);
            }
            ge_madd(acc, acc, cur, neg != 0);
            if (e >= end) break;
        }
    }
    if (carry && single && d.x == end) return;  // nothing new for this bucket
    ge_store(single ? bucket_sums + (size_t)id * 128 : seg_sums + (size_t)seg * 128, acc);
}
// buckets cut into several segments: a thread adds up to kHeavySeq segment sums itself, larger buckets
// are left to one warp each (lane-strided partial sums + shuffle tree)
static constexpr uint32_t kHeavySeq = 8;
__global__ void __launch_bounds__(128) msm_heavy_small_kernel(const uint32_t* __restrict__ heavy,
                                                              const uint32_t* __restrict__ heavy_cnt, int group,
                                                              uint32_t heavy_base, const uint32_t* __restrict__ segoff,
                                                              const uint8_t* __restrict__ seg_sums, int carry,
                                                              uint8_t* __restrict__ bucket_sums) {
    uint32_t cnt = heavy_cnt[group];
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
        uint32_t id = heavy[heavy_base + i];
        uint32_t s0 = segoff[id], s1 = segoff[id + 1];
        if (s1 - s0 > kHeavySeq) continue;
        ge_p3 acc;
        ge_load(acc, seg_sums + (size_t)s0 * 128);
        if (carry) {
            ge_p3 old;
            ge_load(old, bucket_sums + (size_t)id * 128);
            ge_add(acc, acc, old);
        }
        for (uint32_t q = s0 + 1; q < s1; q++) {
            ge_p3 x;
            ge_load(x, seg_sums + (size_t)q * 128);
            ge_add(acc, acc, x);
        }
        ge_store(bucket_sums + (size_t)id * 128, acc);
    }
}
__global__ void __launch_bounds__(128) msm_heavy_fix_kernel(const uint32_t* __restrict__ heavy,
                                                            const uint32_t* __restrict__ heavy_cnt, int group,
                                                            uint32_t heavy_base, const uint32_t* __restrict__ segoff,
                                                            const uint8_t* __restrict__ seg_sums, int carry,
                                                            uint8_t* __restrict__ bucket_sums) {
    uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    uint32_t cnt = heavy_cnt[group];
    for (; gw < cnt; gw += (gridDim.x * blockDim.x) >> 5) {
        uint32_t id = heavy[heavy_base + gw];
        uint32_t s0 = segoff[id], s1 = segoff[id + 1];
        if (s1 - s0 <= kHeavySeq) continue;  // warp-uniform
        ge_p3 acc;
        ge_p3_0(acc);
        for (uint32_t q = s0 + lane; q < s1; q += 32) {
            ge_p3 x;
            ge_load(x, seg_sums + (size_t)q * 128);
            ge_add(acc, acc, x);
        }
        ge_warp_sum(acc);
        if (lane == 0) {
            if (carry) {
                ge_p3 old;
                ge_load(old, bucket_sums + (size_t)id * 128);
                ge_add(acc, acc, old);
            }
            ge_store(bucket_sums + (size_t)id * 128, acc);
        }
    }
}

// ---- 6. running-sum reduction -------------------------------------------------------------------
// One level maps n pairs (X_j, Y_j) per window to ceil(n/m) pairs, preserving
//     sum_j (j+1) X_j + sum_j Y_j :
//   R = sum X_j, T = sum_{i=1..m} i X_{tm+i-1}, U = sum Y_j  over the chunk,
//   X'_t = m R,  Y'_t = T + U - m R.      (m is a power of two: m R by doublings)
// Level 0 has no Y (has_y = 0).  Chunks past the end read as identity.
__global__ void __launch_bounds__(128) msm_reduce_level_kernel(const uint8_t* __restrict__ Xin,
                                                               const uint8_t* __restrict__ Yin, uint32_t n_in,
                                                               uint32_t n_out, int W, int has_y, uint32_t in_stride,
                                                               uint32_t out_stride, uint8_t* __restrict__ Xout,
                                                               uint8_t* __restrict__ Yout) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_out * (uint32_t)W) return;
    uint32_t w = g / n_out, t = g % n_out;
    const uint8_t* xb = Xin + (size_t)w * in_stride * 128;
    const uint8_t* yb = Yin + (size_t)w * in_stride * 128;
    ge_p3 run, tot, u;
    ge_p3_0(run);
    ge_p3_0(tot);
    ge_p3_0(u);
    for (int i = kReduceM - 1; i >= 0; i--) {
        uint32_t j = t * kReduceM + i;
        if (j < n_in) {
            ge_p3 x;
            ge_load(x, xb + (size_t)j * 128);
            ge_add(run, run, x);
            if (has_y) {
                ge_p3 y;
                ge_load(y, yb + (size_t)j * 128);
                ge_add(u, u, y);
            }
        }
        ge_add(tot, tot, run);
    }
    ge_p3 mr = run;
#pragma unroll 1
    for (int s = 1; s < kReduceM; s <<= 1) ge_dbl(mr, mr);
    ge_p3 nmr;
    ge_neg(nmr, mr);
    ge_add(tot, tot, u);
    ge_add(tot, tot, nmr);
    ge_store(Xout + ((size_t)w * out_stride + t) * 128, mr);
    ge_store(Yout + ((size_t)w * out_stride + t) * 128, tot);
}

// ---- 6b. warp-cooperative reduction level -------------------------------------------------------------
// Same invariant as msm_reduce_level_kernel, but one WARP consumes 32*m pairs: each lane runs the
// sequential running sum over its m pairs, then the 32 lane results are combined with shuffles:
//   R = sum_l run_l (inclusive suffix scan S_l, R = S_0),  sum_l l*run_l = sum_{l>=1} S_l,
//   T = sum_l tot_l + m * sum_l l*run_l,  U = sum_l u_l.
// Depth ~ 3m + 20 point operations for a 32m-fold reduction: used for the upper (latency-bound) levels.
__global__ void __launch_bounds__(128) msm_reduce_warp_kernel(const uint8_t* __restrict__ Xin,
                                                              const uint8_t* __restrict__ Yin, uint32_t n_in,
                                                              uint32_t n_out, int nwin, int has_y, int m,
                                                              uint32_t in_stride, uint32_t out_stride,
                                                              uint8_t* __restrict__ Xout, uint8_t* __restrict__ Yout) {
    uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (gw >= n_out * (uint32_t)nwin) return;  // whole warp
    uint32_t w = gw / n_out, t = gw % n_out;
    const uint8_t* xb = Xin + (size_t)w * in_stride * 128;
    const uint8_t* yb = Yin + (size_t)w * in_stride * 128;
    ge_p3 run, tot, u;
    ge_p3_0(run);
    ge_p3_0(tot);
    ge_p3_0(u);
    uint32_t j0 = (t * 32 + lane) * (uint32_t)m;
    for (int i = m - 1; i >= 0; i--) {
        uint32_t j = j0 + i;
        if (j < n_in) {
            ge_p3 x;
            ge_load(x, xb + (size_t)j * 128);
            ge_add(run, run, x);
            if (has_y) {
                ge_p3 y;
                ge_load(y, yb + (size_t)j * 128);
                ge_add(u, u, y);
            }
        }
        if (m > 1) ge_add(tot, tot, run);
    }
    if (m == 1) tot = run;
    // inclusive suffix scan of run over the lanes
    ge_p3 S = run;
#pragma unroll 1
    for (int o = 1; o < 32; o <<= 1) {
        ge_p3 other;
        ge_shfl_down(other, S, o);
        if (lane + o < 32) ge_add(S, S, other);
    }
    ge_p3 V = S;
    if (lane == 0) ge_p3_0(V);
    ge_warp_sum(V);  // sum_{l>=1} S_l = sum_l l * run_l
    ge_p3 Csum;
    ge_add(Csum, tot, u);
    ge_warp_sum(Csum);  // sum_l (tot_l + u_l)
    if (lane == 0) {
        for (int q = 1; q < m; q <<= 1) ge_dbl(V, V);  // m * sum l run_l
        ge_add(Csum, Csum, V);                         // T + U
        ge_p3 MR = S;                                  // S_0 = R
        for (int q = 1; q < 32 * m; q <<= 1) ge_dbl(MR, MR);
        ge_p3 nMR;
        ge_neg(nMR, MR);
        ge_add(Csum, Csum, nMR);
        ge_store(Xout + ((size_t)w * out_stride + t) * 128, MR);
        ge_store(Yout + ((size_t)w * out_stride + t) * 128, Csum);
    }
}

// ---- 6c. quad-cooperative versions of the two reduction kernels -------------------------------------------
// For the small window groups at the end of the pipeline the reduction is pure latency (a few thousand
// threads, ~90 dependent point operations, and it is the exposed tail of the whole MSM).  These variants
// run every point operation on a quad of lanes (ge_add_quad / ge_dbl_quad: 3 resp. 2 multiplication depths
// instead of 9 / 8; measured 0.99 / 0.72 us against 2.21 / 1.67 us per dependent operation).
// Logical thread = quad; all 32 lanes of a warp stay converged (full-mask shuffles inside the quad ops).
// Measured and rejected: out-of-line (__noinline__) copies of the two quad operations to shrink these kernels
// (120 -> 36 KB of SASS) — the argument traffic through local memory costs more than the cold instruction
// fetches save: 2.32 vs 2.27 ms per 2^20 MSM, 0.92 vs 0.86 ms at 2^16 (same box, alternating runs).
__global__ void __launch_bounds__(128) msm_reduce_level_quad_kernel(const uint8_t* __restrict__ Xin,
                                                                    const uint8_t* __restrict__ Yin, uint32_t n_in,
                                                                    uint32_t n_out, int W, int has_y, uint32_t in_stride,
                                                                    uint32_t out_stride, uint8_t* __restrict__ Xout,
                                                                    uint8_t* __restrict__ Yout) {
    uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 2;  // logical thread
    const uint32_t total = n_out * (uint32_t)W;
    const bool live = g < total;
    if (!live) g = total - 1;  // keep the warp converged; results are discarded
    uint32_t w = g / n_out, t = g % n_out;
    const uint8_t* xb = Xin + (size_t)w * in_stride * 128;
    const uint8_t* yb = Yin + (size_t)w * in_stride * 128;
    ge_p3 run, tot, u, ident;
    ge_p3_0(run);
    ge_p3_0(tot);
    ge_p3_0(u);
    ge_p3_0(ident);
#pragma unroll 1
    for (int i = kReduceM - 1; i >= 0; i--) {
        uint32_t j = t * kReduceM + i;
        ge_p3 x = ident, y = ident;
        if (j < n_in) {
            ge_load(x, xb + (size_t)j * 128);
            if (has_y) ge_load(y, yb + (size_t)j * 128);
        }
        ge_add_quad(run, run, x);
        if (has_y) ge_add_quad(u, u, y);
        ge_add_quad(tot, tot, run);
    }
    ge_p3 mr = run;
#pragma unroll 1
    for (int s = 1; s < kReduceM; s <<= 1) ge_dbl_quad(mr, mr);
    ge_p3 nmr;
    ge_neg(nmr, mr);
    ge_add_quad(tot, tot, u);
    ge_add_quad(tot, tot, nmr);
    if (live && (threadIdx.x & 3) == 0) {
        ge_store(Xout + ((size_t)w * out_stride + t) * 128, mr);
        ge_store(Yout + ((size_t)w * out_stride + t) * 128, tot);
    }
}
// one warp = 8 quads consumes 8*m pairs
__global__ void __launch_bounds__(128) msm_reduce_warp_quad_kernel(const uint8_t* __restrict__ Xin,
                                                                   const uint8_t* __restrict__ Yin, uint32_t n_in,
                                                                   uint32_t n_out, int nwin, int has_y, int m,
                                                                   uint32_t in_stride, uint32_t out_stride,
                                                                   uint8_t* __restrict__ Xout,
                                                                   uint8_t* __restrict__ Yout) {
    uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31, L = lane >> 2;  // logical lane 0..7
    if (gw >= n_out * (uint32_t)nwin) return;           // whole warp
    uint32_t w = gw / n_out, t = gw % n_out;
    const uint8_t* xb = Xin + (size_t)w * in_stride * 128;
    const uint8_t* yb = Yin + (size_t)w * in_stride * 128;
    ge_p3 run, tot, u, ident;
    ge_p3_0(run);
    ge_p3_0(tot);
    ge_p3_0(u);
    ge_p3_0(ident);
    uint32_t j0 = (t * 8 + (uint32_t)L) * (uint32_t)m;
#pragma unroll 1
    for (int i = m - 1; i >= 0; i--) {
        uint32_t j = j0 + i;
        ge_p3 x = ident, y = ident;
        if (j < n_in) {
            ge_load(x, xb + (size_t)j * 128);
            if (has_y) ge_load(y, yb + (size_t)j * 128);
        }
        ge_add_quad(run, run, x);
        if (has_y) ge_add_quad(u, u, y);
        if (m > 1) ge_add_quad(tot, tot, run);
    }
    if (m == 1) tot = run;
    // inclusive suffix scan of run over the 8 logical lanes (shuffle distance 4 lanes per logical lane)
    ge_p3 S = run;
#pragma unroll 1
    for (int o = 1; o < 8; o <<= 1) {
        ge_p3 other, sum;
        ge_shfl_down(other, S, 4 * o);
        ge_add_quad(sum, S, other);
        if (L + o < 8) S = sum;
    }
    ge_p3 V = S;
    if (L == 0) ge_p3_0(V);
    ge_p3 Csum;
    ge_add_quad(Csum, tot, u);
#pragma unroll 1
    for (int o = 4; o > 0; o >>= 1) {  // sums over the 8 logical lanes, result in logical lane 0
        ge_p3 other;
        ge_shfl_down(other, V, 4 * o);
        ge_add_quad(V, V, other);
        ge_shfl_down(other, Csum, 4 * o);
        ge_add_quad(Csum, Csum, other);
    }
    // logical lane 0 holds the results; the remaining operations still run on every lane (converged warp)
    for (int q = 1; q < m; q <<= 1) ge_dbl_quad(V, V);
    ge_add_quad(Csum, Csum, V);
    ge_p3 MR = S;
    for (int q = 1; q < 8 * m; q <<= 1) ge_dbl_quad(MR, MR);
    ge_p3 nMR;
    ge_neg(nMR, MR);
    ge_add_quad(Csum, Csum, nMR);
    if (lane == 0) {
        ge_store(Xout + ((size_t)w * out_stride + t) * 128, MR);
        ge_store(Yout + ((size_t)w * out_stride + t) * 128, Csum);
    }
}

// ---- 6c. low-latency reduction for the exposed tail ----------------------------------------------------
// The running-sum levels above are work-efficient but ~110 dependent point operations deep (21 in level 0,
// ~31 in each of the three warp levels), and for the LAST window group (and for small MSMs throughout) that
// depth is the exposed tail of the call.  Here the bucket index j (weight j + 1) is split as j = hi * 2^lbits + lo:
//     sum_j (j+1) X_j = 2^lbits * sum_hi hi R_hi + sum_lo lo C_lo + T,
// R_hi / C_lo the row / column sums of the (hi, lo) grid (msm_reduce2d_partial_kernel + msm_reduce2d_sums_kernel),
// T the grand total.  The two short weighted sums are taken bit by bit,
//     sum_i i Z_i = sum_k 2^k S_k,   S_k = sum of the Z_i whose index has bit k set,
// one CTA per bit (msm_reduce2d_bits_kernel, every S_k a plain tree sum again; T by one more CTA), which leaves
// ONE point Q_p per bit position p of the bucket weight: a Horner chain of c - 2 doublings and c - 1 additions
// on a quad (msm_reduce2d_horner_kernel).  About 2x the additions of the running-sum scheme, at a depth of
// ~12 + ~12 + 2c quad operations.
// These kernels are a few dozen dependent additions executed ONCE: cold instruction fetch is a large part of
// their time, so the sums share one loop with exactly one inlined copy of the (quad-cooperative) addition.
//
// Sum of the points src[(first + i * stride)], i in [0, count) with (i & mask) == mask (mask = 0 or one bit), by
// one CTA of 128 threads = 32 quads: strided quad sums, shuffle tree over the 8 quads of a warp, the 4 warp
// results through shared memory.  Result in every lane of quad 0 of warp 0.
// NWARP = 1: the same by one warp (8 quads), no shared memory; result in quad 0 of that warp.
template <int NWARP>
__device__ __forceinline__ void block_quad_point_sum(ge_p3& acc, const uint8_t* __restrict__ src, uint32_t first,
                                                     uint32_t stride, uint32_t count, uint32_t mask,
                                                     uint8_t (*sh)[128]) {
    constexpr uint32_t NQ = 8u * NWARP;  // quads that share the sum
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t quad = NWARP == 1 ? lane >> 2 : threadIdx.x >> 2;
    const uint32_t nsel = mask ? count >> 1 : count;  // selected elements, enumerated by t
    const uint32_t low = mask ? mask - 1u : 0u;
    const int nseq = (int)((nsel + NQ - 1u) / NQ);
    ge_p3_0(acc);
#pragma unroll 1
    for (int step = 0; step < nseq + (NWARP == 1 ? 3 : 5); step++) {
        ge_p3 x;
        if (step < nseq) {
            const uint32_t t = quad + NQ * (uint32_t)step;
            ge_p3_0(x);
            if (t < nsel) {
                const uint32_t i = mask ? (((t & ~low) << 1) | mask | (t & low)) : t;
                ge_load(x, src + (size_t)(first + i * stride) * 128);
            }
        } else {
            if (NWARP > 1 && step == nseq + 3) {  // the four warp sums meet in warp 0 (the other warps idle along)
                if (lane == 0) ge_store(sh[warp], acc);
                __syncthreads();
                ge_load(acc, sh[(lane >> 2) & 3]);
            }
            const int s2 = step - nseq;  // quad distances 4, 2, 1 within the warp, then 2, 1 over the warp sums
            ge_shfl_down(x, acc, s2 < 3 ? (16 >> s2) : (8 >> (s2 - 3)));
        }
        ge_add_quad(acc, acc, x);
    }
}
// Row and column sums in two steps.  (A single step — one CTA per row or column, strided quad sums and a
// shuffle tree — is issue-bound: a quad addition costs ~1000 warp instructions for 8 additions, a thread-level
// one ~2000 for 32; measured 54 us for two windows of 2^15 buckets.)
// (1) work-efficient, thread-level: every thread adds 8 consecutive buckets of a row (PR) resp. of a column (PC)
__global__ void __launch_bounds__(128) msm_reduce2d_partial_kernel(const uint8_t* __restrict__ X, uint32_t B, int lbits,
                                                                   int nwin, uint32_t out_stride,
                                                                   uint8_t* __restrict__ PR, uint8_t* __restrict__ PC) {
    const uint32_t n8 = B >> 3, Lo = 1u << lbits;
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= 2u * n8 * (uint32_t)nwin) return;
    const uint32_t w = g / (2u * n8), r = g % (2u * n8);
    const bool row = r < n8;
    const uint32_t t = row ? r : r - n8;
    // row chunk t: buckets 8t .. 8t+7;  column chunk t = hi8 * Lo + lo: buckets (8 hi8 + j) * Lo + lo
    const uint32_t first = row ? t * 8u : ((t >> lbits) * 8u << lbits) + (t & (Lo - 1u));
    const uint32_t stride = row ? 1u : Lo;
    const uint8_t* xb = X + (size_t)w * B * 128;
    ge_p3 acc, x;
    ge_load(acc, xb + (size_t)first * 128);
#pragma unroll 1
    for (uint32_t j = 1; j < 8; j++) {
        ge_load(x, xb + (size_t)(first + j * stride) * 128);
        ge_add(acc, acc, x);
    }
    ge_store((row ? PR : PC) + ((size_t)w * out_stride + t) * 128, acc);
}
// (2) one warp per (window, row or column): the Lo/8 resp. H/8 partial sums, on quads
__global__ void __launch_bounds__(128) msm_reduce2d_sums_kernel(const uint8_t* __restrict__ PR,
                                                                const uint8_t* __restrict__ PC, uint32_t B, int lbits,
                                                                int nwin, uint32_t stride, uint8_t* __restrict__ sums) {
    const uint32_t Lo = 1u << lbits, H = B >> lbits, per = H + Lo;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (gw >= per * (uint32_t)nwin) return;  // whole warp
    const uint32_t w = gw / per, r = gw % per;
    ge_p3 acc;
    const bool row = r < H;  // row r: its Lo/8 chunks are adjacent; column r - H: chunk hi8 at hi8 * Lo + lo
    const uint8_t* src = (row ? PR : PC) + (size_t)w * stride * 128;
    block_quad_point_sum<1>(acc, src, row ? r * (Lo >> 3) : r - H, row ? 1u : Lo, row ? Lo >> 3 : H >> 3, 0u, nullptr);
    if ((threadIdx.x & 31) == 0) ge_store(sums + ((size_t)w * stride + r) * 128, acc);
}
// one CTA per (window, bit of the bucket index) and one per window for the total T = sum of all rows;
// output Q[w][p], p = bit position (column bits first), Q[w][nb] = T
__global__ void __launch_bounds__(128) msm_reduce2d_bits_kernel(const uint8_t* __restrict__ sums, uint32_t in_stride,
                                                                uint32_t B, int lbits, int hbits,
                                                                uint8_t* __restrict__ Q) {
    __shared__ __align__(16) uint8_t sh[4][128];
    const int nb = lbits + hbits;
    const uint32_t w = blockIdx.x / (uint32_t)(nb + 1);
    const int role = (int)(blockIdx.x % (uint32_t)(nb + 1));
    const uint32_t Lo = 1u << lbits, H = B >> lbits;
    const uint8_t* rows = sums + (size_t)w * in_stride * 128;
    const uint8_t* cols = rows + (size_t)H * 128;
    ge_p3 acc;
    const bool col = role < lbits;
    const uint32_t mask = col ? 1u << role : (role < nb ? 1u << (role - lbits) : 0u);
    block_quad_point_sum<4>(acc, col ? cols : rows, 0u, 1u, col ? Lo : H, mask, sh);
    if (threadIdx.x == 0) ge_store(Q + ((size_t)w * 32 + role) * 128, acc);
}
// one warp per window: Horner over the bit positions, acc = 2 acc + Q_p, finally + T — a dependent chain, run in
// octet form (fe8.cuh: the warp shares every point operation; 0.48 / 0.62 us per doubling / addition against
// 0.72 / 0.99 us with one lane per product)
__global__ void __launch_bounds__(32) msm_reduce2d_horner_kernel(const uint8_t* __restrict__ Q, int nb,
                                                                 uint8_t* __restrict__ winX,
                                                                 uint8_t* __restrict__ winY) {
    const Fe8Lane L = fe8_lane();
    const uint32_t w = blockIdx.x;
    const uint8_t* q = Q + (size_t)w * 32 * 128;
    ge8 acc, x;
    ge8_load(acc, q + (size_t)(nb - 1) * 128, L);
#pragma unroll 1
    for (int pbit = nb - 2; pbit >= -1; pbit--) {
        if (pbit >= 0) ge8_dbl(acc, acc, L);
        ge8_load(x, q + (size_t)(pbit >= 0 ? pbit : nb) * 128, L);
        ge8_add(acc, acc, x, L);
    }
    ge8_store(winX + (size_t)w * 128, acc, L);
    ge8_identity(x, L);  // winY: unused by the window combine after a 2-D reduction, kept defined
    ge8_store(winY + (size_t)w * 128, x, L);
}

// ---- 7. window combine + normalise --------------------------------------------------------------
// Horner chain over the windows, top down, kept in `state` between calls so that the chain for the
// upper windows runs (on a second stream) while the lower windows are still being accumulated:
//   for w = w_hi .. w_lo:  R += X_w + Y_w;  if (w > 0) R = 2^c R
// The doublings after the last window of a call do not depend on the next group's sums.
__global__ void __launch_bounds__(32) msm_horner_kernel(const uint8_t* __restrict__ X, const uint8_t* __restrict__ Y,
                                                        int w_hi, int w_lo, int c, int first, int normalize,
                                                        int has_y, uint8_t* __restrict__ state,
                                                        uint8_t* __restrict__ result) {
    // one warp, octet form (fe8.cuh): the whole warp shares every operation of the chain
    const Fe8Lane L = fe8_lane();
    ge8 acc;
    if (first) ge8_identity(acc, L);
    else ge8_load(acc, state, L);
#pragma unroll 1
    for (int w = w_hi; w >= w_lo; w--) {
        ge8 x, y;
        ge8_load(x, X + (size_t)w * 128, L);
        if (has_y) {  // the running-sum reduction leaves two points per window, the 2-D one a single point
            ge8_load(y, Y + (size_t)w * 128, L);
            ge8_add(x, x, y, L);
        }
        ge8_add(acc, acc, x, L);
        if (w > 0) {
#pragma unroll 1
            for (int s = 0; s < c; s++) ge8_dbl(acc, acc, L);
        }
    }
    if (w_lo == 0) {
        if (normalize) ge8_store_normalized(result, acc, L);
        else ge8_store(result, acc, L);
    } else {
        ge8_store(state, acc, L);
    }
}

// ---- 8. small n: Straus, three launches, no sort ------------------------------------------------------------
// The reference's production calls into this library are MSMs of 16 / 64 points (bulletproof_range_proof.cu:724,728)
// and it ships a dedicated n <= 64 kernel for them (cuda_bulletproof_kernels.cu:119-207: one thread per pair runs a
// 256-step double-and-add, then a tree).  A Pippenger call costs ~0.5 ms whatever n is (~20 dependent launches and a
// 2^14-bucket reduction per window), so up to kStrausMax points take this path instead:
//   straus_prepare   thread per point: k = s mod p -> 64 signed 4-bit digits in [-7, 8]; multiples 1P .. 8P
//   straus_sums      CTA per window w: S_w = sum_i sign(d_iw) T_i[|d_iw|], lane-strided sums + shuffle / shared tree
//   straus_combine   one CTA: sum_w 16^w S_w as a binary tree over the windows (pairs: 4 doublings + 1 addition, then
//                    8, 16, ... 128 doublings), every step a warp in octet form (fe8.cuh): 252 doublings + 6 additions
//                    deep instead of the 252 + 63 of a plain Horner chain
// Exact for every curve point (unified additions, all 255 scalar bits), like the Pippenger path.
static constexpr int kStrausWindows = 64, kStrausTable = 8;
// measured crossover with Pippenger (ms, Straus / Pippenger c = 15): 2^10 0.30 / 0.47, 2^12 0.35 / 0.47, 2^13 0.42 / 0.45,
// 2^14 0.53 / 0.46, 2^15 0.78 / 0.49 (the window sums are n * 64 unified additions against n * 18 mixed ones)
static constexpr size_t kStrausMaxDefault = (size_t)1 << 13;
static constexpr uint32_t kStrausSlice = 256;                // points per warp of the window-sum kernel
__global__ void __launch_bounds__(128) straus_prepare_kernel(const uint8_t* __restrict__ scalars,
                                                             const uint8_t* __restrict__ points, uint32_t n,
                                                             uint8_t* __restrict__ tables, int8_t* __restrict__ digits) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    load_scalar_canon(k, scalars + (size_t)i * 32);
    uint32_t carry = 0;
#pragma unroll 1
    for (int w = 0; w < kStrausWindows; w++) {
        int d = (int)((k[w >> 3] >> ((w & 7) * 4)) & 15u) + (int)carry;
        carry = d > 8;
        if (d > 8) d -= 16;
        digits[(size_t)w * n + i] = (int8_t)d;  // k < 2^255: the top digit is at most 7 + 1, no carry out
    }
    ge_p3 P, M[kStrausTable];
    ge_load(P, points + (size_t)i * 128);
    M[0] = P;
    ge_dbl(M[1], P);
    ge_add(M[2], M[1], P);
    ge_dbl(M[3], M[1]);
    ge_add(M[4], M[3], P);
    ge_dbl(M[5], M[2]);
    ge_add(M[6], M[5], P);
    ge_dbl(M[7], M[3]);
#pragma unroll
    for (int m = 0; m < kStrausTable; m++) ge_store(tables + ((size_t)i * kStrausTable + m) * 128, M[m]);
}
// grid (64 windows, S slices of the point range), ONE WARP per CTA; slice sums land in sums[(slice * 64 + window)].
// Lane-strided additions, then a shuffle tree.  (CTAs of 256 threads over 1024-point slices spent most of their time
// in the trees — 4 additions per thread, then 5 shuffle levels and 7 serial additions by one lane.)
__global__ void __launch_bounds__(32) straus_sums_kernel(const uint8_t* __restrict__ tables,
                                                         const int8_t* __restrict__ digits, uint32_t n,
                                                         uint32_t slice_len, uint8_t* __restrict__ sums) {
    const uint32_t w = blockIdx.x, lane = threadIdx.x;
    const uint32_t lo = blockIdx.y * slice_len, hi = lo + slice_len < n ? lo + slice_len : n;
    ge_p3 acc;
    ge_p3_0(acc);
    for (uint32_t i = lo + lane; i < hi; i += 32) {
        const int d = digits[(size_t)w * n + i];
        if (d == 0) continue;
        ge_p3 t;
        ge_load(t, tables + ((size_t)i * kStrausTable + (uint32_t)((d < 0 ? -d : d) - 1)) * 128);
        if (d < 0) ge_neg(t, t);
        ge_add(acc, acc, t);
    }
    ge_warp_sum(acc);
    if (lane == 0) ge_store(sums + ((size_t)blockIdx.y * kStrausWindows + w) * 128, acc);
}
// more than one slice: one warp per window adds the slice sums up (lane = slice), result in sums[window]
__global__ void __launch_bounds__(32) straus_slices_kernel(uint8_t* sums, uint32_t nslices) {
    const uint32_t w = blockIdx.x, lane = threadIdx.x;
    ge_p3 acc;
    ge_p3_0(acc);
    for (uint32_t sl = lane; sl < nslices; sl += 32) {
        ge_p3 t;
        ge_load(t, sums + ((size_t)sl * kStrausWindows + w) * 128);
        ge_add(acc, acc, t);
    }
    ge_warp_sum(acc);
    if (lane == 0) ge_store(sums + (size_t)w * 128, acc);
}
__global__ void __launch_bounds__(1024) straus_combine_kernel(const uint8_t* __restrict__ sums, int normalize,
                                                              uint8_t* __restrict__ result) {
    __shared__ __align__(16) uint8_t sh[2][32][128];
    const Fe8Lane L = fe8_lane();
    const uint32_t warp = threadIdx.x >> 5;
    int pp = 0, shift = 4;
    for (uint32_t live = kStrausWindows / 2; live >= 1; live >>= 1, shift <<= 1, pp ^= 1) {
        if (warp < live) {  // warp-uniform
            const uint8_t* src = live == kStrausWindows / 2 ? sums : &sh[pp ^ 1][0][0];
            ge8 lo, hi;
            ge8_load(lo, src + (size_t)(2 * warp) * 128, L);
            ge8_load(hi, src + (size_t)(2 * warp + 1) * 128, L);
#pragma unroll 1
            for (int s = 0; s < shift; s++) ge8_dbl(hi, hi, L);
            ge8_add(hi, hi, lo, L);
            if (live > 1) {
                ge8_store(&sh[pp][warp][0], hi, L);
            } else if (normalize) {
                ge8_store_normalized(result, hi, L);
            } else {
                ge8_store(result, hi, L);
            }
        }
        __syncthreads();
    }
}

// ---- host side ----------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int msm_pick_window(size_t n) {
    // Measured on B200 (tools/probe_c.py, ms per MSM), with the slotted front end and the 2-D reduction:
    //          c = 12     13     14     15     16     17
    //   2^6            0.459  0.549  0.523  0.605            (c = 8: 0.469, 10: 0.476)
    //   2^10    0.600  0.516  0.634  0.529  0.612            (c = 8: 0.566)
    //   2^12    0.600  0.591  0.629  0.527  0.609
    //   2^14    0.624  0.609  0.654  0.556  0.627
    //   2^16    0.907  0.692  0.786  0.650  0.678  0.826
    //   2^18    2.558  1.640  1.301  0.923  0.957  1.063
    //   2^19           2.894  2.297  1.419  1.298  1.463
    //   2^20           5.457  3.941  2.401  2.006  2.377
    //   2^22          20.60  14.34   8.304  6.435  7.987
    //   2^24          79.4   55.5   31.4   24.39  30.23   (c = 17 was ahead, 25.8 against 25.3, before the slots)
    // Every call is bound by dependent chains (window combine, reductions, one inversion) up to ~2^17 points, and
    // their length falls with the number of windows: wide windows win long before the operation count says so.
    if (n >= ((size_t)1 << 19)) return 16;
    if (n >= ((size_t)1 << 11)) return 15;
    return 13;
}

static_assert(kMaxGroups * kSegBinsPerGroup <= kSegBins, "bin table too small");

void msm_make_plan(MsmPlan* p, size_t n, int c) {
    p->n = n;
    p->c = c > 0 ? c : msm_pick_window(n);
    p->W = (256 + p->c - 1) / p->c;
    p->B = 1u << (p->c - 1);
    p->nbuckets = (uint32_t)p->W * p->B;
    // segment length: a power of two, at least 64 and at least twice the mean run length n / B
    p->seg_shift = 6;
    while (((size_t)1 << p->seg_shift) < 2 * (n / p->B + 1) && p->seg_shift < 20) p->seg_shift++;
    // 2^15 < n <= 2^17: the accumulation launches are partial waves of latency-bound threads; segments of 16 put 4x
    // the threads on the machine (tools/probe_seg.py, c = 15, ms with 64 / 16: 2^16 0.535 / 0.508, 2^17 0.659 / 0.624,
    // 2^18 0.820 / 0.961, 2^14 0.435 / 0.441)
    if (n > ((size_t)1 << 15) && n <= ((size_t)1 << 17) && p->seg_shift == 6) p->seg_shift = 4;
    if (const int o = options().msm_seg_shift; o >= 2 && o <= 20) p->seg_shift = o;  // measurement switch
    int min_shift = p->seg_shift >= 6 ? p->seg_shift - 2 : (p->seg_shift >= 4 ? 4 : p->seg_shift);  // small window groups use shorter segments
    p->max_segs = (size_t)p->nbuckets + ((n * (size_t)p->W) >> min_shift) + 1;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    };
    p->off_table = take(n * 96);
    p->off_counts = take((size_t)p->nbuckets * 4);
    p->off_offsets = take(((size_t)p->nbuckets + 1) * 4);
    p->off_cursors = take((size_t)p->nbuckets * 4);
    p->off_tiles = take(2 * 2048 * 4);  // tile sums of the entry scan | of the segment scan
    p->off_ends = take((size_t)p->nbuckets * 4);
    p->off_segoff = take(((size_t)p->nbuckets + 1) * 4);
    p->off_desc = take(p->max_segs * 8);
    p->off_order = take(p->max_segs * 4);
    p->off_bins = take((2 * kSegBins + 2 + kMaxGroups) * 4);  // hist/cursors | binstart (+1) | heavy counters | slot overflow flag
    p->off_heavy = take((size_t)p->nbuckets * 4);
    // Fixed slots for the first digit pass (msm_digits_kernel<0>): every window but the top one, mean + 8 sigma
    // (+8) entries per bucket, a multiple of 8 so that a slot starts on a sector boundary.  From 2^19 points
    // (CBP_MSM_SLOTS=0/1 forces); measured two-pass / slotted: 2^18 1.06 / 1.11 ms, 2^19 1.43 / 1.41,
    // 2^20 2.27 / 2.20, 2^22 7.02 / 6.76.
    p->cap = 0;
    p->w_exact = 0;
    bool slots = n >= ((size_t)1 << 19);
    if (options().msm_slots >= 0) slots = options().msm_slots != 0;  // measurement / test switch
    size_t entry_words = n * (size_t)p->W;
    if (slots && p->W > 1) {
        double mean = (double)n / (double)p->B;
        size_t cap = (size_t)(mean + 8.0 * sqrt(mean) + 8.0);
        cap = (cap + 7) & ~(size_t)7;
        size_t slot_words = (size_t)(p->W - 1) * p->B * cap;
        if (slot_words + n * (size_t)p->W < ((size_t)1 << 32)) {  // entry offsets are 32-bit
            p->cap = (uint32_t)cap;
            p->w_exact = p->W - 1;
            if (slot_words + n > entry_words) entry_words = slot_words + n;
        }
    }
    p->off_entries = take(entry_words * 4 + 4);
    p->off_toprank = take(p->cap ? n * 8 : 0);
    p->off_buckets = take((size_t)p->nbuckets * 128);
    p->off_segsums = take(p->max_segs * 128);
    uint32_t n1 = (p->B + kReduceM - 1) / kReduceM;
    p->off_redX[0] = take((size_t)p->W * n1 * 128);
    p->off_redY[0] = take((size_t)p->W * n1 * 128);
    // every window owns a private slice of n1 pairs in both ping-pong buffers (group tails run concurrently)
    p->off_redX[1] = take((size_t)p->W * n1 * 128);
    p->off_redY[1] = take((size_t)p->W * n1 * 128);
    p->off_winX = take((size_t)p->W * 128);
    p->off_winY = take((size_t)p->W * 128);
    p->off_state = take(128);
    // the small-n (Straus) path: tables of 8 multiples per point, 64 digits per point, 64 window sums
    const int smax = options().msm_small_max;
    p->small = c <= 0 && n >= 1 && n <= (smax >= 0 ? (size_t)smax : kStrausMaxDefault);
    p->off_stables = take(p->small ? n * kStrausTable * 128 : 0);
    p->off_sdigits = take(p->small ? n * kStrausWindows : 0);
    p->off_ssums = take(p->small ? ((n + kStrausSlice - 1) / kStrausSlice) * kStrausWindows * 128 : 0);
    p->workspace_bytes = off;
}

#define CBP_LAUNCH_CHECK()                        \
    do {                                          \
        cudaError_t e_ = cudaGetLastError();      \
        if (e_ != cudaSuccess) return (int)e_;    \
    } while (0)

// second stream + events for the window-group pipeline (one kit per device, created on first use)
namespace {
struct StreamKit {
    cudaStream_t red[kMaxGroups] = {};  // one reduction stream per window group: group tails overlap each other
    cudaStream_t aux = nullptr;         // the Horner chain, in group order
    cudaEvent_t ev_group[kMaxGroups] = {}, ev_red[kMaxGroups] = {};
    cudaEvent_t ev_ready = nullptr, ev_done = nullptr, ev_front = nullptr;
    bool ok = false;
};
StreamKit g_kits[kMaxDevices][kMsmKits];
// caller holds the device lock
StreamKit* stream_kit(int dev, int idx) {
    if (dev < 0 || dev >= kMaxDevices || idx < 0 || idx >= kMsmKits) return nullptr;
    StreamKit& k = g_kits[dev][idx];
    if (!k.ok) {
        // Group streams carry descending priorities (group 0 = top windows first): with the accumulations of all
        // groups in flight at once (msm_run), the block scheduler drains them in pipeline order and a group's
        // reduction outranks the accumulation of the groups below it.  The Horner chain is the critical path.
        int least = 0, greatest = 0;
        if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithPriority(&k.aux, cudaStreamNonBlocking, greatest) != cudaSuccess) return nullptr;
        for (int i = 0; i < kMaxGroups; i++) {
            int prio = greatest + i < least ? greatest + i : least;
            if (cudaStreamCreateWithPriority(&k.red[i], cudaStreamNonBlocking, prio) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&k.ev_group[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
            if (cudaEventCreateWithFlags(&k.ev_red[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        }
        if (cudaEventCreateWithFlags(&k.ev_ready, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&k.ev_done, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&k.ev_front, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        k.ok = true;
    }
    return &k;
}
}  // namespace

// Measured (round 2, tools/probe_groups_phases.py, 2^20): the accumulation span plus the exposed tail is the same
// ~1.52 ms for every grouping (8,4,2,2: 1.38 + 0.14; 8,4,4: 1.25 + 0.25; 12,4: 1.21 + 0.31; 15,1: 1.23 + 0.30) and also
// with the reduction / chain streams at the highest stream priority (8,4,4: 1.34 + 0.17): the bucket sums, the
// reductions and the chain add up to a fixed amount of multiply-pipe time, the schedule only moves it around.
// Measured and rejected: sorting the lower half of the windows on a side stream while the upper half is being
// accumulated.  The exposed front end shrinks (0.50 -> 0.34 ms) but the accumulation, whose table gathers
// share the L2 with the sort's atomics, slows down by the same amount (1.48 -> 1.69 ms): 2.29 ms either way.
// windows are processed top-down in groups of halving size (.., 4, 2, 2): while the lower groups are
// still being accumulated, the upper groups are reduced and folded into the Horner chain on a second
// stream, so only the last (single-window) group's reduction latency is exposed.
static void make_groups(GroupMap* gm, int W, int c, int seg_shift, bool pipeline, int last_max,
                        const int* preset = nullptr, int npreset = 0) {
    gm->ngroups = 0;
    gm->log2B = c - 1;
    int hi = W - 1;
    if (!pipeline) {
        gm->w_hi[0] = hi;
        gm->w_lo[0] = 0;
        gm->ngroups = 1;
    } else {
        int remaining = W;
        // explicit group sizes (options().groups, top down) for measurements.  With the 2-D reduction on every
        // group (ms at 2^20 / 2^22): 8,4,2,2 1.97 / 6.58; 8,4,4 2.03 / 6.36; 10,4,2 2.01 / 6.42; 9,4,3 2.01 / 6.40;
        // 10,3,3 2.02 / 6.47; 11,3,2 2.02 / 6.44; 12,4 2.02 / 6.38; 10,6 2.02 / 6.41.
        const Options& opt = options();
        const int* glist = opt.ngroups ? opt.groups : preset;
        const int nlist = opt.ngroups ? opt.ngroups : npreset;
        int gi = 0;
        while (remaining > 0) {
            int take = remaining > 1 ? remaining / 2 : 1;
            if (gi < nlist) {
                int v = glist[gi++];
                if (v >= 1 && v <= remaining && gm->ngroups < kMaxGroups - 1) {
                    gm->w_hi[gm->ngroups] = hi;
                    gm->w_lo[gm->ngroups] = hi - v + 1;
                    gm->ngroups++;
                    hi -= v;
                    remaining -= v;
                    continue;
                }
            }
            // no one-window groups (16 windows: 8, 4, 2, 2; 18: 9, 4, 2, 3): a one-window accumulation is a single partial wave
            // (48 % of the multiply pipe in ncu against 84 % for the 8-window group); measured 2.27 vs 2.30 ms at
            // 2^20 and 1.43 vs 1.49 ms at 2^19.  Larger last groups lengthen the exposed tail (last_max: see msm_run).
            if (remaining <= last_max) take = remaining;
            else if (take < 2) take = 2;
            if (gm->ngroups == kMaxGroups - 1) take = remaining;
            gm->w_hi[gm->ngroups] = hi;
            gm->w_lo[gm->ngroups] = hi - take + 1;
            gm->ngroups++;
            hi -= take;
            remaining -= take;
        }
    }
    for (int g = 0; g < gm->ngroups; g++) {
        int nwin = gm->w_hi[g] - gm->w_lo[g] + 1;
        // a group's accumulation has ~nwin * 2^(c-1) threads; with few windows, shorter segments keep the
        // SMs full (the extra partial sums are folded by msm_heavy_small_kernel)
        int shift = seg_shift;
        if (pipeline && gm->ngroups > 1) {
            // (re-measured with the slotted front end, segment shift -1 / -2 / -3 for the 2-window groups and
            // 0 / -1 / -2 for the 4-window group: 2.19 - 2.20 ms at 2^20 for -1..-2 / 0..-1, worse beyond)
            if (nwin <= 1) shift -= 2;
            else if (nwin <= 3) shift -= 1;
        }
        if (shift < 4) shift = seg_shift < 4 ? seg_shift : 4;
        for (int w = gm->w_lo[g]; w <= gm->w_hi[g]; w++) {
            gm->group_of_window[w] = (uint8_t)g;
            gm->seg_shift_of_window[w] = (uint8_t)shift;
        }
    }
}

// Cooperative kernels whose CTAs wait for each other at grid barriers must never hold SM slots that another such grid
// is waiting for: every launch of msm_front_tail_kernel on a device is ordered behind the previous one by an event
// (the enqueue is serialised by the device lock), so at most one of them runs at a time, whatever the streams are.
static cudaEvent_t front_tail_fence(int dev) {
    static cudaEvent_t ev[kMaxDevices];
    if (dev < 0 || dev >= kMaxDevices) return nullptr;
    if (!ev[dev] && cudaEventCreateWithFlags(&ev[dev], cudaEventDisableTiming) != cudaSuccess) ev[dev] = nullptr;
    return ev[dev];
}
// CTAs of msm_front_tail_kernel that are resident at once on this device (0: no cooperative launch), cached
static int front_tail_grid(int dev, bool disable = false) {
    static int cached[kMaxDevices];
    static bool known[kMaxDevices];
    if (dev < 0 || dev >= kMaxDevices) return 0;
    if (disable) {  // a cooperative launch was refused at run time: separate kernels from now on
        cached[dev] = 0;
        known[dev] = true;
        return 0;
    }
    if (!known[dev]) {
        int coop = 0, sms = 0, per_sm = 0;
        if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) == cudaSuccess && coop &&
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, msm_front_tail_kernel, kScanThreads, 0) == cudaSuccess)
            cached[dev] = sms * (per_sm > 2 ? 2 : per_sm);  // two CTAs per SM: barriers get dearer with more
        (void)cudaGetLastError();
        known[dev] = true;
    }
    return cached[dev];
}

// set while msm_run_cached records a launch graph on this thread: the cooperative front kernel (whose per-device fence
// is an event recorded outside the capture) gives way to the separate kernels — inside a graph their launch gaps are gone anyway
static thread_local bool tl_capturing = false;

// d_scalars: n x 32 B, d_points: n x 128 B (reference AoS ge25519), d_result: 128 B
int msm_run(const MsmPlan& p, const void* d_scalars, const void* d_points, void* d_result, void* d_ws,
            int normalize, cudaStream_t st, int* launches, cudaEvent_t points_ready, int kit_index, int flags,
            void* d_front_ws) {
    uint8_t* ws = (uint8_t*)d_ws;
    uint8_t* fw = d_front_ws ? (uint8_t*)d_front_ws : ws;  // front part: everything the scalar side produces + the table
    uint8_t* table = fw + p.off_table;
    uint32_t* counts = (uint32_t*)(fw + p.off_counts);
    uint32_t* offsets = (uint32_t*)(fw + p.off_offsets);
    uint32_t* cursors = (uint32_t*)(fw + p.off_cursors);
    uint32_t* tiles = (uint32_t*)(fw + p.off_tiles);
    uint32_t* ends_buf = (uint32_t*)(fw + p.off_ends);
    uint32_t* segoff = (uint32_t*)(fw + p.off_segoff);
    uint2* desc = (uint2*)(fw + p.off_desc);
    uint32_t* order = (uint32_t*)(fw + p.off_order);
    uint32_t* bins = (uint32_t*)(fw + p.off_bins);
    uint32_t* binstart = bins + kSegBins;
    uint32_t* heavy_cnt = binstart + kSegBins + 1;
    uint32_t* overflow = heavy_cnt + kMaxGroups;
    uint32_t* heavy = (uint32_t*)(fw + p.off_heavy);
    uint32_t* entries = (uint32_t*)(fw + p.off_entries);
    uint8_t* buckets = ws + p.off_buckets;
    uint8_t* segsums = ws + p.off_segsums;
    uint8_t* winX = ws + p.off_winX;
    uint8_t* winY = ws + p.off_winY;
    uint8_t* state = ws + p.off_state;
    int nl = 0;
    size_t n = p.n;
    if (n == 0) {  // empty sum = identity (the reference would cudaMalloc(0) and copy garbage)
        static const uint64_t ident[16] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
        cudaError_t e = cudaMemcpyAsync(d_result, ident, 128, cudaMemcpyHostToDevice, st);
        if (launches) *launches = 0;
        return (int)e;
    }
    if (p.small && flags == 0) {  // (chunked and affine inputs always take the bucket path)
        cudaError_t e;
        if (points_ready && (e = cudaStreamWaitEvent(st, points_ready, 0)) != cudaSuccess) return (int)e;
        uint8_t* tables = ws + p.off_stables;
        int8_t* digits = (int8_t*)(ws + p.off_sdigits);
        uint8_t* sums = ws + p.off_ssums;
        prof_begin(BPK_PROF_MSM_TOTAL, st);
        straus_prepare_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>((const uint8_t*)d_scalars, (const uint8_t*)d_points,
                                                                          (uint32_t)n, tables, digits);
        CBP_LAUNCH_CHECK(); nl++;
        const unsigned nslices = (unsigned)((n + kStrausSlice - 1) / kStrausSlice);
        straus_sums_kernel<<<dim3(kStrausWindows, nslices), 32, 0, st>>>(tables, digits, (uint32_t)n, kStrausSlice, sums);
        CBP_LAUNCH_CHECK(); nl++;
        if (nslices > 1) {
            straus_slices_kernel<<<kStrausWindows, 32, 0, st>>>(sums, nslices);
            CBP_LAUNCH_CHECK(); nl++;
        }
        straus_combine_kernel<<<1, 1024, 0, st>>>(sums, normalize, (uint8_t*)d_result);
        CBP_LAUNCH_CHECK(); nl++;
        prof_end(BPK_PROF_MSM_TOTAL, st);
        if (launches) *launches = nl;
        return 0;
    }
    // bucket offsets, cursors and entry indices are 32-bit (index << 1 | sign; prefix sums over n * W entries)
    if (n >= ((size_t)1 << 31) || n * (size_t)p.W >= ((size_t)1 << 32)) return (int)cudaErrorInvalidValue;
    const int carry = (flags & kMsmCarryIn) ? 1 : 0;
    const bool no_tail = (flags & kMsmNoTail) != 0;
    // the side streams and events of a kit are shared by every call on this device: serialise the enqueue
    DeviceLock dlock;
    if (!dlock.ok()) return (int)cudaErrorInvalidDevice;
    const bool do_front = !(flags & kMsmBackOnly), do_back = !(flags & kMsmFrontOnly);
    // a chunk that only adds into the buckets has no tails to overlap: one group, everything on `st`
    // (a front-only call and its back-only call pass the same kMsmNoTail: the window groups shape the segments)
    // (the last chunk of a chunked input may be small, but its tail reduces the buckets of the whole input)
    const bool pipeline = (n >= (1u << 15) || carry) && !no_tail;
    StreamKit* kit = pipeline ? stream_kit(dlock.dev, kit_index) : nullptr;
    if (pipeline && !kit) return (int)cudaErrorInitializationError;
    GroupMap gm;
    // Window groups.  With every group's accumulation on its own stream (below) the grouping no longer costs
    // multiply-pipe utilisation — a small group's partial last wave is filled by the next group's blocks — so from
    // 2^19 points the windows go in pairs (the last group's reduction, the exposed tail, is then the smallest):
    // tools/probe_groups_phases.py, ms per MSM, groups in turn on one stream / all at once on their own streams:
    //   2^20  8,4,4 1.917 / 1.894   8,4,2,2 1.947 / 1.892   4,4,4,4 2.050 / 1.859   2 x 8 2.311 / 1.829
    //   2^22  8,8 6.282 / 6.219     8,4,4 6.473 / 6.200     2 x 8 - / 6.143          2^19  9,4,3 1.196   2 x 8 1.138
    // Below 2^19 points (c = 15, 18 windows) every grouping measures the same within 2 % (2^16 0.52-0.53 ms, 2^17
    // 0.63-0.64, 2^18 0.81-0.83): those calls are bound by the chain front end -> top group -> its reduction -> 270
    // doublings of the window combine, not by the accumulation.
    static const int kPairs[kMaxGroups] = {2, 2, 2, 2, 2, 2, 2, 2};
    int pairs[kMaxGroups];
    int npairs = 0;
    if (p.W >= 4 && p.W <= 2 * kMaxGroups && (n >= ((size_t)1 << 19) || carry) && options().msm_acc_streams != 0) {
        npairs = (p.W + 1) / 2;
        for (int i = 0; i < npairs; i++) pairs[i] = kPairs[i];
        if (p.W & 1) pairs[0] = 3;
    }
    make_groups(&gm, p.W, p.c, p.seg_shift, kit != nullptr, n >= ((size_t)1 << 20) ? 4 : 3, npairs ? pairs : nullptr, npairs);

    prof_begin(BPK_PROF_MSM_TOTAL, st);
    cudaError_t e = cudaSuccess;
    if (do_front) {
        if ((e = cudaMemsetAsync(counts, 0, (size_t)p.nbuckets * 4, st)) != cudaSuccess) return (int)e;
        if ((e = cudaMemsetAsync(bins, 0, (2 * kSegBins + 2 + kMaxGroups) * 4, st)) != cudaSuccess) return (int)e;
    }
    auto build_table = [&]() -> int {
        // the affine table depends only on the points: build it on a side stream while the scalar-only
        // front end (recoding, histogram, sort) runs on the main stream.  Measured at 2^20 / 2^22 with a temporary switch (ms per MSM,
        // profiles/r02_summary.md): forked before the first digit pass 1.818 / 6.123, in line before it 1.822 / 6.128 (the two kernels
        // gain nothing from running together: one is bound by load/store issue, the other streams 224 B per point),
        // forked behind the first digit pass, under the scans and the placing pass, 1.807 / 6.084 — the default.
        cudaStream_t ps = st;
        if (kit) {
            ps = kit->aux;
            if ((e = cudaEventRecord(kit->ev_done, st)) != cudaSuccess) return (int)e;  // order after earlier work on st
            if ((e = cudaStreamWaitEvent(ps, kit->ev_done, 0)) != cudaSuccess) return (int)e;
        }
        if (points_ready && (e = cudaStreamWaitEvent(ps, points_ready, 0)) != cudaSuccess) return (int)e;
        size_t threads = (n + kPreChunk - 1) / kPreChunk;
        prof_begin(BPK_PROF_MSM_PRECOMPUTE, ps);
        if (flags & kMsmAffineXY)
            msm_precompute_affine_kernel<<<(unsigned)((n + 127) / 128), 128, 0, ps>>>((const uint8_t*)d_points, n, table);
        else
            msm_precompute_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, ps>>>((const uint8_t*)d_points, n, table);
        prof_end(BPK_PROF_MSM_PRECOMPUTE, ps);
        CBP_LAUNCH_CHECK(); nl++;
        if (kit && (e = cudaEventRecord(kit->ev_ready, ps)) != cudaSuccess) return (int)e;
        return 0;
    };
    const bool table_late = do_front && do_back;
    if (do_back && !table_late)
        if (int rc = build_table()) return rc;
    // the cursors after the placing pass are the run ends — with either front end; the fused one (one cooperative launch,
    // section 4c) runs BEFORE the placing pass and keeps its own copy of the ends for its last phase
    bool fused = front_tail_grid(dlock.dev) > 0 && options().msm_fused_front != 0 && !tl_capturing;
    const uint32_t* ends = cursors;
    if (do_front) {
    prof_begin(BPK_PROF_MSM_FRONT, st);
    unsigned dgrid = (unsigned)((n + 255) / 256);
    const uint32_t slotted_ids = (uint32_t)p.w_exact * p.B;  // buckets placed by the first pass (0: none)
    uint2* toprank = p.cap ? (uint2*)(fw + p.off_toprank) : nullptr;
    msm_digits_kernel<0><<<dgrid, 256, 0, st>>>((const uint8_t*)d_scalars, n, p.c, p.W, p.B, p.cap, p.w_exact, counts,
                                                entries, overflow, offsets, toprank);
    CBP_LAUNCH_CHECK(); nl++;
    if (table_late)
        if (int rc = build_table()) return rc;
    uint32_t ntiles = (p.nbuckets + 1 + kScanTile - 1) / kScanTile;  // +1: the sentinel slot
    if (fused) {
        const uint32_t ftiles = (p.nbuckets + 1 + kFtTile - 1) / kFtTile;
        FrontTail ft;
        ft.counts = counts; ft.overflow = overflow; ft.tilesA = tiles; ft.tilesS = tiles + 2048; ft.offsets = offsets;
        ft.cursors = cursors; ft.ends = ends_buf; ft.segoff = segoff; ft.desc = desc; ft.heavy = heavy; ft.heavy_cnt = heavy_cnt;
        ft.hist = bins; ft.binstart = binstart; ft.order = order;
        ft.total = p.nbuckets; ft.B = p.B; ft.slotted_ids = slotted_ids; ft.cap = p.cap;
        ft.ranked_from = p.cap ? slotted_ids : 0xffffffffu;
        // enough CTAs for the tiles and for ~4 segments per thread in the last phase, all of them resident
        unsigned want = ftiles > (unsigned)((p.max_segs + 1023) / 1024) ? ftiles : (unsigned)((p.max_segs + 1023) / 1024);
        unsigned grid = (unsigned)front_tail_grid(dlock.dev);
        if (want < grid) grid = want ? want : 1;
        void* args[] = {&ft, &gm};
        cudaEvent_t fence = front_tail_fence(dlock.dev);
        if (!fence) return (int)cudaErrorInitializationError;
        if ((e = cudaStreamWaitEvent(st, fence, 0)) != cudaSuccess) return (int)e;  // no-op before the first record
        if ((e = cudaLaunchCooperativeKernel((const void*)msm_front_tail_kernel, dim3(grid), dim3(kScanThreads), args, 0, st)) !=
            cudaSuccess) {
            // refused (a partitioned or shared device may not grant the residency the occupancy query promised): nothing has
            // been enqueued, so the separate kernels take over — for this call and every later one on this device
            (void)cudaGetLastError();
            front_tail_grid(dlock.dev, true);
            fused = false;
        } else {
            if ((e = cudaEventRecord(fence, st)) != cudaSuccess) return (int)e;
            nl++;
            msm_digits_kernel<1><<<dgrid, 256, 0, st>>>((const uint8_t*)d_scalars, n, p.c, p.W, p.B, p.cap, p.w_exact, cursors,
                                                        entries, overflow, offsets, toprank);
            CBP_LAUNCH_CHECK(); nl++;
        }
    }
    if (!fused) {
    scan_tile_sums_kernel<false><<<ntiles, kScanThreads, 0, st>>>(counts, p.nbuckets, gm, slotted_ids, overflow, tiles);
    CBP_LAUNCH_CHECK(); nl++;
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
    CBP_LAUNCH_CHECK(); nl++;
    scan_apply_kernel<false><<<ntiles, kScanThreads, 0, st>>>(counts, p.nbuckets, gm, slotted_ids, p.cap,
                                                              p.cap ? slotted_ids : 0xffffffffu, overflow, tiles, offsets,
                                                              cursors);
    CBP_LAUNCH_CHECK(); nl++;
    msm_digits_kernel<1><<<dgrid, 256, 0, st>>>((const uint8_t*)d_scalars, n, p.c, p.W, p.B, p.cap, p.w_exact, cursors,
                                                entries, overflow, offsets, toprank);
    CBP_LAUNCH_CHECK(); nl++;
    // segments
    scan_tile_sums_kernel<true><<<ntiles, kScanThreads, 0, st>>>(counts, p.nbuckets, gm, 0u, overflow, tiles);
    CBP_LAUNCH_CHECK(); nl++;
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
    CBP_LAUNCH_CHECK(); nl++;
    scan_apply_kernel<true><<<ntiles, kScanThreads, 0, st>>>(counts, p.nbuckets, gm, 0u, 0u, 0xffffffffu, overflow, tiles,
                                                             segoff, nullptr);
    CBP_LAUNCH_CHECK(); nl++;
    const uint32_t* nsegs_p = segoff + p.nbuckets;
    unsigned bgrid = (p.nbuckets + 255) / 256, sgrid = (unsigned)((p.max_segs + 255) / 256);
    seg_build_kernel<<<bgrid, 256, 0, st>>>(counts, offsets, segoff, p.nbuckets, p.B, gm, desc, heavy, heavy_cnt);
    CBP_LAUNCH_CHECK(); nl++;
    seg_hist_kernel<<<sgrid < 1184 ? sgrid : 1184, 256, 0, st>>>(desc, ends, nsegs_p, gm, bins);
    CBP_LAUNCH_CHECK(); nl++;
    seg_bin_scan_kernel<<<1, kSegBins, 0, st>>>(bins, binstart);
    CBP_LAUNCH_CHECK(); nl++;
    seg_scatter_kernel<<<sgrid, 256, 0, st>>>(desc, ends, nsegs_p, gm, bins, order);
    CBP_LAUNCH_CHECK(); nl++;
    }

    prof_end(BPK_PROF_MSM_FRONT, st);
    }  // do_front
    if (!do_back) {
        prof_end(BPK_PROF_MSM_TOTAL, st);
        if (launches) *launches = nl;
        return 0;
    }
    if (kit) {  // join the table build (side stream) before the first accumulation
        if ((e = cudaStreamWaitEvent(st, kit->ev_ready, 0)) != cudaSuccess) return (int)e;
    }
    // Accumulation of the groups: in turn on `st` (each group's tail forks off behind its accumulation), or — the
    // groups touch disjoint buckets — all at once, every group on its own stream ahead of its own tail: the next
    // group's thread blocks then fill the SMs while the previous launch drains (no partial last wave per group),
    // and the stream priorities keep the pipeline order.
    const int acc_opt = options().msm_acc_streams;
    const bool acc_streams = kit != nullptr && gm.ngroups > 1 && (acc_opt < 0 ? true : acc_opt != 0);
    if (acc_streams && (e = cudaEventRecord(kit->ev_front, st)) != cudaSuccess) return (int)e;
    for (int g = 0; g < gm.ngroups; g++) {
        cudaStream_t tail = kit ? kit->red[g] : st;
        cudaStream_t as = acc_streams ? tail : st;
        int nwin = gm.w_hi[g] - gm.w_lo[g] + 1, w_lo = gm.w_lo[g];
        // upper bound on this group's segments: its buckets + its share of the entries
        const int gshift = gm.seg_shift_of_window[w_lo];
        size_t seg_bound = (size_t)nwin * p.B + ((n * (size_t)nwin) >> gshift) + 1;
        if (g == 0) prof_begin(BPK_PROF_MSM_ACCUMULATE, st);
        if (acc_streams && (e = cudaStreamWaitEvent(as, kit->ev_front, 0)) != cudaSuccess) return (int)e;
        msm_accumulate_kernel<<<(unsigned)((seg_bound + 127) / 128), 128, 0, as>>>(table, entries, desc, ends, segoff,
                                                                                  gm, order, binstart, g, carry,
                                                                                  buckets, segsums);
        CBP_LAUNCH_CHECK(); nl++;
        if (kit) {
            if ((e = cudaEventRecord(kit->ev_group[g], as)) != cudaSuccess) return (int)e;
            if (acc_streams) {  // `st` follows every accumulation: the phase timers below, and nothing else, run on it
                if ((e = cudaStreamWaitEvent(st, kit->ev_group[g], 0)) != cudaSuccess) return (int)e;
            } else if ((e = cudaStreamWaitEvent(tail, kit->ev_group[g], 0)) != cudaSuccess) return (int)e;
        }
        if (g == gm.ngroups - 1) {
            prof_end(BPK_PROF_MSM_ACCUMULATE, st);
            prof_begin(BPK_PROF_MSM_TAIL, st);
        }
        size_t heavy_bound = ((n * (size_t)nwin) >> gshift) + 1;
        if (heavy_bound > (size_t)nwin * p.B) heavy_bound = (size_t)nwin * p.B;
        unsigned sgrid_h = (unsigned)((heavy_bound + 127) / 128);
        msm_heavy_small_kernel<<<sgrid_h < 4736 ? sgrid_h : 4736, 128, 0, tail>>>(heavy, heavy_cnt, g, (uint32_t)w_lo * p.B,
                                                                                  segoff, segsums, carry, buckets);
        CBP_LAUNCH_CHECK(); nl++;
        unsigned hgrid = (unsigned)((heavy_bound * 32 + 127) / 128);
        msm_heavy_fix_kernel<<<hgrid < 2368 ? hgrid : 2368, 128, 0, tail>>>(heavy, heavy_cnt, g, (uint32_t)w_lo * p.B, segoff,
                                                                            segsums, carry, buckets);
        CBP_LAUNCH_CHECK(); nl++;
        if (no_tail) continue;  // the buckets now hold this chunk too; reduction happens after the last chunk
        // reduction levels for this group's windows
        const uint32_t n1 = (p.B + kReduceM - 1) / kReduceM;  // per-window slice of the ping-pong buffers
        const uint8_t* X = buckets + (size_t)w_lo * p.B * 128;
        const uint8_t* Y = X;
        uint32_t n_in = p.B, in_stride = p.B;
        int has_y = 0, pp = 0, level = 0;
        bool two_d = false;
        // The shallow 2-D reduction (6c), for every group: the reductions of the overlapped groups are not
        // exposed themselves, but the Horner chain runs through them in order, and a group's ~110-operation
        // running-sum levels, starved of SM slots by the next accumulation, delayed the chain to the end
        // (2-D on the last group only / on all groups: 2^18 1.06 / 0.94 ms, 2^20 2.16 / 2.03, 2^22 6.56 / 6.44).
        // The running-sum kernels remain for narrow windows (c < 9) and as a cross-check (CBP_MSM_NO2D).
        if (p.c >= 9 && !options().msm_no2d) {
            const int lbits = (p.c - 1) / 2, hbits = p.c - 1 - lbits;
            const uint32_t per = (p.B >> lbits) + (1u << lbits);  // <= n1 for c >= 9
            uint8_t* sums = ws + p.off_redX[0] + (size_t)w_lo * n1 * 128;
            uint8_t* Q = ws + p.off_redY[0] + (size_t)w_lo * n1 * 128;  // 32 slots per window (n1 >= 32)
            uint8_t* PR = ws + p.off_redX[1] + (size_t)w_lo * n1 * 128;  // B/8 = n1 chunk sums per window each
            uint8_t* PC = ws + p.off_redY[1] + (size_t)w_lo * n1 * 128;
            const uint32_t pthreads = 2u * (p.B >> 3) * (uint32_t)nwin;
            msm_reduce2d_partial_kernel<<<(pthreads + 127) / 128, 128, 0, tail>>>(X, p.B, lbits, nwin, n1, PR, PC);
            CBP_LAUNCH_CHECK(); nl++;
            msm_reduce2d_sums_kernel<<<(per * (uint32_t)nwin * 32u + 127) / 128, 128, 0, tail>>>(PR, PC, p.B, lbits, nwin,
                                                                                                 n1, sums);
            CBP_LAUNCH_CHECK(); nl++;
            msm_reduce2d_bits_kernel<<<(unsigned)(nwin * p.c), 128, 0, tail>>>(sums, n1, p.B, lbits, hbits, Q);
            CBP_LAUNCH_CHECK(); nl++;
            msm_reduce2d_horner_kernel<<<nwin, 32, 0, tail>>>(Q, p.c - 1, winX + (size_t)w_lo * 128,
                                                              winY + (size_t)w_lo * 128);
            CBP_LAUNCH_CHECK(); nl++;
            n_in = 1;
            two_d = true;
        }
        while (n_in > 1) {
            bool seq = level == 0;  // level 0 is work-efficient (thread-sequential), upper levels warp-cooperative
            // small groups are latency-exposed: quad-cooperative kernels; large groups overlap with later
            // accumulations and use the work-efficient one-lane-per-chunk kernels
            const bool quad = kit == nullptr || nwin <= 2;  // small MSMs (single group) are latency-bound throughout
            int m = 1;
            uint32_t n_out;
            if (seq) {
                n_out = (n_in + kReduceM - 1) / kReduceM;
            } else {
                const uint32_t lanes = quad ? 8u : 32u;  // logical lanes per warp
                while (m < 4 && lanes * m < n_in) m <<= 1;
                n_out = (n_in + lanes * m - 1) / (lanes * m);
            }
            const bool last = n_out == 1;
            const uint32_t out_stride = last ? 1u : n1;
            uint8_t* Xo = last ? winX + (size_t)w_lo * 128 : ws + p.off_redX[pp] + (size_t)w_lo * n1 * 128;
            uint8_t* Yo = last ? winY + (size_t)w_lo * 128 : ws + p.off_redY[pp] + (size_t)w_lo * n1 * 128;
            if (seq) {
                uint32_t threads = n_out * (uint32_t)nwin * (quad ? 4u : 1u);
                if (quad)
                    msm_reduce_level_quad_kernel<<<(threads + 127) / 128, 128, 0, tail>>>(X, Y, n_in, n_out, nwin, has_y,
                                                                                         in_stride, out_stride, Xo, Yo);
                else
                    msm_reduce_level_kernel<<<(threads + 127) / 128, 128, 0, tail>>>(X, Y, n_in, n_out, nwin, has_y,
                                                                                    in_stride, out_stride, Xo, Yo);
            } else {
                uint32_t threads = n_out * (uint32_t)nwin * 32;
                if (quad)
                    msm_reduce_warp_quad_kernel<<<(threads + 127) / 128, 128, 0, tail>>>(X, Y, n_in, n_out, nwin, has_y, m,
                                                                                        in_stride, out_stride, Xo, Yo);
                else
                    msm_reduce_warp_kernel<<<(threads + 127) / 128, 128, 0, tail>>>(X, Y, n_in, n_out, nwin, has_y, m,
                                                                                   in_stride, out_stride, Xo, Yo);
            }
            CBP_LAUNCH_CHECK(); nl++;
            X = Xo;
            Y = Yo;
            n_in = n_out;
            in_stride = out_stride;
            has_y = 1;
            pp ^= 1;
            level++;
        }
        cudaStream_t hs = tail;
        if (kit) {
            hs = kit->aux;
            if ((e = cudaEventRecord(kit->ev_red[g], tail)) != cudaSuccess) return (int)e;
            if ((e = cudaStreamWaitEvent(hs, kit->ev_red[g], 0)) != cudaSuccess) return (int)e;
        }
        msm_horner_kernel<<<1, 32, 0, hs>>>(winX, winY, gm.w_hi[g], gm.w_lo[g], p.c, g == 0, normalize, two_d ? 0 : 1,
                                            state, (uint8_t*)d_result);
        CBP_LAUNCH_CHECK(); nl++;
    }
    if (kit) {
        if ((e = cudaEventRecord(kit->ev_done, kit->aux)) != cudaSuccess) return (int)e;
        if ((e = cudaStreamWaitEvent(st, kit->ev_done, 0)) != cudaSuccess) return (int)e;
    }
    prof_end(BPK_PROF_MSM_TAIL, st);
    prof_end(BPK_PROF_MSM_TOTAL, st);
    if (launches) *launches = nl;
    return 0;
}

// ---- cached launch graphs for the latency-bound sizes ----------------------------------------------------------------
namespace {
struct GraphEntry {
    const void *s, *p, *r, *ws;
    size_t n;
    int c, normalize, launches;
    uint64_t epoch, last_use;
    cudaGraphExec_t exec;
};
struct GraphCache {
    std::vector<GraphEntry> entries;
    cudaStream_t cap = nullptr;
    uint64_t tick = 0;
    bool broken = false;  // a capture failed on this device: plain launches from then on
};
GraphCache g_graph_cache[kMaxDevices];
constexpr size_t kGraphCacheMax = 8;
}  // namespace

int msm_run_cached(const MsmPlan& p, const void* d_scalars, const void* d_points, void* d_result, void* d_ws, int normalize,
                   cudaStream_t st, int* launches) {
    const bool eligible = options().msm_graph != 0 && !prof_enabled() && !p.small && p.n > ((size_t)1 << 13) &&
                          p.n < ((size_t)1 << 19);
    if (!eligible) return msm_run(p, d_scalars, d_points, d_result, d_ws, normalize, st, launches, nullptr);
    DeviceLock dlock;  // the cache, and (below) the side streams a capture records on, belong to the device
    if (!dlock.ok()) return (int)cudaErrorInvalidDevice;
    GraphCache& gc = g_graph_cache[dlock.dev];
    if (gc.broken) return msm_run(p, d_scalars, d_points, d_result, d_ws, normalize, st, launches, nullptr);
    const uint64_t epoch = g_options_epoch.load();
    GraphEntry* hit = nullptr;
    for (GraphEntry& e : gc.entries)
        if (e.s == d_scalars && e.p == d_points && e.r == d_result && e.ws == d_ws && e.n == p.n && e.c == p.c &&
            e.normalize == normalize && e.epoch == epoch)
            hit = &e;
    if (!hit) {
        // everything a capture must not do itself: streams, events, occupancy queries
        if (!gc.cap && cudaStreamCreateWithFlags(&gc.cap, cudaStreamNonBlocking) != cudaSuccess) gc.broken = true;
        if (!stream_kit(dlock.dev, 0)) gc.broken = true;
        (void)front_tail_grid(dlock.dev);
        cudaGraph_t graph = nullptr;
        cudaGraphExec_t exec = nullptr;
        int nl = 0, rc = 0;
        if (!gc.broken && cudaStreamBeginCapture(gc.cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            tl_capturing = true;
            rc = msm_run(p, d_scalars, d_points, d_result, d_ws, normalize, gc.cap, &nl, nullptr);
            tl_capturing = false;
            cudaError_t ce = cudaStreamEndCapture(gc.cap, &graph);
            if (rc == 0 && ce == cudaSuccess && graph) ce = cudaGraphInstantiate(&exec, graph, 0);
            if (rc != 0 || ce != cudaSuccess || !exec) gc.broken = true;
            if (graph) cudaGraphDestroy(graph);
        } else {
            gc.broken = true;
        }
        if (gc.broken) {
            (void)cudaGetLastError();
            if (exec) cudaGraphExecDestroy(exec);
            return msm_run(p, d_scalars, d_points, d_result, d_ws, normalize, st, launches, nullptr);
        }
        if (gc.entries.size() >= kGraphCacheMax) {  // evict the least recently used (and any entry of an older options epoch)
            size_t victim = 0;
            for (size_t i = 1; i < gc.entries.size(); i++)
                if (gc.entries[i].epoch != epoch || gc.entries[i].last_use < gc.entries[victim].last_use) victim = i;
            cudaGraphExecDestroy(gc.entries[victim].exec);
            gc.entries.erase(gc.entries.begin() + (long)victim);
        }
        gc.entries.push_back({d_scalars, d_points, d_result, d_ws, p.n, p.c, normalize, nl, epoch, 0, exec});
        hit = &gc.entries.back();
    }
    hit->last_use = ++gc.tick;
    cudaError_t e = cudaGraphLaunch(hit->exec, st);
    if (launches) *launches = hit->launches;
    return (int)e;
}

}  // namespace cbp
