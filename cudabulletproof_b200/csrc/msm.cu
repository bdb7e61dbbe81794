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
// Pipeline (all on one stream, no host synchronisation):
//   1. msm_precompute   AoS ge25519 -> 96-byte affine (y+x, y-x, 2dxy) table   [HBM streaming]
//   2. msm_count        signed c-bit digit recoding + per-(window,bucket) histogram
//   3. scan             exclusive prefix sum over W*B counters
//   4. msm_scatter      second recoding pass, writes (point index, sign) into its bucket run
//   5. msm_accumulate   one thread per bucket: gathers its run, 7M mixed additions  [IMAD-bound, ~90% of time]
//   6. msm_reduce_level running-sum reduction  sum_b b*B_b  per window, m buckets per thread, repeated
//   7. msm_finish       Horner over windows, normalise to Z = 1
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "ge25519.cuh"
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

// ---- 2./4. digit recoding: histogram and scatter ------------------------------------------------
// signed digits d_w in [-(2^(c-1)-1), 2^(c-1)], sum d_w 2^(cw) = k; bucket index |d|-1.
template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_digits_kernel(const uint8_t* __restrict__ scalars, size_t n, int c, int W,
                                                         uint32_t B, uint32_t* __restrict__ counters,
                                                         uint32_t* __restrict__ entries) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    load_scalar_canon(k, scalars + i * 32);
    uint32_t carry = 0;
    const uint32_t half = 1u << (c - 1);
    for (int w = 0; w < W; w++) {
        uint32_t d = raw_digit(k, w, c) + carry;
        uint32_t neg = 0;
        carry = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = 1;
            carry = 1;
        }
        if (d != 0) {
            uint32_t id = (uint32_t)w * B + (d - 1);
            if (SCATTER) {
                uint32_t pos = atomicAdd(&counters[id], 1u);
                entries[pos] = ((uint32_t)i << 1) | neg;
            } else {
                atomicAdd(&counters[id], 1u);
            }
        }
    }
}

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
__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t* __restrict__ in, uint32_t total,
                                                                      uint32_t* __restrict__ tile_sums) {
    __shared__ uint32_t smem[64];
    uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer, s = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++)
        if (base + j < total) s += in[base + j];
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
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t* __restrict__ in, uint32_t total,
                                                                  const uint32_t* __restrict__ tile_sums,
                                                                  uint32_t* __restrict__ offsets,
                                                                  uint32_t* __restrict__ cursors) {
    __shared__ uint32_t smem[64];
    uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanPer;
    uint32_t v[kScanPer], s = 0;
#pragma unroll
    for (int j = 0; j < kScanPer; j++) {
        v[j] = base + j < total ? in[base + j] : 0;
        s += v[j];
    }
    uint32_t bt;
    uint32_t ex = block_exclusive_scan(s, smem, bt) + tile_sums[blockIdx.x];
#pragma unroll
    for (int j = 0; j < kScanPer; j++) {
        if (base + j <= total) offsets[base + j] = ex;  // offsets[total] = grand total (sentinel)
        if (base + j < total) cursors[base + j] = ex;
        ex += v[j];
    }
}

// ---- 5. bucket accumulation ---------------------------------------------------------------------
// One thread per (window, bucket).  order[] (optional) lists bucket ids longest-run first so that the
// 32 buckets of a warp have near-equal run lengths.
__global__ void __launch_bounds__(128, 4)
    msm_accumulate_kernel(const uint8_t* __restrict__ table, const uint32_t* __restrict__ entries,
                          const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ order, uint32_t nbuckets,
                          uint8_t* __restrict__ bucket_sums) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nbuckets) return;
    uint32_t id = order ? order[t] : t;
    uint32_t e = offsets[id], end = offsets[id + 1];
    ge_p3 acc;
    ge_p3_0(acc);
    if (e < end) {
        uint32_t ent = __ldg(entries + e);
        ge_niels q;
        ge_niels_load(q, table + (size_t)(ent >> 1) * 96);
        for (;;) {
            uint32_t neg = ent & 1u;
            ge_niels cur = q;
            ++e;
            if (e < end) {  // prefetch the next operand while this addition runs
                ent = __ldg(entries + e);
                ge_niels_load(q, table + (size_t)(ent >> 1) * 96);
            }
            ge_madd(acc, acc, cur, neg != 0);
            if (e >= end) break;
        }
    }
    ge_store(bucket_sums + (size_t)id * 128, acc);
}

// ---- 6. running-sum reduction -------------------------------------------------------------------
// One level maps n pairs (X_j, Y_j) per window to ceil(n/m) pairs, preserving
//     sum_j (j+1) X_j + sum_j Y_j :
//   R = sum X_j, T = sum_{i=1..m} i X_{tm+i-1}, U = sum Y_j  over the chunk,
//   X'_t = m R,  Y'_t = T + U - m R.      (m is a power of two: m R by doublings)
// Level 0 has no Y (has_y = 0).  Chunks past the end read as identity.
__global__ void __launch_bounds__(128) msm_reduce_level_kernel(const uint8_t* __restrict__ Xin,
                                                               const uint8_t* __restrict__ Yin, uint32_t n_in,
                                                               uint32_t n_out, int W, int has_y,
                                                               uint8_t* __restrict__ Xout, uint8_t* __restrict__ Yout) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_out * (uint32_t)W) return;
    uint32_t w = g / n_out, t = g % n_out;
    const uint8_t* xb = Xin + (size_t)w * n_in * 128;
    const uint8_t* yb = Yin + (size_t)w * n_in * 128;
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
    ge_store(Xout + ((size_t)w * n_out + t) * 128, mr);
    ge_store(Yout + ((size_t)w * n_out + t) * 128, tot);
}

// ---- 7. window combine + normalise --------------------------------------------------------------
// window sum S_w = X_w + Y_w (the single remaining pair, weight 1); result = sum_w 2^(cw) S_w.
__global__ void msm_finish_kernel(const uint8_t* __restrict__ X, const uint8_t* __restrict__ Y, int W, int c,
                                  int normalize, uint8_t* __restrict__ result) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    ge_p3 acc;
    ge_p3_0(acc);
    for (int w = W - 1; w >= 0; w--) {
        for (int s = 0; s < c; s++) ge_dbl(acc, acc);
        ge_p3 x, y;
        ge_load(x, X + (size_t)w * 128);
        ge_load(y, Y + (size_t)w * 128);
        ge_add(acc, acc, x);
        ge_add(acc, acc, y);
    }
    if (normalize) ge_normalize(acc);
    ge_store(result, acc);
}

// ---- host side ----------------------------------------------------------------------------------
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int msm_pick_window(size_t n) {
    // minimise W*(n + 2*2^(c-1)*9/7) over c with W = ceil(256/c); c <= 16 keeps entries in 32 bits
    double best = 1e300;
    int best_c = 4;
    for (int c = 4; c <= 16; c++) {
        int W = (256 + c - 1) / c;
        double cost = (double)W * ((double)n + 2.0 * (double)(1u << (c - 1)) * 9.0 / 7.0) + 1.5 * c * W;
        if (cost < best) {
            best = cost;
            best_c = c;
        }
    }
    return best_c;
}

void msm_make_plan(MsmPlan* p, size_t n, int c) {
    p->n = n;
    p->c = c > 0 ? c : msm_pick_window(n);
    p->W = (256 + p->c - 1) / p->c;
    p->B = 1u << (p->c - 1);
    p->nbuckets = (uint32_t)p->W * p->B;
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
    p->off_tiles = take(1024 * 4);
    p->off_entries = take(n * (size_t)p->W * 4 + 4);
    p->off_buckets = take((size_t)p->nbuckets * 128);
    // reduction ping-pong buffers: level 1 output has ceil(B/m) pairs per window
    uint32_t n1 = (p->B + kReduceM - 1) / kReduceM;
    p->off_redX[0] = take((size_t)p->W * n1 * 128);
    p->off_redY[0] = take((size_t)p->W * n1 * 128);
    uint32_t n2 = (n1 + kReduceM - 1) / kReduceM;
    p->off_redX[1] = take((size_t)p->W * n2 * 128);
    p->off_redY[1] = take((size_t)p->W * n2 * 128);
    p->workspace_bytes = off;
}

#define CBP_LAUNCH_CHECK()                        \
    do {                                          \
        cudaError_t e_ = cudaGetLastError();      \
        if (e_ != cudaSuccess) return (int)e_;    \
    } while (0)

// d_scalars: n x 32 B, d_points: n x 128 B (reference AoS ge25519), d_result: 128 B
int msm_run(const MsmPlan& p, const void* d_scalars, const void* d_points, void* d_result, void* d_ws,
            int normalize, cudaStream_t st, int* launches) {
    uint8_t* ws = (uint8_t*)d_ws;
    uint8_t* table = ws + p.off_table;
    uint32_t* counts = (uint32_t*)(ws + p.off_counts);
    uint32_t* offsets = (uint32_t*)(ws + p.off_offsets);
    uint32_t* cursors = (uint32_t*)(ws + p.off_cursors);
    uint32_t* tiles = (uint32_t*)(ws + p.off_tiles);
    uint32_t* entries = (uint32_t*)(ws + p.off_entries);
    uint8_t* buckets = ws + p.off_buckets;
    int nl = 0;
    size_t n = p.n;
    if (n == 0) {  // empty sum = identity (the reference would cudaMalloc(0) and copy garbage)
        static const uint64_t ident[16] = {0, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0};
        cudaError_t e = cudaMemcpyAsync(d_result, ident, 128, cudaMemcpyHostToDevice, st);
        if (launches) *launches = 0;
        return (int)e;
    }
    prof_begin(BPK_PROF_MSM_TOTAL, st);
    cudaError_t e = cudaMemsetAsync(counts, 0, (size_t)p.nbuckets * 4, st);
    if (e != cudaSuccess) return (int)e;
    {
        size_t threads = (n + kPreChunk - 1) / kPreChunk;
        prof_begin(BPK_PROF_MSM_PRECOMPUTE, st);
        msm_precompute_kernel<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>((const uint8_t*)d_points, n, table);
        prof_end(BPK_PROF_MSM_PRECOMPUTE, st);
        CBP_LAUNCH_CHECK(); nl++;
    }
    unsigned dgrid = (unsigned)((n + 255) / 256);
    msm_digits_kernel<false><<<dgrid, 256, 0, st>>>((const uint8_t*)d_scalars, n, p.c, p.W, p.B, counts, nullptr);
    CBP_LAUNCH_CHECK(); nl++;
    uint32_t ntiles = (p.nbuckets + 1 + kScanTile - 1) / kScanTile;  // +1: the sentinel slot
    scan_tile_sums_kernel<<<ntiles, kScanThreads, 0, st>>>(counts, p.nbuckets, tiles);
    CBP_LAUNCH_CHECK(); nl++;
    scan_tiles_kernel<<<1, 1024, 0, st>>>(tiles, ntiles);
    CBP_LAUNCH_CHECK(); nl++;
    scan_apply_kernel<<<ntiles, kScanThreads, 0, st>>>(counts, p.nbuckets, tiles, offsets, cursors);
    CBP_LAUNCH_CHECK(); nl++;
    msm_digits_kernel<true><<<dgrid, 256, 0, st>>>((const uint8_t*)d_scalars, n, p.c, p.W, p.B, cursors, entries);
    CBP_LAUNCH_CHECK(); nl++;
    prof_begin(BPK_PROF_MSM_ACCUMULATE, st);
    msm_accumulate_kernel<<<(p.nbuckets + 127) / 128, 128, 0, st>>>(table, entries, offsets, nullptr, p.nbuckets, buckets);
    prof_end(BPK_PROF_MSM_ACCUMULATE, st);
    CBP_LAUNCH_CHECK(); nl++;
    // reduction levels
    const uint8_t* X = buckets;
    const uint8_t* Y = buckets;
    uint32_t n_in = p.B;
    int has_y = 0, pp = 0;
    do {
        uint32_t n_out = (n_in + kReduceM - 1) / kReduceM;
        uint8_t* Xo = ws + p.off_redX[pp];
        uint8_t* Yo = ws + p.off_redY[pp];
        uint32_t threads = n_out * (uint32_t)p.W;
        msm_reduce_level_kernel<<<(threads + 127) / 128, 128, 0, st>>>(X, Y, n_in, n_out, p.W, has_y, Xo, Yo);
        CBP_LAUNCH_CHECK(); nl++;
        X = Xo;
        Y = Yo;
        n_in = n_out;
        has_y = 1;
        pp ^= 1;
    } while (n_in > 1);
    msm_finish_kernel<<<1, 32, 0, st>>>(X, Y, p.W, p.c, normalize, (uint8_t*)d_result);
    CBP_LAUNCH_CHECK(); nl++;
    prof_end(BPK_PROF_MSM_TOTAL, st);
    if (launches) *launches = nl;
    return 0;
}

}  // namespace cbp
