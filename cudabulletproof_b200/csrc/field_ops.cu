// field_ops.cu — batched fe25519 arithmetic and mod-l inner products for sm_100a, with their C ABI.
//
// Replaces cuda_field_ops.cu (kernels K5-K13: add/sub/mul/square, the racy "Montgomery" inversion
// K10-K11, the carry-less "SoA" add K12-K13) and cuda_inner_product.cu (K14-K18) of the reference.
// Elementwise ops are HBM-bound (96 B per element moved for 72 IMAD): one element per thread,
// 128-bit loads/stores, grid-stride over a grid sized in multiples of the SM count.
#include <stdio.h>
#include "../../include/cuda_bulletproof.h"
#include "common.h"
#include "fe25519.cuh"
#include "sc25519.cuh"

namespace cbp {

static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

// ---- elementwise --------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256) fe_batch_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ a,
                                                       const uint8_t* __restrict__ b, size_t count) {
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        fe x, y, r;
        fe_load_nc(x, a + i * 32);
        if (OP != BPK_FE_SQR) fe_load_nc(y, b + i * 32);
        if (OP == BPK_FE_ADD) fe_add(r, x, y);
        if (OP == BPK_FE_SUB) fe_sub(r, x, y);
        if (OP == BPK_FE_MUL) fe_mul(r, x, y);
        if (OP == BPK_FE_SQR) fe_sq(r, x);
        fe_canon(r);
        fe_store(out + i * 32, r);
    }
}

// ---- batch inversion: Montgomery's trick, one inversion per CTA tile -------------------------------
// Thread t of a tile owns elements t, t+T, t+2T, ... (coalesced); products are commutative so any
// partition works.  Forward pass stores exclusive prefix products in `out`, the CTA combines the
// per-thread totals with prefix/suffix scans in shared memory, one warp inverts the tile total,
// the backward pass turns prefixes into inverses.  Zero inputs are skipped and map to zero.
static constexpr int kInvThreads = 256, kInvPer = 16, kInvTile = kInvThreads * kInvPer;
__global__ void __launch_bounds__(kInvThreads) fe_batch_invert_kernel(uint8_t* __restrict__ out,
                                                                      const uint8_t* __restrict__ in, size_t in_stride,
                                                                      size_t count) {
    __shared__ fe s_pre[kInvThreads];
    __shared__ fe s_suf[kInvThreads];
    __shared__ fe s_inv;
    const int t = threadIdx.x;
    for (size_t tile = (size_t)blockIdx.x * kInvTile; tile < count; tile += (size_t)gridDim.x * kInvTile) {
        fe acc;
        fe_set1(acc);
#pragma unroll 1
        for (int j = 0; j < kInvPer; j++) {
            size_t i = tile + (size_t)j * kInvThreads + t;
            if (i < count) {
                fe x;
                fe_load_nc(x, in + i * in_stride);
                fe_store(out + i * 32, acc);
                if (!fe_iszero(x)) fe_mul(acc, acc, x);
            }
        }
        s_pre[t] = acc;
        s_suf[t] = acc;
        __syncthreads();
        // inclusive prefix (s_pre) and inclusive suffix (s_suf) products, Hillis-Steele
        for (int o = 1; o < kInvThreads; o <<= 1) {
            fe p, q;
            bool hp = t >= o, hs = t + o < kInvThreads;
            if (hp) p = s_pre[t - o];
            if (hs) q = s_suf[t + o];
            __syncthreads();
            if (hp) {
                fe m = s_pre[t];
                fe_mul(m, m, p);
                s_pre[t] = m;
            }
            if (hs) {
                fe m = s_suf[t];
                fe_mul(m, m, q);
                s_suf[t] = m;
            }
            __syncthreads();
        }
        if (t < 32) {  // whole warp runs the chain (SIMT), lane 0 publishes
            fe total = s_pre[kInvThreads - 1], inv;
            fe_invert(inv, total);
            if (t == 0) s_inv = inv;
        }
        __syncthreads();
        fe inv = s_inv;
        if (t > 0) fe_mul(inv, inv, s_pre[t - 1]);
        if (t + 1 < kInvThreads) fe_mul(inv, inv, s_suf[t + 1]);
        // inv = 1 / (product of this thread's non-zero elements)
#pragma unroll 1
        for (int j = kInvPer - 1; j >= 0; j--) {
            size_t i = tile + (size_t)j * kInvThreads + t;
            if (i < count) {
                fe x, pre, r;
                fe_load_nc(x, in + i * in_stride);
                fe_load(pre, out + i * 32);
                if (fe_iszero(x)) {
                    fe_set0(r);
                } else {
                    fe_mul(r, inv, pre);
                    fe_mul(inv, inv, x);
                    fe_canon(r);
                }
                fe_store(out + i * 32, r);
            }
        }
        __syncthreads();
    }
}

// ---- batch inversion, large arrays: a tree of Montgomery levels, no CTA-wide scans, no thread ever waits ----
// Level k: every thread multiplies its kInvPer (interleaved, coalesced) elements, leaving the exclusive prefix
// products in out_k and its total in x_(k+1) — an array 16x smaller — until at most kInvDirect values are
// left, which are inverted directly, one per thread, in parallel (the only ~50 us of latency in the whole
// operation).  Then the levels unwind: inverse of a thread's total x the stored prefixes -> the inverses.
// 3 multiplications per element per level, levels shrink 16x: 3.2 multiplications per element overall
// (the single-kernel version above spends 4.1 and idles 7 of 8 warps during one inversion per 4096 elements).
// (1024 at first: the two extra levels below 2^14 values were four launches of 4 and 64 CTAs, 90 us of pure latency
// out of 267 us per 2^22 elements; one thread per value with the divsteps inversion takes 27 us for 2^14 of them)
static constexpr size_t kInvDirect = (size_t)1 << 14;
__global__ void __launch_bounds__(kInvThreads) fe_inv_forward_kernel(uint8_t* __restrict__ out,
                                                                     const uint8_t* __restrict__ in, size_t in_stride,
                                                                     size_t count, uint8_t* __restrict__ totals) {
    const size_t tile = (size_t)blockIdx.x * kInvTile;
    const int t = threadIdx.x;
    fe acc;
    fe_set1(acc);
#pragma unroll 1
    for (int j = 0; j < kInvPer; j++) {
        size_t i = tile + (size_t)j * kInvThreads + t;
        if (i < count) {
            fe x;
            fe_load_nc(x, in + i * in_stride);
            fe_store(out + i * 32, acc);
            if (!fe_iszero(x)) fe_mul(acc, acc, x);
        }
    }
    fe_store(totals + ((size_t)blockIdx.x * kInvThreads + t) * 32, acc);  // 1 for threads past the end
}
__global__ void __launch_bounds__(256) fe_inv_direct_kernel(uint8_t* __restrict__ out, const uint8_t* __restrict__ in,
                                                            size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fe x, r;
    fe_load_nc(x, in + i * 32);
    fe_invert(r, x);
    fe_store(out + i * 32, r);
}
__global__ void __launch_bounds__(kInvThreads) fe_inv_backward_kernel(uint8_t* __restrict__ out,
                                                                      const uint8_t* __restrict__ in, size_t in_stride,
                                                                      size_t count,
                                                                      const uint8_t* __restrict__ total_inverses,
                                                                      int canonical) {
    const size_t tile = (size_t)blockIdx.x * kInvTile;
    const int t = threadIdx.x;
    fe inv;
    fe_load_nc(inv, total_inverses + ((size_t)blockIdx.x * kInvThreads + t) * 32);
#pragma unroll 1
    for (int j = kInvPer - 1; j >= 0; j--) {
        size_t i = tile + (size_t)j * kInvThreads + t;
        if (i < count) {
            fe x, pre, r;
            fe_load_nc(x, in + i * in_stride);
            fe_load(pre, out + i * 32);
            if (fe_iszero(x)) {
                fe_set0(r);
            } else {
                fe_mul(r, inv, pre);
                fe_mul(inv, inv, x);
                if (canonical) fe_canon(r);
            }
            fe_store(out + i * 32, r);
        }
    }
}

// ---- inner product mod l -------------------------------------------------------------------------
// Each thread accumulates full 512-bit products into a 576-bit accumulator (no per-element
// reduction), warps and CTAs combine accumulators with shuffles / shared memory, and a single
// Barrett reduction mod l happens once per output.  64 IMAD + ~50 IADD per 64 bytes read.
struct Acc18 {
    uint32_t w[18];
};
__device__ __forceinline__ void acc_zero(Acc18& a) {
#pragma unroll
    for (int i = 0; i < 18; i++) a.w[i] = 0;
}
__device__ __forceinline__ void acc_add_product(Acc18& a, const fe& x, const fe& y) {
    uint32_t p[16];
    mul_wide(p, x, y);
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        c += (uint64_t)a.w[i] + p[i];
        a.w[i] = (uint32_t)c;
        c >>= 32;
    }
    c += a.w[16];
    a.w[16] = (uint32_t)c;
    a.w[17] += (uint32_t)(c >> 32);
}
__device__ __forceinline__ void acc_add(Acc18& a, const Acc18& b) {
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 18; i++) {
        c += (uint64_t)a.w[i] + b.w[i];
        a.w[i] = (uint32_t)c;
        c >>= 32;
    }
}
__device__ __forceinline__ void acc_warp_reduce(Acc18& a) {
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        Acc18 b;
#pragma unroll
        for (int i = 0; i < 18; i++) b.w[i] = __shfl_down_sync(0xffffffffu, a.w[i], o);
        acc_add(a, b);
    }
}
// x (576 bits) mod l = (lo512 mod l) + (hi64 * (2^512 mod l) mod l)
__device__ __forceinline__ void acc_reduce_mod_l(sc& r, const Acc18& a) {
    uint32_t lo[16], hi[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        lo[i] = a.w[i];
        hi[i] = 0;
    }
    // hi = a.w[16..17] * R512 (10 words)
    uint64_t c0 = 0, c1 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c0 += (uint64_t)a.w[16] * kScR512[i];
        hi[i] = (uint32_t)c0;
        c0 >>= 32;
    }
    hi[8] = (uint32_t)c0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c1 += (uint64_t)a.w[17] * kScR512[i] + hi[i + 1];
        hi[i + 1] = (uint32_t)c1;
        c1 >>= 32;
    }
    hi[9] = (uint32_t)c1;
    sc r0, r1;
    sc_reduce512(r0, lo);
    sc_reduce512(r1, hi);
    sc_add(r, r0, r1);
}
// CTA-level combine of the per-thread accumulators, result valid in thread 0
__device__ __forceinline__ void block_acc_reduce(Acc18& acc, Acc18* s_warp) {
    acc_warp_reduce(acc);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    if (lane == 0) s_warp[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        if (lane < nw) acc = s_warp[lane];
        else acc_zero(acc);
        acc_warp_reduce(acc);
    }
}
// block-level accumulate of a[i]*b[i] over [begin, end), result valid in thread 0
__device__ __forceinline__ void block_inner_product(Acc18& acc, const uint8_t* __restrict__ a,
                                                    const uint8_t* __restrict__ b, size_t begin, size_t end,
                                                    size_t stride, Acc18* s_warp) {
    acc_zero(acc);
    for (size_t i = begin; i < end; i += stride) {
        fe x, y;
        fe_load_nc(x, a + i * 32);
        fe_load_nc(y, b + i * 32);
        acc_add_product(acc, x, y);
    }
    block_acc_reduce(acc, s_warp);
}
// stage 1: grid-stride partial sums, one 72-byte accumulator per CTA.
// Measured and rejected (ncu: 0.55 of HBM and 50 % of the multiply pipe at once, long_scoreboard the top stall at 39 %
// achieved occupancy, 54 registers): two elements in flight per thread in registers (spills at 64 registers: 80 -> 99 us
// per 2^22 elements) and a three-stage cp.async pipeline through shared memory (90 us): the 16-word carry chains of the
// 576-bit accumulation, not the loads, are what the warps wait on.  Round 2, again: a register double buffer written so
// that the next pair of elements is requested before the current product (plain loads, then volatile asm loads with
// compiler barriers; 2 / 3 / 4 CTAs per SM) — ptxas sinks the four LDG.128 behind the 64 multiplies either way (they
// sit ~50 instructions ahead of the loop branch, 58 registers) and the kernel takes 85-88 us instead of 75.6.
// Floors at 2^22 elements: HBM 41 us, multiply pipe 30 us (ncu: pipe 50 % busy, issue 44 %, 25 warps / SM resident).
__global__ void __launch_bounds__(256) sc_ip_partial_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                            size_t n, uint32_t* __restrict__ partials) {
    __shared__ Acc18 s_warp[8];
    Acc18 acc;
    block_inner_product(acc, a, b, (size_t)blockIdx.x * blockDim.x + threadIdx.x, n, (size_t)gridDim.x * blockDim.x,
                        s_warp);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < 18; i++) partials[(size_t)blockIdx.x * 18 + i] = acc.w[i];
    }
}
// stage 2: one CTA folds the partials and reduces mod l
__global__ void __launch_bounds__(256) sc_ip_final_kernel(const uint32_t* __restrict__ partials, uint32_t nparts,
                                                          uint8_t* __restrict__ out) {
    __shared__ Acc18 s_warp[8];
    Acc18 acc;
    acc_zero(acc);
    for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
        Acc18 p;
#pragma unroll
        for (int j = 0; j < 18; j++) p.w[j] = partials[(size_t)i * 18 + j];
        acc_add(acc, p);
    }
    acc_warp_reduce(acc);
    int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) s_warp[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        if (lane < 8) acc = s_warp[lane];
        else acc_zero(acc);
        acc_warp_reduce(acc);
        if (lane == 0) {
            sc r;
            acc_reduce_mod_l(r, acc);
            sc_store(out, r);
        }
    }
}
// batched: one CTA per vector pair (vectors contiguous, n elements each)
__global__ void __launch_bounds__(128) sc_ip_batch_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                          size_t n, uint8_t* __restrict__ out) {
    __shared__ Acc18 s_warp[4];
    Acc18 acc;
    size_t base = (size_t)blockIdx.x * n;
    block_inner_product(acc, a, b, base + threadIdx.x, base + n, blockDim.x, s_warp);
    if (threadIdx.x == 0) {
        sc r;
        acc_reduce_mod_l(r, acc);
        sc_store(out + (size_t)blockIdx.x * 32, r);
    }
}

}  // namespace cbp

using namespace cbp;

extern "C" {

int bpk_fe_batch_device(int op, void* d_out, const void* d_a, const void* d_b, size_t count, void* stream) {
    if (op < BPK_FE_ADD || op > BPK_FE_SQR) return fail(BPK_ERR_ARG);
    if (!count) return BPK_OK;
    if (!d_out || !d_a || (op != BPK_FE_SQR && !d_b)) return fail(BPK_ERR_ARG);
    size_t blocks = (count + 255) / 256;
    size_t cap = (size_t)num_sms() * 16;
    unsigned grid = (unsigned)(blocks < cap ? blocks : cap);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* o = (uint8_t*)d_out;
    const uint8_t *a = (const uint8_t*)d_a, *b = (const uint8_t*)d_b;
    switch (op) {
        case BPK_FE_ADD: fe_batch_kernel<BPK_FE_ADD><<<grid, 256, 0, st>>>(o, a, b, count); break;
        case BPK_FE_SUB: fe_batch_kernel<BPK_FE_SUB><<<grid, 256, 0, st>>>(o, a, b, count); break;
        case BPK_FE_MUL: fe_batch_kernel<BPK_FE_MUL><<<grid, 256, 0, st>>>(o, a, b, count); break;
        default: fe_batch_kernel<BPK_FE_SQR><<<grid, 256, 0, st>>>(o, a, a, count); break;
    }
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
}  // extern "C"
namespace cbp {
static size_t inv_level_count(size_t count) { return (count + kInvTile - 1) / kInvTile * kInvThreads; }
// bytes of workspace for the tree version: per level k >= 1 the totals x_k and their prefix/inverse array
static size_t inv_tree_workspace(size_t count) {
    size_t bytes = 0;
    while (count > kInvDirect) {
        count = inv_level_count(count);
        bytes += 2 * count * 32;
    }
    return bytes;
}
static constexpr size_t kInvTreeMin = (size_t)1 << 16;  // below this one kernel is faster than 5+ launches
// out[i] = 1 / fe at (in + i * in_stride), 0 for 0; out is dense (32 B per element) and must not alias in.
// With a workspace of inv_tree_workspace(count) bytes large arrays take the tree path.
int fe_batch_invert_strided(uint8_t* d_out, const uint8_t* d_in, size_t in_stride, size_t count, cudaStream_t st,
                            uint8_t* d_ws, size_t ws_bytes) {
    if (!count) return BPK_OK;
    if (count < kInvTreeMin || !d_ws || ws_bytes < inv_tree_workspace(count)) {
        size_t tiles = (count + kInvTile - 1) / kInvTile;
        size_t cap = (size_t)num_sms() * 4;
        unsigned grid = (unsigned)(tiles < cap ? tiles : cap);
        fe_batch_invert_kernel<<<grid, kInvThreads, 0, st>>>(d_out, d_in, in_stride, count);
        CBP_CHECK_LAUNCH();
        return BPK_OK;
    }
    struct Level {
        uint8_t *out;
        const uint8_t* in;
        size_t stride, count;
    } lv[8];
    int nl = 0;
    lv[0] = {d_out, d_in, in_stride, count};
    uint8_t* w = d_ws;
    while (lv[nl].count > kInvDirect) {  // forward sweeps
        size_t next = inv_level_count(lv[nl].count);
        uint8_t* totals = w;
        uint8_t* next_out = w + next * 32;
        w += 2 * next * 32;
        fe_inv_forward_kernel<<<(unsigned)((lv[nl].count + kInvTile - 1) / kInvTile), kInvThreads, 0, st>>>(
            lv[nl].out, lv[nl].in, lv[nl].stride, lv[nl].count, totals);
        CBP_CHECK_LAUNCH();
        lv[nl + 1] = {next_out, totals, 32, next};
        nl++;
    }
    fe_inv_direct_kernel<<<(unsigned)((lv[nl].count + 255) / 256), 256, 0, st>>>(lv[nl].out, lv[nl].in, lv[nl].count);
    CBP_CHECK_LAUNCH();
    for (int k = nl - 1; k >= 0; k--) {  // unwind
        fe_inv_backward_kernel<<<(unsigned)((lv[k].count + kInvTile - 1) / kInvTile), kInvThreads, 0, st>>>(
            lv[k].out, lv[k].in, lv[k].stride, lv[k].count, lv[k + 1].out, k == 0);
        CBP_CHECK_LAUNCH();
    }
    return BPK_OK;
}
}  // namespace cbp
extern "C" {
int bpk_fe_batch_invert_workspace_bytes(size_t count, size_t* bytes) {
    if (!bytes) return fail(BPK_ERR_ARG);
    // optional: without it (or below 2^16 elements) one kernel does the whole job in the output buffer
    *bytes = count >= kInvTreeMin ? inv_tree_workspace(count) : 0;
    return BPK_OK;
}
int bpk_fe_batch_invert_device(void* d_out, const void* d_in, size_t count, void* d_workspace, size_t workspace_bytes,
                               void* stream) {
    if (!count) return BPK_OK;
    if (!d_out || !d_in || d_out == d_in) return fail(BPK_ERR_ARG);
    return fe_batch_invert_strided((uint8_t*)d_out, (const uint8_t*)d_in, 32, count, (cudaStream_t)stream,
                                   (uint8_t*)d_workspace, workspace_bytes);
}

static unsigned ip_grid(size_t n) {
    size_t blocks = (n + 255) / 256;
    size_t cap = (size_t)num_sms() * 4;
    if (blocks < 1) blocks = 1;
    return (unsigned)(blocks < cap ? blocks : cap);
}
int bpk_sc_inner_product_workspace_bytes(size_t n, size_t* bytes) {
    if (!bytes) return fail(BPK_ERR_ARG);
    *bytes = (size_t)ip_grid(n) * 72;
    return BPK_OK;
}
int bpk_sc_inner_product_device(void* d_out, const void* d_a, const void* d_b, size_t n, void* d_workspace,
                                size_t workspace_bytes, void* stream) {
    if (!d_out || (n && (!d_a || !d_b)) || !d_workspace) return fail(BPK_ERR_ARG);
    unsigned grid = ip_grid(n);
    if (workspace_bytes < (size_t)grid * 72) return fail(BPK_ERR_WORKSPACE);
    cudaStream_t st = (cudaStream_t)stream;
    sc_ip_partial_kernel<<<grid, 256, 0, st>>>((const uint8_t*)d_a, (const uint8_t*)d_b, n, (uint32_t*)d_workspace);
    CBP_CHECK_LAUNCH();
    sc_ip_final_kernel<<<1, 256, 0, st>>>((const uint32_t*)d_workspace, grid, (uint8_t*)d_out);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
int bpk_sc_inner_product_batch_device(void* d_out, const void* d_a, const void* d_b, size_t n, size_t num_vectors,
                                      void* stream) {
    if (!num_vectors) return BPK_OK;
    if (!d_out || (n && (!d_a || !d_b))) return fail(BPK_ERR_ARG);
    sc_ip_batch_kernel<<<(unsigned)num_vectors, 128, 0, (cudaStream_t)stream>>>((const uint8_t*)d_a, (const uint8_t*)d_b,
                                                                                  n, (uint8_t*)d_out);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}

// ---- host-pointer drop-ins (cuda_field_ops.cu:257-462,533-628; cuda_inner_product.cu:97-348) ------
static int host_batch(int op, fe25519* results, const fe25519* a, const fe25519* b, size_t count) {
    if (!count) return BPK_OK;
    uint8_t *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
    cudaError_t e;
    int rc = BPK_OK;
    size_t bytes = count * 32;
    if ((e = cudaMalloc(&d_a, bytes)) != cudaSuccess || (e = cudaMalloc(&d_o, bytes)) != cudaSuccess ||
        (b && (e = cudaMalloc(&d_b, bytes)) != cudaSuccess)) {
        rc = fail(BPK_ERR_CUDA, e);
        goto done;
    }
    if ((e = cudaMemcpyAsync(d_a, a, bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess ||
        (b && (e = cudaMemcpyAsync(d_b, b, bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess)) {
        rc = fail(BPK_ERR_CUDA, e);
        goto done;
    }
    rc = op >= 0 ? bpk_fe_batch_device(op, d_o, d_a, d_b, count, 0) : bpk_fe_batch_invert_device(d_o, d_a, count, 0, 0, 0);
    if (rc == BPK_OK) {
        e = cudaMemcpy(results, d_o, bytes, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(BPK_ERR_CUDA, e);
    }
done:
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_o);
    return rc;
}
void cuda_batch_field_add(fe25519* r, const fe25519* a, const fe25519* b, size_t n) { host_batch(BPK_FE_ADD, r, a, b, n); }
void cuda_batch_field_sub(fe25519* r, const fe25519* a, const fe25519* b, size_t n) { host_batch(BPK_FE_SUB, r, a, b, n); }
void cuda_batch_field_mul(fe25519* r, const fe25519* a, const fe25519* b, size_t n) { host_batch(BPK_FE_MUL, r, a, b, n); }
void cuda_batch_field_mul_karatsuba(fe25519* r, const fe25519* a, const fe25519* b, size_t n) {
    host_batch(BPK_FE_MUL, r, a, b, n);
}
void cuda_batch_field_square(fe25519* r, const fe25519* a, size_t n) { host_batch(BPK_FE_SQR, r, a, nullptr, n); }
void cuda_batch_field_invert(fe25519* r, const fe25519* a, size_t n) { host_batch(-1, r, a, nullptr, n); }
void cuda_soa_field_add(fe25519* r, const fe25519* a, const fe25519* b, size_t n) { host_batch(BPK_FE_ADD, r, a, b, n); }

static int host_inner_product(fe25519* results, const fe25519* a, const fe25519* b, size_t n, size_t num_vectors) {
    uint8_t *d_a = nullptr, *d_b = nullptr, *d_o = nullptr, *d_ws = nullptr;
    cudaError_t e;
    int rc = BPK_OK;
    size_t bytes = n * num_vectors * 32, ws = 0;
    bpk_sc_inner_product_workspace_bytes(n, &ws);
    if ((e = cudaMalloc(&d_a, bytes + 32)) != cudaSuccess || (e = cudaMalloc(&d_b, bytes + 32)) != cudaSuccess ||
        (e = cudaMalloc(&d_o, num_vectors * 32)) != cudaSuccess || (e = cudaMalloc(&d_ws, ws)) != cudaSuccess) {
        rc = fail(BPK_ERR_CUDA, e);
        goto done;
    }
    if (bytes && ((e = cudaMemcpyAsync(d_a, a, bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess ||
                  (e = cudaMemcpyAsync(d_b, b, bytes, cudaMemcpyHostToDevice, 0)) != cudaSuccess)) {
        rc = fail(BPK_ERR_CUDA, e);
        goto done;
    }
    rc = num_vectors == 1 ? bpk_sc_inner_product_device(d_o, d_a, d_b, n, d_ws, ws, 0)
                          : bpk_sc_inner_product_batch_device(d_o, d_a, d_b, n, num_vectors, 0);
    if (rc == BPK_OK) {
        e = cudaMemcpy(results, d_o, num_vectors * 32, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(BPK_ERR_CUDA, e);
    }
done:
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_o);
    cudaFree(d_ws);
    return rc;
}
void cuda_field_vector_inner_product(fe25519* result, const FieldVector* a, const FieldVector* b) {
    if (a->length != b->length) {  // cuda_inner_product.cu:100-103
        fprintf(stderr, "Error: Vector lengths must match for inner product\n");
        fail(BPK_ERR_ARG);
        return;
    }
    host_inner_product(result, a->elements, b->elements, a->length, 1);
}
void cuda_field_vector_inner_product_shared(fe25519* result, const FieldVector* a, const FieldVector* b) {
    cuda_field_vector_inner_product(result, a, b);
}
// cuda_inner_product.cu:302-348: num_vectors independent pairs; the vectors need not be contiguous
// on the host, so they are gathered one by one (all must have the same length, as in the reference)
void cuda_batch_field_vector_inner_product(fe25519* results, const FieldVector* a_vectors, const FieldVector* b_vectors,
                                           size_t num_vectors) {
    if (!num_vectors) return;
    size_t n = a_vectors[0].length;
    for (size_t v = 0; v < num_vectors; v++) {
        if (a_vectors[v].length != n || b_vectors[v].length != n) {
            fprintf(stderr, "Error: Vector lengths must match for inner product\n");
            fail(BPK_ERR_ARG);
            return;
        }
    }
    fe25519* ha = (fe25519*)malloc(n * num_vectors * 32 + 32);
    fe25519* hb = (fe25519*)malloc(n * num_vectors * 32 + 32);
    for (size_t v = 0; v < num_vectors; v++) {
        memcpy(ha + v * n, a_vectors[v].elements, n * 32);
        memcpy(hb + v * n, b_vectors[v].elements, n * 32);
    }
    if (num_vectors == 1) host_inner_product(results, ha, hb, n, 1);
    else host_inner_product(results, ha, hb, n, num_vectors);
    free(ha);
    free(hb);
}

}  // extern "C"
