// capi_msm.cu — C ABI for the MSM path: device-resident bpk_msm_device and the host-pointer
// drop-ins cuda_point_vector_multi_scalar_mul{,_shared} (reference cuda_bulletproof_kernels.cu:62-207).
#include <stdio.h>
#include <stdlib.h>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>
#include <string.h>
#include "../../include/cuda_bulletproof.h"
#include "common.h"
#include "ge25519.cuh"
#include "msm.h"

namespace cbp {
std::atomic<uint64_t> g_launches{0};
std::atomic<uint64_t> g_options_epoch{0};
std::atomic<int> g_last_error{0};
std::atomic<int> g_last_cuda_error{0};

int current_device_index() {
    int dev = -1;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= kMaxDevices) {
        fail(BPK_ERR_CUDA, e != cudaSuccess ? e : cudaErrorInvalidDevice);
        return -1;
    }
    return dev;
}
std::recursive_mutex& device_mutex(int dev) {
    static std::recursive_mutex mu[kMaxDevices];
    return mu[dev];
}
Options& options() {
    static Options opt;
    static std::once_flag once;
    std::call_once(once, [] {
        auto num = [](const char* name, int dflt) {
            const char* v = getenv(name);
            return v && *v ? atoi(v) : dflt;
        };
        opt.msm_slots = num("CBP_MSM_SLOTS", -1);
        opt.msm_no2d = getenv("CBP_MSM_NO2D") ? 1 : 0;
        opt.host_chunk_log2 = num("CBP_HOST_CHUNK_LOG2", 0);
        opt.prover_legacy = getenv("CBP_PROVER_LEGACY") ? 1 : 0;
        opt.msm_small_max = num("CBP_MSM_SMALL_MAX", -1);
        opt.host_register = num("CBP_HOST_REGISTER", 0);
        opt.verify_group = num("CBP_VERIFY_GROUP", -1);
        opt.msm_graph = num("CBP_MSM_GRAPH", 1);
        opt.msm_fused_front = num("CBP_MSM_FUSED_FRONT", 1);
        opt.host_taper_log2 = num("CBP_HOST_TAPER_LOG2", 0);
        opt.msm_acc_streams = num("CBP_MSM_ACC_STREAMS", -1);
        if (const char* g = getenv("CBP_GROUPS")) {
            while (*g && opt.ngroups < 8) {
                int v = atoi(g);
                if (v >= 1) opt.groups[opt.ngroups++] = v;
                while (*g && *g != ',') g++;
                if (*g == ',') g++;
            }
        }
    });
    return opt;
}

// sum of `count` extended points by one warp (lane-strided partial sums, shuffle tree)
__global__ void point_sum_kernel(const uint8_t* __restrict__ pts, size_t count, int normalize,
                                 uint8_t* __restrict__ out) {
    int lane = threadIdx.x;
    ge_p3 acc;
    ge_p3_0(acc);
    for (size_t i = lane; i < count; i += 32) {
        ge_p3 q;
        ge_load(q, pts + i * 128);
        ge_add(acc, acc, q);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_p3 other;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            other.X.v[j] = __shfl_down_sync(0xffffffffu, acc.X.v[j], o);
            other.Y.v[j] = __shfl_down_sync(0xffffffffu, acc.Y.v[j], o);
            other.Z.v[j] = __shfl_down_sync(0xffffffffu, acc.Z.v[j], o);
            other.T.v[j] = __shfl_down_sync(0xffffffffu, acc.T.v[j], o);
        }
        ge_add(acc, acc, other);
    }
    if (lane == 0) {
        if (normalize) ge_normalize(acc);
        ge_store(out, acc);
    }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ ge_p3 basepoint() {
    ge_p3 B;
    B.X = fe{{0x8f25d51au, 0xc9562d60u, 0x9525a7b2u, 0x692cc760u, 0xfdd6dc5cu, 0xc0a4e231u, 0xcd6e53feu, 0x216936d3u}};
    B.Y = fe{{0x66666658u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u}};
    fe_set1(B.Z);
    fe_mul(B.T, B.X, B.Y);
    return B;
}
__global__ void __launch_bounds__(128) synth_points_kernel(uint8_t* __restrict__ pts, uint64_t* __restrict__ ks,
                                                           size_t n, uint64_t seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t k = splitmix64(seed ^ splitmix64(i)) | 1ull;
    uint32_t kw[2] = {(uint32_t)k, (uint32_t)(k >> 32)};
    ge_p3 B = basepoint(), r;
    ge_scalarmult_bits(r, kw, 64, B);
    ge_normalize(r);
    ge_store(pts + i * 128, r);
    if (ks) ks[i] = k;
}
// test hook: re-randomise the projective representation (X, Y, Z, T) -> (zX, zY, zZ, zT) with a per-point z != 0,
// after adding the point *torsion to every torsion_stride-th point.  The group element is unchanged by z, so
// full-size MSM tests can leave the Z = 1 fast path of msm_precompute_kernel without a new expected value.
__global__ void __launch_bounds__(128) projectivize_kernel(uint8_t* __restrict__ pts, size_t n, uint64_t seed,
                                                           const uint8_t* __restrict__ torsion,
                                                           uint32_t torsion_stride) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    ge_p3 p;
    ge_load(p, pts + i * 128);
    if (torsion && torsion_stride && i % torsion_stride == 0) {
        ge_p3 t;
        ge_load(t, torsion);
        ge_add(p, p, t);
    }
    fe z;
    uint64_t s = splitmix64(seed ^ splitmix64(i * 2 + 1));
    for (int j = 0; j < 4; j++) {
        s = splitmix64(s + j);
        z.v[2 * j] = (uint32_t)s;
        z.v[2 * j + 1] = (uint32_t)(s >> 32);
    }
    if (fe_iszero(z)) fe_set1(z);  // z = 0 (mod p) would destroy the point
    fe_mul(p.X, p.X, z);
    fe_mul(p.Y, p.Y, z);
    fe_mul(p.Z, p.Z, z);
    fe_mul(p.T, p.T, z);
    ge_store(pts + i * 128, p);
}
__global__ void synth_scalars_kernel(uint64_t* __restrict__ out, size_t n, uint64_t seed, int bits) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t s = splitmix64(seed ^ (0xA5A5A5A5ull + i * 4));
    for (int j = 0; j < 4; j++) {
        s = splitmix64(s + j);
        int lo = j * 64;
        uint64_t v = s;
        if (bits <= lo) v = 0;
        else if (bits < lo + 64) v &= (~0ull) >> (lo + 64 - bits);
        out[i * 4 + j] = v;
    }
}
}  // namespace cbp

using namespace cbp;

extern "C" {

const char* bpk_version(void) { return "cudabulletproof_b200 0.1 (sm_100a)"; }
int bpk_last_error(void) { return g_last_error.load(); }
int bpk_last_cuda_error(void) { return g_last_cuda_error.load(); }
int bpk_clear_last_error(void) {
    g_last_cuda_error.store(0);
    return g_last_error.exchange(0);
}
uint64_t bpk_kernel_launches(void) { return g_launches.load(); }
int bpk_debug_set_option(int option, long long value) {
    Options& o = options();
    g_options_epoch.fetch_add(1);
    switch (option) {
        case BPK_OPT_MSM_SLOTS: o.msm_slots = (int)value; break;
        case BPK_OPT_MSM_NO2D: o.msm_no2d = value != 0; break;
        case BPK_OPT_HOST_CHUNK_LOG2: o.host_chunk_log2 = (int)value; break;
        case BPK_OPT_PROVER_LEGACY: o.prover_legacy = value != 0; break;
        case BPK_OPT_MSM_SMALL_MAX: o.msm_small_max = (int)value; break;
        case BPK_OPT_HOST_REGISTER: o.host_register = value != 0; break;
        case BPK_OPT_IPA_COMPOSITE_MAX: o.ipa_composite_max = (int)value; break;
        case BPK_OPT_MSM_SEG_SHIFT: o.msm_seg_shift = (int)value; break;
        case BPK_OPT_MSM_ACC_STREAMS: o.msm_acc_streams = (int)value; break;
        case BPK_OPT_HOST_TRACE: o.host_trace = value != 0; break;
        case BPK_OPT_VERIFY_GROUP: o.verify_group = (int)value; break;
        case BPK_OPT_MSM_GRAPH: o.msm_graph = value != 0; break;
        case BPK_OPT_MSM_FUSED_FRONT: o.msm_fused_front = value != 0; break;
        case BPK_OPT_DEBUG_VARIANT: o.debug_variant = (int)value; break;
        case BPK_OPT_HOST_TAPER_LOG2: o.host_taper_log2 = (value >= 10 && value <= 30) ? (int)value : 0; break;
        case BPK_OPT_MSM_GROUPS:  // hex digits, top group first: 0x844 = 8, 4, 4; 0 = automatic
            o.ngroups = 0;
            for (int sh = 28; sh >= 0; sh -= 4) {
                int v = (int)((value >> sh) & 15);
                if (v && o.ngroups < 8) o.groups[o.ngroups++] = v;
            }
            break;
        default: return fail(BPK_ERR_ARG);
    }
    return BPK_OK;
}

int bpk_msm_window_bits(size_t n) { return msm_pick_window(n); }
// entry indices are (point index << 1 | sign) and all offsets 32-bit prefix sums over the n * W entries:
// n < 2^31 and n * ceil(256 / c) < 2^32 (n up to 2^27 at c = 16; larger inputs: split the call, the partial results
// add up with bpk_point_sum_device)
static bool msm_size_ok(size_t n, int window_bits) {
    if (n >= (1ull << 31)) return false;
    const int c = window_bits > 0 ? window_bits : msm_pick_window(n);
    return n * (size_t)((256 + c - 1) / c) < (1ull << 32);
}
int bpk_msm_workspace_bytes(size_t n, int window_bits, size_t* bytes) {
    if (!bytes || window_bits < 0 || window_bits > 17 || (window_bits > 0 && window_bits < 4) || !msm_size_ok(n, window_bits))
        return fail(BPK_ERR_ARG);
    MsmPlan p;
    msm_make_plan(&p, n, window_bits);
    *bytes = p.workspace_bytes;
    return BPK_OK;
}
int bpk_msm_device(const void* d_scalars, const void* d_points, size_t n, void* d_result, void* d_workspace,
                   size_t workspace_bytes, int window_bits, int normalize, void* stream) {
    if (!d_result || (n && (!d_scalars || !d_points || !d_workspace))) return fail(BPK_ERR_ARG);
    if (window_bits < 0 || window_bits > 17 || (window_bits > 0 && window_bits < 4) || !msm_size_ok(n, window_bits))
        return fail(BPK_ERR_ARG);
    MsmPlan p;
    msm_make_plan(&p, n, window_bits);
    if (n && workspace_bytes < p.workspace_bytes) return fail(BPK_ERR_WORKSPACE);
    int launches = 0;
    int rc = msm_run_cached(p, d_scalars, d_points, d_result, d_workspace, normalize, (cudaStream_t)stream, &launches);
    count_launches(launches);
    return fail_cuda(rc);
}
int bpk_msm_device_affine(const void* d_scalars, const void* d_xy, size_t n, void* d_result, void* d_workspace,
                          size_t workspace_bytes, int window_bits, int normalize, void* stream) {
    if (!d_result || (n && (!d_scalars || !d_xy || !d_workspace))) return fail(BPK_ERR_ARG);
    if (window_bits < 0 || window_bits > 17 || (window_bits > 0 && window_bits < 4) || !msm_size_ok(n, window_bits))
        return fail(BPK_ERR_ARG);
    MsmPlan p;
    msm_make_plan(&p, n, window_bits);
    if (n && workspace_bytes < p.workspace_bytes) return fail(BPK_ERR_WORKSPACE);
    int launches = 0;
    int rc = msm_run(p, d_scalars, d_xy, d_result, d_workspace, normalize, (cudaStream_t)stream, &launches, nullptr, 0,
                     kMsmAffineXY);
    count_launches(launches);
    return fail_cuda(rc);
}
int bpk_point_sum_device(const void* d_points, size_t count, void* d_result, int normalize, void* stream) {
    if (!d_result || (count && !d_points)) return fail(BPK_ERR_ARG);
    point_sum_kernel<<<1, 32, 0, (cudaStream_t)stream>>>((const uint8_t*)d_points, count, normalize, (uint8_t*)d_result);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
int bpk_synth_points_device(void* d_points, uint64_t* d_k, size_t n, uint64_t seed, void* stream) {
    if (n && !d_points) return fail(BPK_ERR_ARG);
    if (!n) return BPK_OK;
    synth_points_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>((uint8_t*)d_points, d_k, n, seed);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
int bpk_debug_projectivize_device(void* d_points, size_t n, uint64_t seed, const void* d_torsion,
                                   uint32_t torsion_stride, void* stream) {
    if (n && !d_points) return fail(BPK_ERR_ARG);
    if (!n) return BPK_OK;
    projectivize_kernel<<<(unsigned)((n + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (uint8_t*)d_points, n, seed, (const uint8_t*)d_torsion, torsion_stride);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
int bpk_synth_scalars_device(void* d_scalars, size_t n, uint64_t seed, int bits, void* stream) {
    if ((n && !d_scalars) || bits < 1 || bits > 256) return fail(BPK_ERR_ARG);
    if (!n) return BPK_OK;
    synth_scalars_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>((uint64_t*)d_scalars, n, seed, bits);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}

// ---- host-pointer drop-ins --------------------------------------------------------------------
// The reference wrapper mallocs, copies, synchronises and frees on every call
// (cuda_bulletproof_kernels.cu:77-115).  Here the device buffers are a grow-only cache (one per
// process, guarded by a mutex) and there is one synchronisation.  The upload (160 B per point over PCIe)
// costs more than the whole MSM, so inputs of 2^20 points and more are cut into chunks of kHostChunk
// points: one stream copies chunk after chunk while the compute stream sorts each arrived chunk and adds it
// into the SAME bucket sums (msm_run flags kMsmCarryIn / kMsmNoTail); the bucket reduction and the window
// combine run once, after the last chunk.
namespace {
// (independent MSMs per chunk were measured first — 2^17 / 2^18 / 2^19 / one piece = 7.3 / 5.8 / 4.6 / 5.3 ms at
// n = 2^20: a mid-size MSM is latency-bound, ~1.4 ms — hence the shared buckets)
constexpr size_t kHostChunk = (size_t)1 << 17;
constexpr size_t kHostChunkMin = (size_t)1 << 20;  // below this a single MSM (scalars first) is faster
constexpr int kMaxChunks = 64;  // every chunk keeps its sorted scalars (a front workspace) until its points arrive
struct HostPath {
    uint8_t *d_s = nullptr, *d_p = nullptr, *d_r = nullptr;
    uint8_t* d_ws[1] = {};
    uint8_t* d_front = nullptr;  // chunked path: the sorted scalars (front part of the MSM workspace) of every chunk
    size_t cap_s = 0, cap_p = 0, cap_ws[1] = {}, cap_front = 0;
    static constexpr int kFrontStreams = 4;  // the chunks' sorts are chains of small launches: several run side by side
    cudaStream_t main = nullptr, copy = nullptr, front[kFrontStreams] = {};
    cudaEvent_t ev_points = nullptr;
    std::vector<cudaEvent_t> ev_chunk, ev_scalars, ev_front;
    uint8_t* stage[2] = {};      // pinned staging buffers for pageable inputs (one chunk of scalars + points each)
    size_t cap_stage = 0;
    cudaEvent_t ev_stage[2] = {};  // the copy out of a staging buffer has finished
    bool ok = false;
};
HostPath g_hp[kMaxDevices];  // one per device, used under that device's lock
// Opt-in registration cache (BPK_OPT_HOST_REGISTER / CBP_HOST_REGISTER=1).  The reference's callers hand over
// plain malloc memory (bulletproof_vectors.cu:18,136); a copy from pageable memory is staged by the driver at a
// fraction of the PCIe rate.  With the option on, a pageable buffer of 1 MiB or more is page-locked in place the
// first time it is seen and remembered by address range, so that later calls with the same buffer copy at the
// pinned rate.  The caller must keep such a buffer allocated until bpk_host_release() (or process exit): that
// contract is why this is not the default.
struct Registered {
    const uint8_t* base;
    size_t bytes;
};
std::mutex g_reg_mu;
std::vector<Registered> g_registered;
void maybe_register(const void* ptr, size_t bytes) {
    if (!options().host_register || !ptr || bytes < ((size_t)1 << 20)) return;
    const uint8_t* p = (const uint8_t*)ptr;
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (const Registered& r : g_registered)
        if (p >= r.base && p + bytes <= r.base + r.bytes) return;
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, ptr) == cudaSuccess && attr.type != cudaMemoryTypeUnregistered) return;  // pinned already
    (void)cudaGetLastError();
    if (cudaHostRegister((void*)ptr, bytes, cudaHostRegisterDefault) == cudaSuccess) g_registered.push_back({p, bytes});
    else (void)cudaGetLastError();  // not fatal: the copy falls back to the driver's staging
}
// Pageable caller memory (plain malloc — what the reference's own callers hand over): the driver stages such a copy
// through its own bounce buffers at ~10 GB/s, a fifth of the PCIe rate.  Here a few host threads copy each chunk
// into one of two pinned staging buffers while the previous chunk is on the wire (16.0 -> ~6 ms for 2^20 pairs).
class CopyPool {
  public:
    // dst[0..bytes) = src[0..bytes), split over the pool's threads; returns when done
    void copy(uint8_t* dst, const uint8_t* src, size_t bytes) {
        if (bytes < ((size_t)1 << 20) || !start()) {
            memcpy(dst, src, bytes);
            return;
        }
        std::unique_lock<std::mutex> lk(mu_);
        dst_ = dst; src_ = src; bytes_ = bytes;
        pending_ = (int)threads_.size();
        generation_++;
        cv_.notify_all();
        done_.wait(lk, [&] { return pending_ == 0; });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
            cv_.notify_all();
        }
        for (std::thread& t : threads_) t.join();
    }

  private:
    bool start() {
        std::lock_guard<std::mutex> lk(mu_);
        if (!threads_.empty()) return true;
        unsigned hw = std::thread::hardware_concurrency();
        int nthreads = hw >= 16 ? 8 : (hw >= 4 ? (int)hw / 2 : 1);
        if (nthreads < 2) return false;
        for (int i = 0; i < nthreads; i++) threads_.emplace_back([this, i, nthreads] { work(i, nthreads); });
        return true;
    }
    void work(int idx, int nthreads) {
        uint64_t seen = 0;
        for (;;) {
            const uint8_t* src;
            uint8_t* dst;
            size_t bytes;
            {
                std::unique_lock<std::mutex> lk(mu_);
                cv_.wait(lk, [&] { return stop_ || generation_ != seen; });
                if (stop_) return;
                seen = generation_;
                src = src_; dst = dst_; bytes = bytes_;
            }
            const size_t per = ((bytes + nthreads - 1) / nthreads + 4095) & ~(size_t)4095;
            const size_t lo = per * (size_t)idx, hi = lo + per < bytes ? lo + per : bytes;
            if (lo < hi) memcpy(dst + lo, src + lo, hi - lo);
            {
                std::lock_guard<std::mutex> lk(mu_);
                if (--pending_ == 0) done_.notify_all();
            }
        }
    }
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::vector<std::thread> threads_;
    const uint8_t* src_ = nullptr;
    uint8_t* dst_ = nullptr;
    size_t bytes_ = 0;
    uint64_t generation_ = 0;
    int pending_ = 0;
    bool stop_ = false;
};
CopyPool g_copy_pool;
bool is_pageable(const void* ptr) {
    cudaPointerAttributes attr;
    cudaError_t e = cudaPointerGetAttributes(&attr, ptr);
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return true;
    }
    return attr.type == cudaMemoryTypeUnregistered;
}
cudaError_t grow(uint8_t** p, size_t* cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, need);
    if (e == cudaSuccess) *cap = need;
    return e;
}
}  // namespace

// Timeline of one chunked call (BPK_OPT_HOST_TRACE, measurements only): events on the copy / sort / bucket streams,
// printed relative to the first one after the call has been synchronised.
namespace {
struct TracePoint {
    const char* what;
    size_t chunk;
    cudaEvent_t ev;
    double host_us;
};
std::vector<TracePoint> g_trace;
std::chrono::steady_clock::time_point g_trace_t0;
void trace_mark(const char* what, size_t chunk, cudaStream_t st) {
    if (!options().host_trace) return;
    static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    if (g_trace.empty()) g_trace_t0 = std::chrono::steady_clock::now();
    cudaEvent_t ev;
    if (cudaEventCreate(&ev) != cudaSuccess) return;
    cudaEventRecord(ev, st);
    g_trace.push_back({what, chunk, ev, std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - g_trace_t0).count()});
}
void trace_dump() {
    if (g_trace.empty()) return;
    for (const TracePoint& t : g_trace) {
        float ms = 0;
        cudaEventElapsedTime(&ms, g_trace[0].ev, t.ev);
        fprintf(stderr, "[host-trace] %-12s chunk %2zu  gpu %8.3f ms   enqueued at %8.3f ms\n", t.what, t.chunk, ms, t.host_us / 1e3);
    }
    for (const TracePoint& t : g_trace) cudaEventDestroy(t.ev);
    g_trace.clear();
}
}  // namespace

// point_bytes = 128: the reference's extended ge25519; 64: affine x || y (bpk_msm_host_affine)
static int msm_host(ge25519* result, const fe25519* scalars_h, const void* points_h, size_t n, size_t point_bytes) {
    const int pflag = point_bytes == 64 ? kMsmAffineXY : 0;
    DeviceLock dlock;
    if (!dlock.ok()) return BPK_ERR_CUDA;
    HostPath& hp = g_hp[dlock.dev];
    cudaError_t e;
    if (!hp.ok) {
        // the scalar-side sorts of the chunked path (hp.front) have slack, the bucket sums behind each arriving chunk
        // of points (hp.main) do not
        int least = 0, greatest = 0;
        if ((e = cudaDeviceGetStreamPriorityRange(&least, &greatest)) != cudaSuccess) return fail(BPK_ERR_CUDA, e);
        if ((e = cudaStreamCreateWithPriority(&hp.main, cudaStreamNonBlocking, least > greatest ? least - 1 : least)) != cudaSuccess ||
            (e = cudaStreamCreateWithPriority(&hp.front[0], cudaStreamNonBlocking, least)) != cudaSuccess ||
            (e = cudaStreamCreateWithPriority(&hp.front[1], cudaStreamNonBlocking, least)) != cudaSuccess ||
            (e = cudaStreamCreateWithPriority(&hp.front[2], cudaStreamNonBlocking, least)) != cudaSuccess ||
            (e = cudaStreamCreateWithPriority(&hp.front[3], cudaStreamNonBlocking, least)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&hp.copy, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreateWithFlags(&hp.ev_points, cudaEventDisableTiming)) != cudaSuccess ||
            (e = cudaMalloc(&hp.d_r, 256)) != cudaSuccess)
            return fail(BPK_ERR_CUDA, e);
        hp.ok = true;
    }
    if (n) {
        if ((e = grow(&hp.d_s, &hp.cap_s, n * 32)) != cudaSuccess || (e = grow(&hp.d_p, &hp.cap_p, n * point_bytes)) != cudaSuccess)
            return fail(BPK_ERR_CUDA, e);
    }
    int launches = 0;
    const uint8_t* h_s = (const uint8_t*)scalars_h;
    const uint8_t* h_p = (const uint8_t*)points_h;
    maybe_register(h_s, n * 32);
    maybe_register(h_p, n * point_bytes);
    // pageable inputs take the chunked pipeline from 2^18 points: their upload is staged chunk by chunk anyway
    const bool staged = n >= ((size_t)1 << 18) && (is_pageable(h_s) || is_pageable(h_p));
    if (n < kHostChunkMin && !staged) {
        // one MSM; the scalars go first so that digit recoding / sorting overlaps the (4x larger) point upload
        MsmPlan p;
        msm_make_plan(&p, n, 0);
        if (n) {
            if ((e = grow(&hp.d_ws[0], &hp.cap_ws[0], p.workspace_bytes)) != cudaSuccess ||
                (e = cudaMemcpyAsync(hp.d_s, h_s, n * 32, cudaMemcpyHostToDevice, hp.main)) != cudaSuccess ||
                (e = cudaMemcpyAsync(hp.d_p, h_p, n * point_bytes, cudaMemcpyHostToDevice, hp.copy)) != cudaSuccess ||
                (e = cudaEventRecord(hp.ev_points, hp.copy)) != cudaSuccess)
                return fail(BPK_ERR_CUDA, e);
        }
        int rc = msm_run(p, hp.d_s, hp.d_p, hp.d_r, hp.d_ws[0], 1, hp.main, &launches, n ? hp.ev_points : nullptr, 0, pflag);
        count_launches(launches);
        if (rc) return fail_cuda(rc);
    } else {
        // (affine points halve the bytes per chunk: the bucket sums then need twice the points per chunk to keep up with
        // the wire — 2^20 pairs from pinned memory, chunks of 2^17 / 2^18 / 2^19: 3.08 / 2.63 / 2.80 ms; the reference layout
        // is best at 2^17: 3.70 / 3.75 / 4.61 ms)
        size_t chunk = pflag ? 2 * kHostChunk : kHostChunk;
        if (const int lg = options().host_chunk_log2; lg >= 15 && lg <= 26) chunk = (size_t)1 << lg;  // tuning knob
        size_t nchunks = (n + chunk - 1) / chunk;
        if (nchunks > (size_t)kMaxChunks) {  // keep the event pool bounded for enormous inputs
            nchunks = kMaxChunks;
            chunk = (n + nchunks - 1) / nchunks;
            nchunks = (n + chunk - 1) / chunk;
        }
        // every chunk adds into the SAME buckets (window width of the whole input, one workspace laid out for
        // a full chunk); only the last chunk runs the bucket reduction and the window combine.  The scalars travel
        // first (a fifth of the bytes): every chunk's scalar side — digit recoding, bucket sort, segments — then runs
        // on its own stream while the points are still on the wire, into a front workspace of its own, and what is
        // left behind an arriving chunk of points is the affine table and the bucket sums (measured at 2^20 pairs from
        // pinned memory, PCIe floor 3.0 ms: sort behind each chunk 3.95 ms, sorts ahead 3.5 ms).
        MsmPlan p_full;
        msm_make_plan(&p_full, chunk, msm_pick_window(n));
        // chunk boundaries; the last chunk — the only one whose bucket sums are not hidden under a copy — is halved
        // down to 2^host_taper_log2 points (its sort is off the critical path now, so small chunks cost little)
        std::vector<size_t> lo_of, cnt_of;
        for (size_t c = 0, lo = 0; c < nchunks; c++) {
            size_t cnt = n - lo < chunk ? n - lo : chunk;
            if (c + 1 == nchunks) {
                const size_t floor_sz = (size_t)1 << options().host_taper_log2;
                while (options().host_taper_log2 > 0 && cnt >= 2 * floor_sz && lo_of.size() + 2 < (size_t)kMaxChunks + 32) {
                    lo_of.push_back(lo);
                    cnt_of.push_back(cnt - cnt / 2);
                    lo += cnt - cnt / 2;
                    cnt = cnt / 2;
                }
            }
            lo_of.push_back(lo);
            cnt_of.push_back(cnt);
            lo += cnt;
        }
        nchunks = lo_of.size();
        auto plan_of = [&](size_t c) {
            MsmPlan q = p_full;
            q.n = cnt_of[c];
            return q;
        };
        const size_t front_bytes = (msm_front_bytes(p_full) + 255) & ~(size_t)255;
        if ((e = grow(&hp.d_ws[0], &hp.cap_ws[0], p_full.workspace_bytes)) != cudaSuccess ||
            (e = grow(&hp.d_front, &hp.cap_front, front_bytes * nchunks)) != cudaSuccess)
            return fail(BPK_ERR_CUDA, e);
        for (std::vector<cudaEvent_t>* pool : {&hp.ev_chunk, &hp.ev_scalars, &hp.ev_front})
            while (pool->size() < nchunks) {
                cudaEvent_t ev;
                if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return fail(BPK_ERR_CUDA, e);
                pool->push_back(ev);
            }
        // pageable inputs go through two pinned staging buffers, filled by the copy pool while the previous piece
        // is on the wire
        const size_t stage_bytes = chunk * 128;  // (affine points need half of it)
        if (staged && hp.cap_stage < stage_bytes) {
            for (int b = 0; b < 2; b++) {
                if (hp.stage[b]) cudaFreeHost(hp.stage[b]);
                hp.stage[b] = nullptr;
                if (!hp.ev_stage[b] && (e = cudaEventCreateWithFlags(&hp.ev_stage[b], cudaEventDisableTiming)) != cudaSuccess)
                    return fail(BPK_ERR_CUDA, e);
            }
            hp.cap_stage = 0;
            if ((e = cudaMallocHost(&hp.stage[0], stage_bytes)) != cudaSuccess ||
                (e = cudaMallocHost(&hp.stage[1], stage_bytes)) != cudaSuccess)
                return fail(BPK_ERR_CUDA, e);
            hp.cap_stage = stage_bytes;
        }
        size_t pieces = 0;  // staging pieces issued so far (they alternate between the two buffers)
        auto upload = [&](uint8_t* dst, const uint8_t* src, size_t bytes) -> cudaError_t {
            if (!staged) return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, hp.copy);
            for (size_t off = 0; off < bytes; off += hp.cap_stage, pieces++) {
                const size_t len = bytes - off < hp.cap_stage ? bytes - off : hp.cap_stage;
                uint8_t* buf = hp.stage[pieces & 1];
                cudaError_t r;
                if (pieces >= 2 && (r = cudaEventSynchronize(hp.ev_stage[pieces & 1])) != cudaSuccess) return r;
                g_copy_pool.copy(buf, src + off, len);
                if ((r = cudaMemcpyAsync(dst + off, buf, len, cudaMemcpyHostToDevice, hp.copy)) != cudaSuccess ||
                    (r = cudaEventRecord(hp.ev_stage[pieces & 1], hp.copy)) != cudaSuccess)
                    return r;
            }
            return cudaSuccess;
        };
        auto chunk_flags = [&](size_t c) { return (c > 0 ? kMsmCarryIn : 0) | (c + 1 == nchunks ? 0 : kMsmNoTail); };
        // Enqueue order.  From pinned memory every copy is queued before the first kernel (the launches of a sort
        // take the host ~60 us per chunk, which would otherwise delay the point copies behind them); staged pieces
        // are copied by the calling thread, so there the sorts are queued between the scalars and the points and
        // each chunk's bucket sums right behind its points.
        auto copy_scalars = [&](size_t c) -> cudaError_t {
            cudaError_t r;
            if (c == 0) trace_mark("start", 0, hp.copy);
            if ((r = upload(hp.d_s + lo_of[c] * 32, h_s + lo_of[c] * 32, cnt_of[c] * 32)) != cudaSuccess) return r;
            return cudaEventRecord(hp.ev_scalars[c], hp.copy);
        };
        auto copy_points = [&](size_t c) -> cudaError_t {
            cudaError_t r;
            if ((r = upload(hp.d_p + lo_of[c] * point_bytes, h_p + lo_of[c] * point_bytes, cnt_of[c] * point_bytes)) != cudaSuccess)
                return r;
            r = cudaEventRecord(hp.ev_chunk[c], hp.copy);
            trace_mark("points-in", c, hp.copy);
            return r;
        };
        // the scalar side of chunk c on one of the hp.front streams (it cannot overtake work of an earlier call that still reads the
        // front workspaces: every call ends with a synchronisation)
        auto sort_chunk = [&](size_t c) -> int {
            cudaError_t r;
            cudaStream_t fs = hp.front[c % HostPath::kFrontStreams];
            if ((r = cudaStreamWaitEvent(fs, hp.ev_scalars[c], 0)) != cudaSuccess) return (int)r;
            int nl = 0;
            int rc = msm_run(plan_of(c), hp.d_s + lo_of[c] * 32, nullptr, hp.d_r, hp.d_ws[0], 1, fs, &nl, nullptr, 0,
                             chunk_flags(c) | kMsmFrontOnly, hp.d_front + c * front_bytes);
            launches += nl;
            if (rc == 0) rc = (int)cudaEventRecord(hp.ev_front[c], fs);
            trace_mark("sorted", c, fs);
            return rc;
        };
        // table and bucket sums of chunk c on hp.main; behind the last chunk the reduction and the window combine
        auto sum_chunk = [&](size_t c) -> int {
            cudaError_t r;
            if ((r = cudaStreamWaitEvent(hp.main, hp.ev_chunk[c], 0)) != cudaSuccess ||
                (r = cudaStreamWaitEvent(hp.main, hp.ev_front[c], 0)) != cudaSuccess)
                return (int)r;
            int nl = 0;
            int rc = msm_run(plan_of(c), hp.d_s + lo_of[c] * 32, hp.d_p + lo_of[c] * point_bytes, hp.d_r, hp.d_ws[0], 1, hp.main,
                             &nl, nullptr, 0, chunk_flags(c) | kMsmBackOnly | pflag, hp.d_front + c * front_bytes);
            launches += nl;
            trace_mark("summed", c, hp.main);
            return rc;
        };
        int rc = 0;
        if (!staged) {
            {  // all scalars in two copies: fewer, larger transfers (one per chunk: 3.655 ms, two in all: 3.63 ms at 2^20 pairs);
               // the sorts have slack, the first points do not
                trace_mark("start", 0, hp.copy);
                const size_t half_c = nchunks / 2, split = lo_of[half_c];
                cudaError_t r = cudaMemcpyAsync(hp.d_s, h_s, split * 32, cudaMemcpyHostToDevice, hp.copy);
                for (size_t c = 0; c < half_c && r == cudaSuccess; c++) r = cudaEventRecord(hp.ev_scalars[c], hp.copy);
                if (r == cudaSuccess) r = cudaMemcpyAsync(hp.d_s + split * 32, h_s + split * 32, (n - split) * 32, cudaMemcpyHostToDevice, hp.copy);
                for (size_t c = half_c; c < nchunks && r == cudaSuccess; c++) r = cudaEventRecord(hp.ev_scalars[c], hp.copy);
                rc = (int)r;
            }
            for (size_t c = 0; c < nchunks && rc == 0; c++) rc = (int)copy_points(c);
            for (size_t c = 0; c < nchunks && rc == 0; c++) rc = sort_chunk(c);
            for (size_t c = 0; c < nchunks && rc == 0; c++) rc = sum_chunk(c);
        } else {
            // a helper thread feeds the staging buffers (scalars, then points) while this thread queues the kernels; a
            // kernel may only be made to wait for an event that has been recorded, hence the progress counter
            std::atomic<size_t> uploaded{0};
            std::atomic<int> upload_rc{0};
            const int dev = dlock.dev;
            std::thread feeder([&] {
                cudaError_t r = cudaSetDevice(dev);
                for (size_t u = 0; u < 2 * nchunks && r == cudaSuccess; u++) {
                    r = u < nchunks ? copy_scalars(u) : copy_points(u - nchunks);
                    if (r == cudaSuccess) uploaded.store(u + 1, std::memory_order_release);
                }
                if (r != cudaSuccess) {
                    upload_rc.store((int)r);
                    uploaded.store(2 * nchunks, std::memory_order_release);  // let the waiting thread go; it checks upload_rc
                }
            });
            auto wait_for = [&](size_t units) {
                while (uploaded.load(std::memory_order_acquire) < units) std::this_thread::yield();
                return upload_rc.load();
            };
            for (size_t c = 0; c < nchunks && rc == 0; c++)
                if ((rc = wait_for(c + 1)) == 0) rc = sort_chunk(c);
            for (size_t c = 0; c < nchunks && rc == 0; c++)
                if ((rc = wait_for(nchunks + c + 1)) == 0) rc = sum_chunk(c);
            feeder.join();
            if (rc == 0) rc = upload_rc.load();
        }
        if (rc) {
            count_launches(launches);
            return fail_cuda(rc);
        }
        count_launches(launches);
    }
    ge25519 tmp;
    e = cudaMemcpyAsync(&tmp, hp.d_r, 128, cudaMemcpyDeviceToHost, hp.main);
    if (e == cudaSuccess) e = cudaStreamSynchronize(hp.main);
    if (e != cudaSuccess) return fail(BPK_ERR_CUDA, e);
    trace_dump();
    *result = tmp;
    return BPK_OK;
}

int bpk_msm_host_affine(void* result, const void* scalars, const void* xy, size_t n) {
    if (!result || (n && (!scalars || !xy))) return fail(BPK_ERR_ARG);
    return msm_host((ge25519*)result, (const fe25519*)scalars, xy, n, 64);
}
int bpk_host_release(void) {
    std::lock_guard<std::mutex> lk(g_reg_mu);
    for (const Registered& r : g_registered) cudaHostUnregister((void*)r.base);
    g_registered.clear();
    (void)cudaGetLastError();
    return BPK_OK;
}

void cuda_point_vector_multi_scalar_mul(ge25519* result, const FieldVector* scalars, const PointVector* points) {
    if (scalars->length != points->length) {  // cuda_bulletproof_kernels.cu:65-68: message, result untouched
        fprintf(stderr, "Error: Vector lengths must match for multi-scalar multiplication\n");
        fail(BPK_ERR_ARG);
        return;
    }
    msm_host(result, scalars->elements, points->elements, scalars->length, 128);
}
void cuda_point_vector_multi_scalar_mul_shared(ge25519* result, const FieldVector* scalars, const PointVector* points) {
    cuda_point_vector_multi_scalar_mul(result, scalars, points);
}

}  // extern "C"
