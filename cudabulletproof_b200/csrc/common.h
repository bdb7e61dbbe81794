// common.h — shared host-side state of the library (error codes, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include <mutex>
#include "../../include/bpk.h"

namespace cbp {
extern std::atomic<uint64_t> g_launches;
extern std::atomic<int> g_last_error;
extern std::atomic<int> g_last_cuda_error;

inline int fail(int code, cudaError_t ce = cudaSuccess) {
    g_last_error.store(code);
    if (ce != cudaSuccess) g_last_cuda_error.store((int)ce);
    return code;
}
inline int fail_cuda(int ce) { return ce == 0 ? BPK_OK : fail(BPK_ERR_CUDA, (cudaError_t)ce); }
inline void count_launches(int n) { g_launches.fetch_add((uint64_t)n); }

// ---- per-device host state ------------------------------------------------------------------------
// Every piece of cached host-side state (side streams and events, upload buffers, generator tables) is keyed by
// the CUDA device that is current when an entry point is called, and every enqueue that touches shared side
// streams or events holds that device's lock: two host threads may call any entry points concurrently, on any
// streams, and a thread may switch devices between calls.  cudaStreamWaitEvent binds to the event's most recent
// record at the time of the call, so serialising the ENQUEUE (not the execution) is enough.
constexpr int kMaxDevices = 64;
int current_device_index();             // cudaGetDevice(), or -1 (error recorded) when out of range
std::recursive_mutex& device_mutex(int dev);
struct DeviceLock {
    int dev;
    std::unique_lock<std::recursive_mutex> lock;
    DeviceLock() : dev(current_device_index()) {
        if (dev >= 0) lock = std::unique_lock<std::recursive_mutex>(device_mutex(dev));
    }
    bool ok() const { return dev >= 0; }
};

// ---- measurement / test switches --------------------------------------------------------------------
// Read ONCE per process from the environment (first use), changeable afterwards only through
// bpk_debug_set_option(): nothing on a call path looks at the environment.
struct Options {
    int msm_slots = -1;       // CBP_MSM_SLOTS: -1 auto (slotted first digit pass from 2^19 points), 0 off, 1 on
    int msm_no2d = 0;         // CBP_MSM_NO2D: 1 = running-sum bucket reduction everywhere (cross-check)
    int host_chunk_log2 = 0;  // CBP_HOST_CHUNK_LOG2: chunk size of the host-pointer upload pipeline, 0 = default
    int prover_legacy = 0;    // CBP_PROVER_LEGACY: one-CTA-per-proof prover for every batch size
    int ngroups = 0;          // CBP_GROUPS="8,4,4": explicit window groups of the MSM pipeline, top down
    int groups[8] = {};
    int msm_small_max = -1;   // CBP_MSM_SMALL_MAX: largest n taken by the single-launch small-n MSM (-1 default)
    int ipa_composite_max = -1;  // largest vector length whose IPA rounds run unfolded (-1: default 4096; tests lower it)
    int msm_seg_shift = -1;   // log2 of the accumulation segment length (-1: max(64, 2 * mean run)); measurements
    int msm_acc_streams = -1; // accumulation of every window group on its own stream (see msm_run); -1 auto, 0 / 1
    int host_taper_log2 = 0;  // chunked host path: the last chunk is halved down to 2^this points (0: not at all)
    int host_trace = 0;       // chunked host path: print the timeline of every call to stderr (measurements)
    int debug_variant = 0;    // A/B switch for kernels under measurement (tools/probe_ops.py)
    int verify_group = -1;    // batch verification: proofs per combined identity (-1 auto: 12 from 256 proofs; 0 / 1: one by one)
    int msm_fused_front = 1;  // 1: scans and segment build of the MSM front end in one cooperative launch; 0: eleven launches
    int msm_graph = 1;        // 1: mid-size device MSMs replay a cached CUDA graph of their launch DAG (see msm_run_cached)
    int host_register = 0;    // CBP_HOST_REGISTER: 1 = page-lock large pageable caller buffers once and remember them
};
Options& options();

// optional per-kernel timing: no-ops unless bpk_profile_enable(1)
bool prof_enabled();
extern std::atomic<uint64_t> g_options_epoch;  // bumped by every bpk_debug_set_option: cached launch graphs are stale then
void prof_begin(int kind, cudaStream_t st);
void prof_end(int kind, cudaStream_t st);
}  // namespace cbp

#define CBP_CHECK_LAUNCH()                                                  \
    do {                                                                    \
        cudaError_t e_ = cudaGetLastError();                                \
        if (e_ != cudaSuccess) return cbp::fail(BPK_ERR_CUDA, e_);          \
        cbp::count_launches(1);                                             \
    } while (0)
#define CBP_CUDA(call)                                                      \
    do {                                                                    \
        cudaError_t e_ = (call);                                            \
        if (e_ != cudaSuccess) return cbp::fail(BPK_ERR_CUDA, e_);          \
    } while (0)
