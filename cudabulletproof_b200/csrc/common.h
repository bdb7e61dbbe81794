// common.h — shared host-side state of the library (error codes, launch counter).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
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
// optional per-kernel timing: no-ops unless bpk_profile_enable(1)
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
