// profile.cu — optional CUDA-event timing of the dominant kernels (used by bench.py for the
// roofline's "achieved" figure: events are recorded on the launching stream, inside the timed region).
#include <mutex>
#include <vector>
#include "common.h"

namespace cbp {
namespace {
struct Pair {
    cudaEvent_t a, b;
};
std::mutex g_mu;
bool g_enabled = false;
std::vector<Pair> g_pool[BPK_PROF_KINDS];   // created lazily, reused
size_t g_used[BPK_PROF_KINDS] = {0};
constexpr size_t kMaxPairs = 4096;
}  // namespace

bool prof_enabled() { return g_enabled; }
void prof_begin(int kind, cudaStream_t st) {
    if (!g_enabled) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_used[kind] >= kMaxPairs) return;
    if (g_used[kind] == g_pool[kind].size()) {
        Pair p;
        if (cudaEventCreate(&p.a) != cudaSuccess || cudaEventCreate(&p.b) != cudaSuccess) return;
        g_pool[kind].push_back(p);
    }
    cudaEventRecord(g_pool[kind][g_used[kind]].a, st);
}
void prof_end(int kind, cudaStream_t st) {
    if (!g_enabled) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_used[kind] >= g_pool[kind].size()) return;
    cudaEventRecord(g_pool[kind][g_used[kind]].b, st);
    g_used[kind]++;
}
}  // namespace cbp

using namespace cbp;

extern "C" {
int bpk_profile_enable(int enable) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_enabled = enable != 0;
    return BPK_OK;
}
int bpk_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_mu);
    for (int k = 0; k < BPK_PROF_KINDS; k++) g_used[k] = 0;
    return BPK_OK;
}
int bpk_profile_read(int kind, float* mean_ms, int* samples) {
    if (kind < 0 || kind >= BPK_PROF_KINDS || !mean_ms || !samples) return fail(BPK_ERR_ARG);
    std::lock_guard<std::mutex> lk(g_mu);
    double total = 0;
    int n = 0;
    for (size_t i = 0; i < g_used[kind]; i++) {
        float ms = 0;
        if (cudaEventSynchronize(g_pool[kind][i].b) != cudaSuccess) continue;
        if (cudaEventElapsedTime(&ms, g_pool[kind][i].a, g_pool[kind][i].b) != cudaSuccess) continue;
        total += ms;
        n++;
    }
    *mean_ms = n ? (float)(total / n) : 0.f;
    *samples = n;
    return BPK_OK;
}
}
