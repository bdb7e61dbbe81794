// fe8_debug.cu — test hook for the octet-form arithmetic of fe8.cuh (one 32-bit word per lane): runs single
// operations on arrays so that the GPU parity tests can pin them against the CPU oracle and the thread-level code.
#include "common.h"
#include "fe8.cuh"

namespace cbp {

// field ops: one operation per OCTET (4 per warp); point ops: one per WARP
__global__ void __launch_bounds__(128) fe8_op_kernel(int op, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                      uint8_t* __restrict__ out, size_t count) {
    const Fe8Lane L = fe8_lane();
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (op <= BPK_FE8_SUB) {
        size_t i = warp * 4 + L.oct;
        const bool live = i < count;
        if (!live) i = count - 1;  // keep the warp converged
        const uint32_t x = reinterpret_cast<const uint32_t*>(a + i * 32)[L.j];
        const uint32_t y = reinterpret_cast<const uint32_t*>(b + i * 32)[L.j];
        uint32_t r;
        if (op == BPK_FE8_MUL) {
            r = fe8_mul(x, y, L);  // any 32-bit words
        } else {
            const uint32_t tx = fe8_normalize((uint64_t)x, L), ty = fe8_normalize((uint64_t)y, L);
            r = op == BPK_FE8_ADD ? fe8_add(tx, ty, L) : fe8_sub(tx, ty, L);
        }
        fe full;
        fe8_gather(full, r, L);
        fe_canon(full);
        if (live && L.j == 0) fe_store(out + i * 32, full);
        return;
    }
    size_t i = warp;
    const bool live = i < count;
    if (!live) i = count - 1;
    ge8 p, q, r;
    ge8_load(p, a + i * 128, L);
    if (op == BPK_GE8_DBL) {
        ge8_dbl(r, p, L);
    } else if (op == BPK_GE8_ADD) {
        ge8_load(q, b + i * 128, L);
        ge8_add(r, p, q, L);
    } else if (op == BPK_GE8_ADD_CACHED) {
        ge8_load(q, b + i * 128, L);
        ge8_cached c;
        ge8_to_cached(c, q, L);
        ge8_add_cached(r, p, c, L);
    } else if (op == BPK_GE8_DBL_CHAIN) {  // 2^64 p: a dependent chain, as in the window combine
        r = p;
#pragma unroll 1
        for (int s = 0; s < 64; s++) ge8_dbl(r, r, L);
    } else {
        r = p;
    }
    if (live) ge8_store_normalized(out + i * 128, r, L);
}

}  // namespace cbp

using namespace cbp;

extern "C" int bpk_debug_fe8_op_device(int op, const void* d_a, const void* d_b, void* d_out, size_t count, void* stream) {
    if (!count) return BPK_OK;
    if (!d_a || !d_out || op < 0 || op > BPK_GE8_NORMALIZE) return fail(BPK_ERR_ARG);
    if ((op <= BPK_FE8_SUB || op == BPK_GE8_ADD || op == BPK_GE8_ADD_CACHED) && !d_b) return fail(BPK_ERR_ARG);
    const size_t warps = op <= BPK_FE8_SUB ? (count + 3) / 4 : count;
    fe8_op_kernel<<<(unsigned)((warps * 32 + 127) / 128), 128, 0, (cudaStream_t)stream>>>(op, (const uint8_t*)d_a,
                                                                                        (const uint8_t*)d_b, (uint8_t*)d_out, count);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
