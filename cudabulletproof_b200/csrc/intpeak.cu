// intpeak.cu — the integer roofline denominator, measured in the caller's own run.
//
// MEASURED_PEAKS.json (driver-written) holds HBM and bf16 peaks only; the kernels of this library are bound by
// the issue rate of the 32x32->64 multiply-add (IMAD.WIDE.U32 on the fmaheavy pipe).  bpk_measure_int_peak runs
// the microbenchmarks below for a few milliseconds each on the current device, so that bench.py's roofline.peak
// comes from the same process, GPU and clocks as roofline.achieved:
//   [0] IMAD.WIDE.U32.X carry chains — mad.lo.cc / madc.hi.cc rows, exactly fe_mul's inner pattern  (the denominator)
//   [1] IMAD.WIDE.U32, independent accumulators
//   [2] IMAD (32-bit low product), independent accumulators
//   [3] IMAD.HI.U32, independent accumulators
//   [4] DFMA (FP64), independent accumulators — a 5x51-bit floating-point limb product would issue here
//   [5] the carry rows of [0] with as many DFMA beside them in the same thread: the IMAD.WIDE rate that is left when
//       the FP64 pipe is kept busy too — equal to [0] if the two pipes issue side by side, half of it if they do not
// Rates are lane operations per second (one warp instruction = 32).
#include "common.h"
#include "fe25519.cuh"

namespace cbp {

static constexpr int kPeakInner = 512;

__global__ void __launch_bounds__(256) peak_imad_wide_carry_kernel(uint32_t* out, uint32_t a, uint32_t b, int outer) {
    uint32_t c[9], d[9];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c[i] = threadIdx.x + i;
        d[i] = threadIdx.x * 3 + i;
    }
    uint32_t x0 = a + threadIdx.x, x1 = a ^ 0x55, x2 = a + 7, x3 = a * 3, y = b;
    for (int o = 0; o < outer; o++)
        for (int it = 0; it < kPeakInner; it++) {
            mad_row4(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], x0, x1, x2, x3, y);
            mad_row4(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], x1, x2, x3, x0, y);
            y = c[0] ^ d[1];
        }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) s ^= c[i] ^ d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_imad_wide_kernel(uint32_t* out, uint32_t a, uint32_t b, int outer) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t x = a + threadIdx.x, y = b;
    for (int o = 0; o < outer; o++)
        for (int it = 0; it < kPeakInner; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = (uint64_t)x * y + acc[i];
            x = (uint32_t)acc[0];
        }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
}
__global__ void __launch_bounds__(256) peak_imad_lo_kernel(uint32_t* out, uint32_t a, uint32_t b, int outer) {
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t x = a + threadIdx.x, y = b;
    for (int o = 0; o < outer; o++)
        for (int it = 0; it < kPeakInner; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = x * y + acc[i];
            x = acc[0];
        }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_imad_hi_kernel(uint32_t* out, uint32_t a, uint32_t b, int outer) {
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t x = a + threadIdx.x, y = b | 0x80000001u;
    for (int o = 0; o < outer; o++)
        for (int it = 0; it < kPeakInner; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(acc[i]) : "r"(x), "r"(y));
            x += acc[0];
        }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) peak_dfma_kernel(uint32_t* out, uint32_t a, uint32_t b, int outer) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = (double)(threadIdx.x + i);
    double x = 1.0 + 1e-9 * (double)(a + threadIdx.x), y = 1e-7 * (double)b;
    for (int o = 0; o < outer; o++)
        for (int it = 0; it < kPeakInner; it++) {
#pragma unroll
            for (int i = 0; i < 8; i++) acc[i] = fma(acc[i], x, y);
        }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)__double2ll_rn(s);
}

__global__ void __launch_bounds__(256) peak_imad_wide_beside_dfma_kernel(uint32_t* out, uint32_t a, uint32_t b, int outer) {
    uint32_t c[9], d[9];
    double acc[8];
#pragma unroll
    for (int i = 0; i < 9; i++) {
        c[i] = threadIdx.x + i;
        d[i] = threadIdx.x * 3 + i;
    }
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = (double)(threadIdx.x + i);
    uint32_t x0 = a + threadIdx.x, x1 = a ^ 0x55, x2 = a + 7, x3 = a * 3, y = b;
    const double fx = 1.0 + 1e-9 * (double)(a + threadIdx.x), fy = 1e-7 * (double)b;
    for (int o = 0; o < outer; o++)
        for (int it = 0; it < kPeakInner; it++) {
            mad_row4(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], x0, x1, x2, x3, y);
#pragma unroll
            for (int i = 0; i < 4; i++) acc[i] = fma(acc[i], fx, fy);
            mad_row4(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], x1, x2, x3, x0, y);
#pragma unroll
            for (int i = 4; i < 8; i++) acc[i] = fma(acc[i], fx, fy);
            y = c[0] ^ d[1];
        }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) s ^= c[i] ^ d[i];
    double fs = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) fs += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s ^ (uint32_t)__double2ll_rn(fs);
}

}  // namespace cbp

using namespace cbp;

extern "C" int bpk_measure_int_peak(double target_ms, double* rates, int count) {
    if (!rates || count < 1 || !(target_ms > 0)) return fail(BPK_ERR_ARG);
    int dev = current_device_index(), sms = 0;
    if (dev < 0) return BPK_ERR_CUDA;
    CBP_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int grid = sms * 8;  // 8 CTAs of 256 threads per SM: every sub-partition has 16 warps to pick from
    uint32_t* buf = nullptr;
    cudaStream_t st = nullptr;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    CBP_CUDA(cudaMalloc(&buf, (size_t)grid * 256 * 4));
    cudaError_t e = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    using Kern = void (*)(uint32_t*, uint32_t, uint32_t, int);
    const Kern kern[BPK_PEAK_KINDS] = {peak_imad_wide_carry_kernel, peak_imad_wide_kernel, peak_imad_lo_kernel,
                                       peak_imad_hi_kernel, peak_dfma_kernel, peak_imad_wide_beside_dfma_kernel};
    for (int k = 0; k < BPK_PEAK_KINDS && k < count && e == cudaSuccess; k++) {
        int outer = 1;
        float ms = 0;
        for (int attempt = 0; attempt < 12 && e == cudaSuccess; attempt++) {  // grow the launch until it lasts target_ms
            kern[k]<<<grid, 256, 0, st>>>(buf, 12345u, 6789u, attempt == 0 ? 1 : outer);  // first pass: warm-up
            if (attempt == 0) continue;
            cudaEventRecord(e0, st);
            kern[k]<<<grid, 256, 0, st>>>(buf, 12345u, 6789u, outer);
            cudaEventRecord(e1, st);
            if ((e = cudaEventSynchronize(e1)) != cudaSuccess) break;
            cudaEventElapsedTime(&ms, e0, e1);
            count_launches(2);
            if (ms >= target_ms || outer >= (1 << 14)) break;
            int grow = ms > 0.01f ? (int)(target_ms / ms * 1.2) + 1 : 8;
            outer *= grow < 2 ? 2 : (grow > 16 ? 16 : grow);
        }
        if (e != cudaSuccess) break;
        e = cudaGetLastError();
        // 8 multiply-adds per inner iteration in every kernel (two rows of four in the carry kernel)
        rates[k] = ms > 0 ? (double)grid * 256.0 * 8.0 * kPeakInner * outer / (ms * 1e-3) : 0.0;
    }
    if (e1) cudaEventDestroy(e1);
    if (e0) cudaEventDestroy(e0);
    if (st) cudaStreamDestroy(st);
    cudaFree(buf);
    if (e != cudaSuccess) return fail(BPK_ERR_CUDA, e);
    return BPK_OK;
}
