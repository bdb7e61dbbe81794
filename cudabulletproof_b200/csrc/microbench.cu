// microbench.cu — standalone integer-pipe microbenchmarks for B200 (not part of the library).
// Measures the roofline denominators DESIGN.md uses for the IMAD-bound kernels:
//   imad_wide_indep : independent IMAD.WIDE.U32 chains            -> peak IMAD.WIDE issue rate
//   imad_lo_indep   : independent 32-bit IMAD chains              -> peak IMAD issue rate
//   imad_wide_carry : mad.lo.cc/madc.hi.cc carry chains (fe_mul's inner pattern)
//   fe_mul / fe_sq / ge_madd / ge_dbl : sustained field / group operation rates
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o microbench microbench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "fe8.cuh"
#include "sc25519.cuh"
using namespace cbp;

#define ITERS 4096

__global__ void __launch_bounds__(256) k_imad_wide_indep(uint64_t* out, uint32_t a, uint32_t b) {
    uint64_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = (uint64_t)x * y + acc[i];
        x = (uint32_t)acc[0];  // keep operands live / varying
    }
    uint64_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_imad_lo_indep(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t acc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) acc[i] = threadIdx.x + i;
    uint32_t x = a + threadIdx.x, y = b;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) acc[i] = x * y + acc[i];
        x = acc[0];
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_imad_wide_carry(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t c[9], d[9];
#pragma unroll
    for (int i = 0; i < 9; i++) { c[i] = threadIdx.x + i; d[i] = threadIdx.x * 3 + i; }
    uint32_t x0 = a + threadIdx.x, x1 = a ^ 0x55, x2 = a + 7, x3 = a * 3, y = b;
    for (int it = 0; it < ITERS; it++) {
        mad_row4(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], x0, x1, x2, x3, y);
        mad_row4(d[0], d[1], d[2], d[3], d[4], d[5], d[6], d[7], d[8], x1, x2, x3, x0, y);
        y = c[0] ^ d[1];
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) s ^= c[i] ^ d[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_fe_mul(uint32_t* out, uint32_t seed) {
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.v[i] = seed * (i + 1) + threadIdx.x; b.v[i] = seed ^ (i * 77 + blockIdx.x); }
    for (int it = 0; it < ITERS / 4; it++) {
        fe_mul(a, a, b);
        fe_mul(b, b, a);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a.v[0] ^ b.v[3];
}
__global__ void __launch_bounds__(256) k_fe_sq(uint32_t* out, uint32_t seed) {
    fe a, b;
#pragma unroll
    for (int i = 0; i < 8; i++) { a.v[i] = seed * (i + 1) + threadIdx.x; b.v[i] = seed ^ (i * 77 + blockIdx.x); }
    for (int it = 0; it < ITERS / 4; it++) {
        fe_sq(a, a);
        fe_sq(b, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a.v[0] ^ b.v[3];
}
__global__ void __launch_bounds__(128, 4) k_ge_madd(uint32_t* out, uint32_t seed) {
    ge_p3 p;
    ge_niels q;
    ge_p3_0(p);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        p.X.v[i] = seed + i + threadIdx.x;
        q.yplusx.v[i] = seed * 3 + i;
        q.yminusx.v[i] = seed * 5 + i + blockIdx.x;
        q.xy2d.v[i] = seed * 7 + i;
    }
    for (int it = 0; it < ITERS / 16; it++) ge_madd(p, p, q, (it & 1) != 0);
    out[blockIdx.x * blockDim.x + threadIdx.x] = p.X.v[0] ^ p.T.v[1] ^ p.Y.v[2] ^ p.Z.v[3];
}
__global__ void __launch_bounds__(128, 4) k_ge_dbl(uint32_t* out, uint32_t seed) {
    ge_p3 p;
    ge_p3_0(p);
#pragma unroll
    for (int i = 0; i < 8; i++) { p.X.v[i] = seed + i + threadIdx.x; p.Y.v[i] = seed * 9 + blockIdx.x; }
    for (int it = 0; it < ITERS / 16; it++) ge_dbl(p, p);
    out[blockIdx.x * blockDim.x + threadIdx.x] = p.X.v[0] ^ p.T.v[1] ^ p.Y.v[2] ^ p.Z.v[3];
}

// latency probes: ONE warp per SM, dependent chain of point operations
__global__ void __launch_bounds__(32) k_lat_add(uint32_t* out, uint32_t seed, int quad) {
    ge_p3 p, q;
    ge_p3_0(p);
    ge_p3_0(q);
#pragma unroll
    for (int i = 0; i < 8; i++) { p.X.v[i] = seed + i + (threadIdx.x >> 2); q.X.v[i] = seed * 3 + i; q.T.v[i] = seed * 5 + blockIdx.x; }
    for (int it = 0; it < 256; it++) {
        if (quad) ge_add_quad(p, p, q); else ge_add(p, p, q);
    }
    out[blockIdx.x * 32 + threadIdx.x] = p.X.v[0] ^ p.T.v[1] ^ p.Y.v[2] ^ p.Z.v[3];
}
__global__ void __launch_bounds__(32) k_lat_dbl(uint32_t* out, uint32_t seed, int quad) {
    ge_p3 p;
    ge_p3_0(p);
#pragma unroll
    for (int i = 0; i < 8; i++) { p.X.v[i] = seed + i + (threadIdx.x >> 2); p.Y.v[i] = seed * 9 + blockIdx.x; }
    for (int it = 0; it < 256; it++) {
        if (quad) ge_dbl_quad(p, p); else ge_dbl(p, p);
    }
    out[blockIdx.x * 32 + threadIdx.x] = p.X.v[0] ^ p.T.v[1] ^ p.Y.v[2] ^ p.Z.v[3];
}
// what: 0 divsteps fe_invert, same value in every lane; 1 Fermat chain; 2 divsteps, a different value per lane
// (divergent); 3 / 4 / 5 the same for scalars mod l
__global__ void __launch_bounds__(32) k_lat_inv(uint32_t* out, uint32_t seed, int what) {
    fe a, r;
    const uint32_t vary = (what == 2 || what == 5) ? threadIdx.x * 0x9E3779B9u : 0u;
#pragma unroll
    for (int i = 0; i < 8; i++) a.v[i] = (seed + i) * 0x85EBCA6Bu + vary * (i + 1);
    a.v[7] &= 0x0FFFFFFFu;
    if (what == 0 || what == 2) fe_invert(r, a);
    else if (what == 1) fe_invert_fermat(r, a);
    else {
        sc x, y;
#pragma unroll
        for (int i = 0; i < 8; i++) x.v[i] = a.v[i];
        if (what == 4) sc_invert_fermat(y, x);
        else sc_invert(y, x);
#pragma unroll
        for (int i = 0; i < 8; i++) r.v[i] = y.v[i];
    }
    out[blockIdx.x * 32 + threadIdx.x] = r.v[0];
}

// the same chains in octet form (fe8.cuh: one word per lane, the warp shares every point operation); one kernel
// per operation so that no measurement shares its loop body (and its instruction fetches) with another
template <int WHAT>
__global__ void __launch_bounds__(32) k_lat8(uint32_t* out, uint32_t seed) {
    const Fe8Lane L = fe8_lane();
    ge8 p, q;
    ge8_identity(p, L);
    ge8_identity(q, L);
    p.X = seed + L.j;
    p.Y = seed * 9 + L.j + blockIdx.x;
    q.X = seed * 3 + L.j;
    q.T = seed * 5 + blockIdx.x;
    ge8_cached c;
    ge8_to_cached(c, q, L);
    uint32_t z = p.X;
    if (WHAT == 3) {
        z = fe8_invert(p.Y, L);
    } else if (WHAT == 4) {
        uint32_t y = p.Y;
#pragma unroll 1
        for (int it = 0; it < 256; it++) z = fe8_mul(z, y, L);
    } else if (WHAT == 5) {
#pragma unroll 1
        for (int it = 0; it < 256; it++) z = fe8_sub(z, p.Y, L);
    } else if (WHAT == 6) {
#pragma unroll 1
        for (int it = 0; it < 256; it++) z = fe8_mul(z, z, L);
    } else {
#pragma unroll 1
        for (int it = 0; it < 256; it++) {
            if (WHAT == 0) ge8_dbl(p, p, L);
            else if (WHAT == 1) ge8_add(p, p, q, L);
            else ge8_add_cached(p, p, c, L);
        }
    }
    out[blockIdx.x * 32 + threadIdx.x] = p.X ^ p.T ^ p.Y ^ p.Z ^ z;
}
template <int WHAT>
static void lat8_row(const char* name, int sms, void* buf);

template <typename F>
static double time_ms(F launch, int reps) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / reps;
}

template <int WHAT>
static void lat8_row(const char* name, int sms, void* buf) {
    double ms = time_ms([&] { k_lat8<WHAT><<<sms, 32>>>((uint32_t*)buf, 99); }, 5);
    if (WHAT != 3) printf("{\"bench\": \"%s\", \"us_per_op\": %.3f}\n", name, ms * 1e3 / 256);
    else printf("{\"bench\": \"%s\", \"us\": %.3f}\n", name, ms * 1e3);
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    void* buf;
    cudaMalloc(&buf, (size_t)sms * 64 * 256 * 8);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d}\n", prop.name, sms, prop.clockRate);
    for (int bps = 1; bps <= 8; bps *= 2) {
        int grid = sms * bps;
        double threads = (double)grid * 256;
        double ms;
        ms = time_ms([&] { k_imad_wide_indep<<<grid, 256>>>((uint64_t*)buf, 12345, 6789); }, 5);
        printf("{\"bench\": \"imad_wide_indep\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Tops\": %.3f}\n", bps, ms,
               threads * ITERS * 8 / ms / 1e9);
        ms = time_ms([&] { k_imad_lo_indep<<<grid, 256>>>((uint32_t*)buf, 12345, 6789); }, 5);
        printf("{\"bench\": \"imad_lo_indep\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Tops\": %.3f}\n", bps, ms,
               threads * ITERS * 8 / ms / 1e9);
        ms = time_ms([&] { k_imad_wide_carry<<<grid, 256>>>((uint32_t*)buf, 12345, 6789); }, 5);
        printf("{\"bench\": \"imad_wide_carry\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Tops\": %.3f}\n", bps, ms,
               threads * ITERS * 8 / ms / 1e9);  // 2 rows x 4 wide multiply-adds
        ms = time_ms([&] { k_fe_mul<<<grid, 256>>>((uint32_t*)buf, 99); }, 5);
        printf("{\"bench\": \"fe_mul\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.2f, \"imad_Tops\": %.3f}\n", bps, ms,
               threads * (ITERS / 2) / ms / 1e6, threads * (ITERS / 2) * 72 / ms / 1e9);
        ms = time_ms([&] { k_fe_sq<<<grid, 256>>>((uint32_t*)buf, 99); }, 5);
        printf("{\"bench\": \"fe_sq\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.2f}\n", bps, ms,
               threads * (ITERS / 2) / ms / 1e6);
    }
    for (int bps = 1; bps <= 4; bps *= 2) {
        int grid = sms * bps;
        double threads = (double)grid * 128;
        double ms = time_ms([&] { k_ge_madd<<<grid, 128>>>((uint32_t*)buf, 99); }, 5);
        printf("{\"bench\": \"ge_madd\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.3f, \"imad_Tops\": %.3f}\n", bps, ms,
               threads * (ITERS / 16) / ms / 1e6, threads * (ITERS / 16) * 504 / ms / 1e9);
        ms = time_ms([&] { k_ge_dbl<<<grid, 128>>>((uint32_t*)buf, 99); }, 5);
        printf("{\"bench\": \"ge_dbl\", \"blocks_per_sm\": %d, \"ms\": %.4f, \"Gops\": %.3f}\n", bps, ms,
               threads * (ITERS / 16) / ms / 1e6);
    }
    for (int quad = 0; quad < 2; quad++) {
        double ms = time_ms([&] { k_lat_add<<<sms, 32>>>((uint32_t*)buf, 99, quad); }, 5);
        printf("{\"bench\": \"latency_ge_add%s\", \"us_per_op\": %.3f}\n", quad ? "_quad" : "", ms * 1e3 / 256);
        ms = time_ms([&] { k_lat_dbl<<<sms, 32>>>((uint32_t*)buf, 99, quad); }, 5);
        printf("{\"bench\": \"latency_ge_dbl%s\", \"us_per_op\": %.3f}\n", quad ? "_quad" : "", ms * 1e3 / 256);
    }
    {
        lat8_row<0>("latency_ge8_dbl", sms, buf);
        lat8_row<1>("latency_ge8_add", sms, buf);
        lat8_row<2>("latency_ge8_add_cached", sms, buf);
        lat8_row<3>("latency_fe8_invert", sms, buf);
        lat8_row<4>("latency_fe8_mul", sms, buf);
        lat8_row<5>("latency_fe8_sub", sms, buf);
        lat8_row<6>("latency_fe8_sq", sms, buf);
    }
    {
        const char* inames[6] = {"latency_fe_invert_divsteps", "latency_fe_invert_fermat", "latency_fe_invert_divsteps_divergent",
                                 "latency_sc_invert_divsteps", "latency_sc_invert_fermat", "latency_sc_invert_divsteps_divergent"};
        for (int what = 0; what < 6; what++) {
            double ms = time_ms([&] { k_lat_inv<<<sms, 32>>>((uint32_t*)buf, 99, what); }, 5);
            printf("{\"bench\": \"%s\", \"us\": %.3f}\n", inames[what], ms * 1e3);
        }
    }
    {  // PCIe: contiguous upload of 2^20 points (128 B each) against a pitched upload of X,Y,Z only (96 of 128 B)
        const size_t n = (size_t)1 << 20;
        void *h = nullptr, *d = nullptr;
        if (cudaMallocHost(&h, n * 128) == cudaSuccess && cudaMalloc(&d, n * 128) == cudaSuccess) {
            memset(h, 1, n * 128);
            double ms = time_ms([&] { cudaMemcpyAsync(d, h, n * 128, cudaMemcpyHostToDevice, 0); }, 5);
            printf("{\"bench\": \"h2d_contiguous_128B\", \"ms\": %.4f, \"GBps\": %.2f}\n", ms, n * 128 / ms / 1e6);
            ms = time_ms([&] { cudaMemcpy2DAsync(d, 96, h, 128, 96, n, cudaMemcpyHostToDevice, 0); }, 5);
            printf("{\"bench\": \"h2d_pitched_96_of_128B\", \"ms\": %.4f, \"useful_GBps\": %.2f}\n", ms, n * 96 / ms / 1e6);
            ms = time_ms([&] { cudaMemcpy2DAsync(d, 128, h, 128, 96, n, cudaMemcpyHostToDevice, 0); }, 5);
            printf("{\"bench\": \"h2d_pitched_96_of_128B_dst128\", \"ms\": %.4f, \"useful_GBps\": %.2f}\n", ms, n * 96 / ms / 1e6);
        }
    }
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    return 0;
}
