// bench_api.cu — cuda_benchmark_* entry points.  The reference declares them
// (cuda_bulletproof.h:81-84) but never defines them; here each one times `iterations` runs of the
// corresponding device-resident operation on synthetic inputs with CUDA events and prints one line.
#include <stdio.h>
#include "../../include/cuda_bulletproof.h"
#include "common.h"

namespace {
struct Timer {
    cudaEvent_t e0, e1;
    Timer() {
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
    }
    ~Timer() {
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    }
    void start() { cudaEventRecord(e0, 0); }
    float stop_ms() {
        cudaEventRecord(e1, 0);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        return ms;
    }
};
}  // namespace

extern "C" {

void cuda_benchmark_multi_scalar_mul(int iterations, size_t n) {
    if (iterations <= 0 || n == 0) return;
    size_t ws = 0;
    if (bpk_msm_workspace_bytes(n, 0, &ws) != BPK_OK) return;
    uint8_t *d_s = nullptr, *d_p = nullptr, *d_ws = nullptr, *d_r = nullptr;
    if (cudaMalloc(&d_s, n * 32) || cudaMalloc(&d_p, n * 128) || cudaMalloc(&d_ws, ws) || cudaMalloc(&d_r, 128)) goto done;
    bpk_synth_points_device(d_p, nullptr, n, 1, 0);
    bpk_synth_scalars_device(d_s, n, 2, 252, 0);
    bpk_msm_device(d_s, d_p, n, d_r, d_ws, ws, 0, 1, 0);
    {
        Timer t;
        t.start();
        for (int i = 0; i < iterations; i++) bpk_msm_device(d_s, d_p, n, d_r, d_ws, ws, 0, 1, 0);
        float ms = t.stop_ms() / iterations;
        printf("multi_scalar_mul n=%zu: %.4f ms/iter, %.3f Mpoints/s\n", n, ms, n / ms / 1e3);
    }
done:
    cudaFree(d_s);
    cudaFree(d_p);
    cudaFree(d_ws);
    cudaFree(d_r);
}

void cuda_benchmark_inner_product(int iterations, size_t n) {
    if (iterations <= 0 || n == 0) return;
    size_t ws = 0;
    bpk_sc_inner_product_workspace_bytes(n, &ws);
    uint8_t *d_a = nullptr, *d_b = nullptr, *d_ws = nullptr, *d_r = nullptr;
    if (cudaMalloc(&d_a, n * 32) || cudaMalloc(&d_b, n * 32) || cudaMalloc(&d_ws, ws) || cudaMalloc(&d_r, 32)) goto done;
    bpk_synth_scalars_device(d_a, n, 3, 252, 0);
    bpk_synth_scalars_device(d_b, n, 4, 252, 0);
    bpk_sc_inner_product_device(d_r, d_a, d_b, n, d_ws, ws, 0);
    {
        Timer t;
        t.start();
        for (int i = 0; i < iterations; i++) bpk_sc_inner_product_device(d_r, d_a, d_b, n, d_ws, ws, 0);
        float ms = t.stop_ms() / iterations;
        printf("inner_product n=%zu: %.4f ms/iter, %.2f GB/s\n", n, ms, n * 64.0 / ms / 1e6);
    }
done:
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_ws);
    cudaFree(d_r);
}

void cuda_benchmark_field_operations(int iterations, size_t count) {
    if (iterations <= 0 || count == 0) return;
    uint8_t *d_a = nullptr, *d_b = nullptr, *d_o = nullptr;
    if (cudaMalloc(&d_a, count * 32) || cudaMalloc(&d_b, count * 32) || cudaMalloc(&d_o, count * 32)) goto done;
    bpk_synth_scalars_device(d_a, count, 5, 255, 0);
    bpk_synth_scalars_device(d_b, count, 6, 255, 0);
    {
        const char* names[4] = {"add", "sub", "mul", "square"};
        for (int op = 0; op < 4; op++) {
            bpk_fe_batch_device(op, d_o, d_a, d_b, count, 0);
            Timer t;
            t.start();
            for (int i = 0; i < iterations; i++) bpk_fe_batch_device(op, d_o, d_a, d_b, count, 0);
            float ms = t.stop_ms() / iterations;
            printf("field %s count=%zu: %.4f ms/iter, %.2f GB/s\n", names[op], count, ms,
                   count * (op == 3 ? 64.0 : 96.0) / ms / 1e6);
        }
        bpk_fe_batch_invert_device(d_o, d_a, count, nullptr, 0, 0);
        Timer t;
        t.start();
        for (int i = 0; i < iterations; i++) bpk_fe_batch_invert_device(d_o, d_a, count, nullptr, 0, 0);
        float ms = t.stop_ms() / iterations;
        printf("field invert count=%zu: %.4f ms/iter, %.2f Melem/s\n", count, ms, count / ms / 1e3);
    }
done:
    cudaFree(d_a);
    cudaFree(d_b);
    cudaFree(d_o);
}

void cuda_benchmark_range_proof(int iterations, size_t bits) {
    if (iterations <= 0 || bits == 0 || bits > 64 || (bits & (bits - 1))) return;
    const size_t num = 1024;
    size_t gws = 0, vws = 0, rec = bpk_proof_record_bytes(bits);
    if (bpk_gens_workspace_bytes(bits, &gws) != BPK_OK) return;
    bpk_range_verify_workspace_bytes(bits, num, &vws);
    uint8_t *d_gens = nullptr, *d_pts = nullptr, *d_proofs = nullptr, *d_vws = nullptr, *d_acc = nullptr, *d_gam = nullptr;
    uint64_t *d_vals = nullptr, *d_seeds = nullptr;
    size_t nb = 2 * bits + 2;
    if (cudaMalloc(&d_gens, gws) || cudaMalloc(&d_pts, nb * 128) || cudaMalloc(&d_proofs, num * rec) ||
        cudaMalloc(&d_vws, vws) || cudaMalloc(&d_acc, num) || cudaMalloc(&d_gam, num * 32) ||
        cudaMalloc(&d_vals, num * 8) || cudaMalloc(&d_seeds, num * 8))
        goto done;
    bpk_synth_points_device(d_pts, nullptr, nb, 7, 0);
    bpk_gens_init_device(d_gens, gws, d_pts, d_pts + bits * 128, d_pts + 2 * bits * 128, d_pts + 2 * bits * 128 + 128, bits, 0);
    bpk_synth_scalars_device(d_gam, num, 8, 252, 0);
    bpk_synth_scalars_device(d_vals, num / 4, 9, (int)bits, 0);  // 4 x u64 per 32 B: masks only the first word group
    cudaMemset(d_vals, 0, num * 8);
    bpk_synth_scalars_device(d_seeds, num / 4, 10, 256, 0);
    bpk_range_prove_batch_device(d_gens, d_vals, d_gam, d_seeds, bits, num, d_proofs, nullptr, 0, 0);
    bpk_range_verify_batch_device(d_gens, d_proofs, nullptr, bits, num, d_acc, d_vws, vws, 0);
    {
        Timer t;
        t.start();
        for (int i = 0; i < iterations; i++)
            bpk_range_verify_batch_device(d_gens, d_proofs, nullptr, bits, num, d_acc, d_vws, vws, 0);
        float ms = t.stop_ms() / iterations;
        printf("range_proof verify bits=%zu batch=%zu: %.4f ms/iter, %.1f verifies/s\n", bits, num, ms, num / ms * 1e3);
    }
done:
    cudaFree(d_gens);
    cudaFree(d_pts);
    cudaFree(d_proofs);
    cudaFree(d_vws);
    cudaFree(d_acc);
    cudaFree(d_gam);
    cudaFree(d_vals);
    cudaFree(d_seeds);
}

}  // extern "C"
