// Probe: do the carry-chain field operations compute correctly when every input is a compile-time /
// warp-uniform value?  (ptxas then places the whole computation on the uniform datapath: UIMAD.WIDE.U32.X
// with UP carry predicates.)  Prints the results next to the expected values.
#include <stdio.h>
#include "../cudabulletproof_b200/csrc/fe25519.cuh"
using namespace cbp;
__global__ void k_const(uint32_t* out) {
    fe one, two, r;
    fe_set1(one);
    fe_sq(r, one);            // 1
    for (int i = 0; i < 8; i++) out[i] = r.v[i];
    fe_mul(r, one, one);      // 1
    for (int i = 0; i < 8; i++) out[8 + i] = r.v[i];
    fe_add(two, one, one);    // 2
    for (int i = 0; i < 8; i++) out[16 + i] = two.v[i];
    fe_sq(r, two);            // 4
    for (int i = 0; i < 8; i++) out[24 + i] = r.v[i];
    fe_sub(r, one, two);      // -1 = 2^256 - 1 - 38 -> weak repr
    for (int i = 0; i < 8; i++) out[32 + i] = r.v[i];
}
__global__ void k_param(uint32_t* out, fe a) {  // uniform but not constant: kernel parameter
    fe r;
    fe_sq(r, a);
    for (int i = 0; i < 8; i++) out[i] = r.v[i];
    fe_mul(r, a, a);
    for (int i = 0; i < 8; i++) out[8 + i] = r.v[i];
}
__global__ void k_vec(uint32_t* out, const uint32_t* in) {  // per-thread values from memory
    fe a, r;
    for (int i = 0; i < 8; i++) a.v[i] = in[i] + threadIdx.x * 0;
    fe_sq(r, a);
    for (int i = 0; i < 8; i++) out[i] = r.v[i];
    fe_mul(r, a, a);
    for (int i = 0; i < 8; i++) out[8 + i] = r.v[i];
}
static void show(const char* name, const uint32_t* w) {
    printf("%-22s", name);
    for (int i = 7; i >= 0; i--) printf("%08x", w[i]);
    printf("\n");
}
int main() {
    uint32_t *d, *din, h[64];
    cudaMalloc(&d, 256);
    cudaMalloc(&din, 32);
    cudaMemset(d, 0, 256);
    k_const<<<1, 32>>>(d);
    cudaMemcpy(h, d, 160, cudaMemcpyDeviceToHost);
    show("const sq(1)", h); show("const mul(1,1)", h + 8); show("const 1+1", h + 16); show("const sq(2)", h + 24);
    show("const 1-2", h + 32);
    fe a;
    for (int i = 0; i < 8; i++) a.v[i] = 0;
    a.v[0] = 1;
    k_param<<<1, 32>>>(d, a);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    show("param sq(1)", h); show("param mul(1,1)", h + 8);
    for (int i = 0; i < 8; i++) a.v[i] = 0x9abcdef1u * (i + 3);
    k_param<<<1, 32>>>(d, a);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    show("param sq(x)", h); show("param mul(x,x)", h + 8);
    cudaMemcpy(din, a.v, 32, cudaMemcpyHostToDevice);
    k_vec<<<1, 32>>>(d, din);
    cudaMemcpy(h, d, 64, cudaMemcpyDeviceToHost);
    show("vector sq(x)", h); show("vector mul(x,x)", h + 8);
    cudaError_t e = cudaDeviceSynchronize();
    printf("status %s\n", cudaGetErrorString(e));
    return 0;
}
