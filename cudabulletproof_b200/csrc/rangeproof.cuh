// rangeproof.cuh — shared device definitions for range-proof verification / proving (rangeproof.cu,
// prover.cu): generator-table layout, proof-record layout, fixed-base accumulation.
#pragma once
#include <stdint.h>
#include "ge25519.cuh"
#include "sc25519.cuh"

namespace cbp {

constexpr int kMaxK = 6;          // log2 of the largest supported range width (n <= 64)
constexpr int kMaxN = 64;
constexpr int kFixWin = 32;       // fixed-base tables: 32 signed 8-bit windows
constexpr int kFixEntries = 128;  // multiples 1..128 per window
constexpr uint64_t kGensMagic = 0x62706b47454e5331ull;  // "bpkGENS1"

// header of the generator workspace built by bpk_gens_init_device
struct GensHeader {
    uint64_t magic;
    uint32_t n;
    uint32_t nbases;     // 2n + 2: G[0..n), H[0..n), g, h
    uint64_t table_off;  // niels table: ((base*32 + win)*128 + (d-1)) * 96 bytes
    uint64_t scratch_off;
    uint64_t bases_off;  // nbases normalised ge25519 (128 B each)
};

// proof record offsets (bytes), include/bpk.h
constexpr int kRecV = 0, kRecA = 128, kRecS = 256, kRecT1 = 384, kRecT2 = 512, kRecTaux = 640, kRecMu = 672,
              kRecT = 704, kRecIpA = 736, kRecIpB = 768, kRecIpC = 800, kRecIpX = 832, kRecL = 864;
__host__ __device__ inline size_t proof_record_bytes(int k) { return (size_t)kRecL + 256 * (size_t)k; }

struct ge_cached {  // extended precomputed: (Y+X, Y-X, 2Z, 2dT)
    fe YplusX, YminusX, Z2, T2d;
};
__device__ __forceinline__ void ge_to_cached(ge_cached& r, const ge_p3& p) {
    fe_add(r.YplusX, p.Y, p.X);
    fe_sub(r.YminusX, p.Y, p.X);
    fe_dbl(r.Z2, p.Z);
    fe_mul(r.T2d, p.T, fe_const_2d());
}
// r = p + q (q cached) or p - q.  8M.
__device__ __forceinline__ void ge_add_cached(ge_p3& r, const ge_p3& p, const ge_cached& q, bool neg) {
    fe A, B, C, D, E, F, G, H, t;
    fe qa = q.YminusX, qb = q.YplusX;
    fe_cswap(qa, qb, neg);
    fe_sub(t, p.Y, p.X);
    fe_mul(A, t, qa);
    fe_add(t, p.Y, p.X);
    fe_mul(B, t, qb);
    fe_mul(C, p.T, q.T2d);
    fe_mul(D, p.Z, q.Z2);
    fe_sub(E, B, A);
    fe_add(H, B, A);
    fe Fm, Fp;
    fe_sub(Fm, D, C);
    fe_add(Fp, D, C);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        F.v[i] = neg ? Fp.v[i] : Fm.v[i];
        G.v[i] = neg ? Fm.v[i] : Fp.v[i];
    }
    fe_mul(r.X, E, F);
    fe_mul(r.Y, G, H);
    fe_mul(r.Z, F, G);
    fe_mul(r.T, E, H);
}

// signed base-2^w recoding of a scalar < 2^253 into ndig digits in [-(2^(w-1)-1), 2^(w-1)]
template <int WBITS>
__device__ __forceinline__ void sc_recode_signed(int8_t* out, const sc& k, int ndig) {
    uint32_t carry = 0;
    constexpr uint32_t half = 1u << (WBITS - 1);
    for (int j = 0; j < ndig; j++) {
        int bit = j * WBITS;
        uint32_t d = (k.v[bit >> 5] >> (bit & 31)) & ((1u << WBITS) - 1u);
        d += carry;
        carry = 0;
        int v = (int)d;
        if (d > half) {
            v = (int)d - (1 << WBITS);
            carry = 1;
        }
        out[j] = (int8_t)v;  // 128 is stored as -128 when WBITS == 8; see fixed_digit()
    }
}
// decode an 8-bit window digit written by sc_recode_signed<8>: +128 wraps to int8 -128 but real
// negative digits only reach -127, so -128 means +128
__device__ __forceinline__ void fixed_digit(int8_t raw, uint32_t& mag, bool& neg) {
    int v = raw;
    if (v == -128) {
        mag = 128;
        neg = false;
    } else if (v < 0) {
        mag = (uint32_t)(-v);
        neg = true;
    } else {
        mag = (uint32_t)v;
        neg = false;
    }
}

// acc += digit * 2^(8 win) * Base[base]  from the precomputed table
__device__ __forceinline__ void fixed_base_madd(ge_p3& acc, const uint8_t* __restrict__ table, uint32_t base, int win,
                                                int8_t raw) {
    uint32_t mag;
    bool neg;
    fixed_digit(raw, mag, neg);
    if (mag == 0) return;
    ge_niels q;
    ge_niels_load(q, table + (((size_t)base * kFixWin + win) * kFixEntries + (mag - 1)) * 96);
    ge_madd(acc, acc, q, neg);
}

// CTA-wide sum of one point per thread through shared memory; result valid in thread 0.
// red must hold blockDim.x points; blockDim.x must be a power of two.
__device__ __forceinline__ void cta_point_sum(ge_p3& acc, ge_p3* red) {
    int t = threadIdx.x;
    red[t] = acc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (t < o) {
            ge_p3 a = red[t], b = red[t + o];
            ge_add(a, a, b);
            red[t] = a;
        }
        __syncthreads();
    }
    acc = red[0];
    __syncthreads();
}

}  // namespace cbp
