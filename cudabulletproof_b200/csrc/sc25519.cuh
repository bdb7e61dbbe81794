// sc25519.cuh — integers mod l = 2^252 + 27742317777372353535851937790883648493 (the Ed25519 group
// order), 8 x 32-bit limbs, for sm_100a.
//
// The reference does every scalar computation with fe25519_* (mod p) — defect D11: inner products
// (bulletproof_vectors.cu:101-114), the a/b fold (:488-500), delta(y,z) (bulletproof_range_proof.cu:315-374).
// These are the mod-l replacements, kept in the same 32-byte little-endian container.
#pragma once
#include <stdint.h>
#include "fe25519.cuh"

namespace cbp {

struct sc {
    uint32_t v[8];
};

static __device__ __constant__ const uint32_t kScL[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu,
                                                  0x00000000u, 0x00000000u, 0x00000000u, 0x10000000u};
// mu = floor(2^512 / l), 9 words (Barrett, HAC 14.42 with b = 2^32, k = 8)
static __device__ __constant__ const uint32_t kScMu[9] = {0x0a2c131bu, 0xed9ce5a3u, 0x086329a7u, 0x2106215du, 0xffffffebu,
                                                   0xffffffffu, 0xffffffffu, 0xffffffffu, 0x0000000fu};
// 2^512 mod l
static __device__ __constant__ const uint32_t kScR512[8] = {0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u,
                                                     0x17f5be65u, 0xceec73d2u, 0x7c309a3du, 0x0399411bu};

__device__ __forceinline__ void sc_set0(sc& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
__device__ __forceinline__ void sc_set1(sc& r) {
    sc_set0(r);
    r.v[0] = 1;
}
__device__ __forceinline__ void sc_load(sc& r, const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
}
__device__ __forceinline__ void sc_store(void* p, const sc& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// r >= l ?  (9-word r, r[8] may be nonzero)
__device__ __forceinline__ bool sc_geq_l9(const uint32_t (&r)[9]) {
    if (r[8]) return true;
#pragma unroll
    for (int i = 7; i >= 0; i--) {
        if (r[i] != kScL[i]) return r[i] > kScL[i];
    }
    return true;
}
__device__ __forceinline__ void sc_sub_l9(uint32_t (&r)[9]) {
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)r[i] - kScL[i] - borrow;
        r[i] = (uint32_t)d;
        borrow = (d >> 63) & 1;
    }
    r[8] -= (uint32_t)borrow;
}

// r = x mod l for a 512-bit x (16 words).  Barrett: q3 = floor(floor(x / b^7) * mu / b^9),
// r = (x - q3 l) mod b^9, then at most two subtractions of l.
__device__ __forceinline__ void sc_reduce512(sc& out, const uint32_t (&x)[16]) {
    // q2 = q1 * mu, only words >= 9 are needed (q3); compute full columns from 7 up for exact carries
    uint32_t q3[9];
    {
        uint64_t acc_lo = 0, acc_hi = 0;  // 128-bit column accumulator
#pragma unroll
        for (int col = 0; col < 18; col++) {
#pragma unroll
            for (int i = 0; i < 9; i++) {
                int j = col - i;
                if (j < 0 || j > 8) continue;
                uint64_t pr = (uint64_t)x[7 + i] * kScMu[j];
                acc_lo += pr;
                acc_hi += (acc_lo < pr);
            }
            if (col >= 9) q3[col - 9] = (uint32_t)acc_lo;
            acc_lo = (acc_lo >> 32) | (acc_hi << 32);
            acc_hi >>= 32;
        }
    }
    // r2 = (q3 * l) mod b^9
    uint32_t r2[9];
    {
        uint64_t acc_lo = 0, acc_hi = 0;
#pragma unroll
        for (int col = 0; col < 9; col++) {
#pragma unroll
            for (int i = 0; i < 9; i++) {
                int j = col - i;
                if (j < 0 || j > 7) continue;
                if (j >= 4 && j <= 6) continue;  // zero words of l
                uint64_t pr = (uint64_t)q3[i] * kScL[j];
                acc_lo += pr;
                acc_hi += (acc_lo < pr);
            }
            r2[col] = (uint32_t)acc_lo;
            acc_lo = (acc_lo >> 32) | (acc_hi << 32);
            acc_hi >>= 32;
        }
    }
    uint32_t r[9];
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 9; i++) {
        uint64_t d = (uint64_t)x[i] - r2[i] - borrow;
        r[i] = (uint32_t)d;
        borrow = (d >> 63) & 1;
    }
    if (sc_geq_l9(r)) sc_sub_l9(r);
    if (sc_geq_l9(r)) sc_sub_l9(r);
#pragma unroll
    for (int i = 0; i < 8; i++) out.v[i] = r[i];
}

// any 256-bit value -> [0, l)
__device__ __forceinline__ void sc_reduce(sc& r, const sc& a) {
    uint32_t x[16];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = a.v[i];
        x[i + 8] = 0;
    }
    sc_reduce512(r, x);
}
// inputs < l
__device__ __forceinline__ void sc_add(sc& r, const sc& a, const sc& b) {
    uint32_t t[9];
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a.v[i] + b.v[i];
        t[i] = (uint32_t)c;
        c >>= 32;
    }
    t[8] = (uint32_t)c;
    if (sc_geq_l9(t)) sc_sub_l9(t);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
}
__device__ __forceinline__ bool sc_iszero(const sc& a) {
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= a.v[i];
    return o == 0;
}
__device__ __forceinline__ void sc_neg(sc& r, const sc& a) {  // a < l
    bool z = sc_iszero(a);
    uint64_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)kScL[i] - a.v[i] - borrow;
        r.v[i] = z ? 0u : (uint32_t)d;
        borrow = (d >> 63) & 1;
    }
}
__device__ __forceinline__ void sc_sub(sc& r, const sc& a, const sc& b) {
    sc nb;
    sc_neg(nb, b);
    sc_add(r, a, nb);
}
__device__ __forceinline__ void sc_mul(sc& r, const sc& a, const sc& b) {
    uint32_t w[16];
    fe fa, fb;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        fa.v[i] = a.v[i];
        fb.v[i] = b.v[i];
    }
    mul_wide(w, fa, fb);
    sc_reduce512(r, w);
}
__device__ __forceinline__ void sc_sq(sc& r, const sc& a) { sc_mul(r, a, a); }
// a^(l-2); inv(0) = 0
__device__ __noinline__ static void sc_invert(sc& r, const sc& a) {
    const uint32_t e[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u};
    sc acc;
    sc_set1(acc);
    for (int bit = 252; bit >= 0; bit--) {
        sc_sq(acc, acc);
        if ((e[bit >> 5] >> (bit & 31)) & 1) sc_mul(acc, acc, a);
    }
    r = acc;
}

}  // namespace cbp
