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

// r = x mod l for a 512-bit x (16 words), using l = 2^252 + delta with delta < 2^125 (the structure
// ref10's sc_reduce exploits): 2^252 = -delta (mod l), so with x = lo + 2^252 hi
//     x = lo - hi delta,  and the 385-bit product hi*delta is folded the same way twice more:
//     x = lo - lo1 + lo2 - P3 (mod l),   P1 = hi delta, P2 = (P1 >> 252) delta, P3 = (P2 >> 252) delta.
// 36 + 20 + 4 = 60 word products instead of Barrett's 126, no quotient estimate.
// (A first Barrett version in plain C compiled to ~640 instructions per reduction.)
__device__ __forceinline__ void sc_mul_delta(uint32_t* out, int nout, const uint32_t* h, int nh) {
    // out[0..nout) = h[0..nh) * delta (4 words); row-wise schoolbook, every partial sum fits 64 bits
    const uint32_t d[4] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu};
#pragma unroll
    for (int i = 0; i < nout; i++) out[i] = 0;
#pragma unroll
    for (int i = 0; i < nh; i++) {
        uint32_t carry = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            if (i + j < nout) {
                uint64_t m = (uint64_t)h[i] * d[j] + out[i + j] + carry;
                out[i + j] = (uint32_t)m;
                carry = (uint32_t)(m >> 32);
            }
        }
        if (i + 4 < nout) out[i + 4] = carry;
    }
}
// words of (v >> 252) for a little-endian word array v of nv words; nres result words
__device__ __forceinline__ void sc_shr252(uint32_t* res, int nres, const uint32_t* v, int nv) {
#pragma unroll
    for (int i = 0; i < nres; i++) {
        uint32_t lo = 7 + i < nv ? v[7 + i] : 0, hi = 8 + i < nv ? v[8 + i] : 0;
        res[i] = (lo >> 28) | (hi << 4);
    }
}
__device__ __forceinline__ void sc_reduce512(sc& out, const uint32_t (&x)[16]) {
    uint32_t hi[9], P1[13], hi1[5], P2[9], hi2[1], P3[5];
    sc_shr252(hi, 9, x, 16);       // 260 bits
    sc_mul_delta(P1, 13, hi, 9);   // 385 bits
    sc_shr252(hi1, 5, P1, 13);     // 133 bits
    sc_mul_delta(P2, 9, hi1, 5);   // 258 bits
    sc_shr252(hi2, 1, P2, 9);      // 6 bits
    sc_mul_delta(P3, 5, hi2, 1);   // 131 bits
    // v = lo + lo2 + 2l - lo1 - P3  in (0, 4l): signed 64-bit column arithmetic
    const uint32_t two_l[8] = {0xb9eba7dau, 0xb024c634u, 0x45ef39acu, 0x29bdf3bdu, 0u, 0u, 0u, 0x20000000u};
    uint32_t r[9];
    int64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t lo = x[i], lo1 = P1[i], lo2 = P2[i];
        if (i == 7) {
            lo &= 0x0fffffffu;
            lo1 &= 0x0fffffffu;
            lo2 &= 0x0fffffffu;
        }
        c += (int64_t)lo + (int64_t)lo2 + (int64_t)two_l[i] - (int64_t)lo1 - (int64_t)(i < 5 ? P3[i] : 0u);
        r[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: floor division
    }
    r[8] = (uint32_t)c;
    if (sc_geq_l9(r)) sc_sub_l9(r);
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
// dedicated squaring: 36 instead of 64 wide multiplies before the reduction (sq_wide of the field layer);
// the inversion chain is 253 of these
__device__ __forceinline__ void sc_sq(sc& r, const sc& a) {
    uint32_t w[16];
    fe fa;
#pragma unroll
    for (int i = 0; i < 8; i++) fa.v[i] = a.v[i];
    sq_wide(w, fa);
    sc_reduce512(r, w);
}
// a^(l-2); inv(0) = 0: the Fermat chain, 253 squarings + ~130 multiplications with a Barrett reduction each (~200 us
// for a lone thread).  Kept as the cross-check of sc_invert below.
__device__ __noinline__ static void sc_invert_fermat(sc& r, const sc& a) {
    const uint32_t e[8] = {0x5cf5d3ebu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u};
    sc acc;
    sc_set1(acc);
#pragma unroll 1
    for (int bit = 252; bit >= 0; bit--) {
        sc_sq(acc, acc);
        if ((e[bit >> 5] >> (bit & 31)) & 1) sc_mul(acc, acc, a);
    }
    r = acc;
}
// 1 / a mod l in [0, l); inv(0) = 0.  Divsteps (modinv.cuh): every scalar inverted on this path is a public
// Fiat-Shamir challenge (y, the round challenges u_j), so variable time is fine.
__device__ __noinline__ static void sc_invert(sc& r, const sc& a) {
    const uint32_t lw[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u};
    const ModInfo mi = modinfo_from_words(lw);
    uint32_t out[8];
    modinv_words(out, a.v, mi);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = out[i];
}

}  // namespace cbp
