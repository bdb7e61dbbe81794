// ge25519.cuh — twisted-Edwards (a = -1) group arithmetic for sm_100a on top of fe25519.cuh.
//
// Replaces device_curve25519_ops.cuh:188-290 (device_ge25519_add / _normalize / _scalarmult).
// Same addition law as the reference's ge25519_add (curve25519_ops.cu:326-378, add-2008-hwcd-3)
// with the constant fixed to 2d (defect D5), plus the cheaper special cases the reference never
// used: mixed addition with a precomputed affine point (7M) and dedicated doubling (4S+4M).
#pragma once
#include "fe25519.cuh"

namespace cbp {

struct ge_p3 {  // extended (X:Y:Z:T), T = XY/Z — the reference's ge25519
    fe X, Y, Z, T;
};
struct ge_niels {  // affine, precomputed: (y+x, y-x, 2d*x*y); 96 bytes in memory
    fe yplusx, yminusx, xy2d;
};

__device__ __forceinline__ void ge_p3_0(ge_p3& r) {
    fe_set0(r.X);
    fe_set1(r.Y);
    fe_set1(r.Z);
    fe_set0(r.T);
}

// r = p + q, q affine precomputed; sign: add -q instead (swap y+x / y-x, negate 2dxy).  7M.
__device__ __forceinline__ void ge_madd(ge_p3& r, const ge_p3& p, const ge_niels& q, bool neg) {
    fe A, B, C, D, E, F, G, H, t;
    fe qa = q.yminusx, qb = q.yplusx;
    fe_cswap(qa, qb, neg);
    fe_sub(t, p.Y, p.X);
    fe_mul(A, t, qa);
    fe_add(t, p.Y, p.X);
    fe_mul(B, t, qb);
    fe_mul(C, p.T, q.xy2d);
    fe_dbl(D, p.Z);
    fe_sub(E, B, A);
    fe_add(H, B, A);
    fe Fp, Fm;
    fe_sub(Fm, D, C);
    fe_add(Fp, D, C);
    // +q: F = D - C, G = D + C ;  -q: C -> -C
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

// r = p + q, both extended (unified: also valid for p == q).  9M, the reference's formula with k = 2d.
__device__ __forceinline__ void ge_add(ge_p3& r, const ge_p3& p, const ge_p3& q) {
    fe A, B, C, D, E, F, G, H, t, u;
    fe_sub(t, p.Y, p.X);
    fe_sub(u, q.Y, q.X);
    fe_mul(A, t, u);
    fe_add(t, p.Y, p.X);
    fe_add(u, q.Y, q.X);
    fe_mul(B, t, u);
    fe_mul(C, p.T, q.T);
    fe_mul(C, C, fe_const_2d());
    fe_mul(D, p.Z, q.Z);
    fe_dbl(D, D);
    fe_sub(E, B, A);
    fe_sub(F, D, C);
    fe_add(G, D, C);
    fe_add(H, B, A);
    fe_mul(r.X, E, F);
    fe_mul(r.Y, G, H);
    fe_mul(r.Z, F, G);
    fe_mul(r.T, E, H);
}

// r = 2p (dbl-2008-hwcd, a = -1).  4S + 4M.
__device__ __forceinline__ void ge_dbl(ge_p3& r, const ge_p3& p) {
    fe XX, YY, ZZ2, S, E, F, G, H;
    fe_sq(XX, p.X);
    fe_sq(YY, p.Y);
    fe_sq(ZZ2, p.Z);
    fe_dbl(ZZ2, ZZ2);
    fe_add(S, p.X, p.Y);
    fe_sq(S, S);
    fe_add(H, YY, XX);
    fe_sub(G, YY, XX);
    fe_sub(E, S, H);
    fe_sub(F, ZZ2, G);
    fe_mul(r.X, E, F);
    fe_mul(r.Y, G, H);
    fe_mul(r.Z, F, G);
    fe_mul(r.T, E, H);
}

// r = 2p without the T coordinate (4S + 3M): for runs of doublings, where only the last one feeds an addition
__device__ __forceinline__ void ge_dbl_xyz(ge_p3& r, const ge_p3& p) {
    fe XX, YY, ZZ2, S, E, F, G, H;
    fe_sq(XX, p.X);
    fe_sq(YY, p.Y);
    fe_sq(ZZ2, p.Z);
    fe_dbl(ZZ2, ZZ2);
    fe_add(S, p.X, p.Y);
    fe_sq(S, S);
    fe_add(H, YY, XX);
    fe_sub(G, YY, XX);
    fe_sub(E, S, H);
    fe_sub(F, ZZ2, G);
    fe_mul(r.X, E, F);
    fe_mul(r.Y, G, H);
    fe_mul(r.Z, F, G);
}

// ---- quad-cooperative point operations --------------------------------------------------------------
// A dependent chain of point operations run by ONE lane is bound by the latency of its 8-9 sequential
// field multiplications (~600 cycles each in a lone warp).  Here four adjacent lanes (a "quad",
// role = lane & 3) hold identical copies of the operands and each computes one of the four independent
// products of a stage; the results are exchanged with shuffles.  Depth per addition: 3 multiplications
// instead of 9; per doubling: 2 instead of 8.  Used by the latency-bound tails (bucket-reduction levels,
// Horner chain); every lane of the quad ends with the full result.
__device__ __forceinline__ void fe_quad_exchange(fe& r0, fe& r1, fe& r2, fe& r3, const fe& mine) {
    const int base = (threadIdx.x & 31) & ~3;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r0.v[i] = __shfl_sync(0xffffffffu, mine.v[i], base + 0);
        r1.v[i] = __shfl_sync(0xffffffffu, mine.v[i], base + 1);
        r2.v[i] = __shfl_sync(0xffffffffu, mine.v[i], base + 2);
        r3.v[i] = __shfl_sync(0xffffffffu, mine.v[i], base + 3);
    }
}
__device__ __forceinline__ void fe_select4(fe& r, int role, const fe& a0, const fe& a1, const fe& a2, const fe& a3) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t lo = (role & 1) ? a1.v[i] : a0.v[i];
        uint32_t hi = (role & 1) ? a3.v[i] : a2.v[i];
        r.v[i] = (role & 2) ? hi : lo;
    }
}
// r = p + q (unified); all four lanes of the quad must call this with identical p, q
__device__ __forceinline__ void ge_add_quad(ge_p3& r, const ge_p3& p, const ge_p3& q) {
    const int role = threadIdx.x & 3;
    fe s1, a1, s2, a2, x, y, prod;
    fe_sub(s1, p.Y, p.X);
    fe_add(a1, p.Y, p.X);
    fe_sub(s2, q.Y, q.X);
    fe_add(a2, q.Y, q.X);
    fe_select4(x, role, s1, a1, p.T, p.Z);
    fe_select4(y, role, s2, a2, q.T, q.Z);
    fe_mul(prod, x, y);  // role 0: A, 1: B, 2: T1 T2, 3: Z1 Z2
    fe A, B, C, D;
    fe_quad_exchange(A, B, C, D, prod);
    fe_mul(C, C, fe_const_2d());
    fe_dbl(D, D);
    fe E, F, G, H;
    fe_sub(E, B, A);
    fe_sub(F, D, C);
    fe_add(G, D, C);
    fe_add(H, B, A);
    fe_select4(x, role, E, G, F, E);
    fe_select4(y, role, F, H, G, H);
    fe_mul(prod, x, y);  // role 0: X3 = E F, 1: Y3 = G H, 2: Z3 = F G, 3: T3 = E H
    fe_quad_exchange(r.X, r.Y, r.Z, r.T, prod);
}
// r = 2p
__device__ __forceinline__ void ge_dbl_quad(ge_p3& r, const ge_p3& p) {
    const int role = threadIdx.x & 3;
    fe xy, x, prod;
    fe_add(xy, p.X, p.Y);
    fe_select4(x, role, p.X, p.Y, p.Z, xy);
    fe_sq(prod, x);  // XX, YY, ZZ, (X+Y)^2
    fe XX, YY, ZZ, S;
    fe_quad_exchange(XX, YY, ZZ, S, prod);
    fe E, F, G, H, y;
    fe_dbl(ZZ, ZZ);
    fe_add(H, YY, XX);
    fe_sub(G, YY, XX);
    fe_sub(E, S, H);
    fe_sub(F, ZZ, G);
    fe_select4(x, role, E, G, F, E);
    fe_select4(y, role, F, H, G, H);
    fe_mul(prod, x, y);
    fe_quad_exchange(r.X, r.Y, r.Z, r.T, prod);
}

__device__ __forceinline__ void ge_neg(ge_p3& r, const ge_p3& p) {
    fe_neg(r.X, p.X);
    r.Y = p.Y;
    r.Z = p.Z;
    fe_neg(r.T, p.T);
}

// affine precomputed form of a point with Z == 1
__device__ __forceinline__ void ge_to_niels_affine(ge_niels& r, const fe& x, const fe& y) {
    fe_add(r.yplusx, y, x);
    fe_sub(r.yminusx, y, x);
    fe t;
    fe_mul(t, x, y);
    fe_mul(r.xy2d, t, fe_const_2d());
}

// (X/Z, Y/Z, 1, XY/Z^2), canonical limbs — what the reference's ge25519_normalize intends
// (curve25519_ops.cu:574-605, broken there by D4)
__device__ __forceinline__ void ge_normalize(ge_p3& p) {
    fe zi, x, y;
    fe_invert(zi, p.Z);
    fe_mul(x, p.X, zi);
    fe_mul(y, p.Y, zi);
    fe_canon(x);
    fe_canon(y);
    p.X = x;
    p.Y = y;
    fe_set1(p.Z);
    fe_mul(p.T, x, y);
    fe_canon(p.T);
}

// -X^2 + Y^2 == Z^2 + d T^2, X Y == Z T, Z != 0 (the reference's check is a stub, D8)
__device__ __forceinline__ bool ge_is_on_curve(const ge_p3& p) {
    fe x2, y2, z2, t2, lhs, rhs, a, b;
    fe_sq(x2, p.X);
    fe_sq(y2, p.Y);
    fe_sq(z2, p.Z);
    fe_sq(t2, p.T);
    fe_sub(lhs, y2, x2);
    fe_mul(rhs, t2, fe_const_d());
    fe_add(rhs, rhs, z2);
    fe_mul(a, p.X, p.Y);
    fe_mul(b, p.Z, p.T);
    return fe_equal(lhs, rhs) && fe_equal(a, b) && !fe_iszero(p.Z);
}
__device__ __forceinline__ bool ge_is_identity(const ge_p3& p) {
    return fe_iszero(p.X) && fe_equal(p.Y, p.Z) && !fe_iszero(p.Z);
}

// memory I/O in the reference's AoS ge25519 layout (X@0, Y@32, Z@64, T@96)
__device__ __forceinline__ void ge_load(ge_p3& r, const void* p) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    fe_load(r.X, b);
    fe_load(r.Y, b + 32);
    fe_load(r.Z, b + 64);
    fe_load(r.T, b + 96);
}
__device__ __forceinline__ void ge_store(void* p, const ge_p3& a) {
    uint8_t* b = reinterpret_cast<uint8_t*>(p);
    fe_store(b, a.X);
    fe_store(b + 32, a.Y);
    fe_store(b + 64, a.Z);
    fe_store(b + 96, a.T);
}
__device__ __forceinline__ void ge_niels_load(ge_niels& r, const void* p) {
    const uint8_t* b = reinterpret_cast<const uint8_t*>(p);
    fe_load_nc(r.yplusx, b);
    fe_load_nc(r.yminusx, b + 32);
    fe_load_nc(r.xy2d, b + 64);
}
__device__ __forceinline__ void ge_niels_store(void* p, const ge_niels& a) {
    uint8_t* b = reinterpret_cast<uint8_t*>(p);
    fe_store(b, a.yplusx);
    fe_store(b + 32, a.yminusx);
    fe_store(b + 64, a.xy2d);
}

// r = k * p, plain MSB-first double-and-add over `bits` bits (setup / small cases only)
__device__ __forceinline__ void ge_scalarmult_bits(ge_p3& r, const uint32_t* k, int bits, const ge_p3& p) {
    ge_p3 acc;
    ge_p3_0(acc);
    for (int i = bits - 1; i >= 0; i--) {
        ge_dbl(acc, acc);
        if ((k[i >> 5] >> (i & 31)) & 1) ge_add(acc, acc, p);
    }
    r = acc;
}

}  // namespace cbp
