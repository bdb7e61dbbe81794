// fe8.cuh — GF(2^255-19) and ge25519 arithmetic with ONE 32-bit word per lane: 8 lanes ("octet") per field
// element, a warp = 4 octets = the four independent products of one stage of a point operation.
//
// Why: a dependent chain of point operations (the 255 doublings of the MSM's window combine, a Straus ladder, the
// per-output chains of the IPA's G/H fold) runs in ONE warp and is bound by instruction latency, not throughput.
// With one thread per product (ge_add_quad / ge_dbl_quad in ge25519.cuh) a doubling is ~490 dependent instructions:
// 0.72 us.  Here every lane holds one word, so a multiplication is 8 broadcast + 8 rotate shuffles and 16 wide
// multiply-adds per lane, an addition is one IADD, and carries are resolved across lanes with two ballots — about
// 260 instructions per doubling with far more instruction-level parallelism.
//
// Representation ("tight"): value = sum_j w_j 2^(32 j), lane j holds w_j, w_7 <= 2^31 + 2^8; any representative
// of the residue class mod p = 2^255 - 19 (weakly reduced, like fe25519.cuh; fe_canon on output).  Every function
// here must be called by all 32 lanes of a converged warp; the four octets compute independently.
// The algorithm, its bounds and its carry logic are modelled word for word in tools/fe8_model.py.
#pragma once
#include "ge25519.cuh"

namespace cbp {

struct Fe8Lane {
    uint32_t j;     // word index of this lane inside its octet
    uint32_t base;  // first lane of the octet
    uint32_t oct;   // octet index 0..3
    uint32_t k1;    // word j of p - 37:      a - b     = a + ~b + K1        (mod p)
    uint32_t k2;    // word j of 2p - 74:     a - b - c = a + ~b + ~c + K2   (mod p)
    uint32_t d2;    // word j of 2d
    uint32_t one;   // word j of 1
};
__device__ __forceinline__ Fe8Lane fe8_lane() {
    Fe8Lane L;
    const uint32_t lane = threadIdx.x & 31u;
    L.j = lane & 7u;
    L.base = lane & ~7u;
    L.oct = lane >> 3;
    L.k1 = L.j == 0 ? 0xFFFFFFC8u : (L.j == 7 ? 0x7FFFFFFFu : 0xFFFFFFFFu);
    L.k2 = L.j == 0 ? 0xFFFFFF90u : 0xFFFFFFFFu;
    const fe d2 = fe_const_2d();
    uint32_t w = d2.v[0];
#pragma unroll
    for (int i = 1; i < 8; i++) w = L.j == (uint32_t)i ? d2.v[i] : w;
    L.d2 = w;
    L.one = L.j == 0 ? 1u : 0u;
    return L;
}
__device__ __forceinline__ uint32_t fe8_shfl(uint32_t v, uint32_t src) { return __shfl_sync(0xffffffffu, v, (int)src); }

// lane sums s_j < 2^58 of a lazy linear combination of tight values  ->  tight words, same value mod p.
// Lane 7 keeps 31 bits, its excess re-enters lane 0 times 19 (2^255 = 19); the other lanes hand bits >= 32 to the
// next lane.  What is left is at most one carry bit per lane: lanes whose word is all ones propagate, resolved for
// the whole octet at once by adding the generate and propagate masks as integers (tools/fe8_model.py: ripple).
__device__ __forceinline__ uint32_t fe8_normalize(uint64_t s, const Fe8Lane& L) {
    const bool top = L.j == 7;
    const uint32_t lo = (uint32_t)s & (top ? 0x7FFFFFFFu : 0xFFFFFFFFu);
    const uint32_t hi = (uint32_t)(s >> (top ? 31 : 32));  // < 2^26
    const uint32_t r = fe8_shfl(hi, L.base + ((L.j - 1u) & 7u));
    const uint64_t v = (uint64_t)r * (L.j == 0 ? 19u : 1u) + lo;
    const uint32_t vlo = (uint32_t)v;
    // generate / propagate masks of all four octets at once: lane 7 of an octet neither generates (its word has
    // 31 bits) nor propagates, so bit 7 of every byte of both masks is clear and the integer addition below cannot
    // carry from one octet into the next
    const uint32_t g = __ballot_sync(0xffffffffu, (uint32_t)(v >> 32) != 0);
    const uint32_t x = g | __ballot_sync(0xffffffffu, vlo == 0xFFFFFFFFu);
    const uint32_t cin = (((x + g) ^ x ^ g) >> (threadIdx.x & 31u)) & 1u;
    return vlo + cin;
}
// product of two elements given as 32-bit words (any values < 2^256) -> tight.
// Lane j accumulates the folded column T_j = sum_i a_i b_((j-i) mod 8) (i <= j ? 1 : 38); b is split into 16-bit
// halves so that two plain 64-bit accumulators suffice (each term < 2^53.3).
__device__ __forceinline__ uint32_t fe8_mul(uint32_t a, uint32_t b, const Fe8Lane& L) {
#ifndef FE8_ACC4
#define FE8_ACC4 0
#endif
#if FE8_ACC4
    // even and odd i accumulate separately: four dependent wide multiply-adds deep instead of eight
    uint64_t e0 = 0, e1 = 0, o0 = 0, o1 = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t ai = fe8_shfl(a, L.base + (uint32_t)i);
        const uint32_t br = fe8_shfl(b, L.base + ((L.j - (uint32_t)i) & 7u));
        const uint32_t f = (uint32_t)i <= L.j ? 1u : 38u;
        if (i & 1) {
            o0 += (uint64_t)ai * ((br & 0xFFFFu) * f);
            o1 += (uint64_t)ai * ((br >> 16) * f);
        } else {
            e0 += (uint64_t)ai * ((br & 0xFFFFu) * f);
            e1 += (uint64_t)ai * ((br >> 16) * f);
        }
    }
    const uint64_t acc0 = e0 + o0, acc1 = e1 + o1;  // each < 2^56.3
#else
    uint64_t acc0 = 0, acc1 = 0;  // each < 2^56.3
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t ai = fe8_shfl(a, L.base + (uint32_t)i);
        const uint32_t br = fe8_shfl(b, L.base + ((L.j - (uint32_t)i) & 7u));
        const uint32_t f = (uint32_t)i <= L.j ? 1u : 38u;
        acc0 += (uint64_t)ai * ((br & 0xFFFFu) * f);
        acc1 += (uint64_t)ai * ((br >> 16) * f);
    }
#endif
    // T = acc0 + acc1 2^16 as three words (t2 < 2^10)
    const uint64_t lo64 = acc0 + (acc1 << 16);
    const uint32_t t2 = (uint32_t)(acc1 >> 48) + (lo64 < acc0 ? 1u : 0u);
    const uint32_t t0 = (uint32_t)lo64, t1 = (uint32_t)(lo64 >> 32);
    // word j += t1 of lane j-1 and t2 of lane j-2; what leaves the top wraps around times 38 (2^256 = 38)
    const uint32_t r1 = fe8_shfl(t1, L.base + ((L.j - 1u) & 7u));
    const uint32_t r2 = fe8_shfl(t2, L.base + ((L.j - 2u) & 7u));
    const uint64_t w = (uint64_t)t0 + (uint64_t)r1 * (L.j == 0 ? 38u : 1u) + (uint64_t)(r2 * (L.j < 2 ? 38u : 1u));
    return fe8_normalize(w, L);
}
__device__ __forceinline__ uint32_t fe8_add(uint32_t a, uint32_t b, const Fe8Lane& L) {
    return fe8_normalize((uint64_t)a + b, L);
}
__device__ __forceinline__ uint32_t fe8_sub(uint32_t a, uint32_t b, const Fe8Lane& L) {
    return fe8_normalize((uint64_t)a + (uint32_t)~b + L.k1, L);
}

// ---- points: (X, Y, Z, T) one word of each per lane, every octet holds a full copy ------------------------------
struct ge8 {
    uint32_t X, Y, Z, T;
};
__device__ __forceinline__ void ge8_identity(ge8& r, const Fe8Lane& L) {
    r.X = 0;
    r.Y = L.one;
    r.Z = L.one;
    r.T = 0;
}
// memory (reference ge25519 layout, any weakly reduced limbs) -> tight words
__device__ __forceinline__ void ge8_load(ge8& r, const void* p, const Fe8Lane& L) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    r.X = fe8_normalize((uint64_t)w[L.j], L);
    r.Y = fe8_normalize((uint64_t)w[8 + L.j], L);
    r.Z = fe8_normalize((uint64_t)w[16 + L.j], L);
    r.T = fe8_normalize((uint64_t)w[24 + L.j], L);
}
// octet 0 writes the point (tight words are a valid weakly reduced fe25519)
__device__ __forceinline__ void ge8_store(void* p, const ge8& a, const Fe8Lane& L) {
    if (L.oct != 0) return;
    uint32_t* w = reinterpret_cast<uint32_t*>(p);
    w[L.j] = a.X;
    w[8 + L.j] = a.Y;
    w[16 + L.j] = a.Z;
    w[24 + L.j] = a.T;
}
// the second half of every point operation: from the lane sums of E, F, G, H (lazy) to (EF, GH, FG, EH)
__device__ __forceinline__ void ge8_finish(ge8& r, uint64_t sE, uint64_t sF, uint64_t sG, uint64_t sH, const Fe8Lane& L) {
    // octet 0: E F = X3, octet 1: G H = Y3, octet 2: F G = Z3, octet 3: E H = T3
    const uint64_t sU = (L.oct == 0 || L.oct == 3) ? sE : (L.oct == 1 ? sG : sF);
    const uint64_t sV = L.oct == 0 ? sF : (L.oct == 2 ? sG : sH);
    const uint32_t u = fe8_normalize(sU, L), v = fe8_normalize(sV, L);
    const uint32_t pr = fe8_mul(u, v, L);
    r.X = fe8_shfl(pr, L.j);
    r.Y = fe8_shfl(pr, 8u + L.j);
    r.Z = fe8_shfl(pr, 16u + L.j);
    r.T = fe8_shfl(pr, 24u + L.j);
}
// r = 2p (dbl-2008-hwcd, a = -1; same formulas as ge_dbl): two multiplication levels.  The fourth product of the first
// level is X Y itself rather than (X + Y)^2, so that no operand needs an addition (a carry round) first:
// E = (X + Y)^2 - X^2 - Y^2 = 2 X Y.
#ifndef FE8_DBL_XY
#define FE8_DBL_XY 1
#endif
__device__ __forceinline__ void ge8_dbl(ge8& r, const ge8& p, const Fe8Lane& L) {
#if FE8_DBL_XY
    const uint32_t u = L.oct == 0 ? p.X : (L.oct == 1 ? p.Y : (L.oct == 2 ? p.Z : p.X));
    const uint32_t v = L.oct == 3 ? p.Y : u;
    const uint32_t sq = fe8_mul(u, v, L);
    const uint64_t XX = fe8_shfl(sq, L.j), YY = fe8_shfl(sq, 8u + L.j), ZZ = fe8_shfl(sq, 16u + L.j),
                   XY = fe8_shfl(sq, 24u + L.j);
    const uint64_t nXX = (uint32_t)~(uint32_t)XX, nYY = (uint32_t)~(uint32_t)YY;
    // H = YY + XX, G = YY - XX, E = 2 XY, F = 2 ZZ - G
    ge8_finish(r, XY + XY, ZZ + ZZ + XX + nYY + L.k1, YY + nXX + L.k1, YY + XX, L);
#else
    const uint32_t xy = fe8_add(p.X, p.Y, L);
    const uint32_t opnd = L.oct == 0 ? p.X : (L.oct == 1 ? p.Y : (L.oct == 2 ? p.Z : xy));
    const uint32_t sq = fe8_mul(opnd, opnd, L);
    const uint64_t XX = fe8_shfl(sq, L.j), YY = fe8_shfl(sq, 8u + L.j), ZZ = fe8_shfl(sq, 16u + L.j),
                   S = fe8_shfl(sq, 24u + L.j);
    const uint64_t nXX = (uint32_t)~(uint32_t)XX, nYY = (uint32_t)~(uint32_t)YY;
    ge8_finish(r, S + nXX + nYY + L.k2, ZZ + ZZ + XX + nYY + L.k1, YY + nXX + L.k1, YY + XX, L);
#endif
}
// r = p + q, both extended (unified add-2008-hwcd-3, same formulas as ge_add): three multiplication levels
// (the second one is the constant 2d on the T1 T2 product; the other octets multiply by one)
__device__ __forceinline__ void ge8_add(ge8& r, const ge8& p, const ge8& q, const Fe8Lane& L) {
    const uint64_t sU = L.oct == 0 ? (uint64_t)p.Y + (uint32_t)~p.X + L.k1
                                   : (L.oct == 1 ? (uint64_t)p.Y + p.X : (uint64_t)(L.oct == 2 ? p.T : p.Z));
    const uint64_t sV = L.oct == 0 ? (uint64_t)q.Y + (uint32_t)~q.X + L.k1
                                   : (L.oct == 1 ? (uint64_t)q.Y + q.X : (uint64_t)(L.oct == 2 ? q.T : q.Z));
    uint32_t pr = fe8_mul(fe8_normalize(sU, L), fe8_normalize(sV, L), L);
    pr = fe8_mul(pr, L.oct == 2 ? L.d2 : L.one, L);
    const uint64_t A = fe8_shfl(pr, L.j), B = fe8_shfl(pr, 8u + L.j), C = fe8_shfl(pr, 16u + L.j),
                   Dh = fe8_shfl(pr, 24u + L.j);  // Dh = Z1 Z2, D = 2 Dh
    const uint64_t nA = (uint32_t)~(uint32_t)A, nC = (uint32_t)~(uint32_t)C;
    // E = B - A, F = D - C, G = D + C, H = B + A
    ge8_finish(r, B + nA + L.k1, Dh + Dh + nC + L.k1, Dh + Dh + C, B + A, L);
}
// r = p + q with q in "cached" form (Y+X, Y-X, 2Z, 2dT — tight words): two multiplication levels
struct ge8_cached {
    uint32_t YpX, YmX, Z2, T2d;
};
__device__ __forceinline__ void ge8_to_cached(ge8_cached& r, const ge8& q, const Fe8Lane& L) {
    r.YpX = fe8_add(q.Y, q.X, L);
    r.YmX = fe8_sub(q.Y, q.X, L);
    r.Z2 = fe8_add(q.Z, q.Z, L);
    r.T2d = fe8_mul(q.T, L.d2, L);
}
__device__ __forceinline__ void ge8_add_cached(ge8& r, const ge8& p, const ge8_cached& q, const Fe8Lane& L) {
    const uint64_t sU = L.oct == 0 ? (uint64_t)p.Y + (uint32_t)~p.X + L.k1
                                   : (L.oct == 1 ? (uint64_t)p.Y + p.X : (uint64_t)(L.oct == 2 ? p.T : p.Z));
    const uint32_t v = L.oct == 0 ? q.YmX : (L.oct == 1 ? q.YpX : (L.oct == 2 ? q.T2d : q.Z2));
    const uint32_t pr = fe8_mul(fe8_normalize(sU, L), v, L);
    const uint64_t A = fe8_shfl(pr, L.j), B = fe8_shfl(pr, 8u + L.j), C = fe8_shfl(pr, 16u + L.j),
                   D = fe8_shfl(pr, 24u + L.j);
    const uint64_t nA = (uint32_t)~(uint32_t)A, nC = (uint32_t)~(uint32_t)C;
    ge8_finish(r, B + nA + L.k1, D + nC + L.k1, D + C, B + A, L);
}

// z^(p-2) in octet form (every octet computes the same chain): 254 squarings + 11 multiplications
__device__ __forceinline__ uint32_t fe8_sqn(uint32_t a, int n, const Fe8Lane& L) {
#pragma unroll 1
    for (int i = 0; i < n; i++) a = fe8_mul(a, a, L);
    return a;
}
static __device__ __noinline__ uint32_t fe8_invert(uint32_t z, const Fe8Lane& L) {
    const uint32_t z2 = fe8_mul(z, z, L);
    const uint32_t z9 = fe8_mul(fe8_sqn(z2, 2, L), z, L);
    const uint32_t z11 = fe8_mul(z9, z2, L);
    const uint32_t z_5_0 = fe8_mul(fe8_mul(z11, z11, L), z9, L);
    const uint32_t z_10_0 = fe8_mul(fe8_sqn(z_5_0, 5, L), z_5_0, L);
    const uint32_t z_20_0 = fe8_mul(fe8_sqn(z_10_0, 10, L), z_10_0, L);
    const uint32_t z_40_0 = fe8_mul(fe8_sqn(z_20_0, 20, L), z_20_0, L);
    const uint32_t z_50_0 = fe8_mul(fe8_sqn(z_40_0, 10, L), z_10_0, L);
    const uint32_t z_100_0 = fe8_mul(fe8_sqn(z_50_0, 50, L), z_50_0, L);
    const uint32_t z_200_0 = fe8_mul(fe8_sqn(z_100_0, 100, L), z_100_0, L);
    const uint32_t z_250_0 = fe8_mul(fe8_sqn(z_200_0, 50, L), z_50_0, L);
    return fe8_mul(fe8_sqn(z_250_0, 5, L), z11, L);
}
// all eight words of an octet-form element into every lane's thread-level fe (for canonical output code)
__device__ __forceinline__ void fe8_gather(fe& r, uint32_t a, const Fe8Lane& L) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = fe8_shfl(a, L.base + (uint32_t)i);
}
// (X/Z, Y/Z, 1, XY/Z^2) with canonical limbs, stored by lane 0 — what ge_normalize + ge_store produce.  The inversion
// is the divsteps one (every lane runs it on the same value: no divergence), the rest thread-level.
__device__ __forceinline__ void ge8_store_normalized(void* out, const ge8& p, const Fe8Lane& L) {
    ge_p3 o;
    fe8_gather(o.X, p.X, L);
    fe8_gather(o.Y, p.Y, L);
    fe8_gather(o.Z, p.Z, L);
    fe_set0(o.T);
    ge_normalize(o);
    if ((threadIdx.x & 31u) == 0) ge_store(out, o);
}

}  // namespace cbp
