// rangeproof.cu — batched range-proof verification for sm_100a and the generator tables it uses.
//
// Replaces cuda_range_proof_verify / cuda_inner_product_verify (reference notebook cell
// cuda_range_proof_verify.cu, nb:6529-6900, pure host code with heuristic comparisons) and the CPU
// verifier range_proof_verify (bulletproof_range_proof.cu:1717-1812) + inner_product_verify
// (bulletproof_vectors.cu:541-762).  Decisions are bit-exact with oracle/ref_corrected.c.
//
// Instead of folding G/H round by round (the reference's 4(n-1) scalar multiplications,
// bulletproof_vectors.cu:641-663) the verifier evaluates the two identities
//   (t - delta) g + taux h                       ==  z^2 V + x T1 + x^2 T2
//   sum (a s_i + z) G_i + sum ((b s_i^-1 - z^2 2^i) y^-i - z) H_i + (mu + ab - t) h
//                                                ==  A + x S + sum u_j^2 L_j + sum u_j^-2 R_j
// as multi-scalar sums: the 2n+2 shared generators through precomputed fixed-base tables (8- or 16-bit
// windows, no doublings), the 2 log n + 5 per-proof points with signed 5-bit windows.
//
// Kernels per batch (see the block comment above verify_coeff_kernel for the middle ones):
//   verify_transcript  1 thread / proof : on-curve checks, SHA-256 challenges, scalar inversions
//   verify_coeff / verify_fixed / verify_vtab / verify_winsum : the two multi-scalar sums, by phase
//   verify_finish      1 thread / (proof, identity): Horner over the 51 windows, equality test
//   verify_combine     accept bits
#include <stdio.h>
#include <string.h>
#include <stdlib.h>
#include <mutex>
#include <unordered_map>
#include "../../include/cuda_bulletproof.h"
#include "common.h"
#include "fe8.cuh"
#include "rangeproof.cuh"
#include "sha256.cuh"

namespace cbp {

// ---- generator tables -----------------------------------------------------------------------------
// pow_out[b][j] = 2^(wbits j) Base_b
__global__ void gens_pow_kernel(const uint8_t* __restrict__ G, const uint8_t* __restrict__ H,
                                const uint8_t* __restrict__ g, const uint8_t* __restrict__ h, uint32_t n, int wbits,
                                uint8_t* __restrict__ pow_out, uint8_t* __restrict__ bases_out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t nb = 2 * n + 2;
    if (b >= nb) return;
    const uint8_t* src = b < n ? G + (size_t)b * 128 : b < 2 * n ? H + (size_t)(b - n) * 128 : b == 2 * n ? g : h;
    ge_p3 P;
    ge_load(P, src);
    ge_p3 Pn = P;
    ge_normalize(Pn);
    ge_store(bases_out + (size_t)b * 128, Pn);
    const int nwin = fix_nwin(wbits);
    for (int j = 0; j < nwin; j++) {
        ge_store(pow_out + ((size_t)b * nwin + j) * 128, P);
        for (int s = 0; s < wbits; s++) ge_dbl(P, P);
    }
}
// One thread per run of kTabRun consecutive multiples of M = 2^(wbits win) Base: the first one by
// double-and-add, the rest by repeated addition of M, then ONE field inversion for the run (Montgomery's
// trick, prefix products kept in local memory) to write the affine precomputed form (y+x, y-x, 2dxy).
static constexpr int kTabRun = 32;
__global__ void __launch_bounds__(64) gens_table_kernel(const uint8_t* __restrict__ pow_in, uint32_t nbases,
                                                        int nwin, uint32_t entries, uint8_t* __restrict__ table) {
    const uint32_t runs = entries / kTabRun;
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= (size_t)nbases * nwin * runs) return;
    const size_t bw = id / runs;
    const uint32_t run_idx = (uint32_t)(id % runs);
    ge_p3 M, acc;
    ge_load(M, pow_in + bw * 128);
    uint32_t m0 = run_idx * kTabRun + 1;
    ge_scalarmult_bits(acc, &m0, 32 - __clz(m0), M);
    uint8_t* slot = table + (bw * entries + (m0 - 1)) * 96;
    fe pre[kTabRun];
    fe run;
    fe_set1(run);
#pragma unroll 1
    for (int d = 0; d < kTabRun; d++) {
        fe_store(slot + d * 96, acc.X);
        fe_store(slot + d * 96 + 32, acc.Y);
        fe_store(slot + d * 96 + 64, acc.Z);
        pre[d] = run;
        fe_mul(run, run, acc.Z);
        ge_add(acc, acc, M);
    }
    fe inv;
    fe_invert(inv, run);
#pragma unroll 1
    for (int d = kTabRun - 1; d >= 0; d--) {
        fe X, Y, Z, zi;
        fe_load(X, slot + d * 96);
        fe_load(Y, slot + d * 96 + 32);
        fe_load(Z, slot + d * 96 + 64);
        fe_mul(zi, inv, pre[d]);
        fe_mul(inv, inv, Z);
        fe_mul(X, X, zi);
        fe_mul(Y, Y, zi);
        ge_niels q;
        ge_to_niels_affine(q, X, Y);
        ge_niels_store(slot + d * 96, q);
    }
}

// ---- per-proof scalars produced by verify_transcript ------------------------------------------------
struct VScal {
    uint32_t valid;
    uint32_t pad[7];
    sc z, z2, x, x2, a, b;
    sc g1;   // identity 1, g coefficient: t - delta
    sc h1;   // identity 1, h coefficient: taux
    sc h2;   // identity 2, h coefficient: mu + a b - t   (Q = h)
    sc s0;   // prod u_j^-1
    sc ypow[kMaxK + 1];  // y^-(2^m)
    sc usq[kMaxK], uinvsq[kMaxK];
};

__device__ __forceinline__ void words_of(uint32_t (&w)[8], const fe& a) {
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = a.v[i];
}
// affine coordinates of a projective point for hashing (fast path when Z == 1)
__device__ __forceinline__ void affine_xy(fe& x, fe& y, const ge_p3& p) {
    fe one;
    fe_set1(one);
    if (fe_equal(p.Z, one)) {
        x = p.X;
        y = p.Y;
    } else {
        fe zi;
        fe_invert(zi, p.Z);
        fe_mul(x, p.X, zi);
        fe_mul(y, p.Y, zi);
    }
    fe_canon(x);
    fe_canon(y);
}
__device__ __forceinline__ void sc_from_words(sc& r, const uint32_t (&w)[8]) {
    sc t;
#pragma unroll
    for (int i = 0; i < 8; i++) t.v[i] = w[i];
    sc_reduce(r, t);
}
__device__ __forceinline__ void sc_load_reduce(sc& r, const void* p) {
    sc t;
    sc_load(t, p);
    sc_reduce(r, t);
}

// out-of-line scalar multiplication for the transcript kernel: ~45 call sites of 9 KB each otherwise
static __device__ __noinline__ void sc_mul_nf(sc& r, const sc& a, const sc& b) { sc_mul(r, a, b); }

__global__ void __launch_bounds__(64) verify_transcript_kernel(const uint8_t* __restrict__ proofs, size_t rec_bytes,
                                                               const uint8_t* __restrict__ Vext, uint32_t n, int k,
                                                               uint32_t num, VScal* __restrict__ out) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num) return;
    const uint8_t* rec = proofs + (size_t)p * rec_bytes;
    VScal& vs = out[p];
    bool valid = true;
    // every point of the record through ONE loop body (on-curve check, affine coordinates for hashing):
    // q = 0..4: V, A, S, T1, T2; then L_0..L_(k-1), R_0..R_(k-1)
    uint32_t ax[5 + 2 * kMaxK][8], ay[5 + 2 * kMaxK][8];
    ge_p3 V;
#pragma unroll 1
    for (int q = 0; q < 5 + 2 * k; q++) {
        ge_p3 P;
        ge_load(P, rec + (q < 5 ? q * 128 : kRecL + (q - 5) * 128));
        valid = valid && ge_is_on_curve(P);
        fe x, y;
        affine_xy(x, y, P);
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ax[q][i] = x.v[i];
            ay[q][i] = y.v[i];
        }
        if (q == 0) V = P;
    }
    if (Vext) {  // bulletproof_range_proof.cu:1729-1740: the caller's V must be the proof's V
        ge_p3 E;
        ge_load(E, Vext + (size_t)p * 128);
        fe a, b, c, d;
        fe_mul(a, E.X, V.Z);
        fe_mul(b, V.X, E.Z);
        fe_mul(c, E.Y, V.Z);
        fe_mul(d, V.Y, E.Z);
        valid = valid && ge_is_on_curve(E) && fe_equal(a, b) && fe_equal(c, d);
    }
    uint32_t yb[8], zb[8], xb[8];
    Sha256 sh;
    // y: bulletproof_challenge.cu:24-44
    sh.init();
    sh.update_str("BulletproofYChal", 16);
#pragma unroll 1
    for (int q = 0; q < 3; q++) {
        sh.update_words(ax[q]);
        sh.update_words(ay[q]);
    }
    sh.update_str("y_ch", 4);
    sh.final_challenge(yb);
    // z: :47-58
    sh.init();
    sh.update_str("BulletproofZChal", 16);
    sh.update_words(yb);
    sh.update_str("z_ch", 4);
    sh.final_challenge(zb);
    // x: :61-77 (only 4 bytes of "xchal" are hashed)
    sh.init();
    sh.update_str("BulletproofXChal", 16);
#pragma unroll 1
    for (int q = 3; q < 5; q++) {
        sh.update_words(ax[q]);
        sh.update_words(ay[q]);
    }
    sh.update_str("xcha", 4);
    sh.final_challenge(xb);

    sc y, z, x, t, taux, mu, a, b, c;
    sc_from_words(y, yb);
    sc_from_words(z, zb);
    sc_from_words(x, xb);
    sc_load_reduce(t, rec + kRecT);
    sc_load_reduce(taux, rec + kRecTaux);
    sc_load_reduce(mu, rec + kRecMu);
    sc_load_reduce(a, rec + kRecIpA);
    sc_load_reduce(b, rec + kRecIpB);
    sc_load_reduce(c, rec + kRecIpC);
    bool c_ok = true;
#pragma unroll
    for (int i = 0; i < 8; i++) c_ok = c_ok && (c.v[i] == t.v[i]);
    valid = valid && c_ok;

    // IPA transcript: bulletproof_range_proof.cu:1668-1676, bulletproof_vectors.cu:448-465
    uint32_t tr[8];
    sh.init();
    sh.update_str("BulletproofIP", 13);
    sh.update_words(t.v);
    sh.update_words(taux.v);
    sh.update_words(mu.v);
    sh.final_challenge(tr);
    sc u[kMaxK];
#pragma unroll 1
    for (int j = 0; j < k; j++) {
        sh.init();
        sh.update_str("InnerProductChal", 16);
        sh.update_words(tr);
        sh.update_words(ax[5 + j]);      // L_j.x
        sh.update_words(ax[5 + k + j]);  // R_j.x
        sh.final_challenge(tr);
        if (j == 0) {  // stored first-round challenge must be the recomputed one (D15)
            fe xs;
            fe_load(xs, rec + kRecIpX);
            fe_canon(xs);
            bool same = true;
#pragma unroll
            for (int i = 0; i < 8; i++) same = same && (xs.v[i] == tr[i]);
            valid = valid && same;
        }
        sc_from_words(u[j], tr);
    }
    // batch inversion of y, u_0 .. u_{k-1}
    sc pre[kMaxK + 1], run, inv;
    sc_set1(run);
    pre[0] = run;
    sc_mul_nf(run, run, y);
    for (int j = 0; j < k; j++) {
        pre[j + 1] = run;
        sc_mul_nf(run, run, u[j]);
    }
    sc_invert(inv, run);
    sc uinv[kMaxK], yinv;
    for (int j = k - 1; j >= 0; j--) {
        sc_mul_nf(uinv[j], inv, pre[j + 1]);
        sc_mul_nf(inv, inv, u[j]);
    }
    yinv = inv;  // pre[0] = 1

    vs.z = z;
    sc_mul_nf(vs.z2, z, z);
    vs.x = x;
    sc_mul_nf(vs.x2, x, x);
    vs.a = a;
    vs.b = b;
    // delta = (z - z^2) sum y^i - z^3 (2^n - 1): bulletproof_range_proof.cu:315-374
    sc sum_y, cur, zmz2, z3, two_n, delta, tmp;
    // sum_{i < 2^k} y^i = prod_{m < k} (1 + y^(2^m))
    sc_set1(sum_y);
    cur = y;
#pragma unroll 1
    for (int m = 0; m < k; m++) {
        sc_set1(tmp);
        sc_add(tmp, tmp, cur);
        sc_mul_nf(sum_y, sum_y, tmp);
        sc_mul_nf(cur, cur, cur);
    }
    sc_sub(zmz2, z, vs.z2);
    sc_mul_nf(z3, vs.z2, z);
    sc_set0(two_n);
    two_n.v[n >> 5] = 1u << (n & 31);  // n <= 64 < 252
    sc one;
    sc_set1(one);
    sc_sub(two_n, two_n, one);
    sc_mul_nf(delta, zmz2, sum_y);
    sc_mul_nf(tmp, z3, two_n);
    sc_sub(delta, delta, tmp);
    sc_sub(vs.g1, t, delta);
    vs.h1 = taux;
    sc_mul_nf(tmp, a, b);
    sc_sub(tmp, tmp, t);
    sc_add(vs.h2, tmp, mu);
    sc s0;
    sc_set1(s0);
    for (int j = 0; j < k; j++) {
        sc_mul_nf(s0, s0, uinv[j]);
        sc_mul_nf(vs.usq[j], u[j], u[j]);
        sc_mul_nf(vs.uinvsq[j], uinv[j], uinv[j]);
    }
    vs.s0 = s0;
    vs.ypow[0] = yinv;
    for (int m = 1; m <= k; m++) sc_mul_nf(vs.ypow[m], vs.ypow[m - 1], vs.ypow[m - 1]);
    vs.valid = valid ? 1u : 0u;
#ifdef CBP_DEBUG_VSCAL
    if (k <= 4) {
        vs.ypow[5] = sum_y; vs.ypow[6] = delta; vs.usq[4] = zmz2; vs.usq[5] = z3; vs.uinvsq[4] = two_n;
        sc dbg; sc_mul_nf(dbg, uinv[0], uinv[1]); vs.uinvsq[5] = dbg;
    }
#endif
}

// ---- the per-proof multi-scalar sums ----------------------------------------------------------------
// Split by phase into small single-purpose kernels (a first monolithic one-CTA-per-proof kernel was
// instruction-fetch and barrier bound: 243 KB of SASS, 16 % issue utilisation, profiles/r01_verify_msm_*):
//   verify_coeff    thread / (proof, slot)        : coefficients -> signed digits (global memory)
//   verify_fixed    WARP   / proof                 : 131 x 32 table additions, lane = window, shuffle tree
//   verify_vtab     thread / (proof, point)        : multiples 1..16 of the 17 per-proof points (addition chain)
//   verify_winsum   thread / (identity, proof, w)  : 5-bit window sums over the per-proof points
static constexpr int kVarMax = 2 + 2 * kMaxK + 3;  // A, S, L_j, R_j | V, T1, T2
// per-proof points: signed kVarBits-bit windows, multiples 1..2^(kVarBits-1) per point
static constexpr int kVarBits = 5;  // measured per 2^14 proofs: 4 bits 6.46 ms, 5 bits 6.12 ms, 6 bits 6.22 ms (table build no longer hides)
static constexpr int kVarWin = (253 + kVarBits) / kVarBits;  // digits of a scalar < 2^253 (+ carry room)
static constexpr int kVarEntries = 1 << (kVarBits - 1);
static_assert(kVarWin <= 64 && kVarWin * kVarBits >= 254, "digit rows are 64 bytes");
static constexpr int kCoeffThreads = 96;           // n + 3 + kVarMax <= 96 for n <= 64
static constexpr int kRowsMax = 2 * kMaxN + 3;     // G_i, H_i, h(identity 2), g, h(identity 1)

__device__ __forceinline__ int var_point_offset(int q, int k) {  // record offset of per-proof point q
    int nvar2 = 2 + 2 * k;
    return q == 0 ? kRecA : q == 1 ? kRecS : q < nvar2 ? kRecL + (q - 2) * 128 : q == nvar2 ? kRecV :
           q == nvar2 + 1 ? kRecT1 : kRecT2;
}

__global__ void __launch_bounds__(kCoeffThreads) verify_coeff_kernel(const uint8_t* __restrict__ gens,
                                                                     const VScal* __restrict__ vscal, uint32_t n, int k,
                                                                     int8_t* __restrict__ digits,
                                                                     int8_t* __restrict__ vdigits) {
    __shared__ sc s_sh[kMaxN], y_sh[kMaxN];
    const uint32_t p = blockIdx.x;
    const int t = threadIdx.x;
    const VScal& vs = vscal[p];
    if (!vs.valid) return;  // whole CTA
    const int wbits = (int)reinterpret_cast<const GensHeader*>(gens)->wbits;
    int8_t* drow = digits + (size_t)p * kRowsMax * kFixRowBytes;
    int8_t* vrow = vdigits + (size_t)p * kVarMax * 64;
    const int nvar2 = 2 + 2 * k, nvar = nvar2 + 3;
    if (t >= (int)n && t < (int)n + 3) {
        int r = t - (int)n;  // row 2n: h (mu + ab - t), 2n+1: g (t - delta), 2n+2: h (taux)
        sc v = r == 0 ? vs.h2 : r == 1 ? vs.g1 : vs.h1;
        fix_recode(drow + (size_t)(2 * n + r) * kFixRowBytes, v, wbits);
    } else if (t >= (int)n + 3 && t - (int)n - 3 < nvar) {
        int q = t - (int)n - 3;  // positive scalars of the per-proof points (the two sides are compared)
        sc sv;
        if (q == 0) sc_set1(sv);
        else if (q == 1) sv = vs.x;
        else if (q < 2 + k) sv = vs.usq[q - 2];
        else if (q < nvar2) sv = vs.uinvsq[q - 2 - k];
        else if (q == nvar2) sv = vs.z2;
        else if (q == nvar2 + 1) sv = vs.x;
        else sv = vs.x2;
        sc_recode_signed<kVarBits>(vrow + (size_t)q * 64, sv, kVarWin);
    }
    // s_t = prod_j u_j^(+-1) = s_0 * prod over the set bits m of t of u_(k-1-m)^2, and y^-t from the
    // y^-(2^m) ladder, by doubling the filled prefix: level m fills [2^m, 2^(m+1)) from [0, 2^m)
    if (t == 0) {
        s_sh[0] = vs.s0;
        sc one;
        sc_set1(one);
        y_sh[0] = one;
    }
    // levels 0..4 live entirely in warp 0 (t < 32): a warp-level barrier orders them; only the step into the
    // second warp and the final cross-warp reads need the CTA barrier
    __syncwarp();
#pragma unroll 1
    for (int m = 0; m < k; m++) {
        const int half = 1 << m;
        if (half >= 32) __syncthreads();
        if (t >= half && t < 2 * half) {
            sc a = s_sh[t - half], b = y_sh[t - half];
            sc_mul_nf(a, a, vs.usq[k - 1 - m]);
            sc_mul_nf(b, b, vs.ypow[m]);
            s_sh[t] = a;
            y_sh[t] = b;
        }
        __syncwarp();
    }
    __syncthreads();
    if (t < (int)n) {
        sc cg, ch, tmp, two_i;
        sc_mul_nf(cg, vs.a, s_sh[t]);
        sc_add(cg, cg, vs.z);
        sc_set0(two_i);
        two_i.v[t >> 5] = 1u << (t & 31);
        sc_mul_nf(tmp, vs.z2, two_i);
        sc_mul_nf(ch, vs.b, s_sh[n - 1 - t]);  // s_(n-1-t) = s_t^-1: the bits complemented
        sc_sub(ch, ch, tmp);
        sc_mul_nf(ch, ch, y_sh[t]);
        sc_sub(ch, ch, vs.z);
        fix_recode(drow + (size_t)t * kFixRowBytes, cg, wbits);
        fix_recode(drow + (size_t)(n + t) * kFixRowBytes, ch, wbits);
    }
}

// One (half-)warp per proof, lane = window: for every row the lanes read their digits with one 32-byte
// access and each adds digit * 2^(wbits lane) * Base from the table (7M, next operand prefetched).
// 2n+3 perfectly balanced additions per lane, no shared memory, no CTA barrier; shuffle-tree at the end.
// WBITS = 8: 32 lanes per proof, table in L2.  WBITS = 16: 16 lanes per proof, 96-byte reads from HBM.
__device__ __forceinline__ void ge_shfl_xor(ge_p3& out, const ge_p3& in, int mask) {
#pragma unroll
    for (int j = 0; j < 8; j++) {
        out.X.v[j] = __shfl_xor_sync(0xffffffffu, in.X.v[j], mask);
        out.Y.v[j] = __shfl_xor_sync(0xffffffffu, in.Y.v[j], mask);
        out.Z.v[j] = __shfl_xor_sync(0xffffffffu, in.Z.v[j], mask);
        out.T.v[j] = __shfl_xor_sync(0xffffffffu, in.T.v[j], mask);
    }
}
static constexpr int kFixAhead = 4;  // rows of L2 prefetch distance (16-bit tables)
template <int WBITS>
__global__ void __launch_bounds__(128, 4) verify_fixed_kernel(const uint8_t* __restrict__ gens,
                                                              const VScal* __restrict__ vscal,
                                                              const int8_t* __restrict__ digits, uint32_t n,
                                                              uint32_t num, uint8_t* __restrict__ fsum) {
    constexpr int LP = 256 / WBITS;  // lanes (= windows) per proof
    constexpr uint32_t E = 1u << (WBITS - 1);
    const uint32_t p = (blockIdx.x * blockDim.x + threadIdx.x) / LP;
    const int win = threadIdx.x & (LP - 1);
    const bool live = p < num && vscal[p < num ? p : 0].valid;
    if (!__any_sync(0xffffffffu, live)) return;  // dead lanes of a live warp idle through the shuffles
    const GensHeader* gh = reinterpret_cast<const GensHeader*>(gens);
    const uint8_t* table = gens + gh->table_off + (size_t)win * E * 96;  // this lane's window
    const int8_t* drow = digits + (size_t)(live ? p : 0) * kRowsMax * kFixRowBytes;
    const int nrows2 = 2 * (int)n + 1, nrows = nrows2 + 2;
    // sequence: the two identity-1 rows (g, h) first, then the 2n+1 identity-2 rows — one accumulator live
    auto row_of = [&](int sidx) { return sidx < 2 ? nrows2 + sidx : sidx - 2; };
    auto base_of = [&](int row) {
        return row < 2 * (int)n ? (uint32_t)row : row == 2 * (int)n ? 2 * n + 1 : 2 * n + (uint32_t)(row - nrows2);
    };
    auto digit_of = [&](int row, uint32_t& mag, bool& neg) {
        if (!live) {
            mag = 0;
            neg = false;
        } else if (WBITS == 8) {
            fixed_digit(drow[(size_t)row * kFixRowBytes + win], mag, neg);
        } else {
            fixed_digit16(reinterpret_cast<const int16_t*>(drow + (size_t)row * kFixRowBytes)[win], mag, neg);
        }
    };
    ge_p3 acc;
    ge_p3_0(acc);
    uint32_t mag;
    bool neg;
    ge_niels q;
    {
        int r0 = row_of(0);
        digit_of(r0, mag, neg);
        if (mag) ge_niels_load(q, table + ((size_t)base_of(r0) * LP * E + (mag - 1)) * 96);
    }
#pragma unroll 1
    for (int sidx = 0; sidx < nrows; sidx++) {
        uint32_t cmag = mag;
        bool cneg = neg;
        ge_niels cur = q;
        if (sidx + 1 < nrows) {  // prefetch the next operand while this addition runs
            int nr = row_of(sidx + 1);
            digit_of(nr, mag, neg);
            if (mag) ge_niels_load(q, table + ((size_t)base_of(nr) * LP * E + (mag - 1)) * 96);
        }
        if (WBITS == 16 && sidx + kFixAhead < nrows) {  // HBM-resident table: pull a later entry into L2 now
            int fr = row_of(sidx + kFixAhead);
            uint32_t fmag;
            bool fneg;
            digit_of(fr, fmag, fneg);
            if (fmag) {
                const uint8_t* e = table + ((size_t)base_of(fr) * LP * E + (fmag - 1)) * 96;
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e));
                asm volatile("prefetch.global.L2 [%0];" ::"l"(e + 64));
            }
        }
        if (cmag) ge_madd(acc, acc, cur, cneg);
        if (sidx == 1 || sidx == nrows - 1) {  // end of an identity: butterfly sum over the windows
#pragma unroll 1
            for (int o = LP / 2; o > 0; o >>= 1) {
                ge_p3 other;
                ge_shfl_xor(other, acc, o);
                ge_add(acc, acc, other);
            }
            if (live && win == 0) ge_store(fsum + ((size_t)p * 2 + (sidx == 1 ? 0 : 1)) * 128, acc);
            ge_p3_0(acc);
        }
    }
}

// multiples 1..kVarEntries of per-proof point q as cached points, one thread per (proof, point): a chain of
// additions of P (the entries are needed in order anyway; an earlier version recomputed every multiple by
// double-and-add in its own thread, 3x the work)
// For one point per identity (S with scalar x, V with scalar z^2: every window digit is non-zero) the multiples
// are also kept in extended form: the window sums START from that entry instead of adding it to the identity.
__global__ void __launch_bounds__(128) verify_vtab_kernel(const uint8_t* __restrict__ proofs, size_t rec_bytes, int k,
                                                          uint32_t num, uint8_t* __restrict__ vtab,
                                                          uint8_t* __restrict__ vseed) {
    const int nvar = 2 + 2 * k + 3;
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= num * (uint32_t)nvar) return;
    // no dependency on the transcript (runs concurrently with it): records that later turn out invalid
    // just produce unused table entries
    const uint32_t p = id / (uint32_t)nvar;
    const int q = (int)(id % (uint32_t)nvar);
    ge_p3 P, acc;
    ge_load(P, proofs + (size_t)p * rec_bytes + var_point_offset(q, k));
    acc = P;
    uint8_t* dst = vtab + ((size_t)p * kVarMax + q) * kVarEntries * 128;
    const int nvar2 = 2 + 2 * k;
    const int seed_slot = q == 1 ? 1 : q == nvar2 ? 0 : -1;  // index as in the window sums: 0 identity 1, 1 identity 2
    uint8_t* sdst = seed_slot >= 0 ? vseed + ((size_t)p * 2 + seed_slot) * kVarEntries * 128 : nullptr;
#pragma unroll 1
    for (int m = 0; m < kVarEntries; m++) {
        if (sdst) ge_store(sdst + m * 128, acc);
        ge_cached c;
        ge_to_cached(c, acc);
        fe_store(dst + m * 128, c.YplusX);
        fe_store(dst + m * 128 + 32, c.YminusX);
        fe_store(dst + m * 128 + 64, c.Z2);
        fe_store(dst + m * 128 + 96, c.T2d);
        if (m + 1 < kVarEntries) ge_add(acc, acc, P);
    }
}
// window sums: threads [0, kVarWin num) identity 2 (A, S, L_j, R_j), [kVarWin num, 2 kVarWin num) identity 1 (V, T1, T2)
__global__ void __launch_bounds__(128) verify_winsum_kernel(const VScal* __restrict__ vscal,
                                                            const int8_t* __restrict__ vdigits,
                                                            const uint8_t* __restrict__ vtab,
                                                            const uint8_t* __restrict__ vseed, int k, uint32_t num,
                                                            uint8_t* __restrict__ winsum) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= num * 2 * kVarWin) return;
    int which = id >= num * kVarWin;  // 1: identity 1
    uint32_t rem = which ? id - num * kVarWin : id, p = rem / kVarWin;
    int w = (int)(rem % kVarWin);
    if (!vscal[p].valid) return;
    const int nvar2 = 2 + 2 * k, nvar = nvar2 + 3;
    int q0 = which ? nvar2 : 0, q1 = which ? nvar : nvar2;
    ge_p3 ws;
    ge_p3_0(ws);
    const int qs = which ? nvar2 : 1;  // the seed point of this identity (V resp. S)
    {
        int d = vdigits[((size_t)p * kVarMax + qs) * 64 + w];
        if (d != 0) {
            int mag = d < 0 ? -d : d;
            ge_load(ws, vseed + (((size_t)p * 2 + (which ? 0 : 1)) * kVarEntries + (mag - 1)) * 128);
            if (d < 0) ge_neg(ws, ws);
        }
    }
#pragma unroll 1
    for (int q = q0; q < q1; q++) {
        if (q == qs) continue;
        int d = vdigits[((size_t)p * kVarMax + q) * 64 + w];
        if (d != 0) {
            int mag = d < 0 ? -d : d;
            const uint8_t* src = vtab + (((size_t)p * kVarMax + q) * kVarEntries + (mag - 1)) * 128;
            ge_cached c;
            fe_load(c.YplusX, src);
            fe_load(c.YminusX, src + 32);
            fe_load(c.Z2, src + 64);
            fe_load(c.T2d, src + 96);
            ge_add_cached(ws, ws, c, d < 0);
        }
    }
    ge_store(winsum + (((size_t)p * 2 + (which ? 0 : 1)) * kVarWin + w) * 128, ws);
}

// Horner over the kVarWin window sums, then F == Var as projective points.  index 0: identity 1, 1: identity 2.
__global__ void __launch_bounds__(64) verify_finish_kernel(const VScal* __restrict__ vscal,
                                                           const uint8_t* __restrict__ fsum,
                                                           const uint8_t* __restrict__ winsum, uint32_t num,
                                                           uint8_t* __restrict__ flags) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= num * 2) return;
    if (!vscal[id >> 1].valid) {
        flags[id] = 0;
        return;
    }
    ge_p3 acc;
    ge_p3_0(acc);
    const uint8_t* ws = winsum + (size_t)id * kVarWin * 128;
#pragma unroll 1
    for (int w = kVarWin - 1; w >= 0; w--) {
#pragma unroll 1  // keep the loop body small: this kernel is latency-bound and was stalling on instruction fetch
        for (int d = 0; d < kVarBits - 1; d++) ge_dbl_xyz(acc, acc);  // T is only needed by the addition
        ge_dbl(acc, acc);
        ge_p3 x;
        ge_load(x, ws + (size_t)w * 128);
        ge_add(acc, acc, x);
    }
    ge_p3 F;
    ge_load(F, fsum + (size_t)id * 128);
    fe a, b, c, d;
    fe_mul(a, F.X, acc.Z);
    fe_mul(b, acc.X, F.Z);
    fe_mul(c, F.Y, acc.Z);
    fe_mul(d, acc.Y, F.Z);
    flags[id] = (fe_equal(a, b) && fe_equal(c, d)) ? 1 : 0;
}
// ---- a handful of proofs: latency instead of throughput -------------------------------------------------------
// The two kernels above give one (half-)warp resp. one thread to a proof: 131 dependent table additions per lane
// (0.40 ms) and a 255-doubling chain per thread (0.51 ms) — fine with 2^14 proofs in flight, but it is all the time
// ONE proof takes, and one proof per call is what the reference's cuda_range_proof_verify hands over
// (complete_bulletproof_test.cu:153).  For small batches:
//   verify_fixed_small   CTA of 1024 threads per proof: the (2n+3) x LP table additions spread over all threads
//                        (warps 0-1: identity 1, warps 2-31: identity 2), shuffle trees, shared-memory combine
//   verify_finish_small  CTA per (proof, identity): the 51 window sums in groups of four (13 warps side by side,
//                        octet form), then one warp runs the Horner chain over the groups: 255 doublings and
//                        15 additions deep at 0.52 / 0.73 us instead of 255 + 51 at 1.67 / 2.2 us
__device__ __forceinline__ void ge_warp_tree_sum(ge_p3& v) {  // result in lane 0
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_p3 other;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            other.X.v[j] = __shfl_down_sync(0xffffffffu, v.X.v[j], o);
            other.Y.v[j] = __shfl_down_sync(0xffffffffu, v.Y.v[j], o);
            other.Z.v[j] = __shfl_down_sync(0xffffffffu, v.Z.v[j], o);
            other.T.v[j] = __shfl_down_sync(0xffffffffu, v.T.v[j], o);
        }
        ge_add(v, v, other);
    }
}
__global__ void __launch_bounds__(1024) verify_fixed_small_kernel(const uint8_t* __restrict__ gens,
                                                                  const VScal* __restrict__ vscal,
                                                                  const int8_t* __restrict__ digits, uint32_t n,
                                                                  uint8_t* __restrict__ fsum) {
    __shared__ __align__(16) uint8_t sh[32][128];
    const uint32_t p = blockIdx.x;
    if (!vscal[p].valid) return;  // whole CTA
    const FixTab ft = fixtab_of(gens);
    const int LP = ft.nwin;
    const int8_t* drow = digits + (size_t)p * kRowsMax * kFixRowBytes;
    const int nrows2 = 2 * (int)n + 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    ge_p3 acc;
    ge_p3_0(acc);
    if (warp < 2) {  // identity 1: rows 2n+1 (g) and 2n+2 (h), bases 2n and 2n+1
        for (int item = (int)threadIdx.x; item < 2 * LP; item += 64) {
            const int r = item / LP, win = item % LP;
            fixed_base_madd(acc, ft, 2 * n + (uint32_t)r, win, drow + (size_t)(nrows2 + r) * kFixRowBytes);
        }
    } else {  // identity 2: rows 0..2n (G_i, H_i, h with base 2n+1)
        for (int item = (int)threadIdx.x - 64; item < nrows2 * LP; item += 960) {
            const int r = item / LP, win = item % LP;
            fixed_base_madd(acc, ft, r < 2 * (int)n ? (uint32_t)r : 2 * n + 1, win, drow + (size_t)r * kFixRowBytes);
        }
    }
    ge_warp_tree_sum(acc);
    if (lane == 0) ge_store(sh[warp], acc);
    __syncthreads();
    if (warp == 0) {  // lanes 0-1 hold identity 1's two warp sums
        ge_p3_0(acc);
        if (lane < 2) ge_load(acc, sh[lane]);
        ge_warp_tree_sum(acc);
        if (lane == 0) ge_store(fsum + ((size_t)p * 2 + 0) * 128, acc);
    } else if (warp == 1) {  // lanes 0-29: identity 2's thirty warp sums
        ge_p3_0(acc);
        if (lane < 30) ge_load(acc, sh[2 + lane]);
        ge_warp_tree_sum(acc);
        if (lane == 0) ge_store(fsum + ((size_t)p * 2 + 1) * 128, acc);
    }
}
static constexpr int kFinGroup = 4, kFinGroups = (kVarWin + kFinGroup - 1) / kFinGroup;  // 51 windows -> 13 groups
__global__ void __launch_bounds__(32 * kFinGroups) verify_finish_small_kernel(const VScal* __restrict__ vscal,
                                                                              const uint8_t* __restrict__ fsum,
                                                                              const uint8_t* __restrict__ winsum,
                                                                              uint8_t* __restrict__ flags) {
    __shared__ __align__(16) uint8_t sh[kFinGroups][128];
    const uint32_t id = blockIdx.x;  // proof * 2 + identity
    if (!vscal[id >> 1].valid) {
        if (threadIdx.x == 0) flags[id] = 0;
        return;
    }
    const Fe8Lane L = fe8_lane();
    const int warp = threadIdx.x >> 5;
    const uint8_t* ws = winsum + (size_t)id * kVarWin * 128;
    {  // group g = windows 4g .. 4g+3 (the top group is shorter): T_g = sum_i 2^(5 i) S_(4g+i), Horner inside the group
        const int w0 = warp * kFinGroup, w1 = w0 + kFinGroup < kVarWin ? w0 + kFinGroup : kVarWin;
        ge8 acc, x;
        ge8_load(acc, ws + (size_t)(w1 - 1) * 128, L);
#pragma unroll 1
        for (int w = w1 - 2; w >= w0; w--) {
#pragma unroll 1
            for (int d = 0; d < kVarBits; d++) ge8_dbl(acc, acc, L);
            ge8_load(x, ws + (size_t)w * 128, L);
            ge8_add(acc, acc, x, L);
        }
        ge8_store(sh[warp], acc, L);
    }
    __syncthreads();
    if (warp != 0) return;
    ge8 acc, x;
    ge8_load(acc, sh[kFinGroups - 1], L);
#pragma unroll 1
    for (int g = kFinGroups - 2; g >= 0; g--) {
#pragma unroll 1
        for (int d = 0; d < kVarBits * kFinGroup; d++) ge8_dbl(acc, acc, L);
        ge8_load(x, sh[g], L);
        ge8_add(acc, acc, x, L);
    }
    // F == acc as projective points: X_F Z == X Z_F and Y_F Z == Y Z_F (thread-level, every lane the same values)
    ge_p3 V, F;
    fe8_gather(V.X, acc.X, L);
    fe8_gather(V.Y, acc.Y, L);
    fe8_gather(V.Z, acc.Z, L);
    ge_load(F, fsum + (size_t)id * 128);
    fe a, b, c, d;
    fe_mul(a, F.X, V.Z);
    fe_mul(b, V.X, F.Z);
    fe_mul(c, F.Y, V.Z);
    fe_mul(d, V.Y, F.Z);
    if (threadIdx.x == 0) flags[id] = (fe_equal(a, b) && fe_equal(c, d)) ? 1 : 0;
}
__global__ void verify_combine_kernel(const uint8_t* __restrict__ flags, uint32_t num, uint8_t* __restrict__ accept) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < num) accept[p] = flags[2 * p] & flags[2 * p + 1];
}


// ---- grouped verification: one combined identity per group of proofs ---------------------------------------------
// Both identities of every proof of a group are checked as ONE multi-scalar identity
//     sum_p [ w1_p (LHS1_p - RHS1_p) + w2_p (LHS2_p - RHS2_p) ] == 0
// with 128-bit weights w1_p, w2_p drawn by SHA-256 from a digest of ALL records of the group (Bulletproofs paper,
// section 6.2; in the random-oracle model a group that contains an invalid proof passes with probability ~2^-128).
// The 2n+2 shared generators then cost one fixed-base sum per GROUP instead of one per proof (their coefficients add up
// mod l), and the 255 doublings of the window combine run once per group; the 2 log n + 5 points of every proof still
// cost their window sums.  Accept bits stay per proof: malformed records are excluded from the sums and rejected on
// the spot, and every member of a group whose combined identity fails is verified again on its own — as a group of one,
// its two identities combined with its own weights — by a second pass of the same kernels over what the first pass
// left of it (transcript scalars, weights, point tables); that pass takes its count and indices from device memory.
// With 1 % of the proofs tampered and groups of 8, ~8 % of the proofs take the second pass.
struct VWeights {
    sc w1, w2;
};
// digest of one record, as a two-level hash tree so that the leaves run side by side:
// leaf_c = H("bpkVRFY0" || c || bytes [192 c, 192 c + 192)),  digest = H("bpkVRFY1" || leaf_0 || leaf_1 || ...)
static constexpr int kLeafWords = 48;  // 192 bytes per leaf; a record has 864 + 256 k bytes, at most 13 leaves
__global__ void __launch_bounds__(64) vg_leaf_kernel(const uint8_t* __restrict__ proofs, size_t rec_bytes, uint32_t num,
                                                     uint32_t nleaves, uint32_t* __restrict__ leaves) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= num * nleaves) return;
    const uint32_t p = id / nleaves, c = id % nleaves;
    const uint32_t* rec = reinterpret_cast<const uint32_t*>(proofs + (size_t)p * rec_bytes);
    const int words = (int)(rec_bytes / 4), off = (int)c * kLeafWords;
    Sha256 sh;
    sh.init();
    sh.update_str("bpkVRFY0", 8);
    sh.put((uint8_t)c);
    sh.put(0);
    sh.put(0);
    sh.put(0);
#pragma unroll 1
    for (int j = 0; j < 6 && off + 8 * j < words; j++) {
        uint32_t w[8];
#pragma unroll
        for (int i = 0; i < 8; i++) w[i] = rec[off + 8 * j + i];  // records are multiples of 32 bytes
        sh.update_words(w);
    }
    uint32_t d[8];
    sh.final_words(d);
#pragma unroll
    for (int i = 0; i < 8; i++) leaves[(size_t)id * 8 + i] = d[i];
}
// digest = chain over the leaf digests of the record (at most 7 per hash: the message buffer holds 256 bytes)
__global__ void __launch_bounds__(64) vg_digest_kernel(const uint32_t* __restrict__ leaves, uint32_t num, uint32_t nleaves,
                                                       uint32_t* __restrict__ digest) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num) return;
    uint32_t d[8];
#pragma unroll
    for (int i = 0; i < 8; i++) d[i] = 0;
    Sha256 sh;
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < nleaves; c0 += 6) {
        sh.init();
        sh.update_str("bpkVRFY1", 8);
        sh.update_words(d);
#pragma unroll 1
        for (uint32_t c = c0; c < c0 + 6 && c < nleaves; c++) {
            uint32_t w[8];
#pragma unroll
            for (int i = 0; i < 8; i++) w[i] = leaves[((size_t)p * nleaves + c) * 8 + i];
            sh.update_words(w);
        }
        sh.final_words(d);
    }
#pragma unroll
    for (int i = 0; i < 8; i++) digest[(size_t)p * 8 + i] = d[i];
}
// weights of proof p: seed = chain over the digests of its group, (w1, w2) = the two halves of H(seed || index)
__global__ void __launch_bounds__(64) vg_weights_kernel(const uint32_t* __restrict__ digest, uint32_t num, uint32_t K,
                                                        VWeights* __restrict__ wts) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num) return;
    const uint32_t g0 = p / K * K;
    uint32_t seed[8];
#pragma unroll
    for (int i = 0; i < 8; i++) seed[i] = 0;
    Sha256 sh;
#pragma unroll 1
    for (uint32_t j = g0; j < g0 + K && j < num; j++) {
        uint32_t d[8];
#pragma unroll
        for (int i = 0; i < 8; i++) d[i] = digest[(size_t)j * 8 + i];
        sh.init();
        sh.update_str("bpkVRFY2", 8);
        sh.update_words(seed);
        sh.update_words(d);
        sh.final_words(seed);
    }
    uint32_t h[8];
    sh.init();
    sh.update_str("bpkVRFY3", 8);
    sh.update_words(seed);
    sh.put((uint8_t)(p - g0));
    sh.final_words(h);
    VWeights w;
    sc_set0(w.w1);
    sc_set0(w.w2);
#pragma unroll
    for (int i = 0; i < 4; i++) {
        w.w1.v[i] = h[i];
        w.w2.v[i] = h[4 + i];
    }
    w.w1.v[0] |= 1u;  // never zero
    w.w2.v[0] |= 1u;
    wts[p] = w;
}
// coefficients, in two steps.  vg_coeff: one CTA per proof (the ladder of verify_coeff_kernel): its weighted coefficients
// of the shared generators go to memory as scalars (row t < n: G_t, n + t: H_t, 2n: g, 2n + 1: h), the weighted scalars
// of its own points are recoded.  vg_rowsum: thread per (group, row) adds the members' scalars mod l and recodes the sum.
// (A first version looped over the members inside one CTA per group: 0.72 ms per 2^14 proofs, the members' ladders
// in sequence.)
__global__ void __launch_bounds__(kCoeffThreads) vg_coeff_kernel(const VScal* __restrict__ vscal,
                                                                 const VWeights* __restrict__ wts, uint32_t n, int k,
                                                                 sc* __restrict__ cw, int8_t* __restrict__ vdigits) {
    __shared__ sc s_sh[kMaxN], y_sh[kMaxN];
    __shared__ sc wsc[4];  // w2 a, w2 b, w2 z, w2 z^2: the weight enters every G_t / H_t coefficient through these
    const uint32_t p = blockIdx.x;
    const int t = threadIdx.x;
    const VScal& vs = vscal[p];
    if (!vs.valid) return;  // whole CTA
    const int nvar2 = 2 + 2 * k, nvar = nvar2 + 3;
    const sc w1 = wts[p].w1, w2 = wts[p].w2;
    sc* row = cw + (size_t)p * kRowsMax;
    if (t == (int)n) {  // h: w1 taux + w2 (mu + ab - t)
        sc a, b;
        sc_mul_nf(a, w1, vs.h1);
        sc_mul_nf(b, w2, vs.h2);
        sc_add(a, a, b);
        row[2 * n + 1] = a;
    } else if (t == (int)n + 1) {  // g: w1 (t - delta)
        sc a;
        sc_mul_nf(a, w1, vs.g1);
        row[2 * n] = a;
    } else if (t == (int)n + 2) {  // under the ladder below; visible after its closing barrier
        sc a;
        sc_mul_nf(a, w2, vs.a);
        wsc[0] = a;
        sc_mul_nf(a, w2, vs.b);
        wsc[1] = a;
        sc_mul_nf(a, w2, vs.z);
        wsc[2] = a;
        sc_mul_nf(a, w2, vs.z2);
        wsc[3] = a;
    } else if (t >= (int)n + 3 && t - (int)n - 3 < nvar) {
        int q = t - (int)n - 3;
        sc sv;
        if (q == 0) sv = w2;
        else {
            if (q == 1) sv = vs.x;
            else if (q < 2 + k) sv = vs.usq[q - 2];
            else if (q < nvar2) sv = vs.uinvsq[q - 2 - k];
            else if (q == nvar2) sv = vs.z2;
            else if (q == nvar2 + 1) sv = vs.x;
            else sv = vs.x2;
            sc_mul_nf(sv, sv, q < nvar2 ? w2 : w1);
        }
        sc_recode_signed<kVarBits>(vdigits + ((size_t)p * kVarMax + q) * 64, sv, kVarWin);
    }
    if (t == 0) {
        s_sh[0] = vs.s0;
        sc one;
        sc_set1(one);
        y_sh[0] = one;
    }
    __syncwarp();
#pragma unroll 1
    for (int m = 0; m < k; m++) {
        const int half = 1 << m;
        if (half >= 32) __syncthreads();
        if (t >= half && t < 2 * half) {
            sc a = s_sh[t - half], b = y_sh[t - half];
            sc_mul_nf(a, a, vs.usq[k - 1 - m]);
            sc_mul_nf(b, b, vs.ypow[m]);
            s_sh[t] = a;
            y_sh[t] = b;
        }
        __syncwarp();
    }
    __syncthreads();
    if (t < (int)n) {
        const sc wa = wsc[0], wb = wsc[1], wz = wsc[2], wz2 = wsc[3];
        sc cg, ch, tmp, two_i;
        sc_mul_nf(cg, wa, s_sh[t]);
        sc_add(cg, cg, wz);
        sc_set0(two_i);
        two_i.v[t >> 5] = 1u << (t & 31);
        sc_mul_nf(tmp, wz2, two_i);
        sc_mul_nf(ch, wb, s_sh[n - 1 - t]);
        sc_sub(ch, ch, tmp);
        sc_mul_nf(ch, ch, y_sh[t]);
        sc_sub(ch, ch, wz);
        row[t] = cg;
        row[n + t] = ch;
    }
}
__global__ void __launch_bounds__(128) vg_rowsum_kernel(const uint8_t* __restrict__ gens, const VScal* __restrict__ vscal,
                                                        const sc* __restrict__ cw, uint32_t n, uint32_t K, uint32_t num,
                                                        uint32_t ngroups, int8_t* __restrict__ gdigits,
                                                        const uint32_t* __restrict__ dyn_groups = nullptr,
                                                        const uint32_t* __restrict__ index = nullptr) {
    const uint32_t nrows = 2 * n + 2;
    const uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (dyn_groups) ngroups = *dyn_groups;
    if (id >= ngroups * nrows) return;
    const uint32_t g = id / nrows, r = id % nrows;
    const int wbits = (int)reinterpret_cast<const GensHeader*>(gens)->wbits;
    sc sum;
    sc_set0(sum);
#pragma unroll 1
    for (uint32_t m = g * K; m < g * K + K && (index || m < num); m++) {
        const uint32_t p = index ? index[m] : m;
        if (!vscal[p].valid) continue;
        sc_add(sum, sum, cw[(size_t)p * kRowsMax + r]);
    }
    fix_recode(gdigits + ((size_t)g * kRowsMax + r) * kFixRowBytes, sum, wbits);
}
// fixed-base sum of a group: one CTA of 128 threads, lane = window, the 2n+2 rows dealt round-robin to the 128 / LP
// row slices (16-bit tables: 8 slices of 16 lanes); warp butterflies, then the four warp sums through shared memory
template <int WBITS>
__global__ void __launch_bounds__(128) vg_fixed_kernel(const uint8_t* __restrict__ gens, const int8_t* __restrict__ gdigits,
                                                       uint32_t n, uint8_t* __restrict__ gfsum,
                                                       const uint32_t* __restrict__ dyn_groups = nullptr) {
    if (dyn_groups && blockIdx.x >= *dyn_groups) return;  // whole CTA
    __shared__ __align__(16) uint8_t sh[4][128];
    constexpr int LP = 256 / WBITS;
    constexpr int S = 128 / LP;
    constexpr uint32_t E = 1u << (WBITS - 1);
    const uint32_t g = blockIdx.x;
    const int win = threadIdx.x & (LP - 1), slice = threadIdx.x / LP;
    const GensHeader* gh = reinterpret_cast<const GensHeader*>(gens);
    const uint8_t* table = gens + gh->table_off + (size_t)win * E * 96;
    const int8_t* drow = gdigits + (size_t)g * kRowsMax * kFixRowBytes;
    const int nrows = 2 * (int)n + 2;
    auto digit_of = [&](int row, uint32_t& mag, bool& neg) {
        if (WBITS == 8) fixed_digit(drow[(size_t)row * kFixRowBytes + win], mag, neg);
        else fixed_digit16(reinterpret_cast<const int16_t*>(drow + (size_t)row * kFixRowBytes)[win], mag, neg);
    };
    ge_p3 acc;
    ge_p3_0(acc);
    uint32_t mag = 0;
    bool neg = false;
    ge_niels q;
    if (slice < nrows) {
        digit_of(slice, mag, neg);
        if (mag) ge_niels_load(q, table + ((size_t)slice * LP * E + (mag - 1)) * 96);
    }
#pragma unroll 1
    for (int row = slice; row < nrows; row += S) {
        const uint32_t cmag = mag;
        const bool cneg = neg;
        const ge_niels cur = q;
        if (row + S < nrows) {
            digit_of(row + S, mag, neg);
            if (mag) ge_niels_load(q, table + ((size_t)(row + S) * LP * E + (mag - 1)) * 96);
        }
        if (cmag) ge_madd(acc, acc, cur, cneg);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_p3 other;
        ge_shfl_xor(other, acc, o);
        ge_add(acc, acc, other);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) ge_store(sh[warp], acc);
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll 1
        for (int w = 1; w < 4; w++) {
            ge_p3 other;
            ge_load(other, sh[w]);
            ge_add(acc, acc, other);
        }
        ge_store(gfsum + (size_t)g * 128, acc);
    }
}
// window sums of a group: thread per (group, window) over the points of all its valid members
// (measured and dropped: requesting the next operand — a 128-byte gather from the 570 MB of tables — before the current
// addition: 157 instead of 128 registers, 3.34 instead of 3.12 ms per 2^14 proofs in groups of 12)
__global__ void __launch_bounds__(128) vg_winsum_kernel(const VScal* __restrict__ vscal, const int8_t* __restrict__ vdigits,
                                                        const uint8_t* __restrict__ vtab, int k, uint32_t K, uint32_t num,
                                                        uint32_t ngroups, uint8_t* __restrict__ gwinsum,
                                                        const uint32_t* __restrict__ dyn_groups = nullptr,
                                                        const uint32_t* __restrict__ index = nullptr) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (dyn_groups) ngroups = *dyn_groups;
    if (id >= ngroups * kVarWin) return;
    const uint32_t g = id / kVarWin;
    const int w = (int)(id % kVarWin);
    const int nvar = 2 + 2 * k + 3;
    ge_p3 ws;
    ge_p3_0(ws);
#pragma unroll 1
    for (uint32_t m = g * K; m < g * K + K && (index || m < num); m++) {
        const uint32_t p = index ? index[m] : m;
        if (!vscal[p].valid) continue;
#pragma unroll 1
        for (int q = 0; q < nvar; q++) {
            int d = vdigits[((size_t)p * kVarMax + q) * 64 + w];
            if (d != 0) {
                int mag = d < 0 ? -d : d;
                const uint8_t* src = vtab + (((size_t)p * kVarMax + q) * kVarEntries + (mag - 1)) * 128;
                ge_cached c;
                fe_load(c.YplusX, src);
                fe_load(c.YminusX, src + 32);
                fe_load(c.Z2, src + 64);
                fe_load(c.T2d, src + 96);
                ge_add_cached(ws, ws, c, d < 0);
            }
        }
    }
    ge_store(gwinsum + ((size_t)g * kVarWin + w) * 128, ws);
}
// Horner over the window sums of a group and the comparison with its fixed-base sum: one warp per group in octet
// form (255 doublings + 51 additions deep at 0.63 / 0.82 us; all groups of a pass are resident at once)
__global__ void __launch_bounds__(128) vg_finish_kernel(const uint8_t* __restrict__ gfsum, const uint8_t* __restrict__ gwinsum,
                                                        uint32_t ngroups, uint8_t* __restrict__ gflags,
                                                        const uint32_t* __restrict__ dyn_groups = nullptr) {
    const uint32_t g = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (dyn_groups) ngroups = *dyn_groups;
    if (g >= ngroups) return;  // whole warp
    const Fe8Lane L = fe8_lane();
    const uint8_t* ws = gwinsum + (size_t)g * kVarWin * 128;
    ge8 acc, x;
    ge8_load(acc, ws + (size_t)(kVarWin - 1) * 128, L);
#pragma unroll 1
    for (int w = kVarWin - 2; w >= 0; w--) {
#pragma unroll 1
        for (int d = 0; d < kVarBits; d++) ge8_dbl(acc, acc, L);
        ge8_load(x, ws + (size_t)w * 128, L);
        ge8_add(acc, acc, x, L);
    }
    ge_p3 V, F;
    fe8_gather(V.X, acc.X, L);
    fe8_gather(V.Y, acc.Y, L);
    fe8_gather(V.Z, acc.Z, L);
    ge_load(F, gfsum + (size_t)g * 128);
    fe a, b, c, d;
    fe_mul(a, F.X, V.Z);
    fe_mul(b, V.X, F.Z);
    fe_mul(c, F.Y, V.Z);
    fe_mul(d, V.Y, F.Z);
    if ((threadIdx.x & 31) == 0) gflags[g] = (fe_equal(a, b) && fe_equal(c, d)) ? 1 : 0;
}
// per proof: rejected (malformed), accepted with its group, or queued for the second pass
__global__ void vg_decide_kernel(const VScal* __restrict__ vscal, const uint8_t* __restrict__ gflags, uint32_t K,
                                 uint32_t num, uint8_t* __restrict__ accept, uint32_t* __restrict__ fb_count,
                                 uint32_t* __restrict__ fb_index) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num) return;
    if (!vscal[p].valid) {
        accept[p] = 0;
    } else if (gflags[p / K]) {
        accept[p] = 1;
    } else {
        accept[p] = 0;
        fb_index[atomicAdd(fb_count, 1u)] = p;
    }
}
// second pass: the verdict of queue slot s is the verdict of proof fb_index[s]
__global__ void vg_decide_single_kernel(const uint8_t* __restrict__ gflags, const uint32_t* __restrict__ fb_count,
                                        const uint32_t* __restrict__ fb_index, uint8_t* __restrict__ accept) {
    uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s < *fb_count) accept[fb_index[s]] = gflags[s];
}

static constexpr uint32_t kVerifySmallBatch = 64;  // up to this many proofs per call take the latency kernels
static constexpr size_t kVerifyChunk = 16384;  // proofs per pass (~62 KiB of scratch per proof); large passes keep the thread-per-proof kernels busy
static size_t align256(size_t x) { return (x + 255) / 256 * 256; }

}  // namespace cbp

using namespace cbp;

extern "C" {

size_t bpk_proof_record_bytes(size_t n) {
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    return proof_record_bytes(k);
}
static bool wbits_ok(int wbits) { return wbits == 8 || wbits == 16; }
int bpk_gens_workspace_bytes_ex(size_t n, int window_bits, size_t* bytes) {
    if (!bytes || n == 0 || n > kMaxN || (n & (n - 1)) || !wbits_ok(window_bits)) return fail(BPK_ERR_ARG);
    size_t nb = 2 * n + 2, nwin = (size_t)fix_nwin(window_bits);
    size_t table = nb * nwin * fix_entries(window_bits) * 96;
    size_t scratch = nb * nwin * 128;
    *bytes = 256 + align256(table) + align256(scratch) + align256(nb * 128);
    return BPK_OK;
}
int bpk_gens_workspace_bytes(size_t n, size_t* bytes) { return bpk_gens_workspace_bytes_ex(n, 8, bytes); }

// window width of every table built by this process, by device address (the verifier sizes its launch
// from it without reading the device-side header back)
static std::mutex g_gens_mu;
static std::unordered_map<const void*, int> g_gens_wbits;
static int gens_wbits_of(const void* d_gens_ws) {
    std::lock_guard<std::mutex> lk(g_gens_mu);
    auto it = g_gens_wbits.find(d_gens_ws);
    return it == g_gens_wbits.end() ? 0 : it->second;
}
int bpk_gens_window_bits(const void* d_gens_ws) { return gens_wbits_of(d_gens_ws); }

int bpk_gens_init_device_ex(void* d_gens_ws, size_t ws_bytes, const void* d_G, const void* d_H, const void* d_g,
                            const void* d_h, size_t n, int window_bits, void* stream) {
    size_t need = 0;
    if (bpk_gens_workspace_bytes_ex(n, window_bits, &need) != BPK_OK) return BPK_ERR_ARG;
    if (!d_gens_ws || !d_G || !d_H || !d_g || !d_h) return fail(BPK_ERR_ARG);
    if (ws_bytes < need) return fail(BPK_ERR_WORKSPACE);
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t nb = (uint32_t)(2 * n + 2);
    const int nwin = fix_nwin(window_bits);
    const uint32_t entries = fix_entries(window_bits);
    GensHeader hd;
    memset(&hd, 0, sizeof hd);
    hd.magic = kGensMagic;
    hd.n = (uint32_t)n;
    hd.nbases = nb;
    hd.wbits = (uint32_t)window_bits;
    hd.nwin = (uint32_t)nwin;
    hd.entries = entries;
    hd.table_off = 256;
    hd.scratch_off = hd.table_off + align256((size_t)nb * nwin * entries * 96);
    hd.bases_off = hd.scratch_off + align256((size_t)nb * nwin * 128);
    uint8_t* ws = (uint8_t*)d_gens_ws;
    CBP_CUDA(cudaMemcpyAsync(ws, &hd, sizeof hd, cudaMemcpyHostToDevice, st));
    uint8_t* pow = ws + hd.scratch_off;
    gens_pow_kernel<<<(nb + 63) / 64, 64, 0, st>>>((const uint8_t*)d_G, (const uint8_t*)d_H, (const uint8_t*)d_g,
                                                  (const uint8_t*)d_h, (uint32_t)n, window_bits, pow,
                                                  ws + hd.bases_off);
    CBP_CHECK_LAUNCH();
    size_t threads = (size_t)nb * nwin * (entries / kTabRun);
    gens_table_kernel<<<(unsigned)((threads + 63) / 64), 64, 0, st>>>(pow, nb, nwin, entries, ws + hd.table_off);
    CBP_CHECK_LAUNCH();
    {
        std::lock_guard<std::mutex> lk(g_gens_mu);
        g_gens_wbits[d_gens_ws] = window_bits;
    }
    return BPK_OK;
}
int bpk_gens_init_device(void* d_gens_ws, size_t ws_bytes, const void* d_G, const void* d_H, const void* d_g,
                         const void* d_h, size_t n, void* stream) {
    return bpk_gens_init_device_ex(d_gens_ws, ws_bytes, d_G, d_H, d_g, d_h, n, 8, stream);
}

struct VerifyLayout {
    size_t vscal, fsum, winsum, flags, digits, vdigits, vtab, vseed;
    // grouped verification: record digests, weights, per-group digit rows / sums / flags, second-pass queue and copies
    size_t leaves, digest, wts, cw, gdigits, gfsum, gwinsum, gflags, fb_count, fb_index;
    size_t total;
};
static constexpr uint32_t kVerifyGroupMin = 2;     // smallest group the grouped path is used with
static constexpr uint32_t kVerifyGroupDefault = 12;  // measured at 2^14 proofs, 1 % tampered: 4 / 8 / 12 / 16 -> 4.41 / 3.78 / 3.63 / 3.77 ms
static constexpr uint32_t kVerifyGroupBatchMin = 256;  // below this many proofs per call: one by one
// Measured and dropped: cutting a grouped pass into 2 / 3 / 4 sub-batches on prioritised streams of their own, so that
// the latency-bound phases of one (transcript hashes, the 255-doubling chains of the finish) run under the window
// sums of another: 3.54 / 3.66 / 3.73 ms against 3.11 ms per 2^14 proofs in one piece — the latency-bound kernels take
// as long for a quarter of the batch, and four times the launches cost the host more than the overlap returns.
static VerifyLayout verify_layout(size_t chunk, size_t rec_bytes) {  // rec_bytes sizes the leaf digests
    VerifyLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align256(bytes);
        return o;
    };
    L.vscal = take(chunk * sizeof(VScal));
    L.fsum = take(chunk * 2 * 128);
    L.winsum = take(chunk * 2 * kVarWin * 128);
    L.flags = take(chunk * 2);
    L.digits = take(chunk * kRowsMax * kFixRowBytes);
    L.vdigits = take(chunk * kVarMax * 64);
    L.vtab = take(chunk * kVarMax * kVarEntries * 128);
    L.vseed = take(chunk * 2 * kVarEntries * 128);
    const size_t groups = chunk;  // the second pass treats every queued proof as a group of one
    L.leaves = take(chunk * ((rec_bytes + 191) / 192) * 32);
    L.digest = take(chunk * 32);
    L.wts = take(chunk * sizeof(VWeights));
    L.cw = take(chunk * kRowsMax * sizeof(sc));
    L.gdigits = take(groups * kRowsMax * kFixRowBytes);
    L.gfsum = take(groups * 128);
    L.gwinsum = take(groups * kVarWin * 128);
    L.gflags = take(groups);
    L.fb_count = take(4);
    L.fb_index = take(chunk * 4);
    L.total = off;
    return L;
}
int bpk_range_verify_workspace_bytes(size_t n, size_t num_proofs, size_t* bytes) {
    if (!bytes || n == 0 || n > kMaxN || (n & (n - 1))) return fail(BPK_ERR_ARG);
    size_t chunk = num_proofs < kVerifyChunk ? num_proofs : kVerifyChunk;
    if (chunk == 0) chunk = 1;
    *bytes = verify_layout(chunk, bpk_proof_record_bytes(n)).total;
    return BPK_OK;
}
// one side stream per device for the work that does not depend on the transcript
struct VerifySide {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_weights = nullptr;
    bool ok = false;
};
static VerifySide g_vside[kMaxDevices];
// caller holds the device lock
static VerifySide* verify_side(int dev) {
    if (dev < 0 || dev >= kMaxDevices) return nullptr;
    VerifySide& v = g_vside[dev];
    if (!v.ok) {
        if (cudaStreamCreateWithFlags(&v.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&v.ev_fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&v.ev_join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&v.ev_weights, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        v.ok = true;
    }
    return &v;
}

int bpk_range_verify_batch_device(const void* d_gens_ws, const void* d_proofs, const void* d_V, size_t n,
                                  size_t num_proofs, uint8_t* d_accept, void* d_workspace, size_t workspace_bytes,
                                  void* stream) {
    size_t need = 0;
    if (bpk_range_verify_workspace_bytes(n, num_proofs, &need) != BPK_OK) return BPK_ERR_ARG;
    if (!num_proofs) return BPK_OK;
    if (!d_gens_ws || !d_proofs || !d_accept || !d_workspace) return fail(BPK_ERR_ARG);
    if (workspace_bytes < need) return fail(BPK_ERR_WORKSPACE);
    const int wbits = gens_wbits_of(d_gens_ws);
    if (!wbits) return fail(BPK_ERR_ARG);  // not a table built by bpk_gens_init_device[_ex]
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    size_t rec = proof_record_bytes(k);
    size_t chunk = num_proofs < kVerifyChunk ? num_proofs : kVerifyChunk;
    VerifyLayout L = verify_layout(chunk, rec);
    uint8_t* ws = (uint8_t*)d_workspace;
    VScal* vscal = (VScal*)(ws + L.vscal);
    uint8_t *fsum = ws + L.fsum, *winsum = ws + L.winsum, *flags = ws + L.flags, *vtab = ws + L.vtab;
    uint8_t* vseed = ws + L.vseed;
    int8_t *digits = (int8_t*)(ws + L.digits), *vdigits = (int8_t*)(ws + L.vdigits);
    const int nvar = 2 + 2 * k + 3;
    cudaStream_t st = (cudaStream_t)stream;
    DeviceLock dlock;  // the side stream and its events are shared by every call on this device
    if (!dlock.ok()) return BPK_ERR_CUDA;
    VerifySide* side = verify_side(dlock.dev);
    if (!side) return fail(BPK_ERR_CUDA);
    prof_begin(BPK_PROF_VERIFY_TOTAL, st);
    for (size_t done = 0; done < num_proofs; done += chunk) {
        uint32_t cnt = (uint32_t)((num_proofs - done) < chunk ? (num_proofs - done) : chunk);
        const uint8_t* pr = (const uint8_t*)d_proofs + done * rec;
        const uint8_t* ve = d_V ? (const uint8_t*)d_V + done * 128 : nullptr;
        // side stream: the per-proof point tables need only the records, so they are built while the
        // latency-bound transcript / coefficient kernels run.  (Also moving the window sums there, next to the
        // coefficient and fixed-base kernels, was measured: 6.43 vs 6.56 ms — both sides want the same pipe.)
        cudaStream_t ss = side->stream;
        CBP_CUDA(cudaEventRecord(side->ev_fork, st));
        CBP_CUDA(cudaStreamWaitEvent(ss, side->ev_fork, 0));
        verify_transcript_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(pr, rec, ve, (uint32_t)n, k, cnt, vscal);
        CBP_CHECK_LAUNCH();
        // grouped verification (block comment above vg_leaf_kernel)
        uint32_t K = options().verify_group < 0 ? (cnt >= kVerifyGroupBatchMin ? kVerifyGroupDefault : 0u)
                                                : (uint32_t)options().verify_group;
        if (K > 64) K = 64;
        if (K < kVerifyGroupMin) {
            verify_vtab_kernel<<<(cnt * nvar + 127) / 128, 128, 0, ss>>>(pr, rec, k, cnt, vtab, vseed);
            CBP_CHECK_LAUNCH();
        } else {
            uint32_t *leaves = (uint32_t*)(ws + L.leaves), *digest = (uint32_t*)(ws + L.digest);
            VWeights* wts = (VWeights*)(ws + L.wts);
            int8_t* gdigits = (int8_t*)(ws + L.gdigits);
            sc* cw = (sc*)(ws + L.cw);
            const uint32_t nrows = (uint32_t)(2 * n + 2);
            uint8_t *gfsum = ws + L.gfsum, *gwinsum = ws + L.gwinsum, *gflags = ws + L.gflags;
            uint32_t *fb_count = (uint32_t*)(ws + L.fb_count), *fb_index = (uint32_t*)(ws + L.fb_index);
            const uint32_t G = (cnt + K - 1) / K;
            // side stream, under the transcript: record digests and weights (the coefficients wait for them), then
            // the tables of the per-proof points (the window sums wait for those); all need only the records
            const uint32_t nleaves = (uint32_t)((rec + 191) / 192);
            vg_leaf_kernel<<<(cnt * nleaves + 63) / 64, 64, 0, ss>>>(pr, rec, cnt, nleaves, leaves);
            CBP_CHECK_LAUNCH();
            vg_digest_kernel<<<(cnt + 63) / 64, 64, 0, ss>>>(leaves, cnt, nleaves, digest);
            CBP_CHECK_LAUNCH();
            vg_weights_kernel<<<(cnt + 63) / 64, 64, 0, ss>>>(digest, cnt, K, wts);
            CBP_CHECK_LAUNCH();
            CBP_CUDA(cudaEventRecord(side->ev_weights, ss));
            verify_vtab_kernel<<<(cnt * nvar + 127) / 128, 128, 0, ss>>>(pr, rec, k, cnt, vtab, vseed);
            CBP_CHECK_LAUNCH();
            CBP_CUDA(cudaEventRecord(side->ev_join, ss));
            CBP_CUDA(cudaMemsetAsync(fb_count, 0, 4, st));
            CBP_CUDA(cudaStreamWaitEvent(st, side->ev_weights, 0));
            vg_coeff_kernel<<<cnt, kCoeffThreads, 0, st>>>(vscal, wts, (uint32_t)n, k, cw, vdigits);
            CBP_CHECK_LAUNCH();
            vg_rowsum_kernel<<<(G * nrows + 127) / 128, 128, 0, st>>>((const uint8_t*)d_gens_ws, vscal, cw, (uint32_t)n, K, cnt, G,
                                                                      gdigits);
            CBP_CHECK_LAUNCH();
            CBP_CUDA(cudaStreamWaitEvent(st, side->ev_join, 0));
            // the groups' fixed-base sums (few, latency-bound CTAs) on the side stream, under the window sums
            CBP_CUDA(cudaEventRecord(side->ev_fork, st));
            CBP_CUDA(cudaStreamWaitEvent(ss, side->ev_fork, 0));
            if (wbits == 8) vg_fixed_kernel<8><<<G, 128, 0, ss>>>((const uint8_t*)d_gens_ws, gdigits, (uint32_t)n, gfsum);
            else vg_fixed_kernel<16><<<G, 128, 0, ss>>>((const uint8_t*)d_gens_ws, gdigits, (uint32_t)n, gfsum);
            CBP_CHECK_LAUNCH();
            CBP_CUDA(cudaEventRecord(side->ev_join, ss));
            prof_begin(BPK_PROF_VERIFY_MSM, st);
            vg_winsum_kernel<<<(G * kVarWin + 127) / 128, 128, 0, st>>>(vscal, vdigits, vtab, k, K, cnt, G, gwinsum);
            prof_end(BPK_PROF_VERIFY_MSM, st);
            CBP_CHECK_LAUNCH();
            CBP_CUDA(cudaStreamWaitEvent(st, side->ev_join, 0));
            vg_finish_kernel<<<(G * 32 + 127) / 128, 128, 0, st>>>(gfsum, gwinsum, G, gflags);
            CBP_CHECK_LAUNCH();
            vg_decide_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(vscal, gflags, K, cnt, d_accept + done, fb_count, fb_index);
            CBP_CHECK_LAUNCH();
            // second pass: every member of a failed group as a group of ONE (its two identities combined with its own
            // weights), on what the first pass left of it — transcript scalars, weights, point tables; count and indices
            // on the device, grids sized for the whole batch, slots beyond the count exit at once
            // (their coefficient scalars cw[] and weighted digit rows are those of the first pass)
            vg_rowsum_kernel<<<(cnt * nrows + 127) / 128, 128, 0, st>>>((const uint8_t*)d_gens_ws, vscal, cw, (uint32_t)n, 1, cnt,
                                                                        cnt, gdigits, fb_count, fb_index);
            CBP_CHECK_LAUNCH();
            if (wbits == 8) vg_fixed_kernel<8><<<cnt, 128, 0, st>>>((const uint8_t*)d_gens_ws, gdigits, (uint32_t)n, gfsum, fb_count);
            else vg_fixed_kernel<16><<<cnt, 128, 0, st>>>((const uint8_t*)d_gens_ws, gdigits, (uint32_t)n, gfsum, fb_count);
            CBP_CHECK_LAUNCH();
            vg_winsum_kernel<<<(cnt * kVarWin + 127) / 128, 128, 0, st>>>(vscal, vdigits, vtab, k, 1, cnt, cnt, gwinsum, fb_count,
                                                                          fb_index);
            CBP_CHECK_LAUNCH();
            vg_finish_kernel<<<(cnt * 32 + 127) / 128, 128, 0, st>>>(gfsum, gwinsum, cnt, gflags, fb_count);
            CBP_CHECK_LAUNCH();
            vg_decide_single_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(gflags, fb_count, fb_index, d_accept + done);
            CBP_CHECK_LAUNCH();
            continue;
        }
        CBP_CUDA(cudaEventRecord(side->ev_join, ss));
        verify_coeff_kernel<<<cnt, kCoeffThreads, 0, st>>>((const uint8_t*)d_gens_ws, vscal, (uint32_t)n, k, digits,
                                                           vdigits);
        CBP_CHECK_LAUNCH();
        const bool small = cnt <= kVerifySmallBatch;  // latency path: a CTA per proof / per identity
        prof_begin(BPK_PROF_VERIFY_MSM, st);
        if (small)
            verify_fixed_small_kernel<<<cnt, 1024, 0, st>>>((const uint8_t*)d_gens_ws, vscal, digits, (uint32_t)n, fsum);
        else if (wbits == 8)
            verify_fixed_kernel<8><<<(cnt * 32 + 127) / 128, 128, 0, st>>>((const uint8_t*)d_gens_ws, vscal, digits,
                                                                           (uint32_t)n, cnt, fsum);
        else
            verify_fixed_kernel<16><<<(cnt * 16 + 127) / 128, 128, 0, st>>>((const uint8_t*)d_gens_ws, vscal, digits,
                                                                            (uint32_t)n, cnt, fsum);
        prof_end(BPK_PROF_VERIFY_MSM, st);
        CBP_CHECK_LAUNCH();
        CBP_CUDA(cudaStreamWaitEvent(st, side->ev_join, 0));
        verify_winsum_kernel<<<(cnt * 2 * kVarWin + 127) / 128, 128, 0, st>>>(vscal, vdigits, vtab, vseed, k, cnt, winsum);
        CBP_CHECK_LAUNCH();
        if (small) verify_finish_small_kernel<<<cnt * 2, 32 * kFinGroups, 0, st>>>(vscal, fsum, winsum, flags);
        else verify_finish_kernel<<<(cnt * 2 + 63) / 64, 64, 0, st>>>(vscal, fsum, winsum, cnt, flags);
        CBP_CHECK_LAUNCH();
        verify_combine_kernel<<<(cnt + 255) / 256, 256, 0, st>>>(flags, cnt, d_accept + done);
        CBP_CHECK_LAUNCH();
    }
    prof_end(BPK_PROF_VERIFY_TOTAL, st);
    return BPK_OK;
}

}  // extern "C"
