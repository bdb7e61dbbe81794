// prover.cu — batched range-proof generation for sm_100a (one CTA per proof).
//
// Restates generate_range_proof (bulletproof_range_proof.cu:1159-1714) and inner_product_prove
// (bulletproof_vectors.cu:277-538) with the corrections of oracle/ref_corrected.c; given the same
// seed it emits byte-identical proofs.  Used to synthesise benchmark inputs and as the on-device
// prover for the IPA folding path.
//
// B200-first restructuring: the reference folds the generator vectors every round with 4 n' CPU
// scalar multiplications (bulletproof_vectors.cu:641-663) and takes 4 MSMs of size n' over the folded
// points (:390-446).  Here G and H are never folded as points: round r's L and R are written over the
// ORIGINAL generators with composite scalars  a_j * w_i  (w_i = prod_{q<r} u_q^(+-1) is the folding
// weight of generator i), so every point operation is a doubling-free mixed addition out of the
// shared 8-bit fixed-base tables, and the only per-round vector work is the mod-l a/b fold and a
// weight update — all resident in shared memory.
#include <stdio.h>
#include <stdlib.h>
#include "common.h"
#include "rangeproof.cuh"
#include "sha256.cuh"

namespace cbp {

// Out-of-line copies: the prover kernel is one long straight-line program per proof (31 scalar-multiplication
// sites, 14 fixed-base sums); fully inlined it was 870 KB of SASS and bound by instruction fetch.
static __device__ __noinline__ void sc_mul_nf(sc& r, const sc& a, const sc& b) { sc_mul(r, a, b); }
// normalise up to three points with ONE field inversion (Montgomery); same canonical results as ge_normalize
static __device__ __noinline__ void ge_normalize_many(ge_p3* pts, int cnt) {
    fe z01, z012, inv, zi[3];
    if (cnt == 1) {
        fe_invert(zi[0], pts[0].Z);
    } else {
        fe_mul(z01, pts[0].Z, pts[1].Z);
        if (cnt == 3) fe_mul(z012, z01, pts[2].Z);
        else z012 = z01;
        fe_invert(inv, z012);
        if (cnt == 3) {
            fe_mul(zi[2], inv, z01);
            fe_mul(inv, inv, pts[2].Z);
        }
        fe_mul(zi[1], inv, pts[0].Z);
        fe_mul(zi[0], inv, pts[1].Z);
    }
    for (int i = 0; i < cnt; i++) {
        fe x, y;
        fe_mul(x, pts[i].X, zi[i]);
        fe_mul(y, pts[i].Y, zi[i]);
        fe_canon(x);
        fe_canon(y);
        pts[i].X = x;
        pts[i].Y = y;
        fe_set1(pts[i].Z);
        fe_mul(pts[i].T, x, y);
        fe_canon(pts[i].T);
    }
}

static constexpr int kPThreads = 64;  // = kMaxN: one thread per vector element; small CTAs keep more proofs per SM in
                                      // flight, which is what hides the serial sections (hashes, inversions)

__device__ __forceinline__ uint64_t sm64_at(uint64_t seed, uint64_t idx) {
    uint64_t z = seed + (idx + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// Where a proof's blinding values and nonces (alpha, rho, tau1, tau2, sL, sR) come from:
//  * seeds: the SplitMix64 stream oracle/ref_corrected.c draws from, keyed by a 64-bit seed.  NOT cryptographic
//    (invertible, 64 bits of state): for parity tests and benchmark inputs only.
//  * keys:  one 32-byte secret per proof from the caller's CSPRNG (the reference draws every value from OpenSSL
//    RAND_bytes, bulletproof_range_proof.cu:153); draw j = SHA-256("cbp-bp-nonce" || key || j_le32), i.e. SHA-256
//    in counter mode as the PRF, so hiding rests on the 256-bit key.
struct NonceSource {
    const uint64_t* seeds;
    const uint8_t* keys;
};
struct Nonce {
    uint64_t seed;
    const uint8_t* key;  // 32 bytes, or nullptr: seeded stream
};
__device__ __forceinline__ Nonce nonce_of(const NonceSource& src, uint32_t proof) {
    Nonce nn;
    nn.seed = src.keys ? 0 : src.seeds[proof];
    nn.key = src.keys ? src.keys + (size_t)proof * 32 : nullptr;
    return nn;
}
static __device__ __noinline__ void keyed_draw(uint32_t (&out)[8], const uint8_t* key, uint32_t j) {
    Sha256 sh;
    sh.init();
    sh.update_str("cbp-bp-nonce", 12);
    sh.update(key, 32);
    const uint8_t ctr[4] = {(uint8_t)j, (uint8_t)(j >> 8), (uint8_t)(j >> 16), (uint8_t)(j >> 24)};
    sh.update(ctr, 4);
    uint32_t h[8];
    sh.final_words(h);
#pragma unroll
    for (int i = 0; i < 8; i++) out[i] = __byte_perm(h[i], 0, 0x0123);  // digest bytes as a little-endian integer
}
// the j-th 32-byte draw of the stream, clamped like generate_random_scalar
// (bulletproof_range_proof.cu:153-159), then reduced mod l
__device__ __forceinline__ void draw_scalar(sc& r, const Nonce& nn, uint32_t j) {
    sc t;
    if (nn.key) {
        keyed_draw(t.v, nn.key, j);
    } else {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            uint64_t v = sm64_at(nn.seed, (uint64_t)j * 4 + q);
            t.v[2 * q] = (uint32_t)v;
            t.v[2 * q + 1] = (uint32_t)(v >> 32);
        }
    }
    t.v[7] &= 0x7FFFFFFFu;
    t.v[0] &= 0xFFFFFFF8u;
    t.v[7] |= 0x40000000u;
    sc_reduce(r, t);
}
__device__ __forceinline__ void cta_sc_sum(sc& v, sc* sred) {
    int t = threadIdx.x;
    sred[t] = v;
    __syncthreads();
    for (int o = kPThreads >> 1; o > 0; o >>= 1) {
        if (t < o) {
            sc a = sred[t];
            sc_add(a, a, sred[t + o]);
            sred[t] = a;
        }
        __syncthreads();
    }
    v = sred[0];
    __syncthreads();
}
// sum over all rows of digits * table, NOT normalised, valid in thread 0
static __device__ __noinline__ void cta_fixed_msm(ge_p3& result, const FixTab& table, int nrows,
                                              int8_t (*digits)[kFixRowBytes], ge_p3* red) {
    ge_p3 acc;
    ge_p3_0(acc);
    for (int item = threadIdx.x; item < nrows * table.nwin; item += kPThreads) {
        int row = item / table.nwin, win = item % table.nwin;
        fixed_base_madd(acc, table, (uint32_t)row, win, digits[row]);
    }
    cta_point_sum(acc, red);
    result = acc;
}
__device__ __forceinline__ void zero_row(int8_t* row) {
#pragma unroll
    for (int i = 0; i < kFixRowBytes; i++) row[i] = 0;
}
__device__ __forceinline__ void hash_xy(Sha256& sh, const ge_p3& P) {  // P normalised, canonical
    sh.update_words(P.X.v);
    sh.update_words(P.Y.v);
}

__global__ void __launch_bounds__(kPThreads, 8) range_prove_kernel(const uint8_t* __restrict__ gens,
                                                                const uint64_t* __restrict__ values,
                                                                const uint8_t* __restrict__ gammas,
                                                                NonceSource nsrc, uint32_t n, int k,
                                                                uint8_t* __restrict__ proofs, size_t rec_bytes) {
    __shared__ sc sa[kMaxN], sb[kMaxN], swG[kMaxN], swH[kMaxN];
    __shared__ sc sred[kPThreads];
    __shared__ __align__(16) int8_t digits[2 * kMaxN + 2][kFixRowBytes];
    __shared__ ge_p3 red[kPThreads];
    __shared__ sc sh_z, sh_x, sh_u, sh_uinv;
    __shared__ sc sh_ypow[kMaxK + 1], sh_yinvpow[kMaxK + 1];
    __shared__ uint32_t sh_tr[8];
    __shared__ ge_p3 sh_pts[3];

    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    const FixTab table = fixtab_of(gens);
    uint8_t* rec = proofs + (size_t)p * rec_bytes;
    const uint64_t v = values[p];
    const Nonce seed = nonce_of(nsrc, p);
    const int nrows = 2 * (int)n + 2, row_g = 2 * (int)n, row_h = 2 * (int)n + 1;

    if (n < 64 && (v >> n) != 0) {  // validate_range_input (:238-263): initialised, invalid proof (D20)
        ge_p3 O;
        ge_p3_0(O);
        for (int q = t; q < 5 + 2 * k; q += kPThreads) ge_store(rec + (q < 5 ? q * 128 : kRecL + (q - 5) * 128), O);
        if (t < 7) {
            fe zf;
            fe_set0(zf);
            fe_store(rec + kRecTaux + t * 32, zf);
        }
        return;
    }
    sc gamma, alpha, rho, tau1, tau2, aL, aR, sL, sR, one;
    sc_set1(one);
    {
        sc g0;
        sc_load(g0, gammas + (size_t)p * 32);
        sc_reduce(gamma, g0);
    }
    draw_scalar(alpha, seed, 2 * n);
    draw_scalar(rho, seed, 2 * n + 1);
    draw_scalar(tau1, seed, 2 * n + 2);
    draw_scalar(tau2, seed, 2 * n + 3);
    sc_set0(aL);
    sc_set0(aR);
    sc_set0(sL);
    sc_set0(sR);
    if (t < (int)n) {
        if ((v >> t) & 1) sc_set1(aL);
        sc_sub(aR, aL, one);
        draw_scalar(sL, seed, 2 * t);
        draw_scalar(sR, seed, 2 * t + 1);
    }
    ge_p3 P;
    // ---- V = v g + gamma h (pedersen_commit, :277-296) ----
    for (int r = t; r < nrows; r += kPThreads) zero_row(digits[r]);
    __syncthreads();
    if (t == 0) {
        sc vs;
        sc_set0(vs);
        vs.v[0] = (uint32_t)v;
        vs.v[1] = (uint32_t)(v >> 32);
        fix_recode(digits[row_g], vs, table.wbits);
        fix_recode(digits[row_h], gamma, table.wbits);
    }
    __syncthreads();
    cta_fixed_msm(P, table, nrows, digits, red);
    if (t == 0) sh_pts[0] = P;
    __syncthreads();
    // ---- A = alpha h + <aL, G> + <aR, H> (:1267-1276) ----
    if (t < (int)n) {
        fix_recode(digits[t], aL, table.wbits);
        fix_recode(digits[n + t], aR, table.wbits);
    }
    if (t == 0) {
        zero_row(digits[row_g]);
        fix_recode(digits[row_h], alpha, table.wbits);
    }
    __syncthreads();
    cta_fixed_msm(P, table, nrows, digits, red);
    if (t == 0) sh_pts[1] = P;
    __syncthreads();
    // ---- S = rho h + <sL, G> + <sR, H> (:1279-1288) ----
    if (t < (int)n) {
        fix_recode(digits[t], sL, table.wbits);
        fix_recode(digits[n + t], sR, table.wbits);
    }
    if (t == 0) fix_recode(digits[row_h], rho, table.wbits);
    __syncthreads();
    cta_fixed_msm(P, table, nrows, digits, red);
    if (t == 0) {
        sh_pts[2] = P;
        ge_normalize_many(sh_pts, 3);  // V, A, S with one inversion
        ge_store(rec + kRecV, sh_pts[0]);
        ge_store(rec + kRecA, sh_pts[1]);
        ge_store(rec + kRecS, sh_pts[2]);
        // y, z challenges
        Sha256 sh;
        uint32_t yb[8], zb[8];
        sh.init();
        sh.update_str("BulletproofYChal", 16);
        hash_xy(sh, sh_pts[0]);
        hash_xy(sh, sh_pts[1]);
        hash_xy(sh, sh_pts[2]);
        sh.update_str("y_ch", 4);
        sh.final_challenge(yb);
        sh.init();
        sh.update_str("BulletproofZChal", 16);
        sh.update_words(yb);
        sh.update_str("z_ch", 4);
        sh.final_challenge(zb);
        sc y, z, yi;
        sc ty, tz;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            ty.v[i] = yb[i];
            tz.v[i] = zb[i];
        }
        sc_reduce(y, ty);
        sc_reduce(z, tz);
        sc_invert(yi, y);
        sh_z = z;
        sh_ypow[0] = y;
        sh_yinvpow[0] = yi;
        for (int m = 1; m <= k; m++) {
            sc_mul_nf(sh_ypow[m], sh_ypow[m - 1], sh_ypow[m - 1]);
            sc_mul_nf(sh_yinvpow[m], sh_yinvpow[m - 1], sh_yinvpow[m - 1]);
        }
    }
    __syncthreads();
    // ---- l(X) = l0 + l1 X, r(X) = r0 + r1 X and t0, t1, t2 ----
    sc z = sh_z, z2, l0, r0, r1, yi_pow, yinv_pow;
    sc_mul_nf(z2, z, z);
    sc_set0(l0);
    sc_set0(r0);
    sc_set0(r1);
    sc_set1(yi_pow);
    sc_set1(yinv_pow);
    if (t < (int)n) {
        for (int m = 0; m < k; m++) {
            if ((t >> m) & 1) {
                sc_mul_nf(yi_pow, yi_pow, sh_ypow[m]);
                sc_mul_nf(yinv_pow, yinv_pow, sh_yinvpow[m]);
            }
        }
        sc two_i, tmp;
        sc_set0(two_i);
        two_i.v[t >> 5] = 1u << (t & 31);
        sc_sub(l0, aL, z);
        sc_add(tmp, aR, z);
        sc_mul_nf(tmp, tmp, yi_pow);
        sc_mul_nf(two_i, z2, two_i);
        sc_add(r0, tmp, two_i);
        sc_mul_nf(r1, yi_pow, sR);
    }
    sc t0, t1, t2, tmp, tmp2;
    sc_mul_nf(t0, l0, r0);
    cta_sc_sum(t0, sred);
    sc_mul_nf(tmp, l0, r1);
    sc_mul_nf(tmp2, sL, r0);
    sc_add(t1, tmp, tmp2);
    cta_sc_sum(t1, sred);
    sc_mul_nf(t2, sL, r1);
    cta_sc_sum(t2, sred);
    // ---- T1 = t1 g + tau1 h, T2 = t2 g + tau2 h ----
    for (int r = t; r < nrows; r += kPThreads) zero_row(digits[r]);
    __syncthreads();
    if (t == 0) {
        fix_recode(digits[row_g], t1, table.wbits);
        fix_recode(digits[row_h], tau1, table.wbits);
    }
    __syncthreads();
    cta_fixed_msm(P, table, nrows, digits, red);
    if (t == 0) {
        sh_pts[0] = P;
        fix_recode(digits[row_g], t2, table.wbits);
        fix_recode(digits[row_h], tau2, table.wbits);
    }
    __syncthreads();
    cta_fixed_msm(P, table, nrows, digits, red);
    if (t == 0) {
        sh_pts[1] = P;
        ge_normalize_many(sh_pts, 2);  // T1, T2 with one inversion
        ge_store(rec + kRecT1, sh_pts[0]);
        ge_store(rec + kRecT2, sh_pts[1]);
        Sha256 sh;
        uint32_t xb[8];
        sh.init();
        sh.update_str("BulletproofXChal", 16);
        hash_xy(sh, sh_pts[0]);
        hash_xy(sh, sh_pts[1]);
        sh.update_str("xcha", 4);
        sh.final_challenge(xb);
        sc tx;
#pragma unroll
        for (int i = 0; i < 8; i++) tx.v[i] = xb[i];
        sc_reduce(sh_x, tx);
    }
    __syncthreads();
    sc x = sh_x, x2, tt, taux, mu;
    sc_mul_nf(x2, x, x);
    sc_mul_nf(tmp, t1, x);
    sc_mul_nf(tmp2, t2, x2);
    sc_add(tt, t0, tmp);
    sc_add(tt, tt, tmp2);
    sc_mul_nf(tmp, tau1, x);
    sc_mul_nf(tmp2, tau2, x2);
    sc_add(taux, tmp, tmp2);
    sc_mul_nf(tmp, z2, gamma);
    sc_add(taux, taux, tmp);
    sc_mul_nf(tmp, rho, x);
    sc_add(mu, alpha, tmp);
    if (t < (int)n) {
        sc a, b;
        sc_mul_nf(tmp, sL, x);
        sc_add(a, l0, tmp);
        sc_mul_nf(tmp, r1, x);
        sc_add(b, r0, tmp);
        sa[t] = a;
        sb[t] = b;
        sc_set1(swG[t]);
        swH[t] = yinv_pow;  // H'_i = y^-i H_i folded into the H weights
    }
    if (t == 0) {
        sc_store(rec + kRecT, tt);
        sc_store(rec + kRecTaux, taux);
        sc_store(rec + kRecMu, mu);
        sc_store(rec + kRecIpC, tt);
        Sha256 sh;
        sh.init();
        sh.update_str("BulletproofIP", 13);
        sh.update_words(tt.v);
        sh.update_words(taux.v);
        sh.update_words(mu.v);
        sh.final_challenge(sh_tr);
    }
    __syncthreads();
    // ---- inner-product argument over (G, H' = y^-i H, Q = h) ----
    for (int r = 0; r < k; r++) {
        const int nr = (int)n >> r, np = nr >> 1, bitpos = k - 1 - r;
        sc cL, cR;
        sc_set0(cL);
        sc_set0(cR);
        if (t < np) {
            sc_mul_nf(cL, sa[t], sb[t + np]);
            sc_mul_nf(cR, sa[t + np], sb[t]);
        }
        cta_sc_sum(cL, sred);
        cta_sc_sum(cR, sred);
        const int m = t & (nr - 1), hi = (t >> bitpos) & 1;
        for (int side = 0; side < 2; side++) {  // 0: L, 1: R
            if (t < (int)n) {
                sc cg, ch;
                sc_set0(cg);
                sc_set0(ch);
                bool g_on = side == 0 ? hi : !hi;  // L uses G_R and H_L; R uses G_L and H_R
                if (g_on) sc_mul_nf(cg, sa[side == 0 ? m - np : m + np], swG[t]);
                else sc_mul_nf(ch, sb[side == 0 ? m + np : m - np], swH[t]);
                fix_recode(digits[t], cg, table.wbits);
                fix_recode(digits[n + t], ch, table.wbits);
            }
            if (t == 0) {
                zero_row(digits[row_g]);
                fix_recode(digits[row_h], side == 0 ? cL : cR, table.wbits);
            }
            __syncthreads();
            cta_fixed_msm(P, table, nrows, digits, red);
            if (t == 0) sh_pts[side] = P;
            __syncthreads();
        }
        if (t == 0) {
            ge_normalize_many(sh_pts, 2);  // L_r, R_r with one inversion
            ge_store(rec + kRecL + (size_t)r * 128, sh_pts[0]);
            ge_store(rec + kRecL + (size_t)(k + r) * 128, sh_pts[1]);
            Sha256 sh;
            uint32_t ub[8];
            sh.init();
            sh.update_str("InnerProductChal", 16);
            sh.update_words(sh_tr);
            sh.update_words(sh_pts[0].X.v);
            sh.update_words(sh_pts[1].X.v);
            sh.final_challenge(ub);
            sc tu, u, ui;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                sh_tr[i] = ub[i];
                tu.v[i] = ub[i];
            }
            if (r == 0) sc_store(rec + kRecIpX, tu);  // raw challenge (:471-474)
            sc_reduce(u, tu);
            sc_invert(ui, u);
            sh_u = u;
            sh_uinv = ui;
        }
        __syncthreads();
        sc u = sh_u, ui = sh_uinv, na, nb;
        if (t < np) {  // a' = u a_L + u^-1 a_R ; b' = u^-1 b_L + u b_R
            sc_mul_nf(tmp, u, sa[t]);
            sc_mul_nf(tmp2, ui, sa[t + np]);
            sc_add(na, tmp, tmp2);
            sc_mul_nf(tmp, ui, sb[t]);
            sc_mul_nf(tmp2, u, sb[t + np]);
            sc_add(nb, tmp, tmp2);
        }
        __syncthreads();
        if (t < np) {
            sa[t] = na;
            sb[t] = nb;
        }
        if (t < (int)n) {  // G' = u^-1 G_L + u G_R ; H' = u H_L + u^-1 H_R as weight updates
            sc wg = swG[t], wh = swH[t];
            sc_mul_nf(wg, wg, hi ? u : ui);
            sc_mul_nf(wh, wh, hi ? ui : u);
            swG[t] = wg;
            swH[t] = wh;
        }
        __syncthreads();
    }
    if (t == 0) {
        sc_store(rec + kRecIpA, sa[0]);
        sc_store(rec + kRecIpB, sb[0]);
    }
}


// =====================================================================================================
// Batched prover: the same protocol as range_prove_kernel, cut into phases that each run over the WHOLE
// batch.  The one-CTA-per-proof kernel above is bound by its serial sections (17 field inversions, 7 scalar
// inversions, 10 hashes per proof, executed by one thread while the rest of the CTA waits at a barrier: ncu
// shows `barrier` as the top stall at 20 % issue utilisation).  Here every inversion is a Montgomery batch
// inversion ACROSS proofs, every fixed-base sum runs one window per lane as in the verifier, and the
// per-proof vector work keeps one 64-thread CTA per proof.  State lives in a caller-provided workspace.
// Bytes of the proofs are identical to the kernel above (and to the CPU oracle).
// =====================================================================================================
struct PScal {  // per-proof scalars
    sc gamma, alpha, rho, tau1, tau2, y, z, x, t0, t1, t2, u, uinv, yinv;
    uint32_t tr[8];
    uint32_t valid, pad[7];
};
static constexpr int kPbRows = 2 * kMaxN + 2;  // digit rows per fixed-base sum: G_i, H_i, g, h

struct PbLayout {
    size_t ps, pts, zinv, inv_in, inv_out, digits, l0, r0, r1, va, vb, wg, wh, tree, total;
};

__device__ __forceinline__ int8_t* pb_digits(int8_t* digits, uint32_t p, int slot) {
    return digits + ((size_t)p * 3 + slot) * kPbRows * kFixRowBytes;
}
__device__ __forceinline__ void pb_zero_rows(int8_t* rows, int nrows, int t) {
    uint32_t* w = reinterpret_cast<uint32_t*>(rows);
    for (int i = t; i < nrows * kFixRowBytes / 4; i += kPThreads) w[i] = 0;
}

// phase 1: validity, blinding draws, digits of V (slot 0), A (slot 1), S (slot 2)
__global__ void __launch_bounds__(kPThreads) pb_init_kernel(const uint8_t* __restrict__ gens,
                                                            const uint64_t* __restrict__ values,
                                                            const uint8_t* __restrict__ gammas,
                                                            NonceSource nsrc, uint32_t n, int k,
                                                            uint8_t* __restrict__ proofs, size_t rec_bytes,
                                                            PScal* __restrict__ ps, int8_t* __restrict__ digits) {
    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    const int wbits = (int)reinterpret_cast<const GensHeader*>(gens)->wbits;
    uint8_t* rec = proofs + (size_t)p * rec_bytes;
    const uint64_t v = values[p];
    const Nonce seed = nonce_of(nsrc, p);
    const int nrows = 2 * (int)n + 2, row_g = 2 * (int)n, row_h = 2 * (int)n + 1;
    if (n < 64 && (v >> n) != 0) {  // validate_range_input (:238-263): initialised, invalid proof (D20)
        ge_p3 O;
        ge_p3_0(O);
        for (int q = t; q < 5 + 2 * k; q += kPThreads) ge_store(rec + (q < 5 ? q * 128 : kRecL + (q - 5) * 128), O);
        if (t < 7) {
            fe zf;
            fe_set0(zf);
            fe_store(rec + kRecTaux + t * 32, zf);
        }
        if (t == 0) ps[p].valid = 0;
        return;
    }
    for (int slot = 0; slot < 3; slot++) pb_zero_rows(pb_digits(digits, p, slot), nrows, t);
    __syncthreads();
    int8_t* dV = pb_digits(digits, p, 0);
    int8_t* dA = pb_digits(digits, p, 1);
    int8_t* dS = pb_digits(digits, p, 2);
    if (t == 0) {
        PScal& s = ps[p];
        s.valid = 1;
        sc g0, gamma, alpha, rho, tau1, tau2, vs;
        sc_load(g0, gammas + (size_t)p * 32);
        sc_reduce(gamma, g0);
        draw_scalar(alpha, seed, 2 * n);
        draw_scalar(rho, seed, 2 * n + 1);
        draw_scalar(tau1, seed, 2 * n + 2);
        draw_scalar(tau2, seed, 2 * n + 3);
        s.gamma = gamma;
        s.alpha = alpha;
        s.rho = rho;
        s.tau1 = tau1;
        s.tau2 = tau2;
        sc_set0(vs);
        vs.v[0] = (uint32_t)v;
        vs.v[1] = (uint32_t)(v >> 32);
        fix_recode(dV + row_g * kFixRowBytes, vs, wbits);     // V = v g + gamma h
        fix_recode(dV + row_h * kFixRowBytes, gamma, wbits);
        fix_recode(dA + row_h * kFixRowBytes, alpha, wbits);  // A = alpha h + <aL, G> + <aR, H>
        fix_recode(dS + row_h * kFixRowBytes, rho, wbits);    // S = rho h + <sL, G> + <sR, H>
    }
    if (t < (int)n) {
        sc aL, aR, sL, sR, one;
        sc_set1(one);
        sc_set0(aL);
        if ((v >> t) & 1) sc_set1(aL);
        sc_sub(aR, aL, one);
        draw_scalar(sL, seed, 2 * t);
        draw_scalar(sR, seed, 2 * t + 1);
        fix_recode(dA + t * kFixRowBytes, aL, wbits);
        fix_recode(dA + (n + t) * kFixRowBytes, aR, wbits);
        fix_recode(dS + t * kFixRowBytes, sL, wbits);
        fix_recode(dS + (n + t) * kFixRowBytes, sR, wbits);
    }
}

// fixed-base sums: LP = 256 / WBITS lanes per (proof, slot), lane = window, as verify_fixed_kernel
template <int WBITS>
__global__ void __launch_bounds__(128, 4) pb_fixed_msm_kernel(const uint8_t* __restrict__ gens,
                                                              const PScal* __restrict__ ps,
                                                              const int8_t* __restrict__ digits, int nslots, int nrows,
                                                              int3 first_row, uint32_t num, uint8_t* __restrict__ out) {
    constexpr int LP = 256 / WBITS;
    constexpr uint32_t E = 1u << (WBITS - 1);
    const uint32_t unit = (blockIdx.x * blockDim.x + threadIdx.x) / LP;
    const int win = threadIdx.x & (LP - 1);
    // slot-major: the lanes of a warp then work on the SAME sum of different proofs, whose zero rows coincide
    // (in an IPA round L uses half of the G rows and R the other half: proof-major order ran every row at
    // half the lanes)
    const int slot = (int)(unit / num);
    const uint32_t p = unit % num;
    const bool live = slot < nslots && ps[p].valid;
    if (!__any_sync(0xffffffffu, live)) return;
    const GensHeader* gh = reinterpret_cast<const GensHeader*>(gens);
    const uint8_t* table = gens + gh->table_off + (size_t)win * E * 96;
    const int8_t* drow = digits + ((size_t)(live ? p : 0) * 3 + slot) * kPbRows * kFixRowBytes;
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
    // V, T1, T2 only have the g and h rows: their sums start there instead of walking 2n rows of zero digits
    const int row0 = slot == 0 ? first_row.x : slot == 1 ? first_row.y : first_row.z;
    digit_of(row0, mag, neg);
    if (mag) ge_niels_load(q, table + ((size_t)row0 * LP * E + (mag - 1)) * 96);
#pragma unroll 1
    for (int row = row0; row < nrows; row++) {
        uint32_t cmag = mag;
        bool cneg = neg;
        ge_niels cur = q;
        if (row + 1 < nrows) {
            digit_of(row + 1, mag, neg);
            if (mag) ge_niels_load(q, table + ((size_t)(row + 1) * LP * E + (mag - 1)) * 96);
        }
        if (cmag) ge_madd(acc, acc, cur, cneg);
    }
#pragma unroll 1
    for (int o = LP / 2; o > 0; o >>= 1) {
        ge_p3 other;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            other.X.v[j] = __shfl_xor_sync(0xffffffffu, acc.X.v[j], o);
            other.Y.v[j] = __shfl_xor_sync(0xffffffffu, acc.Y.v[j], o);
            other.Z.v[j] = __shfl_xor_sync(0xffffffffu, acc.Z.v[j], o);
            other.T.v[j] = __shfl_xor_sync(0xffffffffu, acc.T.v[j], o);
        }
        ge_add(acc, acc, other);
    }
    if (live && win == 0) ge_store(out + ((size_t)p * 3 + slot) * 128, acc);
}

// points -> (X/Z, Y/Z, 1, XY/Z^2) canonical, given 1/Z
__global__ void __launch_bounds__(128) pb_apply_norm_kernel(uint8_t* __restrict__ pts, const uint8_t* __restrict__ zinv,
                                                            size_t count) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fe X, Y, zi, x, y, tt, one;
    fe_load(X, pts + i * 128);
    fe_load(Y, pts + i * 128 + 32);
    fe_load_nc(zi, zinv + i * 32);
    fe_mul(x, X, zi);
    fe_mul(y, Y, zi);
    fe_canon(x);
    fe_canon(y);
    fe_mul(tt, x, y);
    fe_canon(tt);
    fe_set1(one);
    fe_store(pts + i * 128, x);
    fe_store(pts + i * 128 + 32, y);
    fe_store(pts + i * 128 + 64, one);
    fe_store(pts + i * 128 + 96, tt);
}

// scalar inversions mod l, one per thread: the batch's independent Fermat chains run side by side in ~0.2 ms (the
// chain's latency).  A tile-wise Montgomery kernel was measured too: 0.43 ms per call — its one inversion per tile
// plus the serial combines are a longer chain than the plain inversion, and the multiplications it saves are free.
__global__ void __launch_bounds__(64) pb_sc_invert_each_kernel(sc* __restrict__ out, const sc* __restrict__ in,
                                                               uint32_t count) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    sc x = in[i], r;
    sc_invert(r, x);
    out[i] = r;
}

// phase 2: V, A, S normalised -> record; y, z
__global__ void __launch_bounds__(64) pb_yz_kernel(uint8_t* __restrict__ proofs, size_t rec_bytes,
                                                   const uint8_t* __restrict__ pts, uint32_t num, PScal* __restrict__ ps,
                                                   sc* __restrict__ inv_in) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num) return;
    if (!ps[p].valid) {
        sc one;
        sc_set1(one);
        inv_in[p] = one;
        return;
    }
    uint8_t* rec = proofs + (size_t)p * rec_bytes;
    ge_p3 V, A, S;
    ge_load(V, pts + ((size_t)p * 3 + 0) * 128);
    ge_load(A, pts + ((size_t)p * 3 + 1) * 128);
    ge_load(S, pts + ((size_t)p * 3 + 2) * 128);
    ge_store(rec + kRecV, V);
    ge_store(rec + kRecA, A);
    ge_store(rec + kRecS, S);
    Sha256 sh;
    uint32_t yb[8], zb[8];
    sh.init();
    sh.update_str("BulletproofYChal", 16);
    hash_xy(sh, V);
    hash_xy(sh, A);
    hash_xy(sh, S);
    sh.update_str("y_ch", 4);
    sh.final_challenge(yb);
    sh.init();
    sh.update_str("BulletproofZChal", 16);
    sh.update_words(yb);
    sh.update_str("z_ch", 4);
    sh.final_challenge(zb);
    sc y, z, ty, tz;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ty.v[i] = yb[i];
        tz.v[i] = zb[i];
    }
    sc_reduce(y, ty);
    sc_reduce(z, tz);
    ps[p].y = y;
    ps[p].z = z;
    inv_in[p] = y;
}

// phase 3: l(X), r(X), t0..t2; digits of T1 (slot 0), T2 (slot 1)
__global__ void __launch_bounds__(kPThreads) pb_poly_kernel(const uint8_t* __restrict__ gens,
                                                            const uint64_t* __restrict__ values,
                                                            NonceSource nsrc, uint32_t n, int k,
                                                            PScal* __restrict__ ps, const sc* __restrict__ inv_out,
                                                            int8_t* __restrict__ digits, sc* __restrict__ vl0,
                                                            sc* __restrict__ vr0, sc* __restrict__ vr1,
                                                            sc* __restrict__ vwh) {
    __shared__ sc sred[kPThreads];
    __shared__ sc sh_ypow[kMaxK + 1], sh_yinvpow[kMaxK + 1];
    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    if (!ps[p].valid) return;
    const int wbits = (int)reinterpret_cast<const GensHeader*>(gens)->wbits;
    const uint64_t v = values[p];
    const Nonce seed = nonce_of(nsrc, p);
    const int nrows = 2 * (int)n + 2, row_g = 2 * (int)n, row_h = 2 * (int)n + 1;
    if (t == 0) {
        sc yi = inv_out[p];
        ps[p].yinv = yi;
        sh_ypow[0] = ps[p].y;
        sh_yinvpow[0] = yi;
        for (int m = 1; m <= k; m++) {
            sc_mul_nf(sh_ypow[m], sh_ypow[m - 1], sh_ypow[m - 1]);
            sc_mul_nf(sh_yinvpow[m], sh_yinvpow[m - 1], sh_yinvpow[m - 1]);
        }
    }
    pb_zero_rows(pb_digits(digits, p, 0), nrows, t);
    pb_zero_rows(pb_digits(digits, p, 1), nrows, t);
    __syncthreads();
    sc z = ps[p].z, z2, l0, r0, r1, yi_pow, yinv_pow, aL, aR, sL, sR, one;
    sc_set1(one);
    sc_mul_nf(z2, z, z);
    sc_set0(l0);
    sc_set0(r0);
    sc_set0(r1);
    sc_set0(aL);
    sc_set0(aR);
    sc_set0(sL);
    sc_set0(sR);
    sc_set1(yi_pow);
    sc_set1(yinv_pow);
    if (t < (int)n) {
        if ((v >> t) & 1) sc_set1(aL);
        sc_sub(aR, aL, one);
        draw_scalar(sL, seed, 2 * t);
        draw_scalar(sR, seed, 2 * t + 1);
        for (int m = 0; m < k; m++) {
            if ((t >> m) & 1) {
                sc_mul_nf(yi_pow, yi_pow, sh_ypow[m]);
                sc_mul_nf(yinv_pow, yinv_pow, sh_yinvpow[m]);
            }
        }
        sc two_i, tmp;
        sc_set0(two_i);
        two_i.v[t >> 5] = 1u << (t & 31);
        sc_sub(l0, aL, z);
        sc_add(tmp, aR, z);
        sc_mul_nf(tmp, tmp, yi_pow);
        sc_mul_nf(two_i, z2, two_i);
        sc_add(r0, tmp, two_i);
        sc_mul_nf(r1, yi_pow, sR);
        vl0[(size_t)p * kMaxN + t] = l0;
        vr0[(size_t)p * kMaxN + t] = r0;
        vr1[(size_t)p * kMaxN + t] = r1;
        vwh[(size_t)p * kMaxN + t] = yinv_pow;  // H'_i = y^-i H_i folded into the H weights
    }
    sc t0, t1, t2, tmp, tmp2;
    sc_mul_nf(t0, l0, r0);
    cta_sc_sum(t0, sred);
    sc_mul_nf(tmp, l0, r1);
    sc_mul_nf(tmp2, sL, r0);
    sc_add(t1, tmp, tmp2);
    cta_sc_sum(t1, sred);
    sc_mul_nf(t2, sL, r1);
    cta_sc_sum(t2, sred);
    if (t == 0) {
        ps[p].t0 = t0;
        ps[p].t1 = t1;
        ps[p].t2 = t2;
        int8_t* d1 = pb_digits(digits, p, 0);
        int8_t* d2 = pb_digits(digits, p, 1);
        fix_recode(d1 + row_g * kFixRowBytes, t1, wbits);
        fix_recode(d1 + row_h * kFixRowBytes, ps[p].tau1, wbits);
        fix_recode(d2 + row_g * kFixRowBytes, t2, wbits);
        fix_recode(d2 + row_h * kFixRowBytes, ps[p].tau2, wbits);
    }
}

// phase 4: T1, T2 -> record; x; t, taux, mu; a, b, weights; IPA transcript seed
__global__ void __launch_bounds__(kPThreads) pb_ipa_init_kernel(NonceSource nsrc, uint32_t n,
                                                                uint8_t* __restrict__ proofs, size_t rec_bytes,
                                                                const uint8_t* __restrict__ pts, PScal* __restrict__ ps,
                                                                const sc* __restrict__ vl0, const sc* __restrict__ vr0,
                                                                const sc* __restrict__ vr1, sc* __restrict__ va,
                                                                sc* __restrict__ vb, sc* __restrict__ vwg) {
    __shared__ sc sh_x;
    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    if (!ps[p].valid) return;
    uint8_t* rec = proofs + (size_t)p * rec_bytes;
    const Nonce seed = nonce_of(nsrc, p);
    if (t == 0) {
        ge_p3 T1, T2;
        ge_load(T1, pts + ((size_t)p * 3 + 0) * 128);
        ge_load(T2, pts + ((size_t)p * 3 + 1) * 128);
        ge_store(rec + kRecT1, T1);
        ge_store(rec + kRecT2, T2);
        Sha256 sh;
        uint32_t xb[8];
        sh.init();
        sh.update_str("BulletproofXChal", 16);
        hash_xy(sh, T1);
        hash_xy(sh, T2);
        sh.update_str("xcha", 4);
        sh.final_challenge(xb);
        sc tx, x;
#pragma unroll
        for (int i = 0; i < 8; i++) tx.v[i] = xb[i];
        sc_reduce(x, tx);
        sh_x = x;
        ps[p].x = x;
    }
    __syncthreads();
    sc x = sh_x, tmp, tmp2;
    if (t < (int)n) {
        sc sL, sR, a, b;
        draw_scalar(sL, seed, 2 * t);
        (void)sR;
        sc_mul_nf(tmp, sL, x);
        sc_add(a, vl0[(size_t)p * kMaxN + t], tmp);
        sc_mul_nf(tmp, vr1[(size_t)p * kMaxN + t], x);
        sc_add(b, vr0[(size_t)p * kMaxN + t], tmp);
        va[(size_t)p * kMaxN + t] = a;
        vb[(size_t)p * kMaxN + t] = b;
        sc one;
        sc_set1(one);
        vwg[(size_t)p * kMaxN + t] = one;
    }
    if (t == 0) {
        const PScal& s = ps[p];
        sc x2, tt, taux, mu, z2;
        sc_mul_nf(x2, x, x);
        sc_mul_nf(z2, s.z, s.z);
        sc_mul_nf(tmp, s.t1, x);
        sc_mul_nf(tmp2, s.t2, x2);
        sc_add(tt, s.t0, tmp);
        sc_add(tt, tt, tmp2);
        sc_mul_nf(tmp, s.tau1, x);
        sc_mul_nf(tmp2, s.tau2, x2);
        sc_add(taux, tmp, tmp2);
        sc_mul_nf(tmp, z2, s.gamma);
        sc_add(taux, taux, tmp);
        sc_mul_nf(tmp, s.rho, x);
        sc_add(mu, s.alpha, tmp);
        sc_store(rec + kRecT, tt);
        sc_store(rec + kRecTaux, taux);
        sc_store(rec + kRecMu, mu);
        sc_store(rec + kRecIpC, tt);
        Sha256 sh;
        sh.init();
        sh.update_str("BulletproofIP", 13);
        sh.update_words(tt.v);
        sh.update_words(taux.v);
        sh.update_words(mu.v);
        sh.final_challenge(ps[p].tr);
    }
}

// round r, part 1: cL, cR and the digits of L (slot 0) and R (slot 1) over the ORIGINAL generators
__device__ __forceinline__ void pb_lr_digits_body(const uint8_t* __restrict__ gens, uint32_t n, int k, int r,
                                                  const PScal* __restrict__ ps, const sc* va, const sc* vb, const sc* vwg,
                                                  const sc* vwh, int8_t* __restrict__ digits, sc* sred) {
    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    const int wbits = (int)reinterpret_cast<const GensHeader*>(gens)->wbits;
    const int row_g = 2 * (int)n, row_h = 2 * (int)n + 1;
    const sc* sa = va + (size_t)p * kMaxN;
    const sc* sb = vb + (size_t)p * kMaxN;
    const int nr = (int)n >> r, np = nr >> 1, bitpos = k - 1 - r;
    sc cL, cR;
    sc_set0(cL);
    sc_set0(cR);
    if (t < np) {
        sc_mul_nf(cL, sa[t], sb[t + np]);
        sc_mul_nf(cR, sa[t + np], sb[t]);
    }
    cta_sc_sum(cL, sred);
    cta_sc_sum(cR, sred);
    const int m = t & (nr - 1), hi = (t >> bitpos) & 1;
    for (int side = 0; side < 2; side++) {  // 0: L, 1: R
        int8_t* d = pb_digits(digits, p, side);
        if (t < (int)n) {
            sc cg, ch;
            sc_set0(cg);
            sc_set0(ch);
            bool g_on = side == 0 ? hi : !hi;  // L uses G_R and H_L; R uses G_L and H_R
            if (g_on) sc_mul_nf(cg, sa[side == 0 ? m - np : m + np], vwg[(size_t)p * kMaxN + t]);
            else sc_mul_nf(ch, sb[side == 0 ? m + np : m - np], vwh[(size_t)p * kMaxN + t]);
            fix_recode(d + t * kFixRowBytes, cg, wbits);
            fix_recode(d + (n + t) * kFixRowBytes, ch, wbits);
        }
        if (t == 0) {
            zero_row(d + row_g * kFixRowBytes);
            fix_recode(d + row_h * kFixRowBytes, side == 0 ? cL : cR, wbits);
        }
    }
}

__global__ void __launch_bounds__(kPThreads) pb_lr_digits_kernel(const uint8_t* __restrict__ gens, uint32_t n, int k, int r,
                                                                 const PScal* __restrict__ ps, const sc* __restrict__ va,
                                                                 const sc* __restrict__ vb, const sc* __restrict__ vwg,
                                                                 const sc* __restrict__ vwh, int8_t* __restrict__ digits) {
    __shared__ sc sred[kPThreads];
    if (!ps[blockIdx.x].valid) return;
    pb_lr_digits_body(gens, n, k, r, ps, va, vb, vwg, vwh, digits, sred);
}

// round r, part 2: L_r, R_r -> record; challenge u
__global__ void __launch_bounds__(64) pb_u_kernel(uint8_t* __restrict__ proofs, size_t rec_bytes, int k, int r,
                                                  const uint8_t* __restrict__ pts, uint32_t num, PScal* __restrict__ ps,
                                                  sc* __restrict__ inv_in) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= num) return;
    if (!ps[p].valid) {
        sc one;
        sc_set1(one);
        inv_in[p] = one;
        return;
    }
    uint8_t* rec = proofs + (size_t)p * rec_bytes;
    ge_p3 Lp, Rp;
    ge_load(Lp, pts + ((size_t)p * 3 + 0) * 128);
    ge_load(Rp, pts + ((size_t)p * 3 + 1) * 128);
    ge_store(rec + kRecL + (size_t)r * 128, Lp);
    ge_store(rec + kRecL + (size_t)(k + r) * 128, Rp);
    Sha256 sh;
    uint32_t ub[8];
    sh.init();
    sh.update_str("InnerProductChal", 16);
    sh.update_words(ps[p].tr);
    sh.update_words(Lp.X.v);
    sh.update_words(Rp.X.v);
    sh.final_challenge(ub);
    sc tu, u;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        ps[p].tr[i] = ub[i];
        tu.v[i] = ub[i];
    }
    if (r == 0) sc_store(rec + kRecIpX, tu);  // raw challenge (:471-474)
    sc_reduce(u, tu);
    ps[p].u = u;
    inv_in[p] = u;
}

// round r, part 3: fold a, b; update the generator weights
__device__ __forceinline__ void pb_fold_body(uint32_t n, int k, int r, uint8_t* __restrict__ proofs, size_t rec_bytes,
                                             const PScal* __restrict__ ps, const sc* __restrict__ inv_out, sc* va, sc* vb,
                                             sc* vwg, sc* vwh) {
    const int t = threadIdx.x;
    const uint32_t p = blockIdx.x;
    sc* sa = va + (size_t)p * kMaxN;
    sc* sb = vb + (size_t)p * kMaxN;
    const int nr = (int)n >> r, np = nr >> 1, bitpos = k - 1 - r;
    const int hi = (t >> bitpos) & 1;
    sc u = ps[p].u, ui = inv_out[p], na, nb, tmp, tmp2;
    if (t < np) {  // a' = u a_L + u^-1 a_R ; b' = u^-1 b_L + u b_R
        sc_mul_nf(tmp, u, sa[t]);
        sc_mul_nf(tmp2, ui, sa[t + np]);
        sc_add(na, tmp, tmp2);
        sc_mul_nf(tmp, ui, sb[t]);
        sc_mul_nf(tmp2, u, sb[t + np]);
        sc_add(nb, tmp, tmp2);
    }
    __syncthreads();
    if (t < np) {
        sa[t] = na;
        sb[t] = nb;
    }
    if (t < (int)n) {  // G' = u^-1 G_L + u G_R ; H' = u H_L + u^-1 H_R as weight updates
        sc wg = vwg[(size_t)p * kMaxN + t], wh = vwh[(size_t)p * kMaxN + t];
        sc_mul_nf(wg, wg, hi ? u : ui);
        sc_mul_nf(wh, wh, hi ? ui : u);
        vwg[(size_t)p * kMaxN + t] = wg;
        vwh[(size_t)p * kMaxN + t] = wh;
    }
    if (r == k - 1 && t == 0) {
        uint8_t* rec = proofs + (size_t)p * rec_bytes;
        sc_store(rec + kRecIpA, na);
        sc_store(rec + kRecIpB, nb);
    }
}

// fold of round r, then (same CTA, vectors still in cache) the digits of round r + 1
__global__ void __launch_bounds__(kPThreads) pb_fold_digits_kernel(const uint8_t* __restrict__ gens, uint32_t n, int k, int r,
                                                                   uint8_t* __restrict__ proofs, size_t rec_bytes,
                                                                   const PScal* __restrict__ ps,
                                                                   const sc* __restrict__ inv_out, sc* va, sc* vb, sc* vwg,
                                                                   sc* vwh, int8_t* __restrict__ digits) {
    __shared__ sc sred[kPThreads];
    if (!ps[blockIdx.x].valid) return;
    pb_fold_body(n, k, r, proofs, rec_bytes, ps, inv_out, va, vb, vwg, vwh);
    if (r + 1 < k) {
        __syncthreads();  // the folded vectors were written by other threads of this CTA
        pb_lr_digits_body(gens, n, k, r + 1, ps, va, vb, vwg, vwh, digits, sred);
    }
}

int fe_batch_invert_strided(uint8_t* d_out, const uint8_t* d_in, size_t in_stride, size_t count, cudaStream_t st,
                            uint8_t* d_ws, size_t ws_bytes);

static size_t pb_align(size_t x) { return (x + 255) / 256 * 256; }
static constexpr size_t kPbChunk = 16384;
static constexpr size_t kPbMinBatch = 64;  // below this the single-kernel prover has the lower latency
static PbLayout pb_layout(size_t cnt) {
    PbLayout L;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += pb_align(bytes);
        return o;
    };
    L.ps = take(cnt * sizeof(PScal));
    L.pts = take(cnt * 3 * 128);
    L.zinv = take(cnt * 3 * 32);
    L.inv_in = take(cnt * sizeof(sc));
    L.inv_out = take(cnt * sizeof(sc));
    L.digits = take(cnt * 3 * (size_t)kPbRows * kFixRowBytes);
    L.l0 = take(cnt * kMaxN * sizeof(sc));
    L.r0 = take(cnt * kMaxN * sizeof(sc));
    L.r1 = take(cnt * kMaxN * sizeof(sc));
    L.va = take(cnt * kMaxN * sizeof(sc));
    L.vb = take(cnt * kMaxN * sizeof(sc));
    L.wg = take(cnt * kMaxN * sizeof(sc));
    L.wh = take(cnt * kMaxN * sizeof(sc));
    L.tree = take(cnt * 3 * 16);  // workspace of the field batch inversion (4.3 B per element needed)
    L.total = off;
    return L;
}

}  // namespace cbp

using namespace cbp;

extern "C" {

int bpk_range_prove_workspace_bytes(size_t n, size_t num_proofs, size_t* bytes) {
    if (!bytes || n == 0 || n > kMaxN || (n & (n - 1))) return fail(BPK_ERR_ARG);
    // optional: without a workspace (or for a handful of proofs) the one-CTA-per-proof kernel is used
    size_t chunk = num_proofs < kPbChunk ? num_proofs : kPbChunk;
    *bytes = num_proofs >= kPbMinBatch ? pb_layout(chunk).total : 0;
    return BPK_OK;
}
static int range_prove_batch(const void* d_gens_ws, const uint64_t* d_values, const void* d_gammas,
                             const uint64_t* d_seeds, const uint8_t* d_keys, size_t n, size_t num_proofs, void* d_proofs,
                             void* d_workspace, size_t workspace_bytes, void* stream) {
    if (n == 0 || n > kMaxN || (n & (n - 1))) return fail(BPK_ERR_ARG);
    if (!num_proofs) return BPK_OK;
    if (!d_gens_ws || !d_values || !d_gammas || (!d_seeds && !d_keys) || !d_proofs) return fail(BPK_ERR_ARG);
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    const size_t rec = proof_record_bytes(k);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t chunk = num_proofs < kPbChunk ? num_proofs : kPbChunk;
    const bool legacy_env = options().prover_legacy != 0;
    const int wbits = bpk_gens_window_bits(d_gens_ws);
    if (legacy_env || n < 2 || num_proofs < kPbMinBatch || !d_workspace || workspace_bytes < pb_layout(chunk).total || !wbits) {
        range_prove_kernel<<<(unsigned)num_proofs, kPThreads, 0, st>>>((const uint8_t*)d_gens_ws, d_values,
                                                                      (const uint8_t*)d_gammas, NonceSource{d_seeds, d_keys},
                                                                      (uint32_t)n, k, (uint8_t*)d_proofs, rec);
        CBP_CHECK_LAUNCH();
        return BPK_OK;
    }
    const PbLayout L = pb_layout(chunk);
    uint8_t* ws = (uint8_t*)d_workspace;
    PScal* ps = (PScal*)(ws + L.ps);
    uint8_t *pts = ws + L.pts, *zinv = ws + L.zinv, *tree = ws + L.tree;
    sc *inv_in = (sc*)(ws + L.inv_in), *inv_out = (sc*)(ws + L.inv_out);
    int8_t* digits = (int8_t*)(ws + L.digits);
    sc *vl0 = (sc*)(ws + L.l0), *vr0 = (sc*)(ws + L.r0), *vr1 = (sc*)(ws + L.r1);
    sc *va = (sc*)(ws + L.va), *vb = (sc*)(ws + L.vb), *vwg = (sc*)(ws + L.wg), *vwh = (sc*)(ws + L.wh);
    const uint8_t* gens = (const uint8_t*)d_gens_ws;
    const int nrows = 2 * (int)n + 2;
    for (size_t done = 0; done < num_proofs; done += chunk) {
        const uint32_t cnt = (uint32_t)((num_proofs - done) < chunk ? (num_proofs - done) : chunk);
        const uint64_t* vals = d_values + done;
        const NonceSource seeds{d_keys ? nullptr : d_seeds + done, d_keys ? d_keys + done * 32 : nullptr};
        const uint8_t* gam = (const uint8_t*)d_gammas + done * 32;
        uint8_t* proofs = (uint8_t*)d_proofs + done * rec;
        auto fixed_msm = [&](int nslots, int3 first_row) -> int {
            const int lp = 256 / wbits;
            unsigned grid = (unsigned)(((size_t)cnt * nslots * lp + 127) / 128);
            if (wbits == 8) pb_fixed_msm_kernel<8><<<grid, 128, 0, st>>>(gens, ps, digits, nslots, nrows, first_row, cnt, pts);
            else pb_fixed_msm_kernel<16><<<grid, 128, 0, st>>>(gens, ps, digits, nslots, nrows, first_row, cnt, pts);
            CBP_CHECK_LAUNCH();
            // normalise every slot of every proof with ONE batch inversion (stale slots are harmless)
            int rc = fe_batch_invert_strided(zinv, pts + 64, 128, (size_t)cnt * 3, st, tree, (size_t)cnt * 3 * 16);
            if (rc != BPK_OK) return rc;
            pb_apply_norm_kernel<<<(unsigned)(((size_t)cnt * 3 + 127) / 128), 128, 0, st>>>(pts, zinv, (size_t)cnt * 3);
            CBP_CHECK_LAUNCH();
            return BPK_OK;
        };
        auto sc_invert_batch = [&]() -> int {
            pb_sc_invert_each_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(inv_out, inv_in, cnt);
            CBP_CHECK_LAUNCH();
            return BPK_OK;
        };
        int rc;
        CBP_CUDA(cudaMemsetAsync(pts, 0, (size_t)cnt * 3 * 128, st));  // defined contents for skipped (invalid) proofs
        pb_init_kernel<<<cnt, kPThreads, 0, st>>>(gens, vals, gam, seeds, (uint32_t)n, k, proofs, rec, ps, digits);
        CBP_CHECK_LAUNCH();
        const int gh = 2 * (int)n;  // first of the two rows g, h
        if ((rc = fixed_msm(3, make_int3(gh, 0, 0))) != BPK_OK) return rc;  // V | A | S
        pb_yz_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(proofs, rec, pts, cnt, ps, inv_in);
        CBP_CHECK_LAUNCH();
        if ((rc = sc_invert_batch()) != BPK_OK) return rc;
        pb_poly_kernel<<<cnt, kPThreads, 0, st>>>(gens, vals, seeds, (uint32_t)n, k, ps, inv_out, digits, vl0, vr0, vr1, vwh);
        CBP_CHECK_LAUNCH();
        if ((rc = fixed_msm(2, make_int3(gh, gh, 0))) != BPK_OK) return rc;  // T1 | T2
        pb_ipa_init_kernel<<<cnt, kPThreads, 0, st>>>(seeds, (uint32_t)n, proofs, rec, pts, ps, vl0, vr0, vr1, va, vb, vwg);
        CBP_CHECK_LAUNCH();
        if (k > 0) {
            pb_lr_digits_kernel<<<cnt, kPThreads, 0, st>>>(gens, (uint32_t)n, k, 0, ps, va, vb, vwg, vwh, digits);
            CBP_CHECK_LAUNCH();
        }
        for (int r = 0; r < k; r++) {
            if ((rc = fixed_msm(2, make_int3(0, 0, 0))) != BPK_OK) return rc;  // L_r | R_r
            pb_u_kernel<<<(cnt + 63) / 64, 64, 0, st>>>(proofs, rec, k, r, pts, cnt, ps, inv_in);
            CBP_CHECK_LAUNCH();
            if ((rc = sc_invert_batch()) != BPK_OK) return rc;
            pb_fold_digits_kernel<<<cnt, kPThreads, 0, st>>>(gens, (uint32_t)n, k, r, proofs, rec, ps, inv_out, va, vb, vwg,
                                                             vwh, digits);
            CBP_CHECK_LAUNCH();
        }
    }
    return BPK_OK;
}
int bpk_range_prove_batch_device(const void* d_gens_ws, const uint64_t* d_values, const void* d_gammas,
                                 const uint64_t* d_seeds, size_t n, size_t num_proofs, void* d_proofs,
                                 void* d_workspace, size_t workspace_bytes, void* stream) {
    return range_prove_batch(d_gens_ws, d_values, d_gammas, d_seeds, nullptr, n, num_proofs, d_proofs, d_workspace,
                             workspace_bytes, stream);
}
int bpk_range_prove_batch_keyed_device(const void* d_gens_ws, const uint64_t* d_values, const void* d_gammas,
                                       const void* d_keys, size_t n, size_t num_proofs, void* d_proofs,
                                       void* d_workspace, size_t workspace_bytes, void* stream) {
    if (num_proofs && !d_keys) return fail(BPK_ERR_ARG);
    return range_prove_batch(d_gens_ws, d_values, d_gammas, nullptr, (const uint8_t*)d_keys, n, num_proofs, d_proofs,
                             d_workspace, workspace_bytes, stream);
}

}  // extern "C"
