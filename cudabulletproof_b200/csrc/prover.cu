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
// the j-th 32-byte draw of the stream, clamped like generate_random_scalar
// (bulletproof_range_proof.cu:153-159), then reduced mod l
__device__ __forceinline__ void draw_scalar(sc& r, uint64_t seed, uint32_t j) {
    sc t;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        uint64_t v = sm64_at(seed, (uint64_t)j * 4 + q);
        t.v[2 * q] = (uint32_t)v;
        t.v[2 * q + 1] = (uint32_t)(v >> 32);
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
                                                                const uint64_t* __restrict__ seeds, uint32_t n, int k,
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
    const uint64_t v = values[p], seed = seeds[p];
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

}  // namespace cbp

using namespace cbp;

extern "C" {

int bpk_range_prove_workspace_bytes(size_t n, size_t num_proofs, size_t* bytes) {
    (void)num_proofs;
    if (!bytes || n == 0 || n > kMaxN || (n & (n - 1))) return fail(BPK_ERR_ARG);
    *bytes = 0;
    return BPK_OK;
}
int bpk_range_prove_batch_device(const void* d_gens_ws, const uint64_t* d_values, const void* d_gammas,
                                 const uint64_t* d_seeds, size_t n, size_t num_proofs, void* d_proofs,
                                 void* d_workspace, size_t workspace_bytes, void* stream) {
    (void)d_workspace;
    (void)workspace_bytes;
    if (n == 0 || n > kMaxN || (n & (n - 1))) return fail(BPK_ERR_ARG);
    if (!num_proofs) return BPK_OK;
    if (!d_gens_ws || !d_values || !d_gammas || !d_seeds || !d_proofs) return fail(BPK_ERR_ARG);
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    range_prove_kernel<<<(unsigned)num_proofs, kPThreads, 0, (cudaStream_t)stream>>>(
        (const uint8_t*)d_gens_ws, d_values, (const uint8_t*)d_gammas, d_seeds, (uint32_t)n, k, (uint8_t*)d_proofs,
        proof_record_bytes(k));
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}

}  // extern "C"
