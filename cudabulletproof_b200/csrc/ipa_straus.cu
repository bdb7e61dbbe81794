// ipa_straus.cu — the rounds of the inner-product argument over at most kIpaCompositeMax generators, without folding
// a single point (inner_product_prove, bulletproof_vectors.cu:348-517, restated).
//
// The reference (and the first version of bpk_ipa_prove_device) folds the generators every round,
// G'_j = u^-1 G_j + u G_(j+n') (bulletproof_vectors.cu:641-663): 2 n' double-scalar multiplications, each a 253-step
// chain — 0.3 ms of pure latency per round however small n' is — and then multiplies the folded points by the folded
// scalars.  Here round r's L and R are written over the BASE generators instead,
//     L_r = sum_i (a_(j-n') wG_i) G_i [j = i mod m >= n'] + sum_i (b_(j+n') wH_i) H_i [j < n'] + c_L Q,
//     R_r = the complementary halves, c_R,
// with wG_i / wH_i the product of the challenges (or their inverses) that folding would have applied to generator i.
// L and R together touch every base generator once, so a round is ONE pass over fixed tables of the multiples
// 1..8 of every generator (built once): composite scalars -> signed 4-bit digits -> 2 x 64 window sums -> the
// binary combine tree in octet form (fe8.cuh), L and R side by side, with the round challenge, its inverse
// (divsteps) and the folded scalars at the end of the same launch sequence: 6 launches per round on one stream.
// Same group elements as the folded formulation whenever the generators have prime order (every generator set
// of a Bulletproofs setup; bpk_gens_derive_device clears the cofactor): identical bytes, see tests.
#include "common.h"
#include "fe8.cuh"
#include "ipa_straus.h"
#include "sc25519.cuh"
#include "sha256.cuh"

namespace cbp {

static constexpr int kWin = 64, kTab = 8;
static constexpr uint32_t kSlice = 256;  // points per warp of the window-sum kernel

// multiples 1..8 of G[0..m), H[0..m), Q in affine precomputed form (y+x, y-x, 2dxy; 96 B, one inversion per
// generator by Montgomery's trick — built once per argument): row t = i | m + i | 2m
__global__ void __launch_bounds__(128) ipa_tables_kernel(const uint8_t* __restrict__ g, const uint8_t* __restrict__ h,
                                                         const uint8_t* __restrict__ Q, uint32_t m,
                                                         uint8_t* __restrict__ tables, uint8_t* __restrict__ wG,
                                                         uint8_t* __restrict__ wH) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > 2 * m) return;
    ge_p3 P, M[kTab];
    ge_load(P, t < m ? g + (size_t)t * 128 : (t < 2 * m ? h + (size_t)(t - m) * 128 : Q));
    M[0] = P;
    ge_dbl(M[1], P);
    ge_add(M[2], M[1], P);
    ge_dbl(M[3], M[1]);
    ge_add(M[4], M[3], P);
    ge_dbl(M[5], M[2]);
    ge_add(M[6], M[5], P);
    ge_dbl(M[7], M[3]);
    fe pre[kTab], run, inv;
    fe_set1(run);
#pragma unroll
    for (int k = 0; k < kTab; k++) {
        pre[k] = run;
        fe_mul(run, run, M[k].Z);
    }
    fe_invert(inv, run);
#pragma unroll
    for (int k = kTab - 1; k >= 0; k--) {
        fe zi, x, y;
        fe_mul(zi, inv, pre[k]);
        fe_mul(inv, inv, M[k].Z);
        fe_mul(x, M[k].X, zi);
        fe_mul(y, M[k].Y, zi);
        ge_niels q;
        ge_to_niels_affine(q, x, y);
        ge_niels_store(tables + ((size_t)t * kTab + k) * 96, q);
    }
    if (t < 2 * m) {  // folding weights start at 1
        sc one;
        sc_set1(one);
        sc_store((t < m ? wG + (size_t)t * 32 : wH + (size_t)(t - m) * 32), one);
    }
}
// c_L = <a_lo, b_hi>, c_R = <a_hi, b_lo> mod l (vectors reduced, np <= kIpaCompositeMax / 2): one CTA
__global__ void __launch_bounds__(256) ipa_cross_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                                        uint32_t np, uint8_t* __restrict__ cL, uint8_t* __restrict__ cR) {
    __shared__ sc red[2][256];
    sc sl, sr;
    sc_set0(sl);
    sc_set0(sr);
    for (uint32_t j = threadIdx.x; j < np; j += blockDim.x) {
        sc al, ar, bl, br, t;
        sc_load(al, a + (size_t)j * 32);
        sc_load(ar, a + (size_t)(j + np) * 32);
        sc_load(bl, b + (size_t)j * 32);
        sc_load(br, b + (size_t)(j + np) * 32);
        sc_mul(t, al, br);
        sc_add(sl, sl, t);
        sc_mul(t, ar, bl);
        sc_add(sr, sr, t);
    }
    red[0][threadIdx.x] = sl;
    red[1][threadIdx.x] = sr;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            sc x = red[0][threadIdx.x], y = red[1][threadIdx.x];
            sc_add(x, x, red[0][threadIdx.x + o]);
            sc_add(y, y, red[1][threadIdx.x + o]);
            red[0][threadIdx.x] = x;
            red[1][threadIdx.x] = y;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        sc_store(cL, red[0][0]);
        sc_store(cR, red[1][0]);
    }
}
// composite scalars of both sides -> digits[(side * 64 + w) * npts + q], rows[side * npts + q];  npts = mb + 1
__global__ void __launch_bounds__(128) ipa_digits_kernel(uint32_t mb, uint32_t mcur, const uint8_t* __restrict__ a,
                                                         const uint8_t* __restrict__ b, const uint8_t* __restrict__ wG,
                                                         const uint8_t* __restrict__ wH, const uint8_t* __restrict__ cL,
                                                         const uint8_t* __restrict__ cR, int8_t* __restrict__ digits,
                                                         uint32_t* __restrict__ rows) {
    const uint32_t npts = mb + 1, p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= 2 * npts) return;
    const uint32_t side = p / npts, q = p % npts, np = mcur >> 1, half = mb >> 1;
    sc s;
    uint32_t row;
    if (q == mb) {
        sc_load(s, side ? cR : cL);
        row = 2 * mb;
    } else {
        const bool is_h = q >= half;
        const uint32_t tau = is_h ? q - half : q;
        // L takes the G's of the upper half of every length-m block and the H's of the lower half; R the others
        const bool upper = (side == 0) != is_h;
        const uint32_t i = (tau / np) * mcur + (tau % np) + (upper ? np : 0u);
        const uint32_t j = i % mcur, other = upper ? j - np : j + np;  // the scalar comes from the opposite half
        sc x, w;
        sc_load(x, (is_h ? b : a) + (size_t)other * 32);
        sc_load(w, (is_h ? wH : wG) + (size_t)i * 32);
        sc_mul(s, x, w);
        row = is_h ? mb + i : i;
    }
    rows[side * npts + q] = row;
    uint32_t carry = 0;
#pragma unroll 1
    for (int w = 0; w < kWin; w++) {
        int d = (int)((s.v[w >> 3] >> ((w & 7) * 4)) & 15u) + (int)carry;
        carry = d > 8;
        if (d > 8) d -= 16;
        digits[((size_t)side * kWin + w) * npts + q] = (int8_t)d;  // s < l < 2^253: no carry out of the top digit
    }
}
// grid (64 windows, slices, 2 sides), ONE WARP per CTA: sums[((side * nslices + slice) * 64 + w)].
// Lane-strided mixed additions (7M) from the affine tables, then a shuffle tree of unified additions.  (A first
// version with 256-thread CTAs over 1024-point slices spent most of its time in the trees: 4 additions per thread,
// then 5 shuffle levels and 7 serial additions by one lane — 182 us per round at n = 4096.)
__global__ void __launch_bounds__(32) ipa_sums_kernel(const uint8_t* __restrict__ tables, const int8_t* __restrict__ digits,
                                                      const uint32_t* __restrict__ rows, uint32_t npts,
                                                      uint32_t slice_len, uint8_t* __restrict__ sums) {
    const uint32_t w = blockIdx.x, side = blockIdx.z, lane = threadIdx.x;
    const uint32_t lo = blockIdx.y * slice_len, hi = lo + slice_len < npts ? lo + slice_len : npts;
    const int8_t* dg = digits + ((size_t)side * kWin + w) * npts;
    const uint32_t* rw = rows + (size_t)side * npts;
    ge_p3 acc;
    ge_p3_0(acc);
    for (uint32_t i = lo + lane; i < hi; i += 32) {
        const int d = dg[i];
        if (d == 0) continue;
        ge_niels q;
        ge_niels_load(q, tables + ((size_t)rw[i] * kTab + (uint32_t)((d < 0 ? -d : d) - 1)) * 96);
        ge_madd(acc, acc, q, d < 0);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_p3 other;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            other.X.v[k] = __shfl_down_sync(0xffffffffu, acc.X.v[k], o);
            other.Y.v[k] = __shfl_down_sync(0xffffffffu, acc.Y.v[k], o);
            other.Z.v[k] = __shfl_down_sync(0xffffffffu, acc.Z.v[k], o);
            other.T.v[k] = __shfl_down_sync(0xffffffffu, acc.T.v[k], o);
        }
        ge_add(acc, acc, other);
    }
    if (lane == 0) ge_store(sums + (((size_t)side * gridDim.y + blockIdx.y) * kWin + w) * 128, acc);
}
// more than one slice: one warp per (window, side) adds the slice sums up into slice 0's slot
__global__ void __launch_bounds__(32) ipa_slices_kernel(uint8_t* sums, uint32_t nslices) {
    const uint32_t w = blockIdx.x, side = blockIdx.y, lane = threadIdx.x;
    ge_p3 acc;
    ge_p3_0(acc);
    for (uint32_t sl = lane; sl < nslices; sl += 32) {
        ge_p3 t;
        ge_load(t, sums + (((size_t)side * nslices + sl) * kWin + w) * 128);
        ge_add(acc, acc, t);
    }
#pragma unroll 1
    for (int o = 16; o > 0; o >>= 1) {
        ge_p3 other;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            other.X.v[k] = __shfl_down_sync(0xffffffffu, acc.X.v[k], o);
            other.Y.v[k] = __shfl_down_sync(0xffffffffu, acc.Y.v[k], o);
            other.Z.v[k] = __shfl_down_sync(0xffffffffu, acc.Z.v[k], o);
            other.T.v[k] = __shfl_down_sync(0xffffffffu, acc.T.v[k], o);
        }
        ge_add(acc, acc, other);
    }
    if (lane == 0) ge_store(sums + ((size_t)side * nslices * kWin + w) * 128, acc);
}
// One CTA: sum_w 16^w S_w for L and R side by side as binary trees over the windows (octet form, fe8.cuh), both
// normalised into the proof, then the round challenge u = H(transcript || L.x || R.x) and u^-1
// (bulletproof_vectors.cu:448-477; ipa_round_challenge of the oracle).
__global__ void __launch_bounds__(1024) ipa_combine_kernel(const uint8_t* __restrict__ sums, uint32_t side_stride,
                                                           int round, uint8_t* __restrict__ tr,
                                                           uint8_t* __restrict__ L_out, uint8_t* __restrict__ R_out,
                                                           uint8_t* __restrict__ x_out, uint8_t* __restrict__ u_out,
                                                           uint8_t* __restrict__ uinv_out) {
    __shared__ __align__(16) uint8_t sh[2][64][128];
    const Fe8Lane L = fe8_lane();
    const uint32_t warp = threadIdx.x >> 5;
    int pp = 0, shift = 4;
    for (uint32_t per_side = kWin / 2; per_side >= 1; per_side >>= 1, shift <<= 1, pp ^= 1) {
        for (uint32_t task = warp; task < 2 * per_side; task += 32) {  // warp-uniform
            const uint32_t side = task / per_side, v = task % per_side;
            const uint8_t* src = per_side == kWin / 2 ? sums + (size_t)side * side_stride * 128 : &sh[pp ^ 1][side * 32][0];
            ge8 lo, hi;
            ge8_load(lo, src + (size_t)(2 * v) * 128, L);
            ge8_load(hi, src + (size_t)(2 * v + 1) * 128, L);
#pragma unroll 1
            for (int s = 0; s < shift; s++) ge8_dbl(hi, hi, L);
            ge8_add(hi, hi, lo, L);
            if (per_side > 1) ge8_store(&sh[pp][side * 32 + v][0], hi, L);
            else ge8_store_normalized(side ? R_out : L_out, hi, L);
        }
        __syncthreads();
    }
    if (threadIdx.x != 0) return;
    __threadfence_block();
    fe lx, rx;
    fe_load(lx, L_out);  // normalised: X is the canonical affine x
    fe_load(rx, R_out);
    sc t0;
    sc_load(t0, tr);
    Sha256 hsh;
    uint32_t ub[8];
    hsh.init();
    hsh.update_str("InnerProductChal", 16);
    hsh.update_words(t0.v);
    hsh.update_words(lx.v);
    hsh.update_words(rx.v);
    hsh.final_challenge(ub);
    sc raw, u, ui;
#pragma unroll
    for (int i = 0; i < 8; i++) raw.v[i] = ub[i];
    sc_store(tr, raw);
    if (round == 0) sc_store(x_out, raw);  // the first raw challenge is stored in the proof (:471-474)
    sc_reduce(u, raw);
    sc_invert(ui, u);
    sc_store(u_out, u);
    sc_store(uinv_out, ui);
}
// a' = u a_L + u^-1 a_R, b' = u^-1 b_L + u b_R (threads < np) and the weights of the base generators:
// wG_i *= (upper ? u : u^-1), wH_i *= (upper ? u^-1 : u) for generator i in the upper / lower half of its block
__global__ void __launch_bounds__(128) ipa_fold_kernel(uint8_t* a, uint8_t* b, uint8_t* wG, uint8_t* wH, uint32_t np,
                                                       uint32_t mcur, uint32_t mb, const uint8_t* __restrict__ u_p,
                                                       const uint8_t* __restrict__ ui_p) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    sc u, ui;
    sc_load(u, u_p);
    sc_load(ui, ui_p);
    if (t < mb) {
        const bool upper = (t % mcur) >= np;
        sc g, h;
        sc_load(g, wG + (size_t)t * 32);
        sc_load(h, wH + (size_t)t * 32);
        sc_mul(g, g, upper ? u : ui);
        sc_mul(h, h, upper ? ui : u);
        sc_store(wG + (size_t)t * 32, g);
        sc_store(wH + (size_t)t * 32, h);
    }
    if (t < np) {
        sc al, ar, bl, br, x, y, na, nb;
        sc_load(al, a + (size_t)t * 32);
        sc_load(ar, a + (size_t)(t + np) * 32);
        sc_load(bl, b + (size_t)t * 32);
        sc_load(br, b + (size_t)(t + np) * 32);
        sc_mul(x, u, al);
        sc_mul(y, ui, ar);
        sc_add(na, x, y);
        sc_mul(x, ui, bl);
        sc_mul(y, u, br);
        sc_add(nb, x, y);
        sc_store(a + (size_t)t * 32, na);  // thread t reads t and t + np, writes t only: safe in place
        sc_store(b + (size_t)t * 32, nb);
    }
}

static size_t align256(size_t x) { return (x + 255) / 256 * 256; }
struct CompositeLayout {
    size_t tables, wG, wH, digits, rows, sums, cross, total;
};
static CompositeLayout composite_layout(size_t mb) {
    CompositeLayout Ly;
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += align256(bytes);
        return o;
    };
    const size_t npts = mb + 1, nslices = (npts + kSlice - 1) / kSlice;
    Ly.tables = take((2 * mb + 1) * kTab * 96);
    Ly.wG = take(mb * 32);
    Ly.wH = take(mb * 32);
    Ly.digits = take(2 * (size_t)kWin * npts);
    Ly.rows = take(2 * npts * 4);
    Ly.sums = take(2 * nslices * kWin * 128);
    Ly.cross = take(64);
    Ly.total = off;
    return Ly;
}
size_t ipa_composite_workspace_bytes(size_t mb) { return composite_layout(mb).total; }

#define IPA_LAUNCHED()                                                  \
    do {                                                                \
        cudaError_t e_ = cudaGetLastError();                            \
        if (e_ != cudaSuccess) return cbp::fail(BPK_ERR_CUDA, e_);      \
        cbp::count_launches(1);                                         \
    } while (0)

// Rounds first_round .. first_round + log2(mb) - 1 of the argument over the mb base generators g, h (device arrays
// of ge25519) with the reduced vectors a, b (folded in place).  tr / u / ui: 32-byte device scratch (tr holds the
// running transcript).  Everything on `st`.
int ipa_prove_composite(uint8_t* a, uint8_t* b, const uint8_t* g, const uint8_t* h, const uint8_t* Q, size_t mb,
                        int first_round, uint8_t* tr, uint8_t* u, uint8_t* ui, uint8_t* d_L, uint8_t* d_R,
                        uint8_t* d_x_out, uint8_t* ws, cudaStream_t st) {
    if (mb < 2 || (mb & (mb - 1)) || mb > kIpaCompositeMax) return fail(BPK_ERR_ARG);
    const CompositeLayout Ly = composite_layout(mb);
    uint8_t *tables = ws + Ly.tables, *wG = ws + Ly.wG, *wH = ws + Ly.wH, *sums = ws + Ly.sums, *cL = ws + Ly.cross,
            *cR = ws + Ly.cross + 32;
    int8_t* digits = (int8_t*)(ws + Ly.digits);
    uint32_t* rows = (uint32_t*)(ws + Ly.rows);
    const uint32_t m = (uint32_t)mb, npts = m + 1, nslices = (npts + kSlice - 1) / kSlice;
    ipa_tables_kernel<<<(2 * m + 1 + 127) / 128, 128, 0, st>>>(g, h, Q, m, tables, wG, wH);
    IPA_LAUNCHED();
    int round = first_round;
    for (uint32_t mcur = m; mcur >= 2; mcur >>= 1, round++) {
        const uint32_t np = mcur >> 1;
        ipa_cross_kernel<<<1, 256, 0, st>>>(a, b, np, cL, cR);
        IPA_LAUNCHED();
        ipa_digits_kernel<<<(2 * npts + 127) / 128, 128, 0, st>>>(m, mcur, a, b, wG, wH, cL, cR, digits, rows);
        IPA_LAUNCHED();
        ipa_sums_kernel<<<dim3(kWin, nslices, 2), 32, 0, st>>>(tables, digits, rows, npts, kSlice, sums);
        IPA_LAUNCHED();
        if (nslices > 1) {
            ipa_slices_kernel<<<dim3(kWin, 2), 32, 0, st>>>(sums, nslices);
            IPA_LAUNCHED();
        }
        ipa_combine_kernel<<<1, 1024, 0, st>>>(sums, nslices * kWin, round, tr, d_L + (size_t)round * 128,
                                               d_R + (size_t)round * 128, d_x_out, u, ui);
        IPA_LAUNCHED();
        const uint32_t fthreads = m > np ? m : np;
        ipa_fold_kernel<<<(fthreads + 127) / 128, 128, 0, st>>>(a, b, wG, wH, np, mcur, m, u, ui);
        IPA_LAUNCHED();
    }
    return BPK_OK;
}

}  // namespace cbp
