// ipa.cu — inner-product-argument folding kernels, stand-alone IPA verification and the host-pointer
// drop-ins cuda_range_proof_verify / cuda_inner_product_verify (cuda_bulletproof.h:61,72).
//
//   fold scalars  bulletproof_vectors.cu:488-500  a' = u a_L + u^-1 a_R ; b' = u^-1 b_L + u b_R  (mod l, D11/D21)
//   fold points   bulletproof_vectors.cu:641-663  G'_j = u^-1 G_j + u G_{j+n'} ; H'_j = u H_j + u^-1 H_{j+n'}
//   verify        bulletproof_vectors.cu:541-762  P + sum(u_j^2 L_j + u_j^-2 R_j) == a G' + b H' + a b Q, evaluated
//                 as ONE Pippenger MSM over 2n + 2k + 2 points that must be the identity (no folding).
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "../../include/cuda_bulletproof.h"
#include "common.h"
#include "ipa_straus.h"
#include "msm.h"
#include "rangeproof.cuh"
#include "sha256.cuh"

namespace cbp {

// ---- folding --------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) ipa_fold_scalars_kernel(uint8_t* a_out, uint8_t* b_out, const uint8_t* a,
                                                               const uint8_t* b, size_t n_half,
                                                               const uint8_t* __restrict__ u_p,
                                                               const uint8_t* __restrict__ ui_p) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n_half) return;
    sc u, ui, al, ar, bl, br, t0, t1, na, nb;
    sc_load(t0, u_p);
    sc_reduce(u, t0);
    sc_load(t0, ui_p);
    sc_reduce(ui, t0);
    sc_load(t0, a + j * 32);
    sc_reduce(al, t0);
    sc_load(t0, a + (j + n_half) * 32);
    sc_reduce(ar, t0);
    sc_load(t0, b + j * 32);
    sc_reduce(bl, t0);
    sc_load(t0, b + (j + n_half) * 32);
    sc_reduce(br, t0);
    sc_mul(t0, u, al);
    sc_mul(t1, ui, ar);
    sc_add(na, t0, t1);
    sc_mul(t0, ui, bl);
    sc_mul(t1, u, br);
    sc_add(nb, t0, t1);
    sc_store(a_out + j * 32, na);  // j < n_half: safe in place (each thread reads j and j+n_half first)
    sc_store(b_out + j * 32, nb);
}
// G'_j = u^-1 G_j + u G_{j+n'} ; H'_j = u H_j + u^-1 H_{j+n'}: one double-scalar multiplication per output with shared
// doublings (Shamir).  Scalars >= l cannot occur (reduced); points may carry torsion: k < l is used as an integer
// exactly like the CPU double-and-add.
// Quad-cooperative: every output is a 253-step double-and-add chain, i.e. pure latency; four lanes share
// each point operation (ge_add_quad / ge_dbl_quad: 0.99 / 0.72 us instead of 2.2 / 1.7 us per step).  The scalars
// are the same for every output of a side, so all quads take the same branches.
__global__ void __launch_bounds__(128) ipa_fold_points_quad_kernel(uint8_t* G_out, uint8_t* H_out, const uint8_t* G,
                                                                   const uint8_t* H, size_t n_half,
                                                                   const uint8_t* __restrict__ u_p,
                                                                   const uint8_t* __restrict__ ui_p) {
    size_t id = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2;
    const bool live = id < 2 * n_half;
    if (!live) id = 2 * n_half - 1;  // keep the warp converged (full-mask shuffles inside the quad operations)
    bool is_h = id >= n_half;
    size_t j = is_h ? id - n_half : id;
    sc u, ui, t0;
    sc_load(t0, u_p);
    sc_reduce(u, t0);
    sc_load(t0, ui_p);
    sc_reduce(ui, t0);
    const uint8_t* src = is_h ? H : G;
    ge_p3 A, B, AB, acc;
    ge_load(A, src + j * 128);             // low half
    ge_load(B, src + (j + n_half) * 128);  // high half
    // G: u^-1 lo + u hi ; H: u lo + u^-1 hi
    const sc k1 = is_h ? u : ui, k2 = is_h ? ui : u;
    ge_add_quad(AB, A, B);
    ge_p3_0(acc);
#pragma unroll 1
    for (int i = 252; i >= 0; i--) {
        ge_dbl_quad(acc, acc);
        uint32_t b1 = (k1.v[i >> 5] >> (i & 31)) & 1, b2 = (k2.v[i >> 5] >> (i & 31)) & 1;
        // b1, b2 differ between the G and the H half only: pick the operand, one addition for the whole warp
        if (__any_sync(0xffffffffu, b1 | b2)) {
            ge_p3 X = (b1 & b2) ? AB : b1 ? A : B, sum;
            ge_add_quad(sum, acc, X);
            if (b1 | b2) acc = sum;
        }
    }
    if (live && (threadIdx.x & 3) == 0) {
        ge_normalize(acc);
        ge_store((is_h ? H_out : G_out) + j * 128, acc);
    }
}

// ---- inner-product argument, prover side, any power-of-two width (inner_product_prove, ------------------
// bulletproof_vectors.cu:375-509 restated; BASELINE config 4 is n = 4096, 12 rounds) -------------------------
// Per round: c_L, c_R (mod-l inner products), L and R as ONE Pippenger MSM each over 2 n' + 1 points
// (a_L x G_R, b_R x H_L, c_L x Q), the round challenge and its inverse, then the a/b and G/H folds above.
__global__ void ipa_prove_setup_kernel(uint8_t* __restrict__ a, uint8_t* __restrict__ b, uint8_t* __restrict__ g,
                                       uint8_t* __restrict__ h, const uint8_t* __restrict__ a_in,
                                       const uint8_t* __restrict__ b_in, const uint8_t* __restrict__ G,
                                       const uint8_t* __restrict__ H, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    sc t, r;
    sc_load(t, a_in + i * 32);
    sc_reduce(r, t);
    sc_store(a + i * 32, r);
    sc_load(t, b_in + i * 32);
    sc_reduce(r, t);
    sc_store(b + i * 32, r);
    const uint4* gs = reinterpret_cast<const uint4*>(G + i * 128);
    const uint4* hs = reinterpret_cast<const uint4*>(H + i * 128);
    uint4* gd = reinterpret_cast<uint4*>(g + i * 128);
    uint4* hd = reinterpret_cast<uint4*>(h + i * 128);
    for (int q = 0; q < 8; q++) {
        gd[q] = gs[q];
        hd[q] = hs[q];
    }
}
// side 0: L = <a_L, G_R> + <b_R, H_L> + c_L Q ; side 1: R = <a_R, G_L> + <b_L, H_R> + c_R Q
__global__ void ipa_prove_gather_kernel(int side, size_t np, const uint8_t* __restrict__ a, const uint8_t* __restrict__ b,
                                        const uint8_t* __restrict__ g, const uint8_t* __restrict__ h,
                                        const uint8_t* __restrict__ Q, const uint8_t* __restrict__ c_side,
                                        uint8_t* __restrict__ scal, uint8_t* __restrict__ pts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > np) return;
    auto copy32 = [](uint8_t* d, const uint8_t* s_) {
        const uint4* sp = reinterpret_cast<const uint4*>(s_);
        uint4* dp = reinterpret_cast<uint4*>(d);
        dp[0] = sp[0];
        dp[1] = sp[1];
    };
    auto copy128 = [](uint8_t* d, const uint8_t* s_) {
        const uint4* sp = reinterpret_cast<const uint4*>(s_);
        uint4* dp = reinterpret_cast<uint4*>(d);
        for (int q = 0; q < 8; q++) dp[q] = sp[q];
    };
    if (i == np) {
        copy32(scal + 2 * np * 32, c_side);
        copy128(pts + 2 * np * 128, Q);
        return;
    }
    copy32(scal + i * 32, a + (side ? np + i : i) * 32);
    copy32(scal + (np + i) * 32, b + (side ? i : np + i) * 32);
    copy128(pts + i * 128, g + (side ? i : np + i) * 128);
    copy128(pts + (np + i) * 128, h + (side ? np + i : i) * 128);
}
__global__ void ipa_prove_challenge_kernel(int round, uint8_t* __restrict__ tr, const uint8_t* __restrict__ Lr,
                                           const uint8_t* __restrict__ Rr, uint8_t* __restrict__ x_out,
                                           uint8_t* __restrict__ u_out, uint8_t* __restrict__ uinv_out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    fe lx, rx;
    fe_load(lx, Lr);  // L, R are normalised: X is the affine x
    fe_load(rx, Rr);
    fe_canon(lx);
    fe_canon(rx);
    sc t0;
    sc_load(t0, tr);
    Sha256 sh;
    uint32_t ub[8];
    sh.init();
    sh.update_str("InnerProductChal", 16);
    sh.update_words(t0.v);
    sh.update_words(lx.v);
    sh.update_words(rx.v);
    sh.final_challenge(ub);
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

// ---- stand-alone IPA verification -------------------------------------------------------------------
struct IpaChal {
    uint32_t valid;
    uint32_t pad[7];
    sc a, b, ab;
    sc u[32], uinv[32], usq[32], uinvsq[32];
};
// one thread: transcript, inversions (k <= 32)
__global__ void ipa_verify_challenges_kernel(const uint8_t* __restrict__ L, const uint8_t* __restrict__ R, int k,
                                             const uint8_t* __restrict__ a_p, const uint8_t* __restrict__ b_p,
                                             const uint8_t* __restrict__ x_p, const uint8_t* __restrict__ tr0,
                                             IpaChal* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    bool valid = true;
    uint32_t tr[8];
    for (int i = 0; i < 8; i++)
        tr[i] = (uint32_t)tr0[4 * i] | ((uint32_t)tr0[4 * i + 1] << 8) | ((uint32_t)tr0[4 * i + 2] << 16) |
                ((uint32_t)tr0[4 * i + 3] << 24);
    Sha256 sh;
    sc run;
    sc_set1(run);
    for (int j = 0; j < k; j++) {
        ge_p3 Lp, Rp;
        ge_load(Lp, L + (size_t)j * 128);
        ge_load(Rp, R + (size_t)j * 128);
        valid = valid && ge_is_on_curve(Lp) && ge_is_on_curve(Rp);
        ge_normalize(Lp);
        ge_normalize(Rp);
        sh.init();
        sh.update_str("InnerProductChal", 16);
        sh.update_words(tr);
        sh.update_words(Lp.X.v);
        sh.update_words(Rp.X.v);
        sh.final_challenge(tr);
        if (j == 0) {
            fe xs;
            fe_load(xs, x_p);
            fe_canon(xs);
            for (int i = 0; i < 8; i++) valid = valid && (xs.v[i] == tr[i]);
        }
        sc t;
        for (int i = 0; i < 8; i++) t.v[i] = tr[i];
        sc_reduce(out->u[j], t);
        out->uinv[j] = run;  // prefix product, fixed up below
        sc_mul(run, run, out->u[j]);
    }
    sc inv;
    sc_invert(inv, run);
    for (int j = k - 1; j >= 0; j--) {
        sc pre = out->uinv[j], ui;
        sc_mul(ui, inv, pre);
        sc_mul(inv, inv, out->u[j]);
        out->uinv[j] = ui;
        sc_mul(out->usq[j], out->u[j], out->u[j]);
        sc_mul(out->uinvsq[j], ui, ui);
    }
    sc t;
    sc_load(t, a_p);
    sc_reduce(out->a, t);
    sc_load(t, b_p);
    sc_reduce(out->b, t);
    sc_mul(out->ab, out->a, out->b);
    out->valid = valid ? 1u : 0u;
}
// builds the MSM instance: [a s_i] G_i, [b s_i^-1] H_i, [ab] Q, [u_j^2] (-L_j), [u_j^-2] (-R_j), [1] (-P)
__global__ void __launch_bounds__(128) ipa_verify_assemble_kernel(const IpaChal* __restrict__ ch, size_t n, int k,
                                                                  const uint8_t* __restrict__ G,
                                                                  const uint8_t* __restrict__ H,
                                                                  const uint8_t* __restrict__ Q,
                                                                  const uint8_t* __restrict__ L,
                                                                  const uint8_t* __restrict__ R,
                                                                  const uint8_t* __restrict__ P,
                                                                  uint8_t* __restrict__ scalars,
                                                                  uint8_t* __restrict__ points) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = 2 * n + 2 * (size_t)k + 2;
    if (id >= total) return;
    sc s;
    ge_p3 pt;
    if (id < 2 * n) {
        bool is_h = id >= n;
        size_t i = is_h ? id - n : id;
        sc_set1(s);
        for (int j = 0; j < k; j++) {
            bool bit = (i >> (k - 1 - j)) & 1;
            sc_mul(s, s, (bit != is_h) ? ch->u[j] : ch->uinv[j]);  // H uses the inverse exponents
        }
        sc_mul(s, s, is_h ? ch->b : ch->a);
        ge_load(pt, (is_h ? H : G) + i * 128);
    } else if (id == 2 * n) {
        s = ch->ab;
        ge_load(pt, Q);
    } else if (id < 2 * n + 1 + 2 * (size_t)k) {
        size_t j = id - 2 * n - 1;
        bool is_r = j >= (size_t)k;
        if (is_r) j -= k;
        s = is_r ? ch->uinvsq[j] : ch->usq[j];
        ge_p3 t;
        ge_load(t, (is_r ? R : L) + j * 128);
        ge_neg(pt, t);
    } else {
        sc_set1(s);
        ge_p3 t;
        ge_load(t, P);
        ge_neg(pt, t);
    }
    sc_store(scalars + id * 32, s);
    ge_store(points + id * 128, pt);
}
__global__ void ipa_verify_decide_kernel(const IpaChal* __restrict__ ch, const uint8_t* __restrict__ result,
                                         uint8_t* __restrict__ accept) {
    if (threadIdx.x || blockIdx.x) return;
    ge_p3 r;
    ge_load(r, result);
    accept[0] = (ch->valid && ge_is_identity(r)) ? 1 : 0;
}

}  // namespace cbp

using namespace cbp;

extern "C" {

int bpk_ipa_fold_scalars_device(void* d_a_out, void* d_b_out, const void* d_a, const void* d_b, size_t n_half,
                                const void* d_u, const void* d_u_inv, void* stream) {
    if (!n_half) return BPK_OK;
    if (!d_a_out || !d_b_out || !d_a || !d_b || !d_u || !d_u_inv) return fail(BPK_ERR_ARG);
    ipa_fold_scalars_kernel<<<(unsigned)((n_half + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (uint8_t*)d_a_out, (uint8_t*)d_b_out, (const uint8_t*)d_a, (const uint8_t*)d_b, n_half, (const uint8_t*)d_u,
        (const uint8_t*)d_u_inv);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}
int bpk_ipa_fold_points_device(void* d_G_out, void* d_H_out, const void* d_G, const void* d_H, size_t n_half,
                               const void* d_u, const void* d_u_inv, void* stream) {
    if (!n_half) return BPK_OK;
    if (!d_G_out || !d_H_out || !d_G || !d_H || !d_u || !d_u_inv) return fail(BPK_ERR_ARG);
    if (d_G_out == d_G || d_H_out == d_H) return fail(BPK_ERR_ARG);  // outputs must not alias inputs
    ipa_fold_points_quad_kernel<<<(unsigned)((8 * n_half + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
        (uint8_t*)d_G_out, (uint8_t*)d_H_out, (const uint8_t*)d_G, (const uint8_t*)d_H, n_half, (const uint8_t*)d_u,
        (const uint8_t*)d_u_inv);
    CBP_CHECK_LAUNCH();
    return BPK_OK;
}

static size_t ipa_align(size_t x) { return (x + 255) / 256 * 256; }
struct IpaSide {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    bool ok = false;
};
static IpaSide g_ipa_side[kMaxDevices];
// caller holds the device lock
static IpaSide* ipa_side(int dev) {
    if (dev < 0 || dev >= kMaxDevices) return nullptr;
    IpaSide& v = g_ipa_side[dev];
    if (!v.ok) {
        if (cudaStreamCreateWithFlags(&v.stream, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&v.ev_fork, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        if (cudaEventCreateWithFlags(&v.ev_join, cudaEventDisableTiming) != cudaSuccess) return nullptr;
        v.ok = true;
    }
    return &v;
}
// longest vector whose rounds run unfolded: kIpaCompositeMax unless a test lowered it (BPK_OPT_IPA_COMPOSITE_MAX)
static size_t ipa_comp_max() {
    const int o = options().ipa_composite_max;
    return (o >= 2 && (size_t)o <= kIpaCompositeMax && !(o & (o - 1))) ? (size_t)o : kIpaCompositeMax;
}
struct IpaProveLayout {
    size_t a, b, g, h, scal[2], pts[2], small, ipws, msmws[2], comp, total;
};
static int ipa_prove_layout(size_t n, IpaProveLayout* L) {
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += ipa_align(bytes);
        return o;
    };
    L->a = take(n * 32);
    L->b = take(n * 32);
    L->g = take(n * 128);
    L->h = take(n * 128);
    for (int side = 0; side < 2; side++) {  // L and R are multiplied concurrently: own inputs, own workspace
        L->scal[side] = take((n + 1) * 32);
        L->pts[side] = take((n + 1) * 128);
    }
    L->small = take(8 * 32);  // tr | u | uinv | cL | cR
    size_t ipb = 0;
    if (bpk_sc_inner_product_workspace_bytes(n / 2 ? n / 2 : 1, &ipb) != BPK_OK) return BPK_ERR_ARG;
    L->ipws = take(ipb + 256);
    MsmPlan p;
    msm_make_plan(&p, n + 1, 0);
    size_t msm_bytes = p.workspace_bytes;
    for (size_t m = n; m >= 2; m >>= 1) {  // every round's plan must fit (window width varies with the size)
        msm_make_plan(&p, m + 1, 0);
        if (p.workspace_bytes > msm_bytes) msm_bytes = p.workspace_bytes;
    }
    if (n <= ipa_comp_max()) msm_bytes = 0;  // every round is a composite round: no MSM workspace
    L->msmws[0] = take(msm_bytes);
    L->msmws[1] = take(msm_bytes);
    L->comp = take(ipa_composite_workspace_bytes(n < ipa_comp_max() ? n : ipa_comp_max()));
    L->total = off;
    return BPK_OK;
}
int bpk_ipa_prove_workspace_bytes(size_t n, size_t* bytes) {
    if (!bytes || n < 2 || (n & (n - 1)) || n > ((size_t)1 << 24)) return fail(BPK_ERR_ARG);
    IpaProveLayout L;
    if (ipa_prove_layout(n, &L) != BPK_OK) return fail(BPK_ERR_ARG);
    *bytes = L.total;
    return BPK_OK;
}
int bpk_ipa_prove_device(const void* d_G, const void* d_H, const void* d_Q, const void* d_a, const void* d_b, size_t n,
                         const uint8_t transcript0[32], void* d_L, void* d_R, void* d_a_out, void* d_b_out, void* d_x_out,
                         void* d_workspace, size_t workspace_bytes, void* stream) {
    if (n < 2 || (n & (n - 1)) || n > ((size_t)1 << 24)) return fail(BPK_ERR_ARG);
    if (!d_G || !d_H || !d_Q || !d_a || !d_b || !transcript0 || !d_L || !d_R || !d_a_out || !d_b_out || !d_x_out ||
        !d_workspace)
        return fail(BPK_ERR_ARG);
    IpaProveLayout Ly;
    if (ipa_prove_layout(n, &Ly) != BPK_OK) return fail(BPK_ERR_ARG);
    if (workspace_bytes < Ly.total) return fail(BPK_ERR_WORKSPACE);
    cudaStream_t st = (cudaStream_t)stream;
    uint8_t* ws = (uint8_t*)d_workspace;
    uint8_t *a = ws + Ly.a, *b = ws + Ly.b, *g = ws + Ly.g, *h = ws + Ly.h;
    DeviceLock dlock;  // side stream / events shared per device
    if (!dlock.ok()) return BPK_ERR_CUDA;
    IpaSide* sd = ipa_side(dlock.dev);
    if (!sd) return fail(BPK_ERR_CUDA);
    uint8_t *tr = ws + Ly.small, *u = tr + 32, *ui = tr + 64, *cL = tr + 96, *cR = tr + 128;
    size_t ipb = 0;
    bpk_sc_inner_product_workspace_bytes(n / 2, &ipb);
    CBP_CUDA(cudaMemcpyAsync(tr, transcript0, 32, cudaMemcpyHostToDevice, st));
    ipa_prove_setup_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(a, b, g, h, (const uint8_t*)d_a, (const uint8_t*)d_b,
                                                                      (const uint8_t*)d_G, (const uint8_t*)d_H, n);
    CBP_CHECK_LAUNCH();
    const size_t comp_max = ipa_comp_max();
    int round = 0;
    // Rounds over more than kIpaCompositeMax generators fold the points and run Pippenger MSMs; from there on the
    // rounds are written over the (by then folded) generators without folding a point again (ipa_straus.cu).
    for (size_t np = n >> 1; np >= 1; np >>= 1, round++) {
        int rc;
        if (2 * np <= comp_max) {
            if ((rc = ipa_prove_composite(a, b, g, h, (const uint8_t*)d_Q, 2 * np, round, tr, u, ui, (uint8_t*)d_L,
                                          (uint8_t*)d_R, (uint8_t*)d_x_out, ws + Ly.comp, st)) != BPK_OK)
                return rc;
            break;
        }
        if ((rc = bpk_sc_inner_product_device(cL, a, b + np * 32, np, ws + Ly.ipws, ipb + 256, st)) != BPK_OK) return rc;
        if ((rc = bpk_sc_inner_product_device(cR, a + np * 32, b, np, ws + Ly.ipws, ipb + 256, st)) != BPK_OK) return rc;
        MsmPlan p;
        msm_make_plan(&p, 2 * np + 1, 0);
        // L on the caller's stream, R on a side stream with its own workspace and side-stream set
        CBP_CUDA(cudaEventRecord(sd->ev_fork, st));
        CBP_CUDA(cudaStreamWaitEvent(sd->stream, sd->ev_fork, 0));
        for (int side = 0; side < 2; side++) {
            cudaStream_t ms = side ? sd->stream : st;
            uint8_t *scal = ws + Ly.scal[side], *pts = ws + Ly.pts[side];
            ipa_prove_gather_kernel<<<(unsigned)((np + 1 + 127) / 128), 128, 0, ms>>>(side, np, a, b, g, h, (const uint8_t*)d_Q,
                                                                                    side ? cR : cL, scal, pts);
            CBP_CHECK_LAUNCH();
            int nl = 0;
            uint8_t* out = (uint8_t*)(side ? d_R : d_L) + (size_t)round * 128;
            int mrc = msm_run(p, scal, pts, out, ws + Ly.msmws[side], 1, ms, &nl, nullptr, side);
            count_launches(nl);
            if (mrc) return fail_cuda(mrc);
        }
        CBP_CUDA(cudaEventRecord(sd->ev_join, sd->stream));
        CBP_CUDA(cudaStreamWaitEvent(st, sd->ev_join, 0));
        ipa_prove_challenge_kernel<<<1, 32, 0, st>>>(round, tr, (const uint8_t*)d_L + (size_t)round * 128,
                                                     (const uint8_t*)d_R + (size_t)round * 128, (uint8_t*)d_x_out, u, ui);
        CBP_CHECK_LAUNCH();
        ipa_fold_scalars_kernel<<<(unsigned)((np + 127) / 128), 128, 0, st>>>(a, b, a, b, np, u, ui);
        CBP_CHECK_LAUNCH();
        // in place: thread j reads g[j], g[j + n'] and writes g[j] only
        ipa_fold_points_quad_kernel<<<(unsigned)((8 * np + 127) / 128), 128, 0, st>>>(g, h, g, h, np, u, ui);
        CBP_CHECK_LAUNCH();
    }
    CBP_CUDA(cudaMemcpyAsync(d_a_out, a, 32, cudaMemcpyDeviceToDevice, st));
    CBP_CUDA(cudaMemcpyAsync(d_b_out, b, 32, cudaMemcpyDeviceToDevice, st));
    return BPK_OK;
}

// ---- host-pointer drop-ins ----------------------------------------------------------------------------
// The reference's wrappers malloc / copy / synchronise / free on every call (cuda_bulletproof_kernels.cu:77-115).
// Here every device keeps one grow-only device buffer, one pinned staging buffer, one stream and the generator
// tables of the last generator set: a call is one upload, the kernels, one byte back, no allocation.
namespace {
struct HostVerify {
    cudaStream_t st = nullptr;
    uint8_t* d_buf = nullptr;
    size_t d_cap = 0;
    uint8_t* h_pin = nullptr;
    size_t h_cap = 0;
    // generator tables are cached across calls: the reference signature passes G, H, g, h every time
    uint8_t* gens_dev = nullptr;
    uint8_t* gens_key = nullptr;
    size_t gens_key_bytes = 0, gens_n = 0;
};
HostVerify g_hv[kMaxDevices];

// caller holds the device lock
bool hv_reserve(HostVerify& hv, size_t dev_bytes, size_t pin_bytes) {
    cudaError_t e = cudaSuccess;
    if (!hv.st) e = cudaStreamCreateWithFlags(&hv.st, cudaStreamNonBlocking);
    if (e == cudaSuccess && dev_bytes > hv.d_cap) {
        if (hv.d_buf) cudaFree(hv.d_buf);
        hv.d_buf = nullptr;
        hv.d_cap = 0;
        if ((e = cudaMalloc(&hv.d_buf, dev_bytes)) == cudaSuccess) hv.d_cap = dev_bytes;
    }
    if (e == cudaSuccess && pin_bytes > hv.h_cap) {
        if (hv.h_pin) cudaFreeHost(hv.h_pin);
        hv.h_pin = nullptr;
        hv.h_cap = 0;
        if ((e = cudaMallocHost(&hv.h_pin, pin_bytes)) == cudaSuccess) hv.h_cap = pin_bytes;
    }
    if (e != cudaSuccess) fail(BPK_ERR_CUDA, e);
    return e == cudaSuccess;
}
}  // namespace

static bool ipa_verify_host(const InnerProductProof* proof, const ge25519* P, const PointVector* G,
                            const PointVector* H, const ge25519* Q, const uint8_t tr0[32]) {
    size_t n = proof->n;
    if (G->length != n || H->length != n) return false;  // bulletproof_vectors.cu:550-556
    if (n == 0 || (n & (n - 1))) return false;
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    if (proof->L_len != (size_t)k || proof->L.length != (size_t)k || proof->R.length != (size_t)k) return false;
    if (proof->a.length < 1 || proof->b.length < 1 || k > 30) return false;
    size_t total = 2 * n + 2 * (size_t)k + 2;
    MsmPlan plan;
    msm_make_plan(&plan, total, 0);
    // one contiguous upload: G | H | Q | P | L | R | a, b, x, transcript
    const size_t off_G = 0, off_H = off_G + n * 128, off_Q = off_H + n * 128, off_P = off_Q + 128, off_L = off_P + 128,
                 off_R = off_L + (size_t)k * 128, off_s = off_R + (size_t)k * 128, up_bytes = off_s + 128;
    const size_t off_ch = ipa_align(up_bytes), off_sc = off_ch + ipa_align(sizeof(IpaChal)), off_pt = off_sc + ipa_align(total * 32),
                 off_res = off_pt + total * 128, off_ws = off_res + 256, bytes = off_ws + plan.workspace_bytes;
    DeviceLock dlock;
    if (!dlock.ok()) return false;
    HostVerify& hv = g_hv[dlock.dev];
    if (!hv_reserve(hv, bytes, up_bytes)) return false;
    uint8_t *d = hv.d_buf, *hp = hv.h_pin;
    memcpy(hp + off_G, G->elements, n * 128);
    memcpy(hp + off_H, H->elements, n * 128);
    memcpy(hp + off_Q, Q, 128);
    memcpy(hp + off_P, P, 128);
    if (k) {
        memcpy(hp + off_L, proof->L.elements, (size_t)k * 128);
        memcpy(hp + off_R, proof->R.elements, (size_t)k * 128);
    }
    memcpy(hp + off_s, proof->a.elements, 32);
    memcpy(hp + off_s + 32, proof->b.elements, 32);
    memcpy(hp + off_s + 64, &proof->x, 32);
    memcpy(hp + off_s + 96, tr0, 32);
    cudaStream_t st = hv.st;
    cudaError_t e;
    bool ok = false;
    do {
        if ((e = cudaMemcpyAsync(d, hp, up_bytes, cudaMemcpyHostToDevice, st)) != cudaSuccess) break;
        IpaChal* ch = (IpaChal*)(d + off_ch);
        ipa_verify_challenges_kernel<<<1, 32, 0, st>>>(d + off_L, d + off_R, k, d + off_s, d + off_s + 32, d + off_s + 64,
                                                       d + off_s + 96, ch);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        ipa_verify_assemble_kernel<<<(unsigned)((total + 127) / 128), 128, 0, st>>>(ch, n, k, d + off_G, d + off_H,
                                                                                   d + off_Q, d + off_L, d + off_R,
                                                                                   d + off_P, d + off_sc, d + off_pt);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        count_launches(2);
        int launches = 0;
        int rc = msm_run(plan, d + off_sc, d + off_pt, d + off_res, d + off_ws, 0, st, &launches, nullptr);
        count_launches(launches);
        if (rc) {
            e = (cudaError_t)rc;
            break;
        }
        ipa_verify_decide_kernel<<<1, 32, 0, st>>>(ch, d + off_res, d + off_res + 128);
        if ((e = cudaGetLastError()) != cudaSuccess) break;
        count_launches(1);
        if ((e = cudaMemcpyAsync(hp, d + off_res + 128, 1, cudaMemcpyDeviceToHost, st)) != cudaSuccess) break;
        if ((e = cudaStreamSynchronize(st)) != cudaSuccess) break;
        ok = hp[0] != 0;
    } while (0);
    if (e != cudaSuccess) fail(BPK_ERR_CUDA, e);
    return ok;
}

bool cuda_inner_product_verify(const InnerProductProof* proof, const ge25519* P, const PointVector* G,
                               const PointVector* H, const ge25519* Q) {
    uint8_t zero[32] = {0};  // bulletproof_vectors.cu:589
    return ipa_verify_host(proof, P, G, H, Q, zero);
}

// caller holds the device lock
static const uint8_t* gens_for(HostVerify& hv, size_t n, const PointVector* G, const PointVector* H, const ge25519* g,
                               const ge25519* h) {
    size_t key_bytes = (2 * n + 2) * 128;
    if (hv.gens_dev && hv.gens_n == n && hv.gens_key_bytes == key_bytes && memcmp(G->elements, hv.gens_key, n * 128) == 0 &&
        memcmp(H->elements, hv.gens_key + n * 128, n * 128) == 0 && memcmp(g, hv.gens_key + 2 * n * 128, 128) == 0 &&
        memcmp(h, hv.gens_key + 2 * n * 128 + 128, 128) == 0)
        return hv.gens_dev;
    if (hv.gens_dev) {
        cudaStreamSynchronize(hv.st);  // no call is in flight (the lock is held), but be explicit
        cudaFree(hv.gens_dev);
    }
    free(hv.gens_key);
    hv.gens_dev = nullptr;
    hv.gens_key = nullptr;
    uint8_t* key = (uint8_t*)malloc(key_bytes);
    if (!key) return nullptr;
    memcpy(key, G->elements, n * 128);
    memcpy(key + n * 128, H->elements, n * 128);
    memcpy(key + 2 * n * 128, g, 128);
    memcpy(key + 2 * n * 128 + 128, h, 128);
    size_t ws = 0;
    uint8_t *d_in = nullptr, *tab = nullptr;
    cudaError_t e = cudaSuccess;
    int rc = bpk_gens_workspace_bytes(n, &ws);
    if (rc == BPK_OK && (e = cudaMalloc(&tab, ws)) == cudaSuccess && (e = cudaMalloc(&d_in, key_bytes)) == cudaSuccess &&
        (e = cudaMemcpyAsync(d_in, key, key_bytes, cudaMemcpyHostToDevice, hv.st)) == cudaSuccess) {
        rc = bpk_gens_init_device(tab, ws, d_in, d_in + n * 128, d_in + 2 * n * 128, d_in + 2 * n * 128 + 128, n, hv.st);
        e = cudaStreamSynchronize(hv.st);
    }
    if (d_in) cudaFree(d_in);
    if (rc != BPK_OK || e != cudaSuccess) {
        if (e != cudaSuccess) fail(BPK_ERR_CUDA, e);
        if (tab) cudaFree(tab);
        free(key);
        return nullptr;
    }
    hv.gens_dev = tab;
    hv.gens_key = key;
    hv.gens_key_bytes = key_bytes;
    hv.gens_n = n;
    return hv.gens_dev;
}

bool cuda_range_proof_verify(const RangeProof* proof, const ge25519* V, size_t n, const PointVector* G,
                             const PointVector* H, const ge25519* g, const ge25519* h) {
    // structural checks of the CPU verifier (nb:6669-6672; bulletproof_vectors.cu:550-556)
    if (G->length != n || H->length != n || proof->ip_proof.n != n) return false;
    if (n == 0 || n > (size_t)kMaxN || (n & (n - 1))) return false;
    int k = 0;
    while (((size_t)1 << k) < n) k++;
    const InnerProductProof* ip = &proof->ip_proof;
    if (ip->L_len != (size_t)k || ip->L.length != (size_t)k || ip->R.length != (size_t)k) return false;
    if (ip->a.length < 1 || ip->b.length < 1) return false;
    size_t rec = proof_record_bytes(k), ws = 0;
    if (bpk_range_verify_workspace_bytes(n, 1, &ws) != BPK_OK) return false;
    // device buffer: record | V | accept | workspace; staging buffer: record | V
    const size_t off_V = ipa_align(rec), off_acc = off_V + 256, off_ws = off_acc + 256;
    DeviceLock dlock;
    if (!dlock.ok()) return false;
    HostVerify& hv = g_hv[dlock.dev];
    if (!hv_reserve(hv, off_ws + ws, off_V + 128)) return false;
    const uint8_t* gens = gens_for(hv, n, G, H, g, h);
    if (!gens) return false;
    uint8_t *hrec = hv.h_pin, *d = hv.d_buf;
    memset(hrec, 0, off_V + 128);
    memcpy(hrec + kRecV, &proof->V, 5 * 128 + 3 * 32);
    memcpy(hrec + kRecIpA, ip->a.elements, 32);
    memcpy(hrec + kRecIpB, ip->b.elements, 32);
    memcpy(hrec + kRecIpC, &ip->c, 32);
    memcpy(hrec + kRecIpX, &ip->x, 32);
    if (k) {
        memcpy(hrec + kRecL, ip->L.elements, (size_t)k * 128);
        memcpy(hrec + kRecL + (size_t)k * 128, ip->R.elements, (size_t)k * 128);
    }
    memcpy(hrec + off_V, V, 128);
    cudaError_t e = cudaMemcpyAsync(d, hrec, off_V + 128, cudaMemcpyHostToDevice, hv.st);
    bool ok = false;
    if (e == cudaSuccess) {
        int rc = bpk_range_verify_batch_device(gens, d, d + off_V, n, 1, d + off_acc, d + off_ws, ws, hv.st);
        if (rc == BPK_OK) e = cudaMemcpyAsync(hrec, d + off_acc, 1, cudaMemcpyDeviceToHost, hv.st);
        if (rc == BPK_OK && e == cudaSuccess) e = cudaStreamSynchronize(hv.st);
        ok = rc == BPK_OK && e == cudaSuccess && hrec[0] != 0;
    }
    if (e != cudaSuccess) fail(BPK_ERR_CUDA, e);
    return ok;
}

}  // extern "C"
