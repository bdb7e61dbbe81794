// modinv.cuh — modular inversion by Bernstein-Yang divsteps ("safegcd"), variable time, one thread, 32-bit words.
//
// The Fermat chains of the first version (254 dependent squarings mod p: 50 us for a lone thread; 253 squarings and
// ~130 multiplications with Barrett reductions mod l: 277 us) were the last step of every normalised MSM and point
// sum and sat on the critical path of every round of the inner-product argument.  Here x^-1 mod M (M odd, < 2^256:
// p = 2^255 - 19 or the group order l) takes at most 20 batches of 30 divsteps on the low 32 bits of (f, g), each
// followed by one 2x2 matrix update of the full-width (f, g) and (d, e) held in nine signed 30-bit limbs.
// Inside a batch the steps are not taken one at a time: runs of even steps are skipped with a count-trailing-zeros,
// and while delta <= 0 up to eight low bits of g are cancelled at once with w = -g / f mod 2^L (g += w f; the
// matrix row follows) — about five loop iterations per 30 steps, bit-identical to the step-by-step recurrence
// (tools/modinv_model.py checks exactly that).
// Inputs are public (Z coordinates of results, Fiat-Shamir challenges): variable time is fine.
// The code compiles for the host as well (tests/c_abi/modinv_host.cpp: pinned against Python big integers on the CPU).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CBP_HD __host__ __device__ __forceinline__
#else
#define CBP_HD inline
#endif

namespace cbp {

static constexpr uint32_t kM30 = 0x3FFFFFFFu;
struct ModInfo {
    int32_t m[9];    // modulus, nine 30-bit limbs
    uint32_t inv30;  // m^-1 mod 2^30
};
// 256-bit little-endian words -> nine 30-bit limbs (limb 8 holds the top 16 bits)
CBP_HD void modinv_load(int32_t (&r)[9], const uint32_t (&w)[8]) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 9; i++) {
        const int bit = 30 * i, word = bit >> 5, sh = bit & 31;
        uint64_t v = w[word];
        if (word + 1 < 8) v |= (uint64_t)w[word + 1] << 32;
        r[i] = (int32_t)((uint32_t)(v >> sh) & kM30);
    }
}
// limbs in [0, 2^30) (value in [0, 2^256)) -> words
CBP_HD void modinv_store(uint32_t (&w)[8], const int32_t (&r)[9]) {
    uint64_t acc = 0;
    int bits = 0, limb = 0;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < 8; i++) {
        while (bits < 32 && limb < 9) {
            acc |= (uint64_t)(uint32_t)r[limb++] << bits;
            bits += 30;
        }
        w[i] = (uint32_t)acc;
        acc >>= 32;
        bits -= 32;
    }
}
CBP_HD ModInfo modinfo_from_words(const uint32_t (&mw)[8]) {
    ModInfo mi;
    modinv_load(mi.m, mw);
    uint32_t x = mw[0];  // m^-1 mod 2^32 by Newton iteration (m odd): correct to 3, 6, 12, 24, 48 bits
    for (int i = 0; i < 4; i++) x *= 2u - mw[0] * x;
    mi.inv30 = x & kM30;
    return mi;
}

struct Trans2x2 {
    int32_t u, v, q, r;
};
CBP_HD int modinv_ctz(uint32_t x) {
#if defined(__CUDA_ARCH__)
    return __ffs((int)x) - 1;
#else
    return __builtin_ctz(x);
#endif
}
// 30 divsteps on the low 32 bits of f (odd) and g; 2^30 [f'; g'] = t [f; g], |entries| <= 2^30
CBP_HD int32_t modinv_divsteps_30(int32_t delta, uint32_t f, uint32_t g, Trans2x2& t) {
    uint32_t u = 1, v = 0, q = 0, r = 1;  // matrix entries modulo 2^32 (they fit in int32)
    int i = 30;
    for (;;) {
        const int zeros = modinv_ctz(g | (1u << i));  // a run of even steps (at most the i that remain)
        g >>= zeros;
        u <<= zeros;
        v <<= zeros;
        delta += zeros;
        i -= zeros;
        if (i == 0) break;
        // g is odd.  delta > 0: the swapping step, written as a relabelling (f, g) <- (g, -f) that costs no step
        // and leaves delta <= 0
        if (delta > 0) {
            const uint32_t tf = f, tu = u, tv = v;
            delta = -delta;
            f = g; g = 0u - tf;
            u = q; v = r;
            q = 0u - tu; r = 0u - tv;
        }
        // delta <= 0: the next 1 - delta steps cannot swap; cancel L low bits of g at once
        int L = 1 - delta;
        L = L < i ? L : i;
        L = L < 8 ? L : 8;
        uint32_t finv = f;            // f^-1 mod 2^3
        finv *= 2u - f * finv;        // mod 2^6
        finv *= 2u - f * finv;        // mod 2^12
        const uint32_t w = (0u - g * finv) & ((1u << L) - 1u);
        g += f * w;
        q += u * w;
        r += v * w;
    }
    t.u = (int32_t)u; t.v = (int32_t)v; t.q = (int32_t)q; t.r = (int32_t)r;
    return delta;
}
// (f, g) <- t (f, g) / 2^30, exact
CBP_HD void modinv_update_fg(int32_t (&f)[9], int32_t (&g)[9], const Trans2x2& t) {
    int64_t cf = (int64_t)t.u * f[0] + (int64_t)t.v * g[0];
    int64_t cg = (int64_t)t.q * f[0] + (int64_t)t.r * g[0];
    cf >>= 30; cg >>= 30;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 9; i++) {
        cf += (int64_t)t.u * f[i] + (int64_t)t.v * g[i];
        cg += (int64_t)t.q * f[i] + (int64_t)t.r * g[i];
        f[i - 1] = (int32_t)((uint32_t)cf & kM30);
        g[i - 1] = (int32_t)((uint32_t)cg & kM30);
        cf >>= 30; cg >>= 30;
    }
    f[8] = (int32_t)cf;
    g[8] = (int32_t)cg;
}
// (d, e) <- t (d, e) / 2^30 mod m; inputs and outputs in (-2m, m)
CBP_HD void modinv_update_de(int32_t (&d)[9], int32_t (&e)[9], const Trans2x2& t, const ModInfo& mi) {
    const int32_t sd = d[8] >> 31, se = e[8] >> 31;
    int32_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    int64_t cd = (int64_t)t.u * d[0] + (int64_t)t.v * e[0];
    int64_t ce = (int64_t)t.q * d[0] + (int64_t)t.r * e[0];
    md -= (int32_t)((mi.inv30 * (uint32_t)cd + (uint32_t)md) & kM30);
    me -= (int32_t)((mi.inv30 * (uint32_t)ce + (uint32_t)me) & kM30);
    cd += (int64_t)mi.m[0] * md;
    ce += (int64_t)mi.m[0] * me;
    cd >>= 30; ce >>= 30;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 1; i < 9; i++) {
        cd += (int64_t)t.u * d[i] + (int64_t)t.v * e[i] + (int64_t)mi.m[i] * md;
        ce += (int64_t)t.q * d[i] + (int64_t)t.r * e[i] + (int64_t)mi.m[i] * me;
        d[i - 1] = (int32_t)((uint32_t)cd & kM30);
        e[i - 1] = (int32_t)((uint32_t)ce & kM30);
        cd >>= 30; ce >>= 30;
    }
    d[8] = (int32_t)cd;
    e[8] = (int32_t)ce;
}
// d += m (limbs renormalised)
CBP_HD void modinv_add_m(int32_t (&d)[9], const ModInfo& mi) {
    int32_t c = 0;
    for (int i = 0; i < 9; i++) {
        c += d[i] + mi.m[i];
        d[i] = i < 8 ? (int32_t)((uint32_t)c & kM30) : c;
        c >>= 30;
    }
}
// out = x^-1 mod m as 8 little-endian words in [0, m); 0 for x = 0 (mod m).  x: any 256-bit value.
CBP_HD void modinv_words(uint32_t (&out)[8], const uint32_t (&x)[8], const ModInfo& mi) {
    int32_t f[9], g[9], d[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, e[9] = {1, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 9; i++) f[i] = mi.m[i];
    modinv_load(g, x);  // g < 2^256 may exceed m: harmless, gcd(m, g) and the cofactors are the same
    int32_t delta = 1;
    for (int batch = 0; batch < 32; batch++) {  // 20 suffice for 256-bit inputs (tools/modinv_model.py); bounded anyway
        uint32_t nz = 0;
        for (int i = 0; i < 9; i++) nz |= (uint32_t)g[i];
        if (nz == 0) break;
        Trans2x2 t;
        delta = modinv_divsteps_30(delta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), t);
        modinv_update_de(d, e, t, mi);
        modinv_update_fg(f, g, t);
    }
    // f = +-1 unless gcd(x, m) != 1, i.e. x = 0 (mod m) for a prime m: the answer is then 0 by convention
    uint32_t rest = 0, rest_m1 = 0;
    for (int i = 1; i < 8; i++) {
        rest |= (uint32_t)f[i];
        rest_m1 |= (uint32_t)f[i] ^ kM30;
    }
    const bool plus = f[0] == 1 && rest == 0 && f[8] == 0;
    const bool minus = (uint32_t)f[0] == kM30 && rest_m1 == 0 && f[8] == -1;
    if (!plus && !minus) {
        for (int i = 0; i < 8; i++) out[i] = 0;
        return;
    }
    // d in (-2m, m): two conditional additions of m bring it into [0, m); then the inverse is sign(f) d mod m
    if (d[8] < 0) modinv_add_m(d, mi);
    if (d[8] < 0) modinv_add_m(d, mi);
    if (minus) {
        int32_t c = 0;
        for (int i = 0; i < 9; i++) {
            c -= d[i];
            d[i] = i < 8 ? (int32_t)((uint32_t)c & kM30) : c;
            c >>= 30;
        }
        if (d[8] < 0) modinv_add_m(d, mi);
    }
    modinv_store(out, d);
}

}  // namespace cbp
