// modinv.cuh — modular inversion by Bernstein-Yang divsteps ("safegcd"), variable time, one thread.
//
// The Fermat chain of fe_invert / sc_invert is 254 dependent squarings: 50 us for a lone thread, and it is the last
// step of every normalised MSM, of every point sum, of every round challenge of the inner-product argument.  Here
// x^-1 mod M (M odd, < 2^256: p = 2^255 - 19 or the group order l) takes at most 9 batches of 62 divsteps on the low
// 64 bits of (f, g), each followed by one 2x2 matrix update of the full-width (f, g) and (d, e): ~6 k instructions.
// Inputs are public (Z coordinates of results, Fiat-Shamir challenges): variable time is fine.
// The code compiles for the host as well (tests/c_abi/modinv_host.cpp: pinned against Python big integers on the CPU);
// tools/modinv_model.py is the same algorithm with every range assertion spelled out.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define CBP_HD __host__ __device__ __forceinline__
#else
#define CBP_HD inline
#endif

namespace cbp {

struct ModInfo {
    int64_t m[5];    // modulus, five 62-bit limbs
    uint64_t inv62;  // m^-1 mod 2^62
};
typedef __int128 cbp_i128;
static constexpr uint64_t kM62 = 0x3FFFFFFFFFFFFFFFull;

// 256-bit little-endian words -> five 62-bit limbs
CBP_HD void modinv_load(int64_t (&r)[5], const uint32_t (&w)[8]) {
    const uint64_t a0 = (uint64_t)w[0] | ((uint64_t)w[1] << 32), a1 = (uint64_t)w[2] | ((uint64_t)w[3] << 32),
                   a2 = (uint64_t)w[4] | ((uint64_t)w[5] << 32), a3 = (uint64_t)w[6] | ((uint64_t)w[7] << 32);
    r[0] = (int64_t)(a0 & kM62);
    r[1] = (int64_t)(((a0 >> 62) | (a1 << 2)) & kM62);
    r[2] = (int64_t)(((a1 >> 60) | (a2 << 4)) & kM62);
    r[3] = (int64_t)(((a2 >> 58) | (a3 << 6)) & kM62);
    r[4] = (int64_t)(a3 >> 56);
}
// limbs in [0, 2^62) (value in [0, 2^256)) -> words
CBP_HD void modinv_store(uint32_t (&w)[8], const int64_t (&r)[5]) {
    const uint64_t l0 = (uint64_t)r[0], l1 = (uint64_t)r[1], l2 = (uint64_t)r[2], l3 = (uint64_t)r[3], l4 = (uint64_t)r[4];
    const uint64_t a0 = l0 | (l1 << 62), a1 = (l1 >> 2) | (l2 << 60), a2 = (l2 >> 4) | (l3 << 58), a3 = (l3 >> 6) | (l4 << 56);
    w[0] = (uint32_t)a0; w[1] = (uint32_t)(a0 >> 32); w[2] = (uint32_t)a1; w[3] = (uint32_t)(a1 >> 32);
    w[4] = (uint32_t)a2; w[5] = (uint32_t)(a2 >> 32); w[6] = (uint32_t)a3; w[7] = (uint32_t)(a3 >> 32);
}
CBP_HD ModInfo modinfo_from_words(const uint32_t (&mw)[8]) {
    ModInfo mi;
    modinv_load(mi.m, mw);
    // Newton iteration for m^-1 mod 2^64 (m odd), then mod 2^62
    const uint64_t m0 = (uint64_t)mw[0] | ((uint64_t)mw[1] << 32);
    uint64_t x = m0;  // correct to 3 bits
    for (int i = 0; i < 5; i++) x *= 2 - m0 * x;
    mi.inv62 = x & kM62;
    return mi;
}

struct Trans2x2 {
    int64_t u, v, q, r;
};
// 62 divsteps on the low 64 bits of f (odd) and g; 2^62 [f'; g'] = t [f; g], |entries| <= 2^62
CBP_HD int64_t modinv_divsteps_62(int64_t delta, uint64_t f0, uint64_t g0, Trans2x2& t) {
    uint64_t u = 1, v = 0, q = 0, r = 1, f = f0, g = g0;  // matrix entries wrap modulo 2^64 (they fit in int64)
#if defined(__CUDA_ARCH__)
#pragma unroll 2
#endif
    for (int i = 0; i < 62; i++) {
        if (g & 1) {
            if (delta > 0) {
                const uint64_t tf = f, tu = u, tv = v;
                delta = 1 - delta;
                f = g; g = g - tf;
                u = q; v = r;
                q = q - tu; r = r - tv;
            } else {
                delta = 1 + delta;
                g += f; q += u; r += v;
            }
        } else {
            delta = 1 + delta;
        }
        g >>= 1;
        u <<= 1; v <<= 1;
    }
    t.u = (int64_t)u; t.v = (int64_t)v; t.q = (int64_t)q; t.r = (int64_t)r;
    return delta;
}
// (f, g) <- t (f, g) / 2^62, exact
CBP_HD void modinv_update_fg(int64_t (&f)[5], int64_t (&g)[5], const Trans2x2& t) {
    cbp_i128 cf = (cbp_i128)t.u * f[0] + (cbp_i128)t.v * g[0];
    cbp_i128 cg = (cbp_i128)t.q * f[0] + (cbp_i128)t.r * g[0];
    cf >>= 62; cg >>= 62;
#pragma unroll
    for (int i = 1; i < 5; i++) {
        cf += (cbp_i128)t.u * f[i] + (cbp_i128)t.v * g[i];
        cg += (cbp_i128)t.q * f[i] + (cbp_i128)t.r * g[i];
        f[i - 1] = (int64_t)((uint64_t)cf & kM62);
        g[i - 1] = (int64_t)((uint64_t)cg & kM62);
        cf >>= 62; cg >>= 62;
    }
    f[4] = (int64_t)cf;
    g[4] = (int64_t)cg;
}
// (d, e) <- t (d, e) / 2^62 mod m; inputs and outputs in (-2m, m)
CBP_HD void modinv_update_de(int64_t (&d)[5], int64_t (&e)[5], const Trans2x2& t, const ModInfo& mi) {
    const int64_t sd = d[4] >> 63, se = e[4] >> 63;
    int64_t md = (t.u & sd) + (t.v & se), me = (t.q & sd) + (t.r & se);
    cbp_i128 cd = (cbp_i128)t.u * d[0] + (cbp_i128)t.v * e[0];
    cbp_i128 ce = (cbp_i128)t.q * d[0] + (cbp_i128)t.r * e[0];
    md -= (int64_t)((mi.inv62 * (uint64_t)cd + (uint64_t)md) & kM62);
    me -= (int64_t)((mi.inv62 * (uint64_t)ce + (uint64_t)me) & kM62);
    cd += (cbp_i128)mi.m[0] * md;
    ce += (cbp_i128)mi.m[0] * me;
    cd >>= 62; ce >>= 62;
#pragma unroll
    for (int i = 1; i < 5; i++) {
        cd += (cbp_i128)t.u * d[i] + (cbp_i128)t.v * e[i] + (cbp_i128)mi.m[i] * md;
        ce += (cbp_i128)t.q * d[i] + (cbp_i128)t.r * e[i] + (cbp_i128)mi.m[i] * me;
        d[i - 1] = (int64_t)((uint64_t)cd & kM62);
        e[i - 1] = (int64_t)((uint64_t)ce & kM62);
        cd >>= 62; ce >>= 62;
    }
    d[4] = (int64_t)cd;
    e[4] = (int64_t)ce;
}
// out = x^-1 mod m as 8 little-endian words in [0, m); 0 for x = 0 (mod m).  x: any 256-bit value.
CBP_HD void modinv_words(uint32_t (&out)[8], const uint32_t (&x)[8], const ModInfo& mi) {
    int64_t f[5], g[5], d[5] = {0, 0, 0, 0, 0}, e[5] = {1, 0, 0, 0, 0};
    for (int i = 0; i < 5; i++) f[i] = mi.m[i];
    modinv_load(g, x);  // g < 2^256 may exceed m: harmless, gcd(m, g) and the cofactors are the same
    int64_t delta = 1;
    for (int batch = 0; batch < 16; batch++) {  // 9 suffice for 256-bit inputs (tools/modinv_model.py); bounded anyway
        if ((g[0] | g[1] | g[2] | g[3] | g[4]) == 0) break;
        Trans2x2 t;
        delta = modinv_divsteps_62(delta, (uint64_t)f[0] | ((uint64_t)f[1] << 62), (uint64_t)g[0] | ((uint64_t)g[1] << 62), t);
        modinv_update_de(d, e, t, mi);
        modinv_update_fg(f, g, t);
    }
    // f = +-1 unless gcd(x, m) != 1, i.e. x = 0 (mod m) for a prime m: the answer is then 0 by convention
    const bool plus = f[0] == 1 && (f[1] | f[2] | f[3] | f[4]) == 0;
    const bool minus = (uint64_t)f[0] == kM62 && (uint64_t)f[1] == kM62 && (uint64_t)f[2] == kM62 && (uint64_t)f[3] == kM62 && f[4] == -1;
    if (!plus && !minus) {
        for (int i = 0; i < 8; i++) out[i] = 0;
        return;
    }
    // d in (-2m, m): two conditional additions of m bring it into [0, m); then the inverse is sign(f) d mod m
    for (int pass = 0; pass < 3; pass++) {
        if (pass == 2 && minus) {
            cbp_i128 c = 0;
            for (int i = 0; i < 5; i++) {
                c -= d[i];
                d[i] = i < 4 ? (int64_t)((uint64_t)c & kM62) : (int64_t)c;
                c >>= 62;
            }
        }
        if (d[4] >> 63) {
            cbp_i128 c = 0;
            for (int i = 0; i < 5; i++) {
                c += (cbp_i128)d[i] + mi.m[i];
                d[i] = i < 4 ? (int64_t)((uint64_t)c & kM62) : (int64_t)c;
                c >>= 62;
            }
        }
    }
    modinv_store(out, d);
}

}  // namespace cbp
