/*
 * oracle/ref_corrected.c — TEST INFRASTRUCTURE ONLY.  CPU oracle for the Bulletproofs hot path.
 *
 * A from-scratch restatement of the reference's functions (same names / signatures / formula
 * shapes; every function cites the reference file:line it follows) with defects D1..D24
 *  fixed.  Plain C, 4x64-bit limbs, unsigned __int128.  Not constant time, not
 * fast, not shipped: the product path is cudabulletproof_b200/csrc and never calls this file.
 *
 * Parity pinning: tests/test_oracle_kat.py (RFC 8032 vectors, [l]B = O, Python big-int model),
 * tests/test_oracle_vs_verbatim.py (algebra-free agreement with the unmodified reference build).
 */
#include "ref_corrected.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------------------------ */
/* constants (re-derived in oracle/pyref.py)                              */
/* ------------------------------------------------------------------------------------------ */
static const uint64_t P25519[4] = {0xFFFFFFFFFFFFFFEDull, 0xFFFFFFFFFFFFFFFFull, 0xFFFFFFFFFFFFFFFFull,
                                   0x7FFFFFFFFFFFFFFFull};
static const uint64_t L25519[4] = {0x5812631a5cf5d3edull, 0x14def9dea2f79cd6ull, 0x0000000000000000ull,
                                   0x1000000000000000ull};
static const fe25519 FE_D = {{0x75eb4dca135978a3ull, 0x00700a4d4141d8abull, 0x8cc740797779e898ull,
                              0x52036cee2b6ffe73ull}};
static const fe25519 FE_2D = {{0xebd69b9426b2f159ull, 0x00e0149a8283b156ull, 0x198e80f2eef3d130ull,
                               0x2406d9dc56dffce7ull}};
static const fe25519 FE_SQRTM1 = {{0xc4ee1b274a0ea0b0ull, 0x2f431806ad2fe478ull, 0x2b4d00993dfbd7a7ull,
                                   0x2b8324804fc1df0bull}};
static const fe25519 FE_BX = {{0xc9562d608f25d51aull, 0x692cc7609525a7b2ull, 0xc0a4e231fdd6dc5cull,
                               0x216936d3cd6e53feull}};
static const fe25519 FE_BY = {{0x6666666666666658ull, 0x6666666666666666ull, 0x6666666666666666ull,
                               0x6666666666666666ull}};

/* ------------------------------------------------------------------------------------------ */
/* fe25519 — reference curve25519_ops.cu:9-315, defects D1-D4, D7                              */
/* ------------------------------------------------------------------------------------------ */
static int limbs_geq(const uint64_t a[4], const uint64_t m[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] != m[i]) return a[i] > m[i];
    }
    return 1;
}
static void limbs_sub(uint64_t a[4], const uint64_t m[4]) { /* a -= m, proper borrow chain (D1) */
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - m[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
}
static void fe_canon(uint64_t a[4]) {
    while (limbs_geq(a, P25519)) limbs_sub(a, P25519);
}
/* fold a 5-limb value (a[4] = bits 256..319) into [0,p): 2^256 = 38 (mod p)  (D3) */
static void fe_reduce5(uint64_t r[4], uint64_t a[5]) {
    while (a[4]) {
        u128 c = (u128)a[4] * 38;
        a[4] = 0;
        for (int i = 0; i < 4; i++) {
            c += a[i];
            a[i] = (uint64_t)c;
            c >>= 64;
        }
        a[4] = (uint64_t)c;
    }
    memcpy(r, a, 32);
    fe_canon(r);
}

void fe25519_0(fe25519 *h) { memset(h->limbs, 0, 32); }                 /* curve25519_ops.cu:9  */
void fe25519_1(fe25519 *h) { memset(h->limbs, 0, 32); h->limbs[0] = 1; } /* :14 */
void fe25519_copy(fe25519 *h, const fe25519 *f) { memmove(h->limbs, f->limbs, 32); } /* :22 */
void fe25519_cswap(fe25519 *f, fe25519 *g, uint8_t b) {                  /* :27 */
    uint64_t mask = (uint64_t)(-(int64_t)(b & 1));
    for (int i = 0; i < 4; i++) {
        uint64_t t = mask & (f->limbs[i] ^ g->limbs[i]);
        f->limbs[i] ^= t;
        g->limbs[i] ^= t;
    }
}
void fe25519_add(fe25519 *h, const fe25519 *f, const fe25519 *g) {       /* :41 (D1) */
    uint64_t a[5];
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)f->limbs[i] + g->limbs[i];
        a[i] = (uint64_t)c;
        c >>= 64;
    }
    a[4] = (uint64_t)c;
    fe_reduce5(h->limbs, a);
}
void fe25519_sub(fe25519 *h, const fe25519 *f, const fe25519 *g) {       /* :71 (D2) */
    uint64_t a[4], b[4];
    memcpy(a, f->limbs, 32);
    memcpy(b, g->limbs, 32);
    fe_canon(a);
    fe_canon(b);
    if (!limbs_geq(a, b)) { /* a < b: a += p (fits: a < p <= 2^255) */
        u128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (u128)a[i] + P25519[i];
            a[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    limbs_sub(a, b);
    memcpy(h->limbs, a, 32);
}
void fe25519_neg(fe25519 *h, const fe25519 *f) {                         /* :210 (D2: neg(0)=0) */
    fe25519 zero;
    fe25519_0(&zero);
    fe25519_sub(h, &zero, f);
}
void fe25519_mul(fe25519 *h, const fe25519 *f, const fe25519 *g) {       /* :93 (D3) */
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; i++) { /* same schoolbook shape as the reference */
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 m = (u128)f->limbs[i] * g->limbs[j] + t[i + j] + carry;
            t[i + j] = (uint64_t)m;
            carry = (uint64_t)(m >> 64);
        }
        t[i + 4] = carry;
    }
    uint64_t a[5];
    u128 c = 0;
    for (int i = 0; i < 4; i++) { /* lo + 38*hi, full width */
        c += (u128)t[i + 4] * 38 + t[i];
        a[i] = (uint64_t)c;
        c >>= 64;
    }
    a[4] = (uint64_t)c;
    fe_reduce5(h->limbs, a);
}
void fe25519_sq(fe25519 *h, const fe25519 *f) { fe25519_mul(h, f, f); }  /* :149 */

static void fe_sqn(fe25519 *h, const fe25519 *f, int n) {
    fe25519_sq(h, f);
    for (int i = 1; i < n; i++) fe25519_sq(h, h);
}
/* z^(2^250-1) and z^11, the shared prefix of the inversion and sqrt chains */
static void fe_pow_2_250_1(fe25519 *z_250_0, fe25519 *z11, const fe25519 *z) {
    fe25519 z2, z9, t, z_5_0, z_10_0, z_20_0, z_40_0, z_50_0, z_100_0, z_200_0;
    fe25519_sq(&z2, z);
    fe_sqn(&t, &z2, 2);
    fe25519_mul(&z9, &t, z);
    fe25519_mul(z11, &z9, &z2);
    fe25519_sq(&t, z11);
    fe25519_mul(&z_5_0, &t, &z9);
    fe_sqn(&t, &z_5_0, 5);
    fe25519_mul(&z_10_0, &t, &z_5_0);
    fe_sqn(&t, &z_10_0, 10);
    fe25519_mul(&z_20_0, &t, &z_10_0);
    fe_sqn(&t, &z_20_0, 20);
    fe25519_mul(&z_40_0, &t, &z_20_0);
    fe_sqn(&t, &z_40_0, 10);
    fe25519_mul(&z_50_0, &t, &z_10_0);
    fe_sqn(&t, &z_50_0, 50);
    fe25519_mul(&z_100_0, &t, &z_50_0);
    fe_sqn(&t, &z_100_0, 100);
    fe25519_mul(&z_200_0, &t, &z_100_0);
    fe_sqn(&t, &z_200_0, 50);
    fe25519_mul(z_250_0, &t, &z_50_0);
}
void fe25519_invert(fe25519 *h, const fe25519 *f) {                      /* :157 (D4): f^(p-2) */
    fe25519 z_250_0, z11, t;
    fe_pow_2_250_1(&z_250_0, &z11, f);
    fe_sqn(&t, &z_250_0, 5);
    fe25519_mul(h, &t, &z11); /* 2^255 - 32 + 11 = 2^255 - 21 */
}
void fe25519_pow2523(fe25519 *h, const fe25519 *f) {                     /* :269 (D7): f^(2^252-3) */
    fe25519 z_250_0, z11, t;
    fe_pow_2_250_1(&z_250_0, &z11, f);
    fe_sqn(&t, &z_250_0, 2);
    fe25519_mul(h, &t, f);
}
void fe25519_tobytes(uint8_t *bytes, const fe25519 *h) {                 /* :220 */
    uint64_t t[4];
    memcpy(t, h->limbs, 32);
    fe_canon(t);
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) bytes[i * 8 + j] = (uint8_t)(t[i] >> (8 * j));
}
void fe25519_frombytes(fe25519 *h, const uint8_t *bytes) {               /* :254 raw load, keeps bit 255 */
    for (int i = 0; i < 4; i++) {
        uint64_t v = 0;
        for (int j = 0; j < 8; j++) v |= (uint64_t)bytes[i * 8 + j] << (8 * j);
        h->limbs[i] = v;
    }
}
static int fe_iszero(const fe25519 *f) {
    uint64_t t[4];
    memcpy(t, f->limbs, 32);
    fe_canon(t);
    return (t[0] | t[1] | t[2] | t[3]) == 0;
}
static int fe_equal(const fe25519 *f, const fe25519 *g) {
    fe25519 d;
    fe25519_sub(&d, f, g);
    return fe_iszero(&d);
}
/* cuda_field_ops.cu:222-254,405-462 (D12): Montgomery's trick; zeros are skipped, inv(0)=0 */
void fe25519_batch_invert(fe25519 *results, const fe25519 *inputs, size_t count) {
    if (count == 0) return;
    fe25519 *prefix = (fe25519 *)malloc(count * sizeof(fe25519));
    fe25519 acc;
    fe25519_1(&acc);
    for (size_t i = 0; i < count; i++) {
        prefix[i] = acc;
        if (!fe_iszero(&inputs[i])) fe25519_mul(&acc, &acc, &inputs[i]);
    }
    fe25519 inv;
    fe25519_invert(&inv, &acc);
    for (size_t i = count; i-- > 0;) {
        fe25519 in = inputs[i];
        if (fe_iszero(&in)) {
            fe25519_0(&results[i]);
        } else {
            fe25519_mul(&results[i], &inv, &prefix[i]);
            fe25519_mul(&inv, &inv, &in);
        }
    }
    free(prefix);
}

/* ------------------------------------------------------------------------------------------ */
/* sc25519 — integers mod l (D11); the reference does these with fe25519_* (mod p)            */
/* ------------------------------------------------------------------------------------------ */
void sc25519_reduce512(fe25519 *r, const uint64_t wide[8]) {
    /* bitwise long division: obviously-correct beats fast here */
    uint64_t acc[4] = {0, 0, 0, 0};
    for (int bit = 511; bit >= 0; bit--) {
        uint64_t in = (wide[bit >> 6] >> (bit & 63)) & 1;
        acc[3] = (acc[3] << 1) | (acc[2] >> 63);
        acc[2] = (acc[2] << 1) | (acc[1] >> 63);
        acc[1] = (acc[1] << 1) | (acc[0] >> 63);
        acc[0] = (acc[0] << 1) | in;
        if (limbs_geq(acc, L25519)) limbs_sub(acc, L25519);
    }
    memcpy(r->limbs, acc, 32);
}
void sc25519_reduce(fe25519 *r, const fe25519 *a) {
    uint64_t w[8] = {a->limbs[0], a->limbs[1], a->limbs[2], a->limbs[3], 0, 0, 0, 0};
    sc25519_reduce512(r, w);
}
void sc25519_frombytes(fe25519 *r, const uint8_t bytes[32]) {
    fe25519 t;
    fe25519_frombytes(&t, bytes);
    sc25519_reduce(r, &t);
}
void sc25519_add(fe25519 *r, const fe25519 *a, const fe25519 *b) {
    fe25519 x, y;
    sc25519_reduce(&x, a);
    sc25519_reduce(&y, b);
    u128 c = 0;
    uint64_t s[4];
    for (int i = 0; i < 4; i++) {
        c += (u128)x.limbs[i] + y.limbs[i];
        s[i] = (uint64_t)c;
        c >>= 64;
    }
    if (limbs_geq(s, L25519)) limbs_sub(s, L25519); /* x+y < 2l < 2^254: no carry out */
    memcpy(r->limbs, s, 32);
}
void sc25519_neg(fe25519 *r, const fe25519 *a) {
    fe25519 x;
    sc25519_reduce(&x, a);
    if ((x.limbs[0] | x.limbs[1] | x.limbs[2] | x.limbs[3]) == 0) {
        fe25519_0(r);
        return;
    }
    uint64_t s[4];
    memcpy(s, L25519, 32);
    limbs_sub(s, x.limbs);
    memcpy(r->limbs, s, 32);
}
void sc25519_sub(fe25519 *r, const fe25519 *a, const fe25519 *b) {
    fe25519 nb;
    sc25519_neg(&nb, b);
    sc25519_add(r, a, &nb);
}
void sc25519_mul(fe25519 *r, const fe25519 *a, const fe25519 *b) {
    uint64_t t[8] = {0};
    for (int i = 0; i < 4; i++) {
        uint64_t carry = 0;
        for (int j = 0; j < 4; j++) {
            u128 m = (u128)a->limbs[i] * b->limbs[j] + t[i + j] + carry;
            t[i + j] = (uint64_t)m;
            carry = (uint64_t)(m >> 64);
        }
        t[i + 4] = carry;
    }
    sc25519_reduce512(r, t);
}
void sc25519_invert(fe25519 *r, const fe25519 *a) { /* a^(l-2), square-and-multiply, inv(0)=0 */
    uint64_t e[4];
    memcpy(e, L25519, 32);
    e[0] -= 2;
    fe25519 acc, base;
    fe25519_1(&acc);
    sc25519_reduce(&base, a);
    for (int bit = 252; bit >= 0; bit--) {
        sc25519_mul(&acc, &acc, &acc);
        if ((e[bit >> 6] >> (bit & 63)) & 1) sc25519_mul(&acc, &acc, &base);
    }
    *r = acc;
}

/* ------------------------------------------------------------------------------------------ */
/* ge25519 — reference curve25519_ops.cu:318-605, defects D5-D9                                */
/* ------------------------------------------------------------------------------------------ */
void ge25519_0(ge25519 *h) {                                             /* :318 */
    fe25519_0(&h->X);
    fe25519_1(&h->Y);
    fe25519_1(&h->Z);
    fe25519_0(&h->T);
}
void ge25519_copy(ge25519 *h, const ge25519 *f) { memmove(h, f, sizeof(ge25519)); } /* :566 */
void oracle_basepoint(ge25519 *out) {                                    /* :418-433 (D6) */
    out->X = FE_BX;
    out->Y = FE_BY;
    fe25519_1(&out->Z);
    fe25519_mul(&out->T, &out->X, &out->Y);
}
/* unified extended addition, same operation order as the reference (:326-378), k = 2d (D5) */
void ge25519_add(ge25519 *r, const ge25519 *p, const ge25519 *q) {
    fe25519 A, B, C, D, E, F, G, H;
    fe25519_sub(&A, &p->Y, &p->X);
    fe25519_sub(&B, &q->Y, &q->X);
    fe25519_mul(&A, &A, &B);
    fe25519_add(&B, &p->Y, &p->X);
    fe25519_add(&C, &q->Y, &q->X);
    fe25519_mul(&B, &B, &C);
    fe25519_mul(&C, &p->T, &q->T);
    fe25519_mul(&C, &C, &FE_2D);
    fe25519_mul(&D, &p->Z, &q->Z);
    fe25519_add(&D, &D, &D);
    fe25519_sub(&E, &B, &A);
    fe25519_sub(&F, &D, &C);
    fe25519_add(&G, &D, &C);
    fe25519_add(&H, &B, &A);
    fe25519_mul(&r->X, &E, &F);
    fe25519_mul(&r->Y, &G, &H);
    fe25519_mul(&r->Z, &F, &G);
    fe25519_mul(&r->T, &E, &H);
}
void ge25519_neg(ge25519 *r, const ge25519 *p) {                         /* :440 */
    fe25519_neg(&r->X, &p->X);
    fe25519_copy(&r->Y, &p->Y);
    fe25519_copy(&r->Z, &p->Z);
    fe25519_neg(&r->T, &p->T);
}
void ge25519_sub(ge25519 *r, const ge25519 *p, const ge25519 *q) {       /* :381 */
    ge25519 nq;
    ge25519_neg(&nq, q);
    ge25519_add(r, p, &nq);
}
void ge25519_double(ge25519 *r, const ge25519 *p) { ge25519_add(r, p, p); } /* :560 */
/* MSB-first double-and-add over all 256 scalar bits (:397-415) */
void ge25519_scalarmult(ge25519 *r, const uint8_t *scalar, const ge25519 *p) {
    ge25519 acc, tmp, base;
    ge25519_copy(&base, p);
    ge25519_0(&acc);
    for (int i = 255; i >= 0; i--) {
        int bit = (scalar[i / 8] >> (i % 8)) & 1;
        ge25519_add(&tmp, &acc, &acc);
        if (bit)
            ge25519_add(&acc, &tmp, &base);
        else
            ge25519_copy(&acc, &tmp);
    }
    ge25519_copy(r, &acc);
}
void ge25519_scalarmult_base(ge25519 *r, const uint8_t *scalar) {        /* :426 (D6) */
    ge25519 B;
    oracle_basepoint(&B);
    ge25519_scalarmult(r, scalar, &B);
}
void ge25519_normalize(ge25519 *p) {                                     /* :574 (D4) */
    fe25519 zi, x, y;
    fe25519_invert(&zi, &p->Z);
    fe25519_mul(&x, &p->X, &zi);
    fe25519_mul(&y, &p->Y, &zi);
    p->X = x;
    p->Y = y;
    fe25519_1(&p->Z);
    fe25519_mul(&p->T, &x, &y);
}
void ge25519_pack(ge25519_compressed *r, const ge25519 *p) {             /* :449 */
    fe25519 zi, x, y;
    uint8_t xb[32];
    fe25519_invert(&zi, &p->Z);
    fe25519_mul(&x, &p->X, &zi);
    fe25519_mul(&y, &p->Y, &zi);
    fe25519_tobytes(r->bytes, &y);
    fe25519_tobytes(xb, &x);
    r->bytes[31] |= (uint8_t)((xb[0] & 1) << 7);
}
int ge25519_unpack(ge25519 *r, const ge25519_compressed *p) {            /* :470 (D7, D8) */
    uint8_t yb[32];
    memcpy(yb, p->bytes, 32);
    int sign = yb[31] >> 7;
    yb[31] &= 0x7F;
    fe25519 y, one, y2, u, v, v3, v7, x, t, chk;
    fe25519_frombytes(&y, yb);
    if (limbs_geq(y.limbs, P25519)) return 0; /* non-canonical y */
    fe25519_1(&one);
    fe25519_sq(&y2, &y);
    fe25519_sub(&u, &y2, &one);      /* u = y^2 - 1 */
    fe25519_mul(&v, &FE_D, &y2);
    fe25519_add(&v, &v, &one);       /* v = d y^2 + 1 */
    fe25519_sq(&v3, &v);
    fe25519_mul(&v3, &v3, &v);       /* v^3 */
    fe25519_sq(&v7, &v3);
    fe25519_mul(&v7, &v7, &v);       /* v^7 */
    fe25519_mul(&t, &u, &v7);
    fe25519_pow2523(&t, &t);         /* (u v^7)^((p-5)/8) */
    fe25519_mul(&x, &u, &v3);
    fe25519_mul(&x, &x, &t);         /* candidate root */
    fe25519_sq(&chk, &x);
    fe25519_mul(&chk, &chk, &v);     /* v x^2 */
    if (!fe_equal(&chk, &u)) {
        fe25519 nu;
        fe25519_neg(&nu, &u);
        if (!fe_equal(&chk, &nu)) return 0;
        fe25519_mul(&x, &x, &FE_SQRTM1);
    }
    if (fe_iszero(&x) && sign) return 0;
    uint8_t xb[32];
    fe25519_tobytes(xb, &x);
    if ((xb[0] & 1) != sign) fe25519_neg(&x, &x);
    r->X = x;
    r->Y = y;
    fe25519_1(&r->Z);
    fe25519_mul(&r->T, &x, &y);
    return 1;
}
int ge25519_is_on_curve(const ge25519 *p) {                              /* :534 (D8) */
    fe25519 x2, y2, z2, t2, lhs, rhs, a, b;
    if (fe_iszero(&p->Z)) return 0;
    fe25519_sq(&x2, &p->X);
    fe25519_sq(&y2, &p->Y);
    fe25519_sq(&z2, &p->Z);
    fe25519_sq(&t2, &p->T);
    fe25519_sub(&lhs, &y2, &x2);
    fe25519_mul(&rhs, &FE_D, &t2);
    fe25519_add(&rhs, &rhs, &z2);
    fe25519_mul(&a, &p->X, &p->Y);
    fe25519_mul(&b, &p->Z, &p->T);
    return fe_equal(&lhs, &rhs) && fe_equal(&a, &b);
}
int ge25519_is_identity(const ge25519 *p) {                              /* :544 (D9) */
    return fe_iszero(&p->X) && fe_equal(&p->Y, &p->Z) && !fe_iszero(&p->Z);
}
int ge25519_equal(const ge25519 *p, const ge25519 *q) {
    fe25519 a, b, c, d;
    fe25519_mul(&a, &p->X, &q->Z);
    fe25519_mul(&b, &q->X, &p->Z);
    fe25519_mul(&c, &p->Y, &q->Z);
    fe25519_mul(&d, &q->Y, &p->Z);
    return fe_equal(&a, &b) && fe_equal(&c, &d);
}

/* ------------------------------------------------------------------------------------------ */
/* SHA-256 (FIPS 180-4) and the Fiat-Shamir builders of bulletproof_challenge.cu:6-77          */
/* ------------------------------------------------------------------------------------------ */
static const uint32_t K256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};
static uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void sha256_block(uint32_t st[8], const uint8_t blk[64]) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) |
               blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = h + S1 + ch + K256[i] + w[i];
        uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
void oracle_sha256(uint8_t out[32], const void *data, size_t len) {
    uint32_t st[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    const uint8_t *p = (const uint8_t *)data;
    size_t rem = len;
    while (rem >= 64) {
        sha256_block(st, p);
        p += 64;
        rem -= 64;
    }
    uint8_t tail[128] = {0};
    memcpy(tail, p, rem);
    tail[rem] = 0x80;
    size_t tl = (rem < 56) ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(st, tail);
    if (tl == 128) sha256_block(st, tail + 64);
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(st[i] >> 24);
        out[4 * i + 1] = (uint8_t)(st[i] >> 16);
        out[4 * i + 2] = (uint8_t)(st[i] >> 8);
        out[4 * i + 3] = (uint8_t)st[i];
    }
}
/* bulletproof_challenge.cu:6-21: SHA-256(domain || data), then clear the top bit */
void generate_challenge(uint8_t *output, const void *data, size_t data_len, const char *domain_sep) {
    size_t dl = strlen(domain_sep);
    uint8_t *buf = (uint8_t *)malloc(dl + data_len);
    memcpy(buf, domain_sep, dl);
    memcpy(buf + dl, data, data_len);
    oracle_sha256(output, buf, dl + data_len);
    free(buf);
    output[31] &= 0x7F;
}
/* affine (x, y) as 64 canonical bytes; the reference hashes X,Y of points it has normalised */
static void point_xy_bytes(uint8_t out[64], const ge25519 *p) {
    ge25519 t = *p;
    ge25519_normalize(&t);
    fe25519_tobytes(out, &t.X);
    fe25519_tobytes(out + 32, &t.Y);
}
void generate_challenge_y(uint8_t *output, const ge25519 *V, const ge25519 *A, const ge25519 *S) { /* :24-44 */
    uint8_t d[196];
    point_xy_bytes(d, V);
    point_xy_bytes(d + 64, A);
    point_xy_bytes(d + 128, S);
    memcpy(d + 192, "y_ch", 4);
    generate_challenge(output, d, sizeof d, "BulletproofYChal");
}
void generate_challenge_z(uint8_t *output, const uint8_t *y_challenge) {  /* :47-58 */
    uint8_t d[36];
    memcpy(d, y_challenge, 32);
    memcpy(d + 32, "z_ch", 4);
    generate_challenge(output, d, sizeof d, "BulletproofZChal");
}
void generate_challenge_x(uint8_t *output, const ge25519 *T1, const ge25519 *T2) { /* :61-77 */
    uint8_t d[132];
    point_xy_bytes(d, T1);
    point_xy_bytes(d + 64, T2);
    memcpy(d + 128, "xchal", 4); /* the reference copies only 4 bytes: "xcha" */
    generate_challenge(output, d, sizeof d, "BulletproofXChal");
}

/* ------------------------------------------------------------------------------------------ */
/* vectors, naive MSM, IPA — bulletproof_vectors.cu                                            */
/* ------------------------------------------------------------------------------------------ */
void field_vector_init(FieldVector *vec, size_t length) {                /* :16 */
    vec->length = length;
    vec->elements = (fe25519 *)calloc(length ? length : 1, sizeof(fe25519));
}
void field_vector_free(FieldVector *vec) {                               /* :27 */
    free(vec->elements);
    vec->elements = NULL;
    vec->length = 0;
}
void point_vector_init(PointVector *vec, size_t length) {                /* :117 */
    vec->length = length;
    vec->elements = (ge25519 *)calloc(length ? length : 1, sizeof(ge25519));
    for (size_t i = 0; i < length; i++) ge25519_0(&vec->elements[i]);
}
void point_vector_free(PointVector *vec) {
    free(vec->elements);
    vec->elements = NULL;
    vec->length = 0;
}
/* :101-114; sum of products mod l (D11) */
void field_vector_inner_product(fe25519 *result, const FieldVector *a, const FieldVector *b) {
    if (a->length != b->length) {
        fprintf(stderr, "Error: Vector lengths must match for inner product\n");
        return;
    }
    fe25519 acc, t;
    fe25519_0(&acc);
    for (size_t i = 0; i < a->length; i++) {
        sc25519_mul(&t, &a->elements[i], &b->elements[i]);
        sc25519_add(&acc, &acc, &t);
    }
    *result = acc;
}
/* :189-224 naive MSM: scalar -> fe25519_tobytes -> 256-bit double-and-add, running sum.
 * The reference normalises after every step; normalisation does not change the group element,
 * so the oracle normalises once at the end (output Z = 1, as the reference returns). */
void point_vector_multi_scalar_mul(ge25519 *result, const FieldVector *scalars, const PointVector *points) {
    if (scalars->length != points->length) {
        fprintf(stderr, "Error: Vector lengths must match for multi-scalar multiplication\n");
        return;
    }
    ge25519 acc, t;
    ge25519_0(&acc);
    for (size_t i = 0; i < scalars->length; i++) {
        uint8_t sb[32];
        fe25519_tobytes(sb, &scalars->elements[i]);
        ge25519_scalarmult(&t, sb, &points->elements[i]);
        ge25519_add(&acc, &acc, &t);
    }
    ge25519_normalize(&acc);
    *result = acc;
}
void inner_product_proof_init(InnerProductProof *proof, size_t n) {      /* :227-249 */
    if ((n & (n - 1)) != 0 || n == 0) {
        fprintf(stderr, "Error: Inner product proof size must be a power of 2\n");
        return;
    }
    proof->n = n;
    field_vector_init(&proof->a, n);
    field_vector_init(&proof->b, n);
    fe25519_0(&proof->c);
    size_t log_n = 0;
    for (size_t t = n; t > 1; t >>= 1) log_n++;
    proof->L_len = log_n;
    point_vector_init(&proof->L, log_n);
    point_vector_init(&proof->R, log_n);
    fe25519_0(&proof->x);
}
void inner_product_proof_free(InnerProductProof *proof) {                /* :252-257 */
    field_vector_free(&proof->a);
    field_vector_free(&proof->b);
    point_vector_free(&proof->L);
    point_vector_free(&proof->R);
}
static void scalarmult_fe(ge25519 *r, const fe25519 *s, const ge25519 *p) {
    uint8_t sb[32];
    fe25519_tobytes(sb, s);
    ge25519_scalarmult(r, sb, p);
}
/* :488-500 with the exponent pattern fixed (D21): a' = u a_L + u^-1 a_R, b' = u^-1 b_L + u b_R */
void ipa_fold_scalars(fe25519 *a_out, fe25519 *b_out, const fe25519 *a, const fe25519 *b, size_t n_half,
                      const fe25519 *u, const fe25519 *u_inv) {
    for (size_t j = 0; j < n_half; j++) {
        fe25519 t0, t1, na, nb;
        sc25519_mul(&t0, u, &a[j]);
        sc25519_mul(&t1, u_inv, &a[j + n_half]);
        sc25519_add(&na, &t0, &t1);
        sc25519_mul(&t0, u_inv, &b[j]);
        sc25519_mul(&t1, u, &b[j + n_half]);
        sc25519_add(&nb, &t0, &t1);
        a_out[j] = na;
        b_out[j] = nb;
    }
}
/* :641-663: G'_j = u^-1 G_j + u G_{j+n'},  H'_j = u H_j + u^-1 H_{j+n'}; outputs normalised */
void ipa_fold_points(ge25519 *G_out, ge25519 *H_out, const ge25519 *G, const ge25519 *H, size_t n_half,
                     const fe25519 *u, const fe25519 *u_inv) {
    for (size_t j = 0; j < n_half; j++) {
        ge25519 t1, t2, g, h;
        scalarmult_fe(&t1, u_inv, &G[j]);
        scalarmult_fe(&t2, u, &G[j + n_half]);
        ge25519_add(&g, &t1, &t2);
        ge25519_normalize(&g);
        scalarmult_fe(&t1, u, &H[j]);
        scalarmult_fe(&t2, u_inv, &H[j + n_half]);
        ge25519_add(&h, &t1, &t2);
        ge25519_normalize(&h);
        G_out[j] = g;
        H_out[j] = h;
    }
}
/* round challenge, :448-465 / :601-626: SHA-256("InnerProductChal" || transcript || L.X || R.X) */
static void ipa_round_challenge(uint8_t out[32], const uint8_t transcript[32], const ge25519 *L, const ge25519 *R) {
    uint8_t d[96], xy[64];
    memcpy(d, transcript, 32);
    point_xy_bytes(xy, L);
    memcpy(d + 32, xy, 32);
    point_xy_bytes(xy, R);
    memcpy(d + 64, xy, 32);
    generate_challenge(out, d, sizeof d, "InnerProductChal");
}
/* :277-538.  Fixes: scalars mod l (D11), G/H folded every round (D16), fold exponents (D21). */
void inner_product_prove(InnerProductProof *proof, const FieldVector *a_in, const FieldVector *b_in,
                         const PointVector *G, const PointVector *H, const ge25519 *Q, const fe25519 *c_in,
                         const uint8_t *initial_transcript) {
    if (a_in->length != b_in->length || a_in->length != G->length || a_in->length != H->length) return;
    size_t n = a_in->length;
    if ((n & (n - 1)) != 0 || n == 0) return;
    inner_product_proof_init(proof, n);
    proof->c = *c_in;
    uint8_t transcript[32];
    memcpy(transcript, initial_transcript, 32);

    fe25519 *a = (fe25519 *)malloc(n * sizeof(fe25519)), *b = (fe25519 *)malloc(n * sizeof(fe25519));
    ge25519 *g = (ge25519 *)malloc(n * sizeof(ge25519)), *h = (ge25519 *)malloc(n * sizeof(ge25519));
    for (size_t i = 0; i < n; i++) {
        sc25519_reduce(&a[i], &a_in->elements[i]);
        sc25519_reduce(&b[i], &b_in->elements[i]);
        g[i] = G->elements[i];
        h[i] = H->elements[i];
    }
    size_t np = n;
    for (size_t round = 0; round < proof->L_len; round++) {
        np >>= 1;
        FieldVector aL = {a, np}, aR = {a + np, np}, bL = {b, np}, bR = {b + np, np};
        PointVector GL = {g, np}, GR = {g + np, np}, HL = {h, np}, HR = {h + np, np};
        fe25519 cL, cR;
        field_vector_inner_product(&cL, &aL, &bR);
        field_vector_inner_product(&cR, &aR, &bL);
        ge25519 L, R, t1, t2, t3;
        point_vector_multi_scalar_mul(&t1, &aL, &GR);
        point_vector_multi_scalar_mul(&t2, &bR, &HL);
        scalarmult_fe(&t3, &cL, Q);
        ge25519_add(&L, &t1, &t2);
        ge25519_add(&L, &L, &t3);
        ge25519_normalize(&L);
        point_vector_multi_scalar_mul(&t1, &aR, &GL);
        point_vector_multi_scalar_mul(&t2, &bL, &HR);
        scalarmult_fe(&t3, &cR, Q);
        ge25519_add(&R, &t1, &t2);
        ge25519_add(&R, &R, &t3);
        ge25519_normalize(&R);
        proof->L.elements[round] = L;
        proof->R.elements[round] = R;

        uint8_t ch[32];
        ipa_round_challenge(ch, transcript, &L, &R);
        memcpy(transcript, ch, 32);
        fe25519 u, u_inv;
        if (round == 0) fe25519_frombytes(&proof->x, ch); /* :471-474 */
        sc25519_frombytes(&u, ch);
        sc25519_invert(&u_inv, &u);
        ipa_fold_scalars(a, b, a, b, np, &u, &u_inv);
        ipa_fold_points(g, h, g, h, np, &u, &u_inv);
    }
    field_vector_free(&proof->a);
    field_vector_free(&proof->b);
    field_vector_init(&proof->a, 1); /* the reference's a/b shrink to length 1 via field_vector_copy (:503-504) */
    field_vector_init(&proof->b, 1);
    proof->a.elements[0] = a[0];
    proof->b.elements[0] = b[0];
    free(a);
    free(b);
    free(g);
    free(h);
}
/* :541-762.  P must already contain c*Q.  Exact check (D18):
 *   P + sum_j (u_j^2 L_j + u_j^-2 R_j) == a*G' + b*H' + (a*b)*Q
 * with G', H' folded round by round exactly as the reference's verifier does (:641-663). */
bool inner_product_verify_transcript(const InnerProductProof *proof, const ge25519 *P, const PointVector *G,
                                     const PointVector *H, const ge25519 *Q, const uint8_t transcript0[32]) {
    if (G->length != proof->n || H->length != proof->n) return false;
    size_t n = proof->n;
    if (n == 0 || (n & (n - 1)) != 0) return false;
    size_t rounds = 0;
    for (size_t t = n; t > 1; t >>= 1) rounds++;
    if (proof->L_len != rounds || proof->L.length != rounds || proof->R.length != rounds) return false;
    if (proof->a.length < 1 || proof->b.length < 1) return false;
    for (size_t i = 0; i < rounds; i++)
        if (!ge25519_is_on_curve(&proof->L.elements[i]) || !ge25519_is_on_curve(&proof->R.elements[i])) return false;

    ge25519 *g = (ge25519 *)malloc(n * sizeof(ge25519)), *h = (ge25519 *)malloc(n * sizeof(ge25519));
    memcpy(g, G->elements, n * sizeof(ge25519));
    memcpy(h, H->elements, n * sizeof(ge25519));
    uint8_t transcript[32];
    memcpy(transcript, transcript0, 32);
    ge25519 acc = *P;
    bool ok = true;
    size_t np = n;
    for (size_t round = 0; round < rounds; round++) {
        np >>= 1;
        uint8_t ch[32];
        ipa_round_challenge(ch, transcript, &proof->L.elements[round], &proof->R.elements[round]);
        memcpy(transcript, ch, 32);
        if (round == 0) { /* the stored first challenge (:597-599) must be the recomputed one (D15) */
            uint8_t xb[32];
            fe25519_tobytes(xb, &proof->x);
            if (memcmp(xb, ch, 32) != 0) ok = false;
        }
        fe25519 u, u_inv, u2, ui2;
        sc25519_frombytes(&u, ch);
        sc25519_invert(&u_inv, &u);
        sc25519_mul(&u2, &u, &u);
        sc25519_mul(&ui2, &u_inv, &u_inv);
        ge25519 t;
        scalarmult_fe(&t, &u2, &proof->L.elements[round]);
        ge25519_add(&acc, &acc, &t);
        scalarmult_fe(&t, &ui2, &proof->R.elements[round]);
        ge25519_add(&acc, &acc, &t);
        ipa_fold_points(g, h, g, h, np, &u, &u_inv);
    }
    fe25519 a, b, ab;
    sc25519_reduce(&a, &proof->a.elements[0]);
    sc25519_reduce(&b, &proof->b.elements[0]);
    sc25519_mul(&ab, &a, &b);
    ge25519 rhs, t;
    scalarmult_fe(&rhs, &a, &g[0]);
    scalarmult_fe(&t, &b, &h[0]);
    ge25519_add(&rhs, &rhs, &t);
    scalarmult_fe(&t, &ab, Q);
    ge25519_add(&rhs, &rhs, &t);
    if (!ge25519_equal(&acc, &rhs)) ok = false;
    free(g);
    free(h);
    return ok;
}
bool inner_product_verify(const InnerProductProof *proof, const ge25519 *P, const PointVector *G,
                          const PointVector *H, const ge25519 *Q) {
    uint8_t zero[32] = {0}; /* :589 */
    return inner_product_verify_transcript(proof, P, G, H, Q, zero);
}

/* ------------------------------------------------------------------------------------------ */
/* range proof — bulletproof_range_proof.cu                                                    */
/* ------------------------------------------------------------------------------------------ */
static uint64_t g_rng_state = 0x9E3779B97F4A7C15ull;
void oracle_seed_rng(uint64_t seed) { g_rng_state = seed; }
static uint64_t splitmix64(void) {
    uint64_t z = (g_rng_state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
/* :153-159 draws RAND_bytes and clamps X25519-style; the oracle draws a deterministic stream and
 * keeps the same clamp so blinding scalars have the reference's shape. */
void generate_random_scalar(uint8_t *output, size_t len) {
    for (size_t i = 0; i < len; i += 8) {
        uint64_t v = splitmix64();
        for (size_t j = 0; j < 8 && i + j < len; j++) output[i + j] = (uint8_t)(v >> (8 * j));
    }
    if (len >= 32) {
        output[31] &= 0x7F;
        output[0] &= 0xF8;
        output[31] |= 0x40;
    }
}
void range_proof_init(RangeProof *proof, size_t n) {                     /* :266 */
    memset(proof, 0, sizeof(RangeProof));
    inner_product_proof_init(&proof->ip_proof, n);
}
void range_proof_free(RangeProof *proof) { inner_product_proof_free(&proof->ip_proof); } /* :272 */
void pedersen_commit(ge25519 *result, const fe25519 *value, const fe25519 *blinding, const ge25519 *g,
                     const ge25519 *h) {                                 /* :277-296 */
    ge25519 t1, t2;
    scalarmult_fe(&t1, value, g);
    scalarmult_fe(&t2, blinding, h);
    ge25519_add(result, &t1, &t2);
    ge25519_normalize(result);
}
void powers_of(FieldVector *result, const fe25519 *base, size_t n) {     /* :299-312, mod l */
    if (result->length != n) {
        field_vector_free(result);
        field_vector_init(result, n);
    }
    if (n == 0) return;
    fe25519_1(&result->elements[0]);
    for (size_t i = 1; i < n; i++) sc25519_mul(&result->elements[i], &result->elements[i - 1], base);
}
/* :315-374: delta = (z - z^2) <1,y^n> - z^3 <1,2^n>   (mod l) */
void compute_precise_delta(fe25519 *delta, const fe25519 *z, const fe25519 *y, size_t n) {
    fe25519 z2, z3, zmz2, sum_y, cur, two, sum2, t1, t2;
    sc25519_mul(&z2, z, z);
    sc25519_mul(&z3, &z2, z);
    sc25519_sub(&zmz2, z, &z2);
    fe25519_1(&sum_y);
    fe25519_1(&cur);
    for (size_t i = 1; i < n; i++) {
        sc25519_mul(&cur, &cur, y);
        sc25519_add(&sum_y, &sum_y, &cur);
    }
    sc25519_mul(&t1, &zmz2, &sum_y);
    fe25519_1(&two);
    sc25519_add(&two, &two, &two);
    fe25519_1(&cur);
    fe25519_1(&sum2);
    for (size_t i = 1; i < n; i++) {
        sc25519_mul(&cur, &cur, &two);
        sc25519_add(&sum2, &sum2, &cur);
    }
    sc25519_mul(&t2, &z3, &sum2);
    sc25519_sub(delta, &t1, &t2);
}
bool validate_range_input(const fe25519 *v, size_t n) {                  /* :238-263: v < 2^n exactly */
    uint8_t b[32];
    fe25519_tobytes(b, v);
    for (size_t bit = n; bit < 256; bit++)
        if ((b[bit / 8] >> (bit % 8)) & 1) return false;
    return true;
}
/* :658-763 corrected (D17).  P = A + x S - z<1,G> + <z y^n + z^2 2^n, H'> - mu h + t Q  with
 * H'_i = y^-i H_i and Q = h (the reference passes h as Q, :1685,1802), i.e. expressed on H:
 *   scalars_G[i] = -z ;  scalars_H[i] = z + z^2 2^i y^-i ;  two MSMs as in the reference (:724,728). */
void calculate_inner_product_point(ge25519 *P, const RangeProof *proof, const fe25519 *x, const fe25519 *y,
                                   const fe25519 *z, const fe25519 *t, const PointVector *G,
                                   const PointVector *H, const ge25519 *g, const ge25519 *h, size_t n) {
    (void)g;
    FieldVector sG, sH;
    field_vector_init(&sG, n);
    field_vector_init(&sH, n);
    fe25519 z2, yinv, yip, two_i, two, tmp, negz;
    sc25519_mul(&z2, z, z);
    sc25519_invert(&yinv, y);
    sc25519_neg(&negz, z);
    fe25519_1(&yip);
    fe25519_1(&two_i);
    fe25519_1(&two);
    sc25519_add(&two, &two, &two);
    for (size_t i = 0; i < n; i++) {
        sG.elements[i] = negz;
        sc25519_mul(&tmp, &z2, &two_i);
        sc25519_mul(&tmp, &tmp, &yip);
        sc25519_add(&sH.elements[i], z, &tmp);
        sc25519_mul(&yip, &yip, &yinv);
        sc25519_mul(&two_i, &two_i, &two);
    }
    ge25519 term1, term2, term3, acc;
    point_vector_multi_scalar_mul(&term1, &sG, G);
    point_vector_multi_scalar_mul(&term2, &sH, H);
    ge25519_add(&acc, &term1, &term2);
    ge25519_add(&acc, &acc, &proof->A);
    scalarmult_fe(&term3, x, &proof->S);
    ge25519_add(&acc, &acc, &term3);
    fe25519 tmmu;
    sc25519_sub(&tmmu, t, &proof->mu); /* (t - mu) h : the -mu h blinding and the +t Q term, Q = h */
    scalarmult_fe(&term3, &tmmu, h);
    ge25519_add(&acc, &acc, &term3);
    ge25519_normalize(&acc);
    *P = acc;
    field_vector_free(&sG);
    field_vector_free(&sH);
}
static bool point_xy_equal(const ge25519 *a, const ge25519 *b) { return ge25519_equal(a, b) != 0; }
/* :1717-1812 corrected: exact checks only (D18).  accept  <=>
 *   V == proof->V,  all proof points on curve,  ip_proof.c == t,
 *   t g + taux h == z^2 V + delta g + x T1 + x^2 T2                     (Bunz et al. (65)),
 *   IPA(P; G, H' = y^-i H, Q = h) with transcript0 = H("BulletproofIP" || t || taux || mu). */
bool range_proof_verify(const RangeProof *proof, const ge25519 *V, size_t n, const PointVector *G,
                        const PointVector *H, const ge25519 *g, const ge25519 *h) {
    if (G->length != n || H->length != n || proof->ip_proof.n != n) return false;
    if (!ge25519_is_on_curve(V) || !ge25519_is_on_curve(&proof->V) || !ge25519_is_on_curve(&proof->A) ||
        !ge25519_is_on_curve(&proof->S) || !ge25519_is_on_curve(&proof->T1) || !ge25519_is_on_curve(&proof->T2))
        return false;
    if (!point_xy_equal(V, &proof->V)) return false; /* :1729-1740 */
    uint8_t yb[32], zb[32], xb[32];
    generate_challenge_y(yb, V, &proof->A, &proof->S);
    generate_challenge_z(zb, yb);
    generate_challenge_x(xb, &proof->T1, &proof->T2);
    fe25519 y, z, x, delta, z2, x2, t, taux;
    sc25519_frombytes(&y, yb);
    sc25519_frombytes(&z, zb);
    sc25519_frombytes(&x, xb);
    compute_precise_delta(&delta, &z, &y, n);
    sc25519_mul(&z2, &z, &z);
    sc25519_mul(&x2, &x, &x);
    sc25519_reduce(&t, &proof->t);
    sc25519_reduce(&taux, &proof->taux);
    bool ok = true;
    { /* ip_proof.c must be the claimed t (the reference passes t as c_in, :1685) */
        fe25519 c;
        sc25519_reduce(&c, &proof->ip_proof.c);
        if (memcmp(&c, &t, 32) != 0) ok = false;
    }
    { /* polynomial identity, :412-655 made exact */
        ge25519 lhs, rhs, tmp;
        fe25519 tmd;
        sc25519_sub(&tmd, &t, &delta);
        scalarmult_fe(&lhs, &tmd, g);
        scalarmult_fe(&tmp, &taux, h);
        ge25519_add(&lhs, &lhs, &tmp);
        scalarmult_fe(&rhs, &z2, V);
        scalarmult_fe(&tmp, &x, &proof->T1);
        ge25519_add(&rhs, &rhs, &tmp);
        scalarmult_fe(&tmp, &x2, &proof->T2);
        ge25519_add(&rhs, &rhs, &tmp);
        if (!ge25519_equal(&lhs, &rhs)) ok = false;
    }
    ge25519 P;
    calculate_inner_product_point(&P, proof, &x, &y, &z, &t, G, H, g, h, n);
    PointVector Hp;
    point_vector_init(&Hp, n);
    fe25519 yinv, yip;
    sc25519_invert(&yinv, &y);
    fe25519_1(&yip);
    for (size_t i = 0; i < n; i++) {
        scalarmult_fe(&Hp.elements[i], &yip, &H->elements[i]);
        sc25519_mul(&yip, &yip, &yinv);
    }
    uint8_t fin[96], tr0[32];
    fe25519_tobytes(fin, &t);
    fe25519_tobytes(fin + 32, &taux);
    fe25519 mu;
    sc25519_reduce(&mu, &proof->mu);
    fe25519_tobytes(fin + 64, &mu);
    generate_challenge(tr0, fin, sizeof fin, "BulletproofIP"); /* :1668-1676 */
    if (!inner_product_verify_transcript(&proof->ip_proof, &P, G, &Hp, h, tr0)) ok = false;
    point_vector_free(&Hp);
    return ok;
}
/* :1159-1714 corrected.  Fixes: scalars mod l (D11); t0,t1 include the z^2 2^n cross terms (D22);
 * taux includes z^2 gamma (D23); real l(x), r(x) go to the IPA (D19); IPA runs on H' (D17);
 * out-of-range input leaves an initialised, invalid proof (D20). */
void generate_range_proof(RangeProof *proof, const fe25519 *v, const fe25519 *gamma, size_t n,
                          const PointVector *G, const PointVector *H, const ge25519 *g, const ge25519 *h) {
    range_proof_init(proof, n);
    ge25519_0(&proof->V);
    ge25519_0(&proof->A);
    ge25519_0(&proof->S);
    ge25519_0(&proof->T1);
    ge25519_0(&proof->T2);
    if (!validate_range_input(v, n)) return;
    fe25519 gam;
    sc25519_reduce(&gam, gamma);
    pedersen_commit(&proof->V, v, &gam, g, h);
    FieldVector aL, aR, sL, sR;
    field_vector_init(&aL, n);
    field_vector_init(&aR, n);
    field_vector_init(&sL, n);
    field_vector_init(&sR, n);
    uint8_t vb[32];
    fe25519_tobytes(vb, v);
    fe25519 one;
    fe25519_1(&one);
    for (size_t i = 0; i < n; i++) { /* :1219-1240 */
        if ((vb[i / 8] >> (i % 8)) & 1) fe25519_1(&aL.elements[i]);
        sc25519_sub(&aR.elements[i], &aL.elements[i], &one);
    }
    for (size_t i = 0; i < n; i++) { /* :1249-1257, same draw order */
        uint8_t l[32], r[32];
        generate_random_scalar(l, 32);
        generate_random_scalar(r, 32);
        sc25519_frombytes(&sL.elements[i], l);
        sc25519_frombytes(&sR.elements[i], r);
    }
    uint8_t ab[32], rb[32];
    generate_random_scalar(ab, 32);
    generate_random_scalar(rb, 32);
    fe25519 alpha, rho;
    sc25519_frombytes(&alpha, ab);
    sc25519_frombytes(&rho, rb);
    ge25519 t1p, t2p, t3p;
    scalarmult_fe(&t1p, &alpha, h); /* :1267-1276 */
    point_vector_multi_scalar_mul(&t2p, &aL, G);
    point_vector_multi_scalar_mul(&t3p, &aR, H);
    ge25519_add(&proof->A, &t1p, &t2p);
    ge25519_add(&proof->A, &proof->A, &t3p);
    ge25519_normalize(&proof->A);
    scalarmult_fe(&t1p, &rho, h); /* :1279-1288 */
    point_vector_multi_scalar_mul(&t2p, &sL, G);
    point_vector_multi_scalar_mul(&t3p, &sR, H);
    ge25519_add(&proof->S, &t1p, &t2p);
    ge25519_add(&proof->S, &proof->S, &t3p);
    ge25519_normalize(&proof->S);

    uint8_t yb[32], zb[32], xb[32];
    generate_challenge_y(yb, &proof->V, &proof->A, &proof->S);
    generate_challenge_z(zb, yb);
    fe25519 y, z, z2, x, x2;
    sc25519_frombytes(&y, yb);
    sc25519_frombytes(&z, zb);
    sc25519_mul(&z2, &z, &z);
    /* l(X) = l0 + l1 X, r(X) = r0 + r1 X  (:1343-1346,1582-1610) */
    FieldVector l0, r0, r1, yn;
    field_vector_init(&l0, n);
    field_vector_init(&r0, n);
    field_vector_init(&r1, n);
    field_vector_init(&yn, 0);
    powers_of(&yn, &y, n);
    fe25519 two_i, two, tmp;
    fe25519_1(&two_i);
    fe25519_1(&two);
    sc25519_add(&two, &two, &two);
    for (size_t i = 0; i < n; i++) {
        sc25519_sub(&l0.elements[i], &aL.elements[i], &z);
        sc25519_add(&tmp, &aR.elements[i], &z);
        sc25519_mul(&tmp, &tmp, &yn.elements[i]);
        fe25519 zz;
        sc25519_mul(&zz, &z2, &two_i);
        sc25519_add(&r0.elements[i], &tmp, &zz);
        sc25519_mul(&r1.elements[i], &yn.elements[i], &sR.elements[i]);
        sc25519_mul(&two_i, &two_i, &two);
    }
    fe25519 t0, t1, t2, u1, u2;
    field_vector_inner_product(&t0, &l0, &r0);
    field_vector_inner_product(&u1, &l0, &r1);
    field_vector_inner_product(&u2, &sL, &r0);
    sc25519_add(&t1, &u1, &u2);
    field_vector_inner_product(&t2, &sL, &r1);
    uint8_t tau1b[32], tau2b[32];
    generate_random_scalar(tau1b, 32); /* :1434-1437 */
    generate_random_scalar(tau2b, 32);
    fe25519 tau1, tau2;
    sc25519_frombytes(&tau1, tau1b);
    sc25519_frombytes(&tau2, tau2b);
    pedersen_commit(&proof->T1, &t1, &tau1, g, h);
    pedersen_commit(&proof->T2, &t2, &tau2, g, h);
    generate_challenge_x(xb, &proof->T1, &proof->T2);
    sc25519_frombytes(&x, xb);
    sc25519_mul(&x2, &x, &x);
    fe25519 t, taux, mu;
    sc25519_mul(&u1, &t1, &x);
    sc25519_mul(&u2, &t2, &x2);
    sc25519_add(&t, &t0, &u1);
    sc25519_add(&t, &t, &u2);
    sc25519_mul(&u1, &tau1, &x);
    sc25519_mul(&u2, &tau2, &x2);
    sc25519_add(&taux, &u1, &u2);
    sc25519_mul(&u1, &z2, &gam);
    sc25519_add(&taux, &taux, &u1);
    sc25519_mul(&u1, &rho, &x);
    sc25519_add(&mu, &alpha, &u1);
    proof->t = t;
    proof->taux = taux;
    proof->mu = mu;
    FieldVector lx, rx;
    field_vector_init(&lx, n);
    field_vector_init(&rx, n);
    for (size_t i = 0; i < n; i++) {
        sc25519_mul(&tmp, &sL.elements[i], &x);
        sc25519_add(&lx.elements[i], &l0.elements[i], &tmp);
        sc25519_mul(&tmp, &r1.elements[i], &x);
        sc25519_add(&rx.elements[i], &r0.elements[i], &tmp);
    }
    /* IPA over (G, H' = y^-i H, Q = h) */
    PointVector Hp;
    point_vector_init(&Hp, n);
    fe25519 yinv, yip;
    sc25519_invert(&yinv, &y);
    fe25519_1(&yip);
    for (size_t i = 0; i < n; i++) {
        scalarmult_fe(&Hp.elements[i], &yip, &H->elements[i]);
        ge25519_normalize(&Hp.elements[i]);
        sc25519_mul(&yip, &yip, &yinv);
    }
    uint8_t fin[96], tr0[32];
    fe25519_tobytes(fin, &t);
    fe25519_tobytes(fin + 32, &taux);
    fe25519_tobytes(fin + 64, &mu);
    generate_challenge(tr0, fin, sizeof fin, "BulletproofIP");
    inner_product_proof_free(&proof->ip_proof);
    inner_product_prove(&proof->ip_proof, &lx, &rx, G, &Hp, h, &t, tr0);
    point_vector_free(&Hp);
    field_vector_free(&aL);
    field_vector_free(&aR);
    field_vector_free(&sL);
    field_vector_free(&sR);
    field_vector_free(&l0);
    field_vector_free(&r0);
    field_vector_free(&r1);
    field_vector_free(&yn);
    field_vector_free(&lx);
    field_vector_free(&rx);
}

/* ------------------------------------------------------------------------------------------ */
/* synthetic generators: complete_bulletproof_test.cu:33-63 labels, mapped onto the curve       */
/* ------------------------------------------------------------------------------------------ */
void oracle_hash_to_point(ge25519 *out, const uint8_t seed[32], uint32_t index) {
    uint8_t in[40];
    memcpy(in, seed, 32);
    in[32] = (uint8_t)(index >> 24); /* big-endian index, as the reference's hash_input[32..35] */
    in[33] = (uint8_t)(index >> 16);
    in[34] = (uint8_t)(index >> 8);
    in[35] = (uint8_t)index;
    for (uint32_t ctr = 0;; ctr++) {
        in[36] = (uint8_t)(ctr >> 24);
        in[37] = (uint8_t)(ctr >> 16);
        in[38] = (uint8_t)(ctr >> 8);
        in[39] = (uint8_t)ctr;
        ge25519_compressed c;
        oracle_sha256(c.bytes, in, ctr == 0 ? 36 : 40); /* counter 0 hashes exactly the reference's 36 bytes */
        ge25519 p;
        if (!ge25519_unpack(&p, &c)) continue;
        ge25519 p2, p4, p8;
        ge25519_double(&p2, &p);
        ge25519_double(&p4, &p2);
        ge25519_double(&p8, &p4); /* clear the cofactor */
        if (ge25519_is_identity(&p8)) continue;
        ge25519_normalize(&p8);
        *out = p8;
        return;
    }
}
