/*
 * oracle/shim/ref_standins.cpp — TEST INFRASTRUCTURE.  Link-time stand-ins for the symbols the
 * verbatim reference host code needs but that live in its CUDA translation units or in OpenSSL:
 *
 *   cuda_point_vector_multi_scalar_mul  -> the reference's own CPU point_vector_multi_scalar_mul
 *   cuda_batch_field_{add,mul,square}   -> loops over the reference's own fe25519_{add,mul,sq}
 *   RAND_bytes                          -> deterministic SplitMix64 stream (refv_seed_rng)
 *   SHA256_Init/Update/Final            -> FIPS 180-4 implementation below (no libcrypto needed
 *                                          at run time; the struct is OpenSSL's SHA256_CTX)
 *
 * Everything here is this repository's own code; no reference source is copied.
 */
#include <openssl/sha.h>
#include <stdint.h>
#include <string.h>
#include "curve25519_ops.h"
#include "bulletproof_vectors.h"
#include "bulletproof_range_proof.h"
#include "bulletproof_challenge.h"

extern "C" {

void cuda_point_vector_multi_scalar_mul(ge25519* result, const FieldVector* scalars, const PointVector* points) {
    point_vector_multi_scalar_mul(result, scalars, points);
}
void cuda_batch_field_add(fe25519* r, const fe25519* a, const fe25519* b, size_t n) {
    for (size_t i = 0; i < n; i++) fe25519_add(&r[i], &a[i], &b[i]);
}
void cuda_batch_field_mul(fe25519* r, const fe25519* a, const fe25519* b, size_t n) {
    for (size_t i = 0; i < n; i++) fe25519_mul(&r[i], &a[i], &b[i]);
}
void cuda_batch_field_square(fe25519* r, const fe25519* a, size_t n) {
    for (size_t i = 0; i < n; i++) fe25519_sq(&r[i], &a[i]);
}

/* sizes/offsets of the reference's structs as its own headers define them (tests compare these
 * with include/cbp_types.h) */
size_t refv_sizeof(int which) {
    switch (which) {
        case 0: return sizeof(fe25519);
        case 1: return sizeof(ge25519);
        case 2: return sizeof(FieldVector);
        case 3: return sizeof(PointVector);
        case 4: return sizeof(InnerProductProof);
        case 5: return sizeof(RangeProof);
        case 6: return offsetof(InnerProductProof, L);
        case 7: return offsetof(InnerProductProof, x);
        case 8: return offsetof(RangeProof, taux);
        case 9: return offsetof(RangeProof, ip_proof);
    }
    return 0;
}


/* extern "C" trampolines: the reference's host functions have C++ linkage when built with g++ */
#define W(ret, name, params, args) ret refv_##name params { return name args; }
W(void, fe25519_add, (fe25519* h, const fe25519* f, const fe25519* g), (h, f, g))
W(void, fe25519_sub, (fe25519* h, const fe25519* f, const fe25519* g), (h, f, g))
W(void, fe25519_mul, (fe25519* h, const fe25519* f, const fe25519* g), (h, f, g))
W(void, fe25519_sq, (fe25519* h, const fe25519* f), (h, f))
W(void, fe25519_invert, (fe25519* h, const fe25519* f), (h, f))
W(void, fe25519_neg, (fe25519* h, const fe25519* f), (h, f))
W(void, fe25519_tobytes, (uint8_t* b, const fe25519* f), (b, f))
W(void, fe25519_frombytes, (fe25519* h, const uint8_t* b), (h, b))
W(void, ge25519_add, (ge25519* r, const ge25519* p, const ge25519* q), (r, p, q))
W(void, ge25519_scalarmult, (ge25519* r, const uint8_t* s, const ge25519* p), (r, s, p))
W(void, ge25519_normalize, (ge25519* p), (p))
W(void, generate_challenge, (uint8_t* o, const void* d, size_t l, const char* dom), (o, d, l, dom))
W(void, generate_challenge_y, (uint8_t* o, const ge25519* V, const ge25519* A, const ge25519* S), (o, V, A, S))
W(void, generate_challenge_z, (uint8_t* o, const uint8_t* y), (o, y))
W(void, generate_challenge_x, (uint8_t* o, const ge25519* T1, const ge25519* T2), (o, T1, T2))
W(void, field_vector_inner_product, (fe25519* r, const FieldVector* a, const FieldVector* b), (r, a, b))
W(void, point_vector_multi_scalar_mul, (ge25519* r, const FieldVector* s, const PointVector* p), (r, s, p))
W(bool, validate_range_input, (const fe25519* v, size_t n), (v, n))
W(void, pedersen_commit, (ge25519* r, const fe25519* v, const fe25519* b, const ge25519* g, const ge25519* h), (r, v, b, g, h))
W(void, generate_range_proof, (RangeProof* pr, const fe25519* v, const fe25519* gm, size_t n, const PointVector* G, const PointVector* H, const ge25519* g, const ge25519* h), (pr, v, gm, n, G, H, g, h))
W(bool, range_proof_verify, (const RangeProof* pr, const ge25519* V, size_t n, const PointVector* G, const PointVector* H, const ge25519* g, const ge25519* h), (pr, V, n, G, H, g, h))
W(void, range_proof_free, (RangeProof* pr), (pr))
#undef W

static uint64_t g_state = 0x9E3779B97F4A7C15ull;
void refv_seed_rng(uint64_t seed) { g_state = seed; }
int RAND_bytes(unsigned char* buf, int num) {
    for (int i = 0; i < num; i += 8) {
        uint64_t z = (g_state += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        for (int j = 0; j < 8 && i + j < num; j++) buf[i + j] = (unsigned char)(z >> (8 * j));
    }
    return 1;
}

static const uint32_t K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98,
    0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786,
    0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8,
    0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13,
    0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819,
    0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a,
    0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7,
    0xc67178f2};
static inline uint32_t ror(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
static void compress(SHA256_CTX* c, const unsigned char* p) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)p[4 * i] << 24) | ((uint32_t)p[4 * i + 1] << 16) | ((uint32_t)p[4 * i + 2] << 8) | p[4 * i + 3];
    for (int i = 16; i < 64; i++)
        w[i] = w[i - 16] + (ror(w[i - 15], 7) ^ ror(w[i - 15], 18) ^ (w[i - 15] >> 3)) + w[i - 7] +
               (ror(w[i - 2], 17) ^ ror(w[i - 2], 19) ^ (w[i - 2] >> 10));
    uint32_t s[8];
    for (int i = 0; i < 8; i++) s[i] = c->h[i];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = s[7] + (ror(s[4], 6) ^ ror(s[4], 11) ^ ror(s[4], 25)) + ((s[4] & s[5]) ^ (~s[4] & s[6])) + K[i] + w[i];
        uint32_t t2 = (ror(s[0], 2) ^ ror(s[0], 13) ^ ror(s[0], 22)) + ((s[0] & s[1]) ^ (s[0] & s[2]) ^ (s[1] & s[2]));
        s[7] = s[6]; s[6] = s[5]; s[5] = s[4]; s[4] = s[3] + t1; s[3] = s[2]; s[2] = s[1]; s[1] = s[0]; s[0] = t1 + t2;
    }
    for (int i = 0; i < 8; i++) c->h[i] += s[i];
}
int SHA256_Init(SHA256_CTX* c) {
    static const uint32_t iv[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    memset(c, 0, sizeof(*c));
    for (int i = 0; i < 8; i++) c->h[i] = iv[i];
    c->md_len = 32;
    return 1;
}
int SHA256_Update(SHA256_CTX* c, const void* data, size_t len) {
    const unsigned char* p = (const unsigned char*)data;
    unsigned char* buf = (unsigned char*)c->data;
    uint64_t bits = ((uint64_t)c->Nh << 32 | c->Nl) + (uint64_t)len * 8;
    c->Nl = (uint32_t)bits;
    c->Nh = (uint32_t)(bits >> 32);
    while (len) {
        size_t take = 64 - c->num;
        if (take > len) take = len;
        memcpy(buf + c->num, p, take);
        c->num += (unsigned)take;
        p += take;
        len -= take;
        if (c->num == 64) {
            compress(c, buf);
            c->num = 0;
        }
    }
    return 1;
}
int SHA256_Final(unsigned char* md, SHA256_CTX* c) {
    unsigned char* buf = (unsigned char*)c->data;
    uint64_t bits = (uint64_t)c->Nh << 32 | c->Nl;
    buf[c->num++] = 0x80;
    if (c->num > 56) {
        memset(buf + c->num, 0, 64 - c->num);
        compress(c, buf);
        c->num = 0;
    }
    memset(buf + c->num, 0, 56 - c->num);
    for (int i = 0; i < 8; i++) buf[63 - i] = (unsigned char)(bits >> (8 * i));
    compress(c, buf);
    for (int i = 0; i < 8; i++) {
        md[4 * i] = (unsigned char)(c->h[i] >> 24);
        md[4 * i + 1] = (unsigned char)(c->h[i] >> 16);
        md[4 * i + 2] = (unsigned char)(c->h[i] >> 8);
        md[4 * i + 3] = (unsigned char)c->h[i];
    }
    return 1;
}
}
