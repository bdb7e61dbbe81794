// sha256.cuh — SHA-256 (FIPS 180-4) for short messages, one thread per message.
//
// The reference derives every Fiat-Shamir challenge on the host with OpenSSL
// (bulletproof_challenge.cu:6-21: SHA-256(domain || data), then out[31] &= 0x7F).  Batched
// verification needs ~10 dependent hashes per proof, so they run on the device, one proof per
// lane; outputs are byte-identical to the host code.
#pragma once
#include <stdint.h>

namespace cbp {

__device__ __forceinline__ uint32_t sha_rotr(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

static __device__ __constant__ uint32_t kSha256K[64] = {
    0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
    0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
    0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
    0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
    0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
    0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
    0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
    0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};

// The whole (short) message is first assembled as big-endian words in local memory; all compressions
// then run from ONE rolled loop in final().  The callers hash ~10 messages per proof from straight-line
// code: with a fully unrolled compression inlined at every update site the verifier's transcript kernel
// was 1.5 MB of SASS and stalled on instruction fetch.
struct Sha256 {
    static constexpr int kMaxBytes = 256;  // longest transcript message: 212 bytes + padding
    uint32_t m[kMaxBytes / 4];
    uint32_t len;  // message bytes so far

    __device__ __forceinline__ void init() {
#pragma unroll 1
        for (int i = 0; i < kMaxBytes / 4; i++) m[i] = 0;
        len = 0;
    }
    __device__ __forceinline__ void put(uint8_t b) {
        if (len < kMaxBytes - 9) m[len >> 2] |= (uint32_t)b << (24 - 8 * (len & 3));  // callers stay far below
        len++;
    }
    __device__ __forceinline__ void update(const uint8_t* p, int n) {
#pragma unroll 1
        for (int i = 0; i < n; i++) put(p[i]);
    }
    // 32 little-endian bytes of a canonical field element / scalar given as 8 words
    __device__ __forceinline__ void update_words(const uint32_t (&v)[8]) {
        if ((len & 3) == 0 && len + 32 < kMaxBytes - 9) {
#pragma unroll
            for (int i = 0; i < 8; i++) m[(len >> 2) + i] = __byte_perm(v[i], 0, 0x0123);
            len += 32;
        } else {
#pragma unroll 1
            for (int i = 0; i < 32; i++) put((uint8_t)(v[i >> 2] >> (8 * (i & 3))));
        }
    }
    __device__ __forceinline__ void update_str(const char* s, int n) { update((const uint8_t*)s, n); }
    // pad, compress every block, leave the digest words (big-endian convention) in h
    __device__ __forceinline__ void final_words(uint32_t (&h)[8]) {
        const uint32_t bits = len * 8;
        put(0x80);
        const int nblocks = (int)((len + 8 + 63) >> 6);
        m[nblocks * 16 - 1] = bits;  // < 2^32 bits; the high length word stays 0
        h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
        h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
#pragma unroll 1
        for (int blk = 0; blk < nblocks; blk++) {
            uint32_t w[16];
#pragma unroll
            for (int i = 0; i < 16; i++) w[i] = m[blk * 16 + i];
            uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll 1
            for (int t = 0; t < 64; t += 16) {
#pragma unroll
                for (int j = 0; j < 16; j++) {
                    if (t > 0) {
                        uint32_t w15 = w[(j + 1) & 15], w2 = w[(j + 14) & 15];
                        uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                        uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                        w[j] = w[j] + s0 + w[(j + 9) & 15] + s1;
                    }
                    uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
                    uint32_t ch = (e & f) ^ (~e & g);
                    uint32_t t1 = hh + S1 + ch + kSha256K[t + j] + w[j];
                    uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
                    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
                    uint32_t t2 = S0 + mj;
                    hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
                }
            }
            h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
        }
    }
    __device__ __forceinline__ void final(uint8_t out[32]) {
        uint32_t h[8];
        final_words(h);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(h[i] >> 24);
            out[4 * i + 1] = (uint8_t)(h[i] >> 16);
            out[4 * i + 2] = (uint8_t)(h[i] >> 8);
            out[4 * i + 3] = (uint8_t)h[i];
        }
    }
    // bulletproof_challenge.cu:6-21: finalise and clear the top bit; returns 8 little-endian words
    __device__ __forceinline__ void final_challenge(uint32_t (&c)[8]) {
        uint32_t h[8];
        final_words(h);
#pragma unroll
        for (int i = 0; i < 8; i++) c[i] = __byte_perm(h[i], 0, 0x0123);
        c[7] &= 0x7FFFFFFFu;  // out[31] &= 0x7F
    }
};

}  // namespace cbp
