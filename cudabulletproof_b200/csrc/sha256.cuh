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

struct Sha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint32_t len;  // total bytes so far

    __device__ void init() {
        h[0] = 0x6a09e667u; h[1] = 0xbb67ae85u; h[2] = 0x3c6ef372u; h[3] = 0xa54ff53au;
        h[4] = 0x510e527fu; h[5] = 0x9b05688cu; h[6] = 0x1f83d9abu; h[7] = 0x5be0cd19u;
        len = 0;
    }
    __device__ void compress() {
        const uint32_t K[64] = {
            0x428a2f98u, 0x71374491u, 0xb5c0fbcfu, 0xe9b5dba5u, 0x3956c25bu, 0x59f111f1u, 0x923f82a4u, 0xab1c5ed5u,
            0xd807aa98u, 0x12835b01u, 0x243185beu, 0x550c7dc3u, 0x72be5d74u, 0x80deb1feu, 0x9bdc06a7u, 0xc19bf174u,
            0xe49b69c1u, 0xefbe4786u, 0x0fc19dc6u, 0x240ca1ccu, 0x2de92c6fu, 0x4a7484aau, 0x5cb0a9dcu, 0x76f988dau,
            0x983e5152u, 0xa831c66du, 0xb00327c8u, 0xbf597fc7u, 0xc6e00bf3u, 0xd5a79147u, 0x06ca6351u, 0x14292967u,
            0x27b70a85u, 0x2e1b2138u, 0x4d2c6dfcu, 0x53380d13u, 0x650a7354u, 0x766a0abbu, 0x81c2c92eu, 0x92722c85u,
            0xa2bfe8a1u, 0xa81a664bu, 0xc24b8b70u, 0xc76c51a3u, 0xd192e819u, 0xd6990624u, 0xf40e3585u, 0x106aa070u,
            0x19a4c116u, 0x1e376c08u, 0x2748774cu, 0x34b0bcb5u, 0x391c0cb3u, 0x4ed8aa4au, 0x5b9cca4fu, 0x682e6ff3u,
            0x748f82eeu, 0x78a5636fu, 0x84c87814u, 0x8cc70208u, 0x90befffau, 0xa4506cebu, 0xbef9a3f7u, 0xc67178f2u};
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; i++)
            w[i] = ((uint32_t)buf[4 * i] << 24) | ((uint32_t)buf[4 * i + 1] << 16) | ((uint32_t)buf[4 * i + 2] << 8) |
                   buf[4 * i + 3];
        uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
        for (int i = 0; i < 64; i++) {
            uint32_t wi;
            if (i < 16) {
                wi = w[i];
            } else {
                uint32_t w15 = w[(i - 15) & 15], w2 = w[(i - 2) & 15];
                uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                wi = w[i & 15] + s0 + w[(i - 7) & 15] + s1;
                w[i & 15] = wi;
            }
            uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
            uint32_t ch = (e & f) ^ (~e & g);
            uint32_t t1 = hh + S1 + ch + K[i] + wi;
            uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
            uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t t2 = S0 + mj;
            hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
    }
    __device__ void update(const uint8_t* p, int n) {
        for (int i = 0; i < n; i++) {
            buf[len & 63] = p[i];
            len++;
            if ((len & 63) == 0) compress();
        }
    }
    // 32 little-endian bytes of a canonical field element / scalar given as 8 words
    __device__ void update_words(const uint32_t (&v)[8]) {
        for (int i = 0; i < 8; i++) {
            uint8_t b[4] = {(uint8_t)v[i], (uint8_t)(v[i] >> 8), (uint8_t)(v[i] >> 16), (uint8_t)(v[i] >> 24)};
            update(b, 4);
        }
    }
    __device__ void update_str(const char* s, int n) { update((const uint8_t*)s, n); }
    __device__ void final(uint8_t out[32]) {
        uint64_t bits = (uint64_t)len * 8;
        uint8_t pad = 0x80;
        update(&pad, 1);
        pad = 0;
        while ((len & 63) != 56) update(&pad, 1);
        uint8_t lb[8];
        for (int i = 0; i < 8; i++) lb[i] = (uint8_t)(bits >> (56 - 8 * i));
        update(lb, 8);
        for (int i = 0; i < 8; i++) {
            out[4 * i] = (uint8_t)(h[i] >> 24);
            out[4 * i + 1] = (uint8_t)(h[i] >> 16);
            out[4 * i + 2] = (uint8_t)(h[i] >> 8);
            out[4 * i + 3] = (uint8_t)h[i];
        }
    }
    // bulletproof_challenge.cu:6-21: finalise and clear the top bit; returns 8 little-endian words
    __device__ void final_challenge(uint32_t (&c)[8]) {
        uint8_t out[32];
        final(out);
        out[31] &= 0x7F;
        for (int i = 0; i < 8; i++)
            c[i] = (uint32_t)out[4 * i] | ((uint32_t)out[4 * i + 1] << 8) | ((uint32_t)out[4 * i + 2] << 16) |
                   ((uint32_t)out[4 * i + 3] << 24);
    }
};

}  // namespace cbp
