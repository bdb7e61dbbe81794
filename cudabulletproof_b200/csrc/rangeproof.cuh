// rangeproof.cuh — shared device definitions for range-proof verification / proving (rangeproof.cu,
// prover.cu): generator-table layout, proof-record layout, fixed-base accumulation.
#pragma once
#include <stdint.h>
#include "ge25519.cuh"
#include "sc25519.cuh"

namespace cbp {

constexpr int kMaxK = 6;          // log2 of the largest supported range width (n <= 64)
constexpr int kMaxN = 64;
constexpr int kFixRowBytes = 32;  // one row of fixed-base digits: 32 x int8 (8-bit windows) or 16 x int16
constexpr uint64_t kGensMagic = 0x62706b47454e5332ull;  // "bpkGENS2"

// Fixed-base tables come in two window widths.  8-bit: 32 windows x 128 multiples per base (51 MB at
// n = 64, lives in L2) — cheap to build, used by the single-proof host drop-ins.  16-bit: 16 windows x
// 32768 multiples (6.5 GB at n = 64, lives in HBM) — HALF the additions per scalar; every addition then
// reads 96 random bytes from HBM, ~0.2 MB per proof, far below what the integer pipe needs to hide.
// (Entries padded to 128 bytes — one cache line each, 37 % less DRAM traffic, 8.7 GB — were measured: 2.78 vs
// 2.81 ms for the verifier's fixed-base kernel, i.e. the traffic is not what limits it; 96 bytes stay.)
__host__ __device__ inline int fix_nwin(int wbits) { return 256 / wbits; }
__host__ __device__ inline uint32_t fix_entries(int wbits) { return 1u << (wbits - 1); }

// header of the generator workspace built by bpk_gens_init_device[_ex]
struct GensHeader {
    uint64_t magic;
    uint32_t n;
    uint32_t nbases;     // 2n + 2: G[0..n), H[0..n), g, h
    uint64_t table_off;  // niels table: ((base*nwin + win)*entries + (d-1)) * 96 bytes
    uint64_t scratch_off;
    uint64_t bases_off;  // nbases normalised ge25519 (128 B each)
    uint32_t wbits;      // 8 or 16
    uint32_t nwin;       // 256 / wbits
    uint32_t entries;    // 2^(wbits-1)
    uint32_t pad;
};
struct FixTab {  // device-side view of the table
    const uint8_t* table;
    int wbits, nwin;
    uint32_t entries;
};
__device__ __forceinline__ FixTab fixtab_of(const uint8_t* gens) {
    const GensHeader* gh = reinterpret_cast<const GensHeader*>(gens);
    FixTab ft;
    ft.table = gens + gh->table_off;
    ft.wbits = (int)gh->wbits;
    ft.nwin = (int)gh->nwin;
    ft.entries = gh->entries;
    return ft;
}

// proof record offsets (bytes), include/bpk.h
constexpr int kRecV = 0, kRecA = 128, kRecS = 256, kRecT1 = 384, kRecT2 = 512, kRecTaux = 640, kRecMu = 672,
              kRecT = 704, kRecIpA = 736, kRecIpB = 768, kRecIpC = 800, kRecIpX = 832, kRecL = 864;
__host__ __device__ inline size_t proof_record_bytes(int k) { return (size_t)kRecL + 256 * (size_t)k; }

struct ge_cached {  // extended precomputed: (Y+X, Y-X, 2Z, 2dT)
    fe YplusX, YminusX, Z2, T2d;
};
__device__ __forceinline__ void ge_to_cached(ge_cached& r, const ge_p3& p) {
    fe_add(r.YplusX, p.Y, p.X);
    fe_sub(r.YminusX, p.Y, p.X);
    fe_dbl(r.Z2, p.Z);
    fe_mul(r.T2d, p.T, fe_const_2d());
}
// r = p + q (q cached) or p - q.  8M.
__device__ __forceinline__ void ge_add_cached(ge_p3& r, const ge_p3& p, const ge_cached& q, bool neg) {
    fe A, B, C, D, E, F, G, H, t;
    fe qa = q.YminusX, qb = q.YplusX;
    fe_cswap(qa, qb, neg);
    fe_sub(t, p.Y, p.X);
    fe_mul(A, t, qa);
    fe_add(t, p.Y, p.X);
    fe_mul(B, t, qb);
    fe_mul(C, p.T, q.T2d);
    fe_mul(D, p.Z, q.Z2);
    fe_sub(E, B, A);
    fe_add(H, B, A);
    fe Fm, Fp;
    fe_sub(Fm, D, C);
    fe_add(Fp, D, C);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        F.v[i] = neg ? Fp.v[i] : Fm.v[i];
        G.v[i] = neg ? Fm.v[i] : Fp.v[i];
    }
    fe_mul(r.X, E, F);
    fe_mul(r.Y, G, H);
    fe_mul(r.Z, F, G);
    fe_mul(r.T, E, H);
}

// signed base-2^w recoding of a scalar < 2^253 into ndig digits in [-(2^(w-1)-1), 2^(w-1)]
template <int WBITS, typename OUT = int8_t>
__device__ __forceinline__ void sc_recode_signed(OUT* out, const sc& k, int ndig) {
    uint32_t carry = 0;
    constexpr uint32_t half = 1u << (WBITS - 1);
    for (int j = 0; j < ndig; j++) {
        int bit = j * WBITS;
        uint32_t d = 0;
        if (bit < 256) {  // windows may straddle a 32-bit word (WBITS = 5)
            int word = bit >> 5, sh = bit & 31;
            uint64_t v = k.v[word];
            if (word + 1 < 8) v |= (uint64_t)k.v[word + 1] << 32;
            d = (uint32_t)(v >> sh) & ((1u << WBITS) - 1u);
        }
        d += carry;
        carry = 0;
        int v = (int)d;
        if (d > half) {
            v = (int)d - (1 << WBITS);
            carry = 1;
        }
        out[j] = (OUT)v;  // +2^(WBITS-1) wraps to the most negative value of OUT; see fixed_digit()
    }
}
// decode an 8-bit window digit written by sc_recode_signed<8>: +128 wraps to int8 -128 but real
// negative digits only reach -127, so -128 means +128
__device__ __forceinline__ void fixed_digit(int8_t raw, uint32_t& mag, bool& neg) {
    int v = raw;
    if (v == -128) {
        mag = 128;
        neg = false;
    } else if (v < 0) {
        mag = (uint32_t)(-v);
        neg = true;
    } else {
        mag = (uint32_t)v;
        neg = false;
    }
}

__device__ __forceinline__ void fixed_digit16(int16_t raw, uint32_t& mag, bool& neg) {
    int v = raw;
    if (v == -32768) {
        mag = 32768;
        neg = false;
    } else if (v < 0) {
        mag = (uint32_t)(-v);
        neg = true;
    } else {
        mag = (uint32_t)v;
        neg = false;
    }
}
// one 32-byte digit row of a scalar < 2^253 in the table's window width
__device__ __forceinline__ void fix_recode(int8_t* row, const sc& k, int wbits) {
    if (wbits == 8) sc_recode_signed<8, int8_t>(row, k, 32);
    else sc_recode_signed<16, int16_t>(reinterpret_cast<int16_t*>(row), k, 16);
}
__device__ __forceinline__ void fix_digit(const int8_t* row, int win, int wbits, uint32_t& mag, bool& neg) {
    if (wbits == 8) fixed_digit(row[win], mag, neg);
    else fixed_digit16(reinterpret_cast<const int16_t*>(row)[win], mag, neg);
}
__device__ __forceinline__ const uint8_t* fix_entry(const FixTab& ft, uint32_t base, int win, uint32_t mag) {
    return ft.table + (((size_t)base * ft.nwin + win) * ft.entries + (mag - 1)) * 96;
}
// acc += digit * 2^(wbits win) * Base[base]  from the precomputed table
__device__ __forceinline__ void fixed_base_madd(ge_p3& acc, const FixTab& ft, uint32_t base, int win,
                                                const int8_t* row) {
    uint32_t mag;
    bool neg;
    fix_digit(row, win, ft.wbits, mag, neg);
    if (mag == 0) return;
    ge_niels q;
    ge_niels_load(q, fix_entry(ft, base, win, mag));
    ge_madd(acc, acc, q, neg);
}

// CTA-wide sum of one point per thread through shared memory; result valid in thread 0.
// red must hold blockDim.x points; blockDim.x must be a power of two.
__device__ __forceinline__ void cta_point_sum(ge_p3& acc, ge_p3* red) {
    int t = threadIdx.x;
    red[t] = acc;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if (t < o) {
            ge_p3 a = red[t], b = red[t + o];
            ge_add(a, a, b);
            red[t] = a;
        }
        __syncthreads();
    }
    acc = red[0];
    __syncthreads();
}

}  // namespace cbp
