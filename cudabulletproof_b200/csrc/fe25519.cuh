// fe25519.cuh — GF(2^255-19) arithmetic for sm_100a.
//
// Replaces the reference's device field layer (device_curve25519_ops.cuh:60-186, a bug-for-bug
// copy of curve25519_ops.cu:41-146) wholesale.  Representation: 8 x 32-bit saturated limbs
// (radix 2^32, little-endian), i.e. the same 32 bytes as the reference's fe25519 (4 x u64 LE), so
// loads/stores need no conversion.  Values are kept only weakly reduced (any residue in
// [0, 2^256)); fe_canon() produces the unique representative < p for output/compare.
//
// The 256x256 product is 64 32x32->64 multiply-adds issued as carry chains
// (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32(.X) on the fmaheavy pipe);
// the reduction uses 2^256 = 38 (mod p): 8 more multiply-adds.  72 IMAD per fe_mul, 44 per fe_sq
// — the figures DESIGN.md's roofline uses.
#pragma once
#include <stdint.h>
#include "modinv.cuh"

namespace cbp {

struct fe {
    uint32_t v[8];
};

// ------------------------------------------------------------------------------------------
// carry-chain building blocks (one asm block per chain: the CC flag is implicit state).
// EVERY written operand is early-clobber ("=&r", "+&r"): each block is several instructions and writes
// its first outputs before it has read its last inputs, so a written operand must never share a register
// with an input.  Without '&' the compiler may coalesce them: for write-only outputs always, for
// read-write accumulators whenever it knows that an input holds the same VALUE as the accumulator's
// initial contents (e.g. both the constant 0) — which silently corrupts the chain.  That second case only
// shows up with compile-time-known operands (squaring the constant 1 returned 39 * (1 + 2^64 + ...));
// tests/test_gpu_codec.py::test_field_ops_on_compile_time_constants guards it.
// ------------------------------------------------------------------------------------------

// acc[0..7] (+carry into acc[8]) += {x0,x1,x2,x3} * b, product k landing on words (2k, 2k+1)
__device__ __forceinline__ void mad_row4(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                                         uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t& c8, uint32_t x0,
                                         uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4), "+&r"(c5), "+&r"(c6), "+&r"(c7), "+&r"(c8)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
}
// same without the carry-out word (used where the bound on the total proves it is zero)
__device__ __forceinline__ void mad_row4_nc(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                                            uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t x0, uint32_t x1,
                                            uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4), "+&r"(c5), "+&r"(c6), "+&r"(c7)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
}
// first row: plain products, no incoming accumulator
__device__ __forceinline__ void mul_row4(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                                         uint32_t& c5, uint32_t& c6, uint32_t& c7, uint32_t x0, uint32_t x1,
                                         uint32_t x2, uint32_t x3, uint32_t b) {
    asm("mul.lo.u32 %0, %8,  %12;\n\t"
        "mul.hi.u32 %1, %8,  %12;\n\t"
        "mul.lo.u32 %2, %9,  %12;\n\t"
        "mul.hi.u32 %3, %9,  %12;\n\t"
        "mul.lo.u32 %4, %10, %12;\n\t"
        "mul.hi.u32 %5, %10, %12;\n\t"
        "mul.lo.u32 %6, %11, %12;\n\t"
        "mul.hi.u32 %7, %11, %12;"
        : "=&r"(c0), "=&r"(c1), "=&r"(c2), "=&r"(c3), "=&r"(c4), "=&r"(c5), "=&r"(c6), "=&r"(c7)
        : "r"(x0), "r"(x1), "r"(x2), "r"(x3), "r"(b));
}

// r[0..7] = lo[0..7] + 38 * hi[0..7]  (mod p, weakly reduced).  2^256 = 38 (mod 2^255-19).
__device__ __forceinline__ void fe_fold(fe& r, const uint32_t (&w)[16]) {
    // even hi words land on (0,1),(2,3),(4,5),(6,7); odd hi words on (1,2),(3,4),(5,6),(7,8)
    uint32_t e0 = w[0], e1 = w[1], e2 = w[2], e3 = w[3], e4 = w[4], e5 = w[5], e6 = w[6], e7 = w[7], e8 = 0;
    mad_row4(e0, e1, e2, e3, e4, e5, e6, e7, e8, w[8], w[10], w[12], w[14], 38u);
    uint32_t o1, o2, o3, o4, o5, o6, o7, o8;
    mul_row4(o1, o2, o3, o4, o5, o6, o7, o8, w[9], w[11], w[13], w[15], 38u);
    uint32_t top;
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, %10;\n\t"
        "addc.cc.u32 %2, %2, %11;\n\t"
        "addc.cc.u32 %3, %3, %12;\n\t"
        "addc.cc.u32 %4, %4, %13;\n\t"
        "addc.cc.u32 %5, %5, %14;\n\t"
        "addc.cc.u32 %6, %6, %15;\n\t"
        "addc.u32    %7, %8, %16;"
        : "+&r"(e1), "+&r"(e2), "+&r"(e3), "+&r"(e4), "+&r"(e5), "+&r"(e6), "+&r"(e7), "=&r"(top)
        : "r"(e8), "r"(o1), "r"(o2), "r"(o3), "r"(o4), "r"(o5), "r"(o6), "r"(o7), "r"(o8));
    // value = e[0..7] + top * 2^256 with top <= 38: fold again, then once more for the final carry
    uint32_t t = top * 38u, c;
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32    %8, 0, 0;"
        : "+&r"(e0), "+&r"(e1), "+&r"(e2), "+&r"(e3), "+&r"(e4), "+&r"(e5), "+&r"(e6), "+&r"(e7), "=&r"(c)
        : "r"(t));
    e0 += c * 38u;  // after a wrap the low words are < 2^12, so this cannot carry
    r.v[0] = e0; r.v[1] = e1; r.v[2] = e2; r.v[3] = e3; r.v[4] = e4; r.v[5] = e5; r.v[6] = e6; r.v[7] = e7;
}

// full 512-bit product into w[0..15]: even-position products accumulate in E (words 0..15),
// odd-position products in O (words 1..15), combined once at the end.
__device__ __forceinline__ void mul_wide(uint32_t (&w)[16], const fe& a, const fe& b) {
    uint32_t E[16], O[15];
#pragma unroll
    for (int i = 0; i < 16; i++) E[i] = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) O[i] = 0;
    const uint32_t* x = a.v;
    const uint32_t* y = b.v;
    mul_row4(E[0], E[1], E[2], E[3], E[4], E[5], E[6], E[7], x[0], x[2], x[4], x[6], y[0]);
    mul_row4(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], x[1], x[3], x[5], x[7], y[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) {
        if (i & 1) {
            // a-even x b_i -> odd positions i+j  -> O[i-1 ..]; a-odd x b_i -> even positions -> E[i+1 ..]
            mad_row4(O[i - 1], O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7], x[0], x[2],
                     x[4], x[6], y[i]);
            if (i < 7)
                mad_row4(E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8], E[i + 9], x[1],
                         x[3], x[5], x[7], y[i]);
            else
                mad_row4_nc(E[8], E[9], E[10], E[11], E[12], E[13], E[14], E[15], x[1], x[3], x[5], x[7], y[7]);
        } else {
            mad_row4(E[i], E[i + 1], E[i + 2], E[i + 3], E[i + 4], E[i + 5], E[i + 6], E[i + 7], E[i + 8], x[0], x[2],
                     x[4], x[6], y[i]);
            mad_row4(O[i], O[i + 1], O[i + 2], O[i + 3], O[i + 4], O[i + 5], O[i + 6], O[i + 7], O[i + 8], x[1], x[3],
                     x[5], x[7], y[i]);
        }
    }
    w[0] = E[0];
    asm("add.cc.u32  %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32    %14, %29, %44;"
        : "=&r"(w[1]), "=&r"(w[2]), "=&r"(w[3]), "=&r"(w[4]), "=&r"(w[5]), "=&r"(w[6]), "=&r"(w[7]), "=&r"(w[8]), "=&r"(w[9]),
          "=&r"(w[10]), "=&r"(w[11]), "=&r"(w[12]), "=&r"(w[13]), "=&r"(w[14]), "=&r"(w[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]), "r"(O[2]),
          "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]),
          "r"(O[12]), "r"(O[13]), "r"(O[14]));
}

// Note (measured, profiles/r01_microbench_int_pipe.jsonl): one level of subtractive Karatsuba (48 wide
// multiplies + ~80 ALU instructions instead of 64) was tried and is SLOWER on B200 (100 vs 112 G fe_mul/s):
// ptxas moves part of the extra carry/negate work onto the same FMA pipe (IMAD.MOV / IMAD.X), so the
// schoolbook carry-chain form below stays.
__device__ __forceinline__ void fe_mul(fe& r, const fe& a, const fe& b) {
    uint32_t w[16];
    mul_wide(w, a, b);
    fe_fold(r, w);
}
// ---- dedicated squaring: 28 off-diagonal products, doubled, plus 8 squares = 36 wide multiplies
// (+8 for the fold) instead of 64 (+8) ----------------------------------------------------------------
__device__ __forceinline__ void mul_row3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                                         uint32_t& c5, uint32_t x0, uint32_t x1, uint32_t x2, uint32_t b) {
    asm("mul.lo.u32 %0, %6, %9;\n\t"
        "mul.hi.u32 %1, %6, %9;\n\t"
        "mul.lo.u32 %2, %7, %9;\n\t"
        "mul.hi.u32 %3, %7, %9;\n\t"
        "mul.lo.u32 %4, %8, %9;\n\t"
        "mul.hi.u32 %5, %8, %9;"
        : "=&r"(c0), "=&r"(c1), "=&r"(c2), "=&r"(c3), "=&r"(c4), "=&r"(c5)
        : "r"(x0), "r"(x1), "r"(x2), "r"(b));
}
// c[0..5] (+carry into c6) += {x0,x1,x2} * b
__device__ __forceinline__ void mad_row3(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                                         uint32_t& c5, uint32_t& c6, uint32_t x0, uint32_t x1, uint32_t x2,
                                         uint32_t b) {
    asm("mad.lo.cc.u32  %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32       %6, %6, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4), "+&r"(c5), "+&r"(c6)
        : "r"(x0), "r"(x1), "r"(x2), "r"(b));
}
__device__ __forceinline__ void mad_row2(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t& c3, uint32_t& c4,
                                         uint32_t x0, uint32_t x1, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32       %4, %4, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2), "+&r"(c3), "+&r"(c4)
        : "r"(x0), "r"(x1), "r"(b));
}
__device__ __forceinline__ void mad_row1(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t x0, uint32_t b) {
    asm("mad.lo.cc.u32  %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32       %2, %2, 0;"
        : "+&r"(c0), "+&r"(c1), "+&r"(c2)
        : "r"(x0), "r"(b));
}
__device__ __forceinline__ void sq_wide(uint32_t (&w)[16], const fe& a) {
    const uint32_t* x = a.v;
    // off-diagonal products x_i x_j (i < j) at word position i + j: even positions in E, odd in O (O[k] = word k+1)
    uint32_t E[16], O[15];
#pragma unroll
    for (int i = 0; i < 16; i++) E[i] = 0;
#pragma unroll
    for (int i = 0; i < 15; i++) O[i] = 0;
    mul_row4(O[0], O[1], O[2], O[3], O[4], O[5], O[6], O[7], x[1], x[3], x[5], x[7], x[0]);  // pos 1,3,5,7
    mul_row3(E[2], E[3], E[4], E[5], E[6], E[7], x[2], x[4], x[6], x[0]);                    // pos 2,4,6
    mad_row3(O[2], O[3], O[4], O[5], O[6], O[7], O[8], x[2], x[4], x[6], x[1]);              // pos 3,5,7
    mad_row3(E[4], E[5], E[6], E[7], E[8], E[9], E[10], x[3], x[5], x[7], x[1]);             // pos 4,6,8
    mad_row3(O[4], O[5], O[6], O[7], O[8], O[9], O[10], x[3], x[5], x[7], x[2]);             // pos 5,7,9
    mad_row2(E[6], E[7], E[8], E[9], E[10], x[4], x[6], x[2]);                               // pos 6,8
    mad_row2(O[6], O[7], O[8], O[9], O[10], x[4], x[6], x[3]);                               // pos 7,9
    mad_row2(E[8], E[9], E[10], E[11], E[12], x[5], x[7], x[3]);                             // pos 8,10
    mad_row2(O[8], O[9], O[10], O[11], O[12], x[5], x[7], x[4]);                             // pos 9,11
    mad_row1(E[10], E[11], E[12], x[6], x[4]);                                               // pos 10
    mad_row1(O[10], O[11], O[12], x[6], x[5]);                                               // pos 11
    mad_row1(E[12], E[13], E[14], x[7], x[5]);                                               // pos 12
    mad_row1(O[12], O[13], O[14], x[7], x[6]);                                               // pos 13
    // S = E + (O << 32)
    uint32_t S[16];
    S[0] = 0;  // no product lands on word 0
    asm("add.cc.u32  %0, %15, %30;\n\t"
        "addc.cc.u32 %1, %16, %31;\n\t"
        "addc.cc.u32 %2, %17, %32;\n\t"
        "addc.cc.u32 %3, %18, %33;\n\t"
        "addc.cc.u32 %4, %19, %34;\n\t"
        "addc.cc.u32 %5, %20, %35;\n\t"
        "addc.cc.u32 %6, %21, %36;\n\t"
        "addc.cc.u32 %7, %22, %37;\n\t"
        "addc.cc.u32 %8, %23, %38;\n\t"
        "addc.cc.u32 %9, %24, %39;\n\t"
        "addc.cc.u32 %10, %25, %40;\n\t"
        "addc.cc.u32 %11, %26, %41;\n\t"
        "addc.cc.u32 %12, %27, %42;\n\t"
        "addc.cc.u32 %13, %28, %43;\n\t"
        "addc.u32    %14, %29, %44;"
        : "=&r"(S[1]), "=&r"(S[2]), "=&r"(S[3]), "=&r"(S[4]), "=&r"(S[5]), "=&r"(S[6]), "=&r"(S[7]), "=&r"(S[8]),
          "=&r"(S[9]), "=&r"(S[10]), "=&r"(S[11]), "=&r"(S[12]), "=&r"(S[13]), "=&r"(S[14]), "=&r"(S[15])
        : "r"(E[1]), "r"(E[2]), "r"(E[3]), "r"(E[4]), "r"(E[5]), "r"(E[6]), "r"(E[7]), "r"(E[8]), "r"(E[9]),
          "r"(E[10]), "r"(E[11]), "r"(E[12]), "r"(E[13]), "r"(E[14]), "r"(E[15]), "r"(O[0]), "r"(O[1]), "r"(O[2]),
          "r"(O[3]), "r"(O[4]), "r"(O[5]), "r"(O[6]), "r"(O[7]), "r"(O[8]), "r"(O[9]), "r"(O[10]), "r"(O[11]),
          "r"(O[12]), "r"(O[13]), "r"(O[14]));
    // w = 2 S (S < 2^511) ...
#pragma unroll
    for (int i = 15; i > 0; i--) w[i] = __funnelshift_l(S[i - 1], S[i], 1);
    w[0] = 0;
    // ... + the squares x_i^2 at words (2i, 2i+1): one carry chain
    asm("mad.lo.cc.u32  %0, %16, %16, %0;\n\t"
        "madc.hi.cc.u32 %1, %16, %16, %1;\n\t"
        "madc.lo.cc.u32 %2, %17, %17, %2;\n\t"
        "madc.hi.cc.u32 %3, %17, %17, %3;\n\t"
        "madc.lo.cc.u32 %4, %18, %18, %4;\n\t"
        "madc.hi.cc.u32 %5, %18, %18, %5;\n\t"
        "madc.lo.cc.u32 %6, %19, %19, %6;\n\t"
        "madc.hi.cc.u32 %7, %19, %19, %7;\n\t"
        "madc.lo.cc.u32 %8, %20, %20, %8;\n\t"
        "madc.hi.cc.u32 %9, %20, %20, %9;\n\t"
        "madc.lo.cc.u32 %10, %21, %21, %10;\n\t"
        "madc.hi.cc.u32 %11, %21, %21, %11;\n\t"
        "madc.lo.cc.u32 %12, %22, %22, %12;\n\t"
        "madc.hi.cc.u32 %13, %22, %22, %13;\n\t"
        "madc.lo.cc.u32 %14, %23, %23, %14;\n\t"
        "madc.hi.u32    %15, %23, %23, %15;"
        : "+&r"(w[0]), "+&r"(w[1]), "+&r"(w[2]), "+&r"(w[3]), "+&r"(w[4]), "+&r"(w[5]), "+&r"(w[6]), "+&r"(w[7]), "+&r"(w[8]),
          "+&r"(w[9]), "+&r"(w[10]), "+&r"(w[11]), "+&r"(w[12]), "+&r"(w[13]), "+&r"(w[14]), "+&r"(w[15])
        : "r"(x[0]), "r"(x[1]), "r"(x[2]), "r"(x[3]), "r"(x[4]), "r"(x[5]), "r"(x[6]), "r"(x[7]));
}
__device__ __forceinline__ void fe_sq(fe& r, const fe& a) {
    uint32_t w[16];
    sq_wide(w, a);
    fe_fold(r, w);
}

// r = a + b (weakly reduced)
__device__ __forceinline__ void fe_add(fe& r, const fe& a, const fe& b) {
    uint32_t c;
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    uint32_t t = c * 38u, c2;
    asm("add.cc.u32  %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32    %8, 0, 0;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]),
          "+&r"(r.v[7]), "=&r"(c2)
        : "r"(t));
    r.v[0] += c2 * 38u;
}

// r = a - b (weakly reduced)
__device__ __forceinline__ void fe_sub(fe& r, const fe& a, const fe& b) {
    uint32_t bw;
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"  // 0 or 0xFFFFFFFF
        : "=&r"(r.v[0]), "=&r"(r.v[1]), "=&r"(r.v[2]), "=&r"(r.v[3]), "=&r"(r.v[4]), "=&r"(r.v[5]), "=&r"(r.v[6]),
          "=&r"(r.v[7]), "=&r"(bw)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // a borrow means the true value is r - 2^256 = r - 38 (mod p)
    uint32_t t = bw & 38u, b2;
    asm("sub.cc.u32  %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32    %8, 0, 0;"
        : "+&r"(r.v[0]), "+&r"(r.v[1]), "+&r"(r.v[2]), "+&r"(r.v[3]), "+&r"(r.v[4]), "+&r"(r.v[5]), "+&r"(r.v[6]),
          "+&r"(r.v[7]), "=&r"(b2)
        : "r"(t));
    r.v[0] -= b2 & 38u;  // after a second wrap the value is >= 2^256 - 38, so this cannot borrow
}

__device__ __forceinline__ void fe_set0(fe& r) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = 0;
}
__device__ __forceinline__ void fe_set1(fe& r) {
    fe_set0(r);
    r.v[0] = 1;
}
__device__ __forceinline__ void fe_neg(fe& r, const fe& a) {
    fe z;
    fe_set0(z);
    fe_sub(r, z, a);
}
// r = 2a
__device__ __forceinline__ void fe_dbl(fe& r, const fe& a) { fe_add(r, a, a); }

// conditional select / negate helpers (branch-free)
__device__ __forceinline__ void fe_cmov(fe& r, const fe& a, bool cond) {
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = cond ? a.v[i] : r.v[i];
}
__device__ __forceinline__ void fe_cswap(fe& a, fe& b, bool cond) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint32_t x = a.v[i], y = b.v[i];
        a.v[i] = cond ? y : x;
        b.v[i] = cond ? x : y;
    }
}

// t += k for a small k (no carry out of word 7 is possible at the call sites)
__device__ __forceinline__ void fe_add_small(fe& t, uint32_t k) {
    asm("add.cc.u32  %0, %0, %8;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.u32    %7, %7, 0;"
        : "+&r"(t.v[0]), "+&r"(t.v[1]), "+&r"(t.v[2]), "+&r"(t.v[3]), "+&r"(t.v[4]), "+&r"(t.v[5]), "+&r"(t.v[6]),
          "+&r"(t.v[7])
        : "r"(k));
}
// unique representative in [0, p).  2^255 = 19 (mod p): fold bit 255 twice, then one conditional
// subtraction of p done as "(t + 19) with bit 255 cleared".
__device__ __forceinline__ void fe_canon(fe& r) {
#pragma unroll
    for (int pass = 0; pass < 2; pass++) {
        uint32_t b = r.v[7] >> 31;
        r.v[7] &= 0x7fffffffu;
        fe_add_small(r, 19u * b);
    }
    fe t = r;
    fe_add_small(t, 19u);
    bool ge = (t.v[7] >> 31) != 0;  // r + 19 >= 2^255  <=>  r >= p
    t.v[7] &= 0x7fffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = ge ? t.v[i] : r.v[i];
}
__device__ __forceinline__ bool fe_iszero(const fe& a) {
    fe t = a;
    fe_canon(t);
    uint32_t o = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) o |= t.v[i];
    return o == 0;
}
__device__ __forceinline__ bool fe_equal(const fe& a, const fe& b) {
    fe d;
    fe_sub(d, a, b);
    return fe_iszero(d);
}

__device__ __forceinline__ void fe_sqn(fe& r, const fe& a, int n) {
    fe_sq(r, a);
#pragma unroll 1  // a dependent chain: unrolling only bloats the code (an inversion was 300 KB of SASS)
    for (int i = 1; i < n; i++) fe_sq(r, r);
}
// z^(2^250-1), z^11: shared prefix of inversion and square-root chains
static __device__ __noinline__ void fe_pow_2_250_1(fe& z_250_0, fe& z11, const fe& z) {
    fe z2, z9, t, z_5_0, z_10_0, z_20_0, z_40_0, z_50_0, z_100_0;
    fe_sq(z2, z);
    fe_sqn(t, z2, 2);
    fe_mul(z9, t, z);
    fe_mul(z11, z9, z2);
    fe_sq(t, z11);
    fe_mul(z_5_0, t, z9);
    fe_sqn(t, z_5_0, 5);
    fe_mul(z_10_0, t, z_5_0);
    fe_sqn(t, z_10_0, 10);
    fe_mul(z_20_0, t, z_10_0);
    fe_sqn(t, z_20_0, 20);
    fe_mul(z_40_0, t, z_20_0);
    fe_sqn(t, z_40_0, 10);
    fe_mul(z_50_0, t, z_10_0);
    fe_sqn(t, z_50_0, 50);
    fe_mul(z_100_0, t, z_50_0);
    fe_sqn(t, z_100_0, 100);
    fe_mul(t, t, z_100_0);  // z_200_0
    fe_sqn(t, t, 50);
    fe_mul(z_250_0, t, z_50_0);
}
// r = a^(p-2); inv(0) = 0.  254 squarings + 11 multiplications: the Fermat chain (replaces the truncated chain of
// curve25519_ops.cu:157-207, defect D4).  Kept as the cross-check of fe_invert below and for the square root.
__device__ __forceinline__ void fe_invert_fermat(fe& r, const fe& a) {
    fe z_250_0, z11, t;
    fe_pow_2_250_1(z_250_0, z11, a);
    fe_sqn(t, z_250_0, 5);
    fe_mul(r, t, z11);
}
// r = 1 / a mod p, canonical; inv(0) = 0.  Bernstein-Yang divsteps (modinv.cuh): at most 9 batches of 62 divsteps
// instead of 254 dependent squarings — 4x less latency for a lone thread (50 us -> 12 us), and every inversion on
// this path is of a public value (Z coordinates of results), so variable time is fine.  One copy per kernel.
static __device__ __noinline__ void fe_invert(fe& r, const fe& a) {
    const uint32_t pw[8] = {0xFFFFFFEDu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0x7FFFFFFFu};
    const ModInfo mi = modinfo_from_words(pw);
    uint32_t out[8];
    modinv_words(out, a.v, mi);
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = out[i];
}
// r = a^((p-5)/8) = a^(2^252-3)
__device__ __forceinline__ void fe_pow2523(fe& r, const fe& a) {
    fe z_250_0, z11, t;
    fe_pow_2_250_1(z_250_0, z11, a);
    fe_sqn(t, z_250_0, 2);
    fe_mul(r, t, a);
}

// 32-byte loads/stores.  The reference's fe25519 is 4 x u64 LE == 8 x u32 LE == 2 x uint4.
__device__ __forceinline__ void fe_load(fe& r, const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = q[0], hi = q[1];
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
}
__device__ __forceinline__ void fe_load_nc(fe& r, const void* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 lo = __ldg(q), hi = __ldg(q + 1);
    r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
    r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
}
__device__ __forceinline__ void fe_store(void* p, const fe& a) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
    q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// curve constants
__device__ __forceinline__ fe fe_const_d() {
    return fe{{0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu}};
}
__device__ __forceinline__ fe fe_const_2d() {
    return fe{{0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu}};
}
__device__ __forceinline__ fe fe_const_sqrtm1() {
    return fe{{0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u}};
}

}  // namespace cbp
