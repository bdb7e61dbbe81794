/*
 * cbp_types.h — the type contract of the drop-in boundary.
 *
 * These are re-declarations (not copies) of the reference's container types so that a
 * caller compiled against the reference headers can link against this library unchanged:
 *
 *   fe25519            curve25519_ops.h:15-17   4 x u64 little-endian limbs, 32 B
 *   ge25519            curve25519_ops.h:20-25   extended twisted-Edwards (X:Y:Z:T), 128 B
 *   ge25519_compressed curve25519_ops.h:28-30   32 B: y little-endian, bit 255 = lsb(x)
 *   FieldVector        bulletproof_vectors.h:8-11
 *   PointVector        bulletproof_vectors.h:14-17
 *   InnerProductProof  bulletproof_vectors.h:65-74
 *   RangeProof         bulletproof_range_proof.h:9-19
 *
 * The layouts are pinned by static asserts below and, in tests/, against the reference's
 * own headers compiled unmodified (oracle/_ref).
 */
#ifndef CBP_TYPES_H
#define CBP_TYPES_H

#include <stdint.h>
#include <stddef.h>
#ifndef __cplusplus
#include <stdbool.h>
#endif

/* If the reference's own headers were included first, reuse its typedefs. */
#ifndef CURVE25519_OPS_H
typedef struct {
    uint64_t limbs[4];
} fe25519;

typedef struct {
    fe25519 X;
    fe25519 Y;
    fe25519 Z;
    fe25519 T;
} ge25519;

typedef struct {
    uint8_t bytes[32];
} ge25519_compressed;
#endif

#ifndef BULLETPROOF_VECTORS_H
typedef struct {
    fe25519* elements;
    size_t length;
} FieldVector;

typedef struct {
    ge25519* elements;
    size_t length;
} PointVector;

typedef struct {
    size_t n;
    FieldVector a;
    FieldVector b;
    fe25519 c;
    PointVector L;
    PointVector R;
    size_t L_len;
    fe25519 x;
} InnerProductProof;
#endif

#ifndef BULLETPROOF_RANGE_PROOF_H
typedef struct {
    ge25519 V;
    ge25519 A;
    ge25519 S;
    ge25519 T1;
    ge25519 T2;
    fe25519 taux;
    fe25519 mu;
    fe25519 t;
    InnerProductProof ip_proof;
} RangeProof;
#endif

#ifdef __cplusplus
#define CBP_STATIC_ASSERT(c, m) static_assert(c, m)
#else
#define CBP_STATIC_ASSERT(c, m) _Static_assert(c, m)
#endif

CBP_STATIC_ASSERT(sizeof(fe25519) == 32, "fe25519 must be 32 bytes");
CBP_STATIC_ASSERT(sizeof(ge25519) == 128, "ge25519 must be 128 bytes");
CBP_STATIC_ASSERT(offsetof(ge25519, Y) == 32 && offsetof(ge25519, Z) == 64 && offsetof(ge25519, T) == 96,
                  "ge25519 field offsets");
CBP_STATIC_ASSERT(sizeof(FieldVector) == 16 && offsetof(FieldVector, length) == 8, "FieldVector layout");
CBP_STATIC_ASSERT(sizeof(PointVector) == 16 && offsetof(PointVector, length) == 8, "PointVector layout");
CBP_STATIC_ASSERT(sizeof(InnerProductProof) == 144, "InnerProductProof must be 144 bytes");
CBP_STATIC_ASSERT(offsetof(InnerProductProof, a) == 8 && offsetof(InnerProductProof, b) == 24 &&
                  offsetof(InnerProductProof, c) == 40 && offsetof(InnerProductProof, L) == 72 &&
                  offsetof(InnerProductProof, R) == 88 && offsetof(InnerProductProof, L_len) == 104 &&
                  offsetof(InnerProductProof, x) == 112, "InnerProductProof offsets");
CBP_STATIC_ASSERT(sizeof(RangeProof) == 880, "RangeProof must be 880 bytes");
CBP_STATIC_ASSERT(offsetof(RangeProof, A) == 128 && offsetof(RangeProof, S) == 256 &&
                  offsetof(RangeProof, T1) == 384 && offsetof(RangeProof, T2) == 512 &&
                  offsetof(RangeProof, taux) == 640 && offsetof(RangeProof, mu) == 672 &&
                  offsetof(RangeProof, t) == 704 && offsetof(RangeProof, ip_proof) == 736, "RangeProof offsets");

#endif /* CBP_TYPES_H */
