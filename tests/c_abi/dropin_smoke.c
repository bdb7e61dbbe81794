/* A plain C caller of the drop-in boundary, the way a reference maintainer's code would use it: only
 * include/cuda_bulletproof.h (the reference's own structs), no Python, no torch.  Built and run by
 * tests/test_c_abi.py.  Checks on the GPU:
 *   - MSM with scalars {1, 1} and points {B, B} equals MSM with scalar {2} and point {B} (the library normalises),
 *     and with scalars {1, l-1} it is the identity;
 *   - batch field multiply 3 * 5 = 15 and batch inversion inv(2) * 2 = 1;
 *   - mismatched vector lengths leave the result untouched.
 * Prints "dropin_smoke ok" and exits 0. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../../include/cuda_bulletproof.h"

static const uint64_t BX[4] = {0xc9562d608f25d51aULL, 0x692cc7609525a7b2ULL, 0xc0a4e231fdd6dc5cULL, 0x216936d3cd6e53feULL};
static const uint64_t BY[4] = {0x6666666666666658ULL, 0x6666666666666666ULL, 0x6666666666666666ULL, 0x6666666666666666ULL};
/* x*y of the base point mod p */
static const uint64_t BT[4] = {0x6dde8ab3a5b7dda3ULL, 0x20f09f80775152f5ULL, 0x66ea4e8e64abe37dULL, 0x67875f0fd78b7665ULL};
static const uint64_t ELL_MINUS_1[4] = {0x5812631a5cf5d3ecULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL};

static void base_point(ge25519* p) {
    memset(p, 0, sizeof *p);
    memcpy(p->X.limbs, BX, 32);
    memcpy(p->Y.limbs, BY, 32);
    p->Z.limbs[0] = 1;
    memcpy(p->T.limbs, BT, 32);
}
static int fail(const char* what) {
    fprintf(stderr, "dropin_smoke FAILED: %s\n", what);
    return 1;
}

int main(void) {
    ge25519 pts[2], r1, r2, r3;
    fe25519 sc[2], one_scalar;
    base_point(&pts[0]);
    base_point(&pts[1]);
    memset(sc, 0, sizeof sc);
    sc[0].limbs[0] = 1;
    sc[1].limbs[0] = 1;
    FieldVector fv = {sc, 2};
    PointVector pv = {pts, 2};
    cuda_point_vector_multi_scalar_mul(&r1, &fv, &pv); /* B + B */
    memset(&one_scalar, 0, sizeof one_scalar);
    one_scalar.limbs[0] = 2;
    FieldVector fv1 = {&one_scalar, 1};
    PointVector pv1 = {pts, 1};
    cuda_point_vector_multi_scalar_mul(&r2, &fv1, &pv1); /* 2 B */
    if (memcmp(&r1, &r2, sizeof r1) != 0) return fail("B + B != 2B");
    if (r1.Z.limbs[0] != 1 || r1.Z.limbs[1] || r1.Z.limbs[2] || r1.Z.limbs[3]) return fail("result not normalised");
    memcpy(sc[1].limbs, ELL_MINUS_1, 32);
    cuda_point_vector_multi_scalar_mul(&r3, &fv, &pv); /* B + (l-1) B = identity */
    if (r3.X.limbs[0] | r3.X.limbs[1] | r3.X.limbs[2] | r3.X.limbs[3]) return fail("l B is not the identity (X)");
    if (r3.Y.limbs[0] != 1 || r3.Y.limbs[1] | r3.Y.limbs[2] | r3.Y.limbs[3]) return fail("l B is not the identity (Y)");

    fe25519 a[2], b[2], out[2];
    memset(a, 0, sizeof a);
    memset(b, 0, sizeof b);
    a[0].limbs[0] = 3;
    b[0].limbs[0] = 5;
    a[1].limbs[0] = 2;
    cuda_batch_field_mul(out, a, b, 1);
    if (out[0].limbs[0] != 15 || out[0].limbs[1] | out[0].limbs[2] | out[0].limbs[3]) return fail("3 * 5 != 15");
    cuda_batch_field_invert(out, a + 1, 1);
    cuda_batch_field_mul(out + 1, out, a + 1, 1);
    if (out[1].limbs[0] != 1 || out[1].limbs[1] | out[1].limbs[2] | out[1].limbs[3]) return fail("inv(2) * 2 != 1");

    ge25519 untouched;
    memset(&untouched, 0xAB, sizeof untouched);
    FieldVector fv_bad = {sc, 1};
    cuda_point_vector_multi_scalar_mul(&untouched, &fv_bad, &pv); /* prints the reference's error line */
    for (size_t i = 0; i < sizeof untouched; i++)
        if (((unsigned char*)&untouched)[i] != 0xAB) return fail("length mismatch touched the result");
    printf("dropin_smoke ok\n");
    return 0;
}
