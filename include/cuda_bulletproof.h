/*
 * cuda_bulletproof.h — the drop-in boundary.
 *
 * Same entry points, argument meaning and struct layouts as the reference's cuda_bulletproof.h
 * (each prototype cites the reference line it replaces), so reference callers
 * (bulletproof_range_proof.cu:724,728; complete_bulletproof_test.cu:153,247,280,286,292) link
 * against libcudabulletproof_b200.so unchanged.  Host pointers in, host pointers out, synchronous,
 * caller owns every buffer — exactly the reference's contract.  Differences, all deliberate
 * (DESIGN.md §3): results are correct GF(2^255-19) / Z_l / Edwards-group values (defects D1-D24);
 * CUDA failures do not exit() the process (they set bpk_last_error() and leave outputs untouched);
 * nothing is printed on the success path.
 *
 * The device-resident API the benchmarks use is in bpk.h.
 */
#ifndef CBP_CUDA_BULLETPROOF_H
#define CBP_CUDA_BULLETPROOF_H

#include "cbp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* cuda_bulletproof.h:13 — result = sum_i scalars[i] * points[i]; length mismatch: message on stderr,
 * *result untouched (cuda_bulletproof_kernels.cu:65-68).  Returned normalised (Z = 1) like the CPU MSM. */
void cuda_point_vector_multi_scalar_mul(ge25519* result, const FieldVector* scalars, const PointVector* points);
/* cuda_bulletproof.h:17 — the reference's n <= 64 shared-memory variant; same result */
void cuda_point_vector_multi_scalar_mul_shared(ge25519* result, const FieldVector* scalars, const PointVector* points);

/* cuda_bulletproof.h:22,26 — result = <a, b> mod l (D11); length mismatch: stderr + untouched
 * (cuda_inner_product.cu:100-103) */
void cuda_field_vector_inner_product(fe25519* result, const FieldVector* a, const FieldVector* b);
void cuda_field_vector_inner_product_shared(fe25519* result, const FieldVector* a, const FieldVector* b);
/* defined extern "C" in cuda_inner_product.cu:302 but absent from the reference header */
void cuda_batch_field_vector_inner_product(fe25519* results, const FieldVector* a_vectors,
                                           const FieldVector* b_vectors, size_t num_vectors);

/* cuda_bulletproof.h:31-50 — elementwise GF(2^255-19); outputs canonical (< p) */
void cuda_batch_field_add(fe25519* results, const fe25519* a, const fe25519* b, size_t count);
void cuda_batch_field_sub(fe25519* results, const fe25519* a, const fe25519* b, size_t count);
void cuda_batch_field_mul(fe25519* results, const fe25519* a, const fe25519* b, size_t count);
void cuda_batch_field_mul_karatsuba(fe25519* results, const fe25519* a, const fe25519* b, size_t count); /* cuda_field_ops.cu:338 */
void cuda_batch_field_square(fe25519* results, const fe25519* inputs, size_t count);
void cuda_batch_field_invert(fe25519* results, const fe25519* inputs, size_t count); /* inv(0) = 0 */
/* cuda_bulletproof.h:55 — the reference's "SoA" add (which drops carries, K13); here a true field add */
void cuda_soa_field_add(fe25519* results, const fe25519* a, const fe25519* b, size_t count);

/* cuda_bulletproof.h:61 — exact verification of one range proof (Q = h as the reference passes it) */
bool cuda_range_proof_verify(const RangeProof* proof, const ge25519* V, size_t n, const PointVector* G,
                             const PointVector* H, const ge25519* g, const ge25519* h);
/* cuda_bulletproof.h:72 — P must contain c*Q; transcript starts at 32 zero bytes (bulletproof_vectors.cu:589) */
bool cuda_inner_product_verify(const InnerProductProof* proof, const ge25519* P, const PointVector* G,
                               const PointVector* H, const ge25519* Q);

/* cuda_bulletproof.h:81-84 — declared by the reference, never defined there; defined here */
void cuda_benchmark_multi_scalar_mul(int iterations, size_t vector_size);
void cuda_benchmark_inner_product(int iterations, size_t vector_size);
void cuda_benchmark_field_operations(int iterations, size_t batch_size);
void cuda_benchmark_range_proof(int iterations, size_t bit_size);

#ifdef __cplusplus
}
#endif
#endif
