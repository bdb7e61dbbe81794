/*
 * oracle/shim/ref_shim.h — force-included (-include) when compiling the UNMODIFIED reference file
 * bulletproof_range_proof.cu as host C++.  That file calls cuda_point_vector_multi_scalar_mul
 * (bulletproof_range_proof.cu:724,728) without including the header that declares it
 * (SURVEY.md §8b), so the shipped file does not compile on its own.  This supplies the one missing
 * prototype; nothing else about the reference is altered.
 */
#ifndef REF_SHIM_H
#define REF_SHIM_H
#include "curve25519_ops.h"
#include "bulletproof_vectors.h"
extern "C" void cuda_point_vector_multi_scalar_mul(ge25519* result, const FieldVector* scalars,
                                                   const PointVector* points);
#endif
