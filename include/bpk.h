/*
 * bpk.h — device-resident C ABI of libcudabulletproof_b200.so ("bpk" = bulletproof kernels).
 *
 * The reference's entry points (cuda_bulletproof.h) take host arrays and malloc/copy/sync per call
 * (cuda_bulletproof_kernels.cu:77-115), which is PCIe-bound at 2^20 points.  These functions are
 * the same operations on buffers already resident in HBM: plain device pointers and sizes, an
 * explicit cudaStream_t (passed as void*), caller-provided workspace, int status, no allocation,
 * no synchronisation, no printing.  The host-pointer drop-ins are thin wrappers over these.
 *
 * Layouts: fe25519 / scalars = 32 B little-endian; ge25519 = 128 B (X,Y,Z,T), arrays are AoS
 * exactly as the reference's FieldVector / PointVector elements.
 */
#ifndef CBP_BPK_H
#define CBP_BPK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BPK_OK 0
#define BPK_ERR_ARG 1
#define BPK_ERR_CUDA 2
#define BPK_ERR_WORKSPACE 3

const char* bpk_version(void);
/* last error recorded by any entry point of this library in this process (0 = none) */
int bpk_last_error(void);
int bpk_last_cuda_error(void);
/* returns the recorded error and resets it (the host-pointer drop-ins cannot return a status) */
int bpk_clear_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
uint64_t bpk_kernel_launches(void);

/* measurement / test switches.  Their defaults come from the environment ONCE per process (CBP_MSM_SLOTS,
 * CBP_MSM_NO2D, CBP_HOST_CHUNK_LOG2, CBP_PROVER_LEGACY, CBP_GROUPS, CBP_MSM_SMALL_MAX); no entry point reads the
 * environment on its call path.  None is needed in production. */
#define BPK_OPT_MSM_SLOTS 0       /* -1 auto, 0 / 1: slotted first digit pass off / on for every size */
#define BPK_OPT_MSM_NO2D 1        /* 1: running-sum bucket reduction everywhere (cross-check of the 2-D one) */
#define BPK_OPT_HOST_CHUNK_LOG2 2 /* log2 of the upload chunk of the host-pointer MSM, 0 = default */
#define BPK_OPT_PROVER_LEGACY 3   /* 1: one-CTA-per-proof prover for every batch size */
#define BPK_OPT_MSM_GROUPS 4      /* window groups of the MSM pipeline as hex digits, top first (0x844); 0 = auto */
#define BPK_OPT_MSM_SMALL_MAX 5   /* largest n routed to the single-launch small-n MSM; -1 default, 0 disables */
#define BPK_OPT_HOST_REGISTER 6   /* 1: page-lock pageable caller buffers >= 1 MiB once and remember them (see below) */
#define BPK_OPT_IPA_COMPOSITE_MAX 7 /* longest vector whose IPA rounds run unfolded (power of two <= 4096); -1 default */
#define BPK_OPT_MSM_SEG_SHIFT 8     /* log2 of the bucket-accumulation segment length; -1 = automatic */
#define BPK_OPT_HOST_TAPER_LOG2 10  /* host-pointer MSM: the last upload chunk is halved down to 2^value points; 0 = not */
#define BPK_OPT_HOST_TRACE 11       /* 1: the chunked host-pointer MSM prints its stream timeline to stderr */
#define BPK_OPT_DEBUG_VARIANT 12    /* A/B switch for kernels under measurement; 0 = the shipped code */
#define BPK_OPT_VERIFY_GROUP 13     /* batch verification: proofs per combined identity; -1 auto, 0 or 1 = one by one */
#define BPK_OPT_MSM_FUSED_FRONT 14   /* 1 (default): scans + segment build of the MSM front end in one cooperative launch; 0: separately */
#define BPK_OPT_MSM_GRAPH 15        /* 1 (default): device MSMs of 2^13 < n < 2^19 pairs replay a cached CUDA graph; 0: plain launches */
#define BPK_OPT_MSM_ACC_STREAMS 9   /* 1 / 0: every window group accumulates on its own stream (groups overlap) / in turn; -1 auto */
int bpk_debug_set_option(int option, long long value);
/* With BPK_OPT_HOST_REGISTER on, the host-pointer MSM page-locks large pageable input buffers in place the first time
 * it sees them, so that repeated calls on the same buffers upload at the pinned PCIe rate.  Such buffers must stay
 * allocated until this call, which unregisters all of them. */
int bpk_host_release(void);

/* ---- per-kernel device timing (bench.py's roofline): CUDA events recorded on the launching stream
 * around the named kernel of every call while enabled; read returns the mean duration in ms ---- */
#define BPK_PROF_MSM_ACCUMULATE 0 /* msm_accumulate_kernel  */
#define BPK_PROF_MSM_TOTAL 1      /* whole bpk_msm_device   */
#define BPK_PROF_VERIFY_MSM 2     /* verify_fixed_kernel    */
#define BPK_PROF_VERIFY_TOTAL 3   /* whole bpk_range_verify_batch_device */
#define BPK_PROF_MSM_PRECOMPUTE 4 /* msm_precompute_kernel (HBM streaming) */
#define BPK_PROF_MSM_FRONT 5      /* recode, histogram, scans, scatter, segment sort (main stream) */
#define BPK_PROF_MSM_TAIL 6       /* end of the last accumulation -> result ready */
#define BPK_PROF_KINDS 7
int bpk_profile_enable(int enable);
int bpk_profile_reset(void);
int bpk_profile_read(int kind, float* mean_ms, int* samples); /* synchronises the recorded events */

/* ---- the integer roofline denominator, measured in the caller's own run (csrc/intpeak.cu) ----
 * Runs issue-rate microbenchmarks for about target_ms each on the current device and returns lane operations per
 * second in rates[0..count): [0] IMAD.WIDE.U32 carry chains (fe_mul's inner pattern; the roofline denominator),
 * [1] IMAD.WIDE.U32 independent, [2] IMAD 32-bit, [3] IMAD.HI.U32, [4] DFMA (FP64).  Synchronises; allocates and
 * frees a small buffer: not for a hot path. */
#define BPK_PEAK_IMAD_WIDE_CARRY 0
#define BPK_PEAK_IMAD_WIDE 1
#define BPK_PEAK_IMAD_LO 2
#define BPK_PEAK_IMAD_HI 3
#define BPK_PEAK_DFMA 4
#define BPK_PEAK_IMAD_WIDE_BESIDE_DFMA 5 /* rate [0] while the same number of DFMA issue from the same threads */
#define BPK_PEAK_KINDS 6
int bpk_measure_int_peak(double target_ms, double* rates, int count);

/* ---- multi-scalar multiplication: replaces cuda_point_vector_multi_scalar_mul (cuda_bulletproof.h:13) ---- */
/* window_bits = 0 picks c(n).  *bytes = workspace needed by bpk_msm_device for that (n, window_bits). */
int bpk_msm_workspace_bytes(size_t n, int window_bits, size_t* bytes);
int bpk_msm_window_bits(size_t n);
/* d_result: one ge25519 (128 B); normalize != 0 returns (x, y, 1, xy) canonical like the CPU MSM */
int bpk_msm_device(const void* d_scalars, const void* d_points, size_t n, void* d_result, void* d_workspace,
                   size_t workspace_bytes, int window_bits, int normalize, void* stream);
/* Extension (no counterpart in the reference): the same sum over AFFINE points, 64 bytes each (x || y as fe25519
   containers, Z = 1 implied, not checked to be on the curve — like every point input of the reference): 96 instead of
   160 bytes per pair, which is what bounds the host-pointer call (PCIe).  Workspace as for bpk_msm_device.
   bpk_msm_host_affine takes host pointers (pinned or pageable) like cuda_point_vector_multi_scalar_mul. */
int bpk_msm_device_affine(const void* d_scalars, const void* d_xy, size_t n, void* d_result, void* d_workspace,
                          size_t workspace_bytes, int window_bits, int normalize, void* stream);
int bpk_msm_host_affine(void* result /* ge25519 */, const void* scalars /* n fe25519 */, const void* xy, size_t n);
/* sum of `count` extended points (multi-GPU partial results), normalised: d_result = sum d_points[i] */
int bpk_point_sum_device(const void* d_points, size_t count, void* d_result, int normalize, void* stream);

/* ---- batched fe25519: replaces cuda_batch_field_{add,sub,mul,square,invert} (cuda_bulletproof.h:31-50) ---- */
#define BPK_FE_ADD 0
#define BPK_FE_SUB 1
#define BPK_FE_MUL 2
#define BPK_FE_SQR 3
int bpk_fe_batch_device(int op, void* d_out, const void* d_a, const void* d_b, size_t count, void* stream);
int bpk_fe_batch_invert_workspace_bytes(size_t count, size_t* bytes);
int bpk_fe_batch_invert_device(void* d_out, const void* d_in, size_t count, void* d_workspace,
                               size_t workspace_bytes, void* stream);

/* ---- scalars mod l: replaces cuda_field_vector_inner_product (cuda_bulletproof.h:22) ---- */
int bpk_sc_inner_product_workspace_bytes(size_t n, size_t* bytes);
int bpk_sc_inner_product_device(void* d_out, const void* d_a, const void* d_b, size_t n, void* d_workspace,
                                size_t workspace_bytes, void* stream);
/* num_vectors independent inner products of length n each (contiguous): cuda_inner_product.cu:302 */
int bpk_sc_inner_product_batch_device(void* d_out, const void* d_a, const void* d_b, size_t n, size_t num_vectors,
                                      void* stream);

/* ---- IPA folding (bulletproof_vectors.cu:488-500 and :641-663) ---- */
/* a' = u a_L + u^-1 a_R ; b' = u^-1 b_L + u b_R  (mod l);  in/out may alias (in place on the low half) */
int bpk_ipa_fold_scalars_device(void* d_a_out, void* d_b_out, const void* d_a, const void* d_b, size_t n_half,
                                const void* d_u, const void* d_u_inv, void* stream);
/* G'_j = u^-1 G_j + u G_{j+n'} ; H'_j = u H_j + u^-1 H_{j+n'} ; outputs normalised */
int bpk_ipa_fold_points_device(void* d_G_out, void* d_H_out, const void* d_G, const void* d_H, size_t n_half,
                               const void* d_u, const void* d_u_inv, void* stream);

/* Inner-product argument, prover side, for any power-of-two n (inner_product_prove,
 * bulletproof_vectors.cu:375-509): d_L / d_R receive log2 n normalised points each, d_a_out / d_b_out the final
 * scalars (32 B), d_x_out the raw first-round challenge the reference stores in the proof (32 B).
 * transcript0: 32 bytes (HOST pointer).  Bit-identical to the CPU oracle for generators of prime order (every
 * Bulletproofs generator set; bpk_gens_derive_device clears the cofactor).  Rounds over at most 4096 generators never
 * fold a point: L and R are multi-scalar multiplications over the base generators with composite scalars
 * (csrc/ipa_straus.cu, 6 launches per round); longer vectors are first folded as the reference does
 * (bulletproof_vectors.cu:641-663) with L and R as Pippenger MSMs. */
int bpk_ipa_prove_workspace_bytes(size_t n, size_t* bytes);
int bpk_ipa_prove_device(const void* d_G, const void* d_H, const void* d_Q, const void* d_a, const void* d_b, size_t n,
                         const uint8_t transcript0[32], void* d_L, void* d_R, void* d_a_out, void* d_b_out, void* d_x_out,
                         void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- point codec and generator derivation (N2, N4) ---- */
/* d_out[i] (32 B) = RFC 8032 encoding of d_points[i] (ge25519, any Z): y little-endian, bit 255 = lsb(x).
 * Replaces ge25519_pack (curve25519_ops.cu:449-468) on arrays; one field inversion per 4096 points. */
int bpk_point_pack_device(void* d_out, const void* d_points, size_t count, void* stream);
/* d_points[i] = decoded point (Z = 1, canonical limbs), d_ok[i] = 1; or the identity and d_ok[i] = 0 when the
 * encoding is invalid (y >= p, no square root, x = 0 with sign bit).  d_ok may be NULL.
 * Replaces ge25519_unpack (curve25519_ops.cu:470-531) with the checks it omits (defects D7, D8). */
int bpk_point_unpack_device(void* d_points, uint8_t* d_ok, const void* d_in, size_t count, void* stream);
/* d_points[i] = generator number first_index + i of the family `seed` (32 bytes, HOST pointer):
 * SHA-256(seed || index_be32 [|| counter_be32]) decoded as a point, counter bumped until it decodes, times 8,
 * never the identity, normalised.  Labels as complete_bulletproof_test.cu:33-41,79-88 (seed byte 1..4). */
int bpk_gens_derive_device(void* d_points, const uint8_t seed[32], uint32_t first_index, size_t count, void* stream);

/* test hook: out[i] = 2 a[i] (op 0), a[i] + b[i] (op 1; b NULL: a[i] + a[i]), 8 a[i] (op 2); extended points,
 * outputs not normalised.  Lets the parity tests pin the group law itself against the CPU oracle. */
int bpk_debug_ge_op_device(int op, const void* d_a, const void* d_b, void* d_out, size_t count, void* stream);
/* test hook: d_points[i] <- the same group element with a pseudo-random Z != 1 (all four coordinates scaled),
 * after adding *d_torsion (one ge25519, may be NULL) to every torsion_stride-th point.  Lets full-size MSM parity
 * tests cover projective and torsion-carrying inputs. */
int bpk_debug_projectivize_device(void* d_points, size_t n, uint64_t seed, const void* d_torsion,
                                  uint32_t torsion_stride, void* stream);
/* test hook for the octet-form arithmetic (csrc/fe8.cuh: one 32-bit word per lane, 8 lanes per field element, used
 * by the latency-bound chains).  Field ops: d_out[i] (32 B, canonical) = a[i] op b[i].  Point ops: d_out[i] (ge25519,
 * normalised) = 2 a[i] | a[i] + b[i] | a[i] + b[i] through the cached form | 2^64 a[i] | a[i]. */
#define BPK_FE8_MUL 0
#define BPK_FE8_ADD 1
#define BPK_FE8_SUB 2
#define BPK_GE8_DBL 3
#define BPK_GE8_ADD 4
#define BPK_GE8_ADD_CACHED 5
#define BPK_GE8_DBL_CHAIN 6
#define BPK_GE8_NORMALIZE 7
int bpk_debug_fe8_op_device(int op, const void* d_a, const void* d_b, void* d_out, size_t count, void* stream);
/* test hook: field operations on compile-time constants; writes 6 field elements (8 words each):
 * 1^2, 1*1, 1+1, 2^2, 1-2 (canonical), 2*(2d) (canonical).  Guards the inline-asm operand constraints. */
int bpk_debug_const_operands_device(uint32_t* d_out48, void* stream);

/* ---- range proofs: replaces cuda_range_proof_verify (cuda_bulletproof.h:61) on batches ---- */
/* Flat proof record, uint64 words (n-bit proof, k = log2 n):
 *   V,A,S,T1,T2 (5 x 128 B) | taux,mu,t (3 x 32 B) | a,b,c,x (4 x 32 B) | L[0..k) (k x 128 B) | R[0..k) (k x 128 B) */
size_t bpk_proof_record_bytes(size_t n);
/* generator set shared by every proof: G[n], H[n], g, h (ge25519, AoS).  Builds device tables of the
 * signed multiples d * 2^(w j) * Base for every window j, so that a fixed-base scalar multiplication is
 * 256/w table additions and no doublings.  window_bits w = 8: 32 windows x 128 multiples (51 MB at n = 64,
 * L2-resident; cheap to build — what the plain entry points use).  w = 16: 16 windows x 32768 multiples
 * (6.5 GB at n = 64, HBM-resident): half the additions per proof, for long-lived batch verifiers/provers.
 * Proof bytes and accept decisions do not depend on w. */
int bpk_gens_workspace_bytes(size_t n, size_t* bytes);
int bpk_gens_init_device(void* d_gens_ws, size_t ws_bytes, const void* d_G, const void* d_H, const void* d_g,
                         const void* d_h, size_t n, void* stream);
/* window width (8 or 16) of a table built in this process, 0 if d_gens_ws is unknown */
int bpk_gens_window_bits(const void* d_gens_ws);
int bpk_gens_workspace_bytes_ex(size_t n, int window_bits, size_t* bytes);
int bpk_gens_init_device_ex(void* d_gens_ws, size_t ws_bytes, const void* d_G, const void* d_H, const void* d_g,
                            const void* d_h, size_t n, int window_bits, void* stream);
/* d_accept[i] = 1 iff proof i verifies (exact checks, bit-exact with the CPU oracle's range_proof_verify).
 * d_V (optional, num_proofs ge25519): the caller's commitments, each must equal its proof's V
 * (bulletproof_range_proof.cu:1729-1740); NULL skips that check. */
int bpk_range_verify_workspace_bytes(size_t n, size_t num_proofs, size_t* bytes);
int bpk_range_verify_batch_device(const void* d_gens_ws, const void* d_proofs, const void* d_V, size_t n,
                                  size_t num_proofs, uint8_t* d_accept, void* d_workspace, size_t workspace_bytes,
                                  void* stream);
/* Batch prover (generate_range_proof, bulletproof_range_proof.cu:1159-1714, restated).  Two sources for the
 * blinding values and nonces (alpha, rho, tau1, tau2, sL, sR); same protocol, same record layout:
 *
 * bpk_range_prove_batch_keyed_device — THE PROVER.  d_keys: 32 secret bytes per proof from the caller's CSPRNG
 *   (the reference draws every value from OpenSSL RAND_bytes, bulletproof_range_proof.cu:153); value j of a proof
 *   is SHA-256("cbp-bp-nonce" || key || j_le32) shaped like generate_random_scalar and reduced mod l.  Never
 *   reuse a key for two different (value, gamma) pairs.
 *
 * bpk_range_prove_batch_device — FOR TESTS AND BENCHMARKS ONLY.  Values come from a SplitMix64 stream seeded by
 *   the 64-bit d_seeds[i]: the stream oracle/ref_corrected.c draws, so proofs are bit-identical to the oracle's.
 *   SplitMix64 is invertible and not cryptographic: a recovered seed yields tau1, tau2, hence gamma and v.  Do not
 *   use it for proofs that must hide anything.
 *
 * The workspace is optional: with bpk_range_prove_workspace_bytes() bytes (28 KB per proof, capped at 2^14
 * proofs) batches of 64+ proofs run as a phase-split pipeline with batch inversions across proofs; with NULL
 * every proof is one CTA.  Same bytes either way. */
int bpk_range_prove_workspace_bytes(size_t n, size_t num_proofs, size_t* bytes);
int bpk_range_prove_batch_keyed_device(const void* d_gens_ws, const uint64_t* d_values,
                                       const void* d_gammas /* 32 B each */, const void* d_keys /* 32 B each */,
                                       size_t n, size_t num_proofs, void* d_proofs, void* d_workspace,
                                       size_t workspace_bytes, void* stream);
int bpk_range_prove_batch_device(const void* d_gens_ws, const uint64_t* d_values, const void* d_gammas /* 32 B each */,
                                 const uint64_t* d_seeds, size_t n, size_t num_proofs, void* d_proofs,
                                 void* d_workspace, size_t workspace_bytes, void* stream);

/* ---- synthetic inputs (bench / tests): P_i = k_i * B for a hash-derived 64-bit k_i, returned
 * normalised, plus k_i itself so that callers can check MSMs against a scalar identity ---- */
int bpk_synth_points_device(void* d_points, uint64_t* d_k, size_t n, uint64_t seed, void* stream);
/* uniform `bits`-bit integers (bits <= 256; 252 gives reduced scalars < l), 32 B little-endian each */
int bpk_synth_scalars_device(void* d_scalars, size_t n, uint64_t seed, int bits, void* stream);

#ifdef __cplusplus
}
#endif
#endif
