"""cudabulletproof_b200 — B200-native (sm_100a) Bulletproofs hot path behind the reference's C API.

The product is the shared library lib/libcudabulletproof_b200.so (hand-written CUDA + a C ABI,
see include/cuda_bulletproof.h and include/bpk.h).  This Python package is only the loader and a
thin ctypes mirror of that ABI for tests, bench.py and torch.distributed plumbing.  There is no CPU
fallback: if the library is missing or no CUDA device is present, calls fail loudly."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libcudabulletproof_b200.so")

BPK_OK = 0
_ERR = {1: "BPK_ERR_ARG", 2: "BPK_ERR_CUDA", 3: "BPK_ERR_WORKSPACE"}


class BpkError(RuntimeError):
    pass


class Fe(C.Structure):
    _fields_ = [("limbs", C.c_uint64 * 4)]


class Ge(C.Structure):
    _fields_ = [("X", Fe), ("Y", Fe), ("Z", Fe), ("T", Fe)]


class FieldVector(C.Structure):
    _fields_ = [("elements", C.c_void_p), ("length", C.c_size_t)]


class PointVector(C.Structure):
    _fields_ = [("elements", C.c_void_p), ("length", C.c_size_t)]


class InnerProductProof(C.Structure):
    _fields_ = [("n", C.c_size_t), ("a", FieldVector), ("b", FieldVector), ("c", Fe),
                ("L", PointVector), ("R", PointVector), ("L_len", C.c_size_t), ("x", Fe)]


class RangeProof(C.Structure):
    _fields_ = [("V", Ge), ("A", Ge), ("S", Ge), ("T1", Ge), ("T2", Ge),
                ("taux", Fe), ("mu", Fe), ("t", Fe), ("ip_proof", InnerProductProof)]


_lib = None

# name -> (restype, argtypes); every symbol include/*.h declares
_vp, _sz, _i, _u64 = C.c_void_p, C.c_size_t, C.c_int, C.c_uint64
SIGNATURES = {
    # include/bpk.h
    "bpk_version": (C.c_char_p, []),
    "bpk_last_error": (_i, []),
    "bpk_last_cuda_error": (_i, []),
    "bpk_clear_last_error": (_i, []),
    "bpk_kernel_launches": (_u64, []),
    "bpk_debug_set_option": (_i, [_i, C.c_longlong]),
    "bpk_host_release": (_i, []),
    "bpk_profile_enable": (_i, [_i]),
    "bpk_profile_reset": (_i, []),
    "bpk_profile_read": (_i, [_i, C.POINTER(C.c_float), C.POINTER(_i)]),
    "bpk_measure_int_peak": (_i, [C.c_double, C.POINTER(C.c_double), _i]),
    "bpk_msm_workspace_bytes": (_i, [_sz, _i, C.POINTER(_sz)]),
    "bpk_msm_window_bits": (_i, [_sz]),
    "bpk_msm_device": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _vp]),
    "bpk_msm_device_affine": (_i, [_vp, _vp, _sz, _vp, _vp, _sz, _i, _i, _vp]),
    "bpk_msm_host_affine": (_i, [_vp, _vp, _vp, _sz]),
    "bpk_point_sum_device": (_i, [_vp, _sz, _vp, _i, _vp]),
    "bpk_fe_batch_device": (_i, [_i, _vp, _vp, _vp, _sz, _vp]),
    "bpk_fe_batch_invert_workspace_bytes": (_i, [_sz, C.POINTER(_sz)]),
    "bpk_fe_batch_invert_device": (_i, [_vp, _vp, _sz, _vp, _sz, _vp]),
    "bpk_sc_inner_product_workspace_bytes": (_i, [_sz, C.POINTER(_sz)]),
    "bpk_sc_inner_product_device": (_i, [_vp, _vp, _vp, _sz, _vp, _sz, _vp]),
    "bpk_sc_inner_product_batch_device": (_i, [_vp, _vp, _vp, _sz, _sz, _vp]),
    "bpk_ipa_fold_scalars_device": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "bpk_ipa_fold_points_device": (_i, [_vp, _vp, _vp, _vp, _sz, _vp, _vp, _vp]),
    "bpk_proof_record_bytes": (_sz, [_sz]),
    "bpk_gens_workspace_bytes": (_i, [_sz, C.POINTER(_sz)]),
    "bpk_gens_init_device": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bpk_point_pack_device": (_i, [_vp, _vp, _sz, _vp]),
    "bpk_point_unpack_device": (_i, [_vp, _vp, _vp, _sz, _vp]),
    "bpk_gens_derive_device": (_i, [_vp, C.c_char_p, C.c_uint32, _sz, _vp]),
    "bpk_debug_ge_op_device": (_i, [_i, _vp, _vp, _vp, _sz, _vp]),
    "bpk_debug_const_operands_device": (_i, [_vp, _vp]),
    "bpk_debug_fe8_op_device": (_i, [_i, _vp, _vp, _vp, _sz, _vp]),
    "bpk_debug_projectivize_device": (_i, [_vp, _sz, _u64, _vp, C.c_uint32, _vp]),
    "bpk_ipa_prove_workspace_bytes": (_i, [_sz, C.POINTER(_sz)]),
    "bpk_ipa_prove_device": (_i, [_vp, _vp, _vp, _vp, _vp, _sz, C.c_char_p, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "bpk_gens_window_bits": (_i, [_vp]),
    "bpk_gens_workspace_bytes_ex": (_i, [_sz, _i, C.POINTER(_sz)]),
    "bpk_gens_init_device_ex": (_i, [_vp, _sz, _vp, _vp, _vp, _vp, _sz, _i, _vp]),
    "bpk_range_verify_workspace_bytes": (_i, [_sz, _sz, C.POINTER(_sz)]),
    "bpk_range_verify_batch_device": (_i, [_vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp]),
    "bpk_range_prove_workspace_bytes": (_i, [_sz, _sz, C.POINTER(_sz)]),
    "bpk_range_prove_batch_device": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp]),
    "bpk_range_prove_batch_keyed_device": (_i, [_vp, _vp, _vp, _vp, _sz, _sz, _vp, _vp, _sz, _vp]),
    "bpk_synth_points_device": (_i, [_vp, _vp, _sz, _u64, _vp]),
    "bpk_synth_scalars_device": (_i, [_vp, _sz, _u64, _i, _vp]),
    # include/cuda_bulletproof.h
    "cuda_point_vector_multi_scalar_mul": (None, [_vp, _vp, _vp]),
    "cuda_point_vector_multi_scalar_mul_shared": (None, [_vp, _vp, _vp]),
    "cuda_field_vector_inner_product": (None, [_vp, _vp, _vp]),
    "cuda_field_vector_inner_product_shared": (None, [_vp, _vp, _vp]),
    "cuda_batch_field_vector_inner_product": (None, [_vp, _vp, _vp, _sz]),
    "cuda_batch_field_add": (None, [_vp, _vp, _vp, _sz]),
    "cuda_batch_field_sub": (None, [_vp, _vp, _vp, _sz]),
    "cuda_batch_field_mul": (None, [_vp, _vp, _vp, _sz]),
    "cuda_batch_field_mul_karatsuba": (None, [_vp, _vp, _vp, _sz]),
    "cuda_batch_field_square": (None, [_vp, _vp, _sz]),
    "cuda_batch_field_invert": (None, [_vp, _vp, _sz]),
    "cuda_soa_field_add": (None, [_vp, _vp, _vp, _sz]),
    "cuda_range_proof_verify": (C.c_bool, [_vp, _vp, _sz, _vp, _vp, _vp, _vp]),
    "cuda_inner_product_verify": (C.c_bool, [_vp, _vp, _vp, _vp, _vp]),
    "cuda_benchmark_multi_scalar_mul": (None, [_i, _sz]),
    "cuda_benchmark_inner_product": (None, [_i, _sz]),
    "cuda_benchmark_field_operations": (None, [_i, _sz]),
    "cuda_benchmark_range_proof": (None, [_i, _sz]),
}


def load():
    """dlopen the CUDA library (no build here: __graft_entry__.build() / build.py does that)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BpkError(f"{LIB_PATH} is missing: run `python cudabulletproof_b200/build.py` "
                           "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        missing = []
        for name, (res, args) in SIGNATURES.items():
            try:
                f = getattr(lib, name)
            except AttributeError:
                missing.append(name)
                continue
            f.restype, f.argtypes = res, args
        lib._missing = missing  # tests/test_abi_exports.py requires this to be empty
        _lib = lib
    return _lib


def check(rc, what=""):
    if rc != BPK_OK:
        lib = load()
        raise BpkError(f"{what} failed: {_ERR.get(rc, rc)} (cuda error {lib.bpk_last_cuda_error()})")


from .host import *  # noqa: E402,F401
