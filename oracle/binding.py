"""oracle/binding.py — TEST INFRASTRUCTURE.  ctypes loaders for the two CPU checkers:

  load_oracle()    -> oracle/liboracle.so            (ref_corrected, this repo's restatement)
  load_verbatim()  -> oracle/_ref/libref_verbatim.so (the unmodified reference host code)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use this."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))


class Fe(C.Structure):
    _fields_ = [("limbs", C.c_uint64 * 4)]


class Ge(C.Structure):
    _fields_ = [("X", Fe), ("Y", Fe), ("Z", Fe), ("T", Fe)]


class FieldVector(C.Structure):
    _fields_ = [("elements", C.POINTER(Fe)), ("length", C.c_size_t)]


class PointVector(C.Structure):
    _fields_ = [("elements", C.POINTER(Ge)), ("length", C.c_size_t)]


class InnerProductProof(C.Structure):
    _fields_ = [("n", C.c_size_t), ("a", FieldVector), ("b", FieldVector), ("c", Fe),
                ("L", PointVector), ("R", PointVector), ("L_len", C.c_size_t), ("x", Fe)]


class RangeProof(C.Structure):
    _fields_ = [("V", Ge), ("A", Ge), ("S", Ge), ("T1", Ge), ("T2", Ge),
                ("taux", Fe), ("mu", Fe), ("t", Fe), ("ip_proof", InnerProductProof)]


assert C.sizeof(Fe) == 32 and C.sizeof(Ge) == 128
assert C.sizeof(InnerProductProof) == 144 and C.sizeof(RangeProof) == 880


def build(force=False):
    """make -C oracle (liboracle.so always; _ref/libref_verbatim.so when /root/reference exists)."""
    lib = os.path.join(HERE, "liboracle.so")
    if force or not os.path.exists(lib) or os.path.getmtime(lib) < os.path.getmtime(os.path.join(HERE, "ref_corrected.c")):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, capture_output=True)
    if os.path.isdir(os.environ.get("REF_DIR", "/root/reference")) and (force or not os.path.exists(os.path.join(HERE, "_ref", "libref_verbatim.so"))):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


def _proto(lib, prefix=""):
    vp, sz, u8p = C.c_void_p, C.c_size_t, C.c_char_p
    sigs = {
        "fe25519_add": (None, [vp, vp, vp]), "fe25519_sub": (None, [vp, vp, vp]),
        "fe25519_mul": (None, [vp, vp, vp]), "fe25519_sq": (None, [vp, vp]),
        "fe25519_invert": (None, [vp, vp]), "fe25519_neg": (None, [vp, vp]),
        "fe25519_tobytes": (None, [vp, vp]), "fe25519_frombytes": (None, [vp, vp]),
        "fe25519_pow2523": (None, [vp, vp]),
        "ge25519_add": (None, [vp, vp, vp]), "ge25519_double": (None, [vp, vp]),
        "ge25519_scalarmult": (None, [vp, vp, vp]), "ge25519_scalarmult_base": (None, [vp, vp]),
        "ge25519_pack": (None, [vp, vp]), "ge25519_unpack": (C.c_int, [vp, vp]),
        "ge25519_normalize": (None, [vp]), "ge25519_is_on_curve": (C.c_int, [vp]),
        "ge25519_is_identity": (C.c_int, [vp]), "ge25519_0": (None, [vp]),
        "generate_challenge": (None, [vp, vp, sz, u8p]),
        "generate_challenge_y": (None, [vp, vp, vp, vp]), "generate_challenge_z": (None, [vp, vp]),
        "generate_challenge_x": (None, [vp, vp, vp]),
        "field_vector_inner_product": (None, [vp, vp, vp]),
        "point_vector_multi_scalar_mul": (None, [vp, vp, vp]),
        "inner_product_prove": (None, [vp, vp, vp, vp, vp, vp, vp, vp]),
        "inner_product_verify": (C.c_bool, [vp, vp, vp, vp, vp]),
        "inner_product_proof_free": (None, [vp]),
        "range_proof_verify": (C.c_bool, [vp, vp, sz, vp, vp, vp, vp]),
        "generate_range_proof": (None, [vp, vp, vp, sz, vp, vp, vp, vp]),
        "range_proof_free": (None, [vp]),
        "pedersen_commit": (None, [vp, vp, vp, vp, vp]),
        "compute_precise_delta": (None, [vp, vp, vp, sz]),
        "validate_range_input": (C.c_bool, [vp, sz]),
        "calculate_inner_product_point": (None, [vp] * 10 + [sz]),
    }
    for name, (res, args) in sigs.items():
        try:
            f = getattr(lib, prefix + name)
        except AttributeError:
            continue
        f.restype, f.argtypes = res, args
        if prefix:
            setattr(lib, name, f)  # expose the reference's C++ functions under their own names
    return lib


_ORACLE = None
_VERBATIM = None


def load_oracle():
    global _ORACLE
    if _ORACLE is None:
        build()
        lib = _proto(C.CDLL(os.path.join(HERE, "liboracle.so")))
        vp, sz = C.c_void_p, C.c_size_t
        for name, res, args in [
            ("fe25519_batch_invert", None, [vp, vp, sz]),
            ("sc25519_reduce", None, [vp, vp]), ("sc25519_add", None, [vp, vp, vp]),
            ("sc25519_sub", None, [vp, vp, vp]), ("sc25519_mul", None, [vp, vp, vp]),
            ("sc25519_invert", None, [vp, vp]), ("sc25519_neg", None, [vp, vp]),
            ("sc25519_frombytes", None, [vp, vp]), ("sc25519_reduce512", None, [vp, vp]),
            ("ge25519_equal", C.c_int, [vp, vp]), ("ge25519_neg", None, [vp, vp]),
            ("oracle_sha256", None, [vp, vp, sz]),
            ("oracle_hash_to_point", None, [vp, vp, C.c_uint32]), ("oracle_basepoint", None, [vp]),
            ("oracle_seed_rng", None, [C.c_uint64]),
            ("ipa_fold_scalars", None, [vp, vp, vp, vp, sz, vp, vp]),
            ("ipa_fold_points", None, [vp, vp, vp, vp, sz, vp, vp]),
            ("inner_product_verify_transcript", C.c_bool, [vp, vp, vp, vp, vp, vp]),
        ]:
            f = getattr(lib, name)
            f.restype, f.argtypes = res, args
        _ORACLE = lib
    return _ORACLE


def load_verbatim():
    """The unmodified reference (oracle/_ref).  Returns None if it was never built."""
    global _VERBATIM
    if _VERBATIM is None:
        path = os.path.join(HERE, "_ref", "libref_verbatim.so")
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            return None
        lib = _proto(C.CDLL(path), prefix="refv_")
        lib.refv_seed_rng.argtypes = [C.c_uint64]
        lib.refv_sizeof.restype = C.c_size_t
        lib.refv_sizeof.argtypes = [C.c_int]
        _VERBATIM = lib
    return _VERBATIM


# ---------- numpy helpers: fe25519 arrays are (n,4) uint64, ge25519 arrays are (n,16) uint64 ----------
def int_to_fe(v):
    return np.frombuffer(int(v).to_bytes(32, "little"), dtype=np.uint64).copy()


def fe_to_int(a):
    return int.from_bytes(np.ascontiguousarray(a, dtype=np.uint64).tobytes(), "little")


def ints_to_fe(vals):
    out = np.empty((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = int_to_fe(v)
    return out


def affine_to_ge(x, y, p=2**255 - 19):
    return np.concatenate([int_to_fe(x), int_to_fe(y), int_to_fe(1), int_to_fe(x * y % p)])


def ge_to_affine(g, p=2**255 - 19):
    g = np.ascontiguousarray(g, dtype=np.uint64).reshape(4, 4)
    X, Y, Z = fe_to_int(g[0]) % p, fe_to_int(g[1]) % p, fe_to_int(g[2]) % p
    zi = pow(Z, p - 2, p)
    return (X * zi % p, Y * zi % p)


def ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def field_vector(a):
    return FieldVector(C.cast(a.ctypes.data, C.POINTER(Fe)), a.shape[0])


def point_vector(a):
    return PointVector(C.cast(a.ctypes.data, C.POINTER(Ge)), a.shape[0])
