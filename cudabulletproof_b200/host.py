"""Host-side mirror of the reference operator interface over the C ABI.

Two layers, same names as the reference where one exists:
  * numpy-in / numpy-out wrappers of the host-pointer drop-ins (cuda_bulletproof.h), used by the
    parity tests so they read like the reference's own call sites;
  * torch-tensor wrappers of the device-resident API (bpk.h): torch is only the allocator and the
    stream provider here — every computation is a kernel of libcudabulletproof_b200.so.
fe25519 arrays are (n, 4) uint64; ge25519 arrays are (n, 16) uint64 (X,Y,Z,T)."""
import ctypes as C

import numpy as np


def _lib():
    from . import load
    return load()


def _check(rc, what):
    from . import check
    check(rc, what)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _guard(fn, *args):
    """Call a host-pointer drop-in (which cannot return a status) and fail loudly on a CUDA error.
    Argument errors keep the reference's behaviour: message on stderr, outputs untouched, no exception."""
    lib = _lib()
    lib.bpk_clear_last_error()
    rv = fn(*args)
    cuda_err = lib.bpk_last_cuda_error()
    if lib.bpk_last_error() == 2:
        from . import BpkError
        raise BpkError(f"{fn.__name__}: CUDA error {cuda_err} (no CPU fallback)")
    return rv


def _fv(a):
    from . import FieldVector
    return FieldVector(a.ctypes.data, a.shape[0])


def _pv(a):
    from . import PointVector
    return PointVector(a.ctypes.data, a.shape[0])


def _c(a, cols):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    assert a.ndim == 2 and a.shape[1] == cols, a.shape
    return a


# ---------------- host-pointer drop-ins (reference cuda_bulletproof.h) ----------------
def cuda_point_vector_multi_scalar_mul(scalars, points, shared=False, result=None):
    """result = sum scalars[i]*points[i]; returns the (16,) uint64 ge25519 (normalised)."""
    s, p = _c(scalars, 4), _c(points, 16)
    out = np.zeros(16, dtype=np.uint64) if result is None else result
    fv, pv = _fv(s), _pv(p)
    fn = _lib().cuda_point_vector_multi_scalar_mul_shared if shared else _lib().cuda_point_vector_multi_scalar_mul
    _guard(fn, _ptr(out), C.byref(fv), C.byref(pv))
    return out


def cuda_field_vector_inner_product(a, b, shared=False, result=None):
    a, b = _c(a, 4), _c(b, 4)
    out = np.zeros(4, dtype=np.uint64) if result is None else result
    fa, fb = _fv(a), _fv(b)
    fn = _lib().cuda_field_vector_inner_product_shared if shared else _lib().cuda_field_vector_inner_product
    _guard(fn, _ptr(out), C.byref(fa), C.byref(fb))
    return out


def cuda_batch_field_vector_inner_product(a_vectors, b_vectors, results=None):
    """results[v] = <a_vectors[v], b_vectors[v]> mod l for separately allocated vectors
    (cuda_inner_product.cu:302-348; defined in the reference, absent from its header)."""
    from . import FieldVector
    a_vectors = [_c(a, 4) for a in a_vectors]
    b_vectors = [_c(b, 4) for b in b_vectors]
    m = len(a_vectors)
    out = np.zeros((m, 4), dtype=np.uint64) if results is None else results
    fa = (FieldVector * max(m, 1))(*[_fv(a) for a in a_vectors])
    fb = (FieldVector * max(m, 1))(*[_fv(b) for b in b_vectors])
    _guard(_lib().cuda_batch_field_vector_inner_product, _ptr(out), fa, fb, m)
    return out


def _batch2(name, a, b):
    a, b = _c(a, 4), _c(b, 4)
    out = np.zeros_like(a)
    _guard(getattr(_lib(), name), _ptr(out), _ptr(a), _ptr(b), a.shape[0])
    return out


def cuda_batch_field_add(a, b):
    return _batch2("cuda_batch_field_add", a, b)


def cuda_batch_field_sub(a, b):
    return _batch2("cuda_batch_field_sub", a, b)


def cuda_batch_field_mul(a, b):
    return _batch2("cuda_batch_field_mul", a, b)


def cuda_soa_field_add(a, b):
    return _batch2("cuda_soa_field_add", a, b)


def cuda_batch_field_square(a):
    a = _c(a, 4)
    out = np.zeros_like(a)
    _guard(_lib().cuda_batch_field_square, _ptr(out), _ptr(a), a.shape[0])
    return out


def cuda_batch_field_invert(a):
    a = _c(a, 4)
    out = np.zeros_like(a)
    _guard(_lib().cuda_batch_field_invert, _ptr(out), _ptr(a), a.shape[0])
    return out


# ---------------- device-resident API (bpk.h) on torch tensors ----------------
def _stream_ptr(stream=None):
    import torch
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def _dev_u8(nbytes, device):
    import torch
    return torch.empty(max(int(nbytes), 16), dtype=torch.uint8, device=device)


class Msm:
    """Pippenger MSM with a reusable workspace: result = sum_i scalars[i] * points[i].

    scalars: cuda uint8 tensor (n, 32); points: cuda uint8 tensor (n, 128) (reference AoS ge25519)."""

    def __init__(self, n, device="cuda", window_bits=0):
        import torch
        self.n, self.device = int(n), torch.device(device)
        lib = _lib()
        self._window_arg = int(window_bits)  # 0 = automatic: lets the library pick the small-n path as well
        self.window_bits = window_bits or lib.bpk_msm_window_bits(self.n)
        nbytes = C.c_size_t(0)
        _check(lib.bpk_msm_workspace_bytes(self.n, self._window_arg, C.byref(nbytes)), "bpk_msm_workspace_bytes")
        self.workspace = _dev_u8(nbytes.value, self.device)
        self.result = torch.zeros(128, dtype=torch.uint8, device=self.device)

    def __call__(self, scalars, points, normalize=True, out=None, stream=None):
        out = self.result if out is None else out
        assert scalars.is_cuda and points.is_cuda and scalars.numel() == self.n * 32 and points.numel() == self.n * 128
        _check(_lib().bpk_msm_device(scalars.data_ptr(), points.data_ptr(), self.n, out.data_ptr(),
                                     self.workspace.data_ptr(), self.workspace.numel(), self._window_arg,
                                     1 if normalize else 0, _stream_ptr(stream)), "bpk_msm_device")
        return out

    def affine(self, scalars, xy, normalize=True, out=None, stream=None):
        """the same over affine points: xy cuda uint8 tensor (n, 64), x || y (bpk_msm_device_affine)"""
        out = self.result if out is None else out
        assert scalars.is_cuda and xy.is_cuda and scalars.numel() == self.n * 32 and xy.numel() == self.n * 64
        _check(_lib().bpk_msm_device_affine(scalars.data_ptr(), xy.data_ptr(), self.n, out.data_ptr(),
                                            self.workspace.data_ptr(), self.workspace.numel(), self._window_arg,
                                            1 if normalize else 0, _stream_ptr(stream)), "bpk_msm_device_affine")
        return out


def point_sum(points, normalize=True, stream=None):
    """Sum of extended points (count, 128) uint8 on device -> (128,) uint8."""
    import torch
    out = torch.zeros(128, dtype=torch.uint8, device=points.device)
    _check(_lib().bpk_point_sum_device(points.data_ptr(), points.numel() // 128, out.data_ptr(),
                                       1 if normalize else 0, _stream_ptr(stream)), "bpk_point_sum_device")
    return out


def synth_points(n, seed, device="cuda", stream=None):
    """P_i = k_i * B (normalised) and the 64-bit k_i: (n,128) uint8, (n,) int64 tensors."""
    import torch
    pts = torch.empty((n, 128), dtype=torch.uint8, device=device)
    ks = torch.empty(n, dtype=torch.int64, device=device)
    _check(_lib().bpk_synth_points_device(pts.data_ptr(), ks.data_ptr(), n, seed, _stream_ptr(stream)),
           "bpk_synth_points_device")
    return pts, ks


def synth_scalars(n, seed, bits=252, device="cuda", stream=None):
    import torch
    sc = torch.empty((n, 32), dtype=torch.uint8, device=device)
    _check(_lib().bpk_synth_scalars_device(sc.data_ptr(), n, seed, bits, _stream_ptr(stream)),
           "bpk_synth_scalars_device")
    return sc


# ---------------- range proofs ----------------
def proof_record_bytes(n):
    return int(_lib().bpk_proof_record_bytes(n))


class Generators:
    """Device-resident generator set (G[n], H[n], g, h) with its fixed-base tables.
    G, H: (n, 16) uint64 numpy arrays or (n,128) uint8 cuda tensors; g, h: (16,) / (128,).
    window_bits: 8 (51 MB of tables at n = 64, L2-resident) or 16 (6.5 GB, HBM-resident, half the point
    additions per fixed-base scalar multiplication); results do not depend on it."""

    def __init__(self, G, H, g, h, device="cuda", stream=None, window_bits=8):
        import torch
        self.device = torch.device(device)
        self.window_bits = int(window_bits)

        def dev(a):
            if isinstance(a, np.ndarray):
                a = torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint64).view(np.uint8).reshape(-1))
            return a.to(self.device).contiguous().view(torch.uint8).reshape(-1)

        dG, dH, dg, dh = dev(G), dev(H), dev(g), dev(h)
        self.n = dG.numel() // 128
        assert dH.numel() == self.n * 128 and dg.numel() == 128 and dh.numel() == 128
        nbytes = C.c_size_t(0)
        _check(_lib().bpk_gens_workspace_bytes_ex(self.n, self.window_bits, C.byref(nbytes)),
               "bpk_gens_workspace_bytes_ex")
        self.workspace = _dev_u8(nbytes.value, self.device)
        _check(_lib().bpk_gens_init_device_ex(self.workspace.data_ptr(), self.workspace.numel(), dG.data_ptr(),
                                              dH.data_ptr(), dg.data_ptr(), dh.data_ptr(), self.n, self.window_bits,
                                              _stream_ptr(stream)),
               "bpk_gens_init_device_ex")
        torch.cuda.current_stream().synchronize()  # inputs may be temporaries
        self.record_bytes = proof_record_bytes(self.n)


def range_prove_batch(gens, values, gammas, seeds=None, stream=None, keys=None):
    """values: (m,) int/uint64, gammas: (m,4) uint64 scalars -> (m, record_bytes) uint8 cuda tensor.
    keys: (m, 32) uint8 secrets from a CSPRNG, one per proof (bpk_range_prove_batch_keyed_device) — the prover.
    seeds: (m,) 64-bit seeds of the oracle's SplitMix64 stream (bpk_range_prove_batch_device) — parity tests and
    benchmark inputs ONLY: that stream is not cryptographic."""
    import torch
    assert (seeds is None) != (keys is None), "exactly one of seeds / keys"
    m = len(values)
    d_v = torch.from_numpy(np.asarray(values, dtype=np.uint64).view(np.int64)).to(gens.device)
    d_g = torch.from_numpy(np.ascontiguousarray(gammas, dtype=np.uint64).view(np.uint8).reshape(-1)).to(gens.device)
    out = torch.zeros((m, gens.record_bytes), dtype=torch.uint8, device=gens.device)
    nbytes = C.c_size_t(0)
    _check(_lib().bpk_range_prove_workspace_bytes(gens.n, m, C.byref(nbytes)), "bpk_range_prove_workspace_bytes")
    ws = _dev_u8(max(1, nbytes.value), gens.device)  # batches of 64+ proofs take the phase-split prover
    if keys is not None:
        d_k = torch.from_numpy(np.ascontiguousarray(keys, dtype=np.uint8).reshape(m, 32)).to(gens.device)
        _check(_lib().bpk_range_prove_batch_keyed_device(gens.workspace.data_ptr(), d_v.data_ptr(), d_g.data_ptr(),
                                                         d_k.data_ptr(), gens.n, m, out.data_ptr(), ws.data_ptr(),
                                                         nbytes.value, _stream_ptr(stream)),
               "bpk_range_prove_batch_keyed_device")
    else:
        d_s = torch.from_numpy(np.asarray(seeds, dtype=np.uint64).view(np.int64)).to(gens.device)
        _check(_lib().bpk_range_prove_batch_device(gens.workspace.data_ptr(), d_v.data_ptr(), d_g.data_ptr(),
                                                   d_s.data_ptr(), gens.n, m, out.data_ptr(), ws.data_ptr(), nbytes.value,
                                                   _stream_ptr(stream)), "bpk_range_prove_batch_device")
    torch.cuda.current_stream().synchronize()
    return out


class RangeVerifier:
    """Batched verification with a reusable workspace: accept[i] in {0,1} per proof record."""

    def __init__(self, gens, max_proofs):
        import torch
        self.gens, self.max_proofs = gens, int(max_proofs)
        nbytes = C.c_size_t(0)
        _check(_lib().bpk_range_verify_workspace_bytes(gens.n, self.max_proofs, C.byref(nbytes)),
               "bpk_range_verify_workspace_bytes")
        self.workspace = _dev_u8(nbytes.value, gens.device)
        self.accept = torch.zeros(self.max_proofs, dtype=torch.uint8, device=gens.device)

    def __call__(self, proofs, V=None, stream=None, out=None):
        m = proofs.numel() // self.gens.record_bytes
        assert m <= self.max_proofs and proofs.is_cuda
        out = self.accept if out is None else out
        _check(_lib().bpk_range_verify_batch_device(self.gens.workspace.data_ptr(), proofs.data_ptr(),
                                                    V.data_ptr() if V is not None else None, self.gens.n, m,
                                                    out.data_ptr(), self.workspace.data_ptr(),
                                                    self.workspace.numel(), _stream_ptr(stream)),
               "bpk_range_verify_batch_device")
        return out[:m]


def cuda_range_proof_verify(proof, V, n, G, H, g, h):
    """Drop-in (cuda_bulletproof.h:61): proof is a ctypes RangeProof (any struct with the reference layout),
    V/g/h (16,) uint64, G/H (n,16) uint64."""
    G, H = _c(G, 16), _c(H, 16)
    gv, hv = _pv(G), _pv(H)
    V = np.ascontiguousarray(V, dtype=np.uint64)
    g = np.ascontiguousarray(g, dtype=np.uint64)
    h = np.ascontiguousarray(h, dtype=np.uint64)
    return bool(_guard(_lib().cuda_range_proof_verify, C.byref(proof), _ptr(V), n, C.byref(gv), C.byref(hv), _ptr(g), _ptr(h)))


def cuda_inner_product_verify(proof, P, G, H, Q):
    G, H = _c(G, 16), _c(H, 16)
    gv, hv = _pv(G), _pv(H)
    P = np.ascontiguousarray(P, dtype=np.uint64)
    Q = np.ascontiguousarray(Q, dtype=np.uint64)
    return bool(_guard(_lib().cuda_inner_product_verify, C.byref(proof), _ptr(P), C.byref(gv), C.byref(hv), _ptr(Q)))


def ipa_fold_scalars(a, b, u, u_inv, stream=None):
    """a, b: cuda uint8 (2n', 32); u, u_inv: cuda uint8 (32,) -> folded (n', 32) tensors."""
    import torch
    nh = a.numel() // 64
    ao = torch.empty((nh, 32), dtype=torch.uint8, device=a.device)
    bo = torch.empty((nh, 32), dtype=torch.uint8, device=a.device)
    _check(_lib().bpk_ipa_fold_scalars_device(ao.data_ptr(), bo.data_ptr(), a.data_ptr(), b.data_ptr(), nh,
                                              u.data_ptr(), u_inv.data_ptr(), _stream_ptr(stream)),
           "bpk_ipa_fold_scalars_device")
    return ao, bo


def ipa_fold_points(G, H, u, u_inv, stream=None):
    import torch
    nh = G.numel() // 256
    Go = torch.empty((nh, 128), dtype=torch.uint8, device=G.device)
    Ho = torch.empty((nh, 128), dtype=torch.uint8, device=G.device)
    _check(_lib().bpk_ipa_fold_points_device(Go.data_ptr(), Ho.data_ptr(), G.data_ptr(), H.data_ptr(), nh,
                                             u.data_ptr(), u_inv.data_ptr(), _stream_ptr(stream)),
           "bpk_ipa_fold_points_device")
    return Go, Ho


# ---- point codec and generator derivation (include/bpk.h) ----------------------
def _as_dev_u8(a, width, device="cuda"):
    """numpy (rows of uint64 / uint8) or cuda tensor -> contiguous (count, width) uint8 cuda tensor"""
    import torch
    if isinstance(a, np.ndarray):
        a = torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1, width))
    return a.to(device).contiguous().view(torch.uint8).reshape(-1, width)


def point_pack(points, stream=None):
    """(count, 128) extended points -> (count, 32) uint8 cuda tensor of RFC 8032 encodings
    (ge25519_pack, curve25519_ops.cu:449-468, batched)."""
    import torch
    d = _as_dev_u8(points, 128)
    out = torch.empty((d.shape[0], 32), dtype=torch.uint8, device=d.device)
    _check(_lib().bpk_point_pack_device(out.data_ptr(), d.data_ptr(), d.shape[0], _stream_ptr(stream)),
           "bpk_point_pack_device")
    return out


def point_unpack(encodings, stream=None):
    """(count, 32) encodings -> ((count, 128) uint8 cuda tensor of points with Z = 1, (count,) uint8 validity mask)
    (ge25519_unpack, curve25519_ops.cu:470-531, batched, with the validity checks of RFC 8032)."""
    import torch
    d = _as_dev_u8(encodings, 32)
    pts = torch.empty((d.shape[0], 128), dtype=torch.uint8, device=d.device)
    ok = torch.empty((d.shape[0],), dtype=torch.uint8, device=d.device)
    _check(_lib().bpk_point_unpack_device(pts.data_ptr(), ok.data_ptr(), d.data_ptr(), d.shape[0], _stream_ptr(stream)),
           "bpk_point_unpack_device")
    return pts, ok


def derive_generators(seed, count, first_index=0, device="cuda", stream=None):
    """Generators first_index .. first_index+count-1 of the family `seed` (32 bytes) as a (count, 128) uint8
    cuda tensor: prime-order points with the reference test's derivation labels
    (complete_bulletproof_test.cu:33-41,79-88), derived on the device."""
    import torch
    seed = bytes(seed)
    assert len(seed) == 32
    out = torch.empty((count, 128), dtype=torch.uint8, device=device)
    _check(_lib().bpk_gens_derive_device(out.data_ptr(), seed, first_index, count, _stream_ptr(stream)),
           "bpk_gens_derive_device")
    torch.cuda.current_stream().synchronize()
    return out


def ipa_prove(G, H, Q, a, b, transcript0=bytes(32), stream=None):
    """Inner-product argument over n = len(a) (a power of two) generators, on the device.
    G, H: (n, 128) uint8 cuda tensors (or (n, 16) uint64 numpy); Q: one point; a, b: (n, 32) / (n, 4).
    Returns (L, R, a_final, b_final, x_raw) as cuda uint8 tensors: L, R (log2 n, 128), the rest (32,)."""
    import torch
    dG, dH, dQ = _as_dev_u8(G, 128), _as_dev_u8(H, 128), _as_dev_u8(Q, 128)
    da, db = _as_dev_u8(a, 32), _as_dev_u8(b, 32)
    n = da.shape[0]
    k = n.bit_length() - 1
    nbytes = C.c_size_t(0)
    _check(_lib().bpk_ipa_prove_workspace_bytes(n, C.byref(nbytes)), "bpk_ipa_prove_workspace_bytes")
    ws = _dev_u8(nbytes.value, dG.device)
    L = torch.zeros((k, 128), dtype=torch.uint8, device=dG.device)
    R = torch.zeros((k, 128), dtype=torch.uint8, device=dG.device)
    small = torch.zeros((3, 32), dtype=torch.uint8, device=dG.device)
    _check(_lib().bpk_ipa_prove_device(dG.data_ptr(), dH.data_ptr(), dQ.data_ptr(), da.data_ptr(), db.data_ptr(), n,
                                       bytes(transcript0), L.data_ptr(), R.data_ptr(), small[0].data_ptr(),
                                       small[1].data_ptr(), small[2].data_ptr(), ws.data_ptr(), nbytes.value,
                                       _stream_ptr(stream)), "bpk_ipa_prove_device")
    torch.cuda.current_stream().synchronize()
    return L, R, small[0], small[1], small[2]


def record_to_range_proof(record, n):
    """flat proof record (include/bpk.h layout, (record_bytes,) uint8) -> (ctypes RangeProof with the reference's
    layout, V as (16,) uint64, keep-alive tuple for the arrays the struct points into)"""
    from . import RangeProof
    k = n.bit_length() - 1
    w = np.ascontiguousarray(np.asarray(record).reshape(-1)).view(np.uint64)
    proof = RangeProof()
    C.memmove(C.byref(proof), w[0:92].tobytes(), 92 * 8)  # V, A, S, T1, T2, taux, mu, t
    a, b = w[92:96].copy().reshape(1, 4), w[96:100].copy().reshape(1, 4)
    Ls = w[108:108 + 16 * k].copy().reshape(k, 16)
    Rs = w[108 + 16 * k:108 + 32 * k].copy().reshape(k, 16)
    ip = proof.ip_proof
    ip.n = n
    ip.a, ip.b = _fv(a), _fv(b)
    C.memmove(C.byref(ip.c), w[100:104].tobytes(), 32)
    C.memmove(C.byref(ip.x), w[104:108].tobytes(), 32)
    ip.L, ip.R = _pv(Ls), _pv(Rs)
    ip.L_len = k
    return proof, w[0:16].copy(), (a, b, Ls, Rs)
