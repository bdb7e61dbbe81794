"""GPU parity: batched fe25519 ops, batch inversion and mod-l inner products, called through the
host-pointer C ABI (cuda_bulletproof.h drop-ins) and compared bit-exactly with the CPU oracle
(oracle/ref_corrected.c) on the same seeded inputs.  Mirrors the reference's only field-op call
sites (complete_bulletproof_test.cu:280,286,292: 10 000 elements)."""
import ctypes as C
import random

import numpy as np
import pytest

from oracle import binding as ob
from oracle import pyref

pytestmark = pytest.mark.gpu

P, L = pyref.P, pyref.L
EDGE = [0, 1, 2, 19, 38, P - 1, P, P + 1, 2 * P, 2 * P + 37, 2**255, 2**256 - 1, 2**256 - 38, 2**255 - 20, 2**128,
        2**128 - 1, 2**192, L, L - 1, 2**252]


def rand_fe(rng, n, edge=True):
    vals = [rng.getrandbits(256) for _ in range(n)]
    if edge:
        for i, e in enumerate(EDGE):
            if i < n:
                vals[i] = e
    return vals


def to_fe(vals):
    return ob.ints_to_fe(vals) if len(vals) else np.zeros((0, 4), np.uint64)


def oracle_batch(oracle, name, *arrs):
    out = np.zeros_like(arrs[0])
    f = getattr(oracle, name)
    for i in range(arrs[0].shape[0]):
        f(ob.ptr(out[i]), *[ob.ptr(a[i]) for a in arrs])
    return out


@pytest.mark.parametrize("count", [1, 7, 10000])
def test_batch_add_sub_mul_square_bit_exact(oracle, count):
    import cudabulletproof_b200 as cbp
    rng = random.Random(100 + count)
    va, vb = rand_fe(rng, count), rand_fe(rng, count)
    rng.shuffle(vb)
    a, b = to_fe(va), to_fe(vb)
    for gpu_fn, name in [(cbp.cuda_batch_field_add, "fe25519_add"), (cbp.cuda_batch_field_sub, "fe25519_sub"),
                         (cbp.cuda_batch_field_mul, "fe25519_mul"), (cbp.cuda_soa_field_add, "fe25519_add")]:
        got = gpu_fn(a, b)
        want = oracle_batch(oracle, name, a, b)
        assert np.array_equal(got, want), name
    got = cbp.cuda_batch_field_square(a)
    assert np.array_equal(got, oracle_batch(oracle, "fe25519_sq", a))
    # spot-check against big-int arithmetic too (independent of the oracle)
    m = cbp.cuda_batch_field_mul(a, b)
    for i in range(0, count, max(1, count // 50)):
        assert ob.fe_to_int(m[i]) == va[i] * vb[i] % P


def test_batch_ops_empty_is_noop():
    import cudabulletproof_b200 as cbp
    e = np.zeros((0, 4), dtype=np.uint64)
    assert cbp.cuda_batch_field_add(e, e).shape == (0, 4)
    assert cbp.cuda_batch_field_invert(e).shape == (0, 4)


@pytest.mark.parametrize("count", [1, 2, 255, 4097, 20000])
def test_batch_invert_bit_exact(oracle, count):
    import cudabulletproof_b200 as cbp
    rng = random.Random(200 + count)
    vals = rand_fe(rng, count)  # includes 0, p, 2p: all map to 0
    a = to_fe(vals)
    got = cbp.cuda_batch_field_invert(a)
    want = np.zeros_like(a)
    oracle.fe25519_batch_invert(ob.ptr(want), ob.ptr(a), count)
    assert np.array_equal(got, want)
    for i in range(0, count, max(1, count // 40)):
        v = vals[i] % P
        assert ob.fe_to_int(got[i]) == (pow(v, P - 2, P) if v else 0)


@pytest.mark.parametrize("n", [0, 1, 16, 64, 513, 4096, 100000])
def test_inner_product_mod_l_bit_exact(oracle, n):
    import cudabulletproof_b200 as cbp
    rng = random.Random(300 + n)
    va, vb = rand_fe(rng, n), rand_fe(rng, n)  # unreduced 256-bit inputs are allowed
    a, b = to_fe(va), to_fe(vb)
    for shared in (False, True):
        got = cbp.cuda_field_vector_inner_product(a, b, shared=shared)
        assert ob.fe_to_int(got) == sum(x * y for x, y in zip(va, vb)) % L
    if 0 < n <= 4096:
        want = np.zeros(4, dtype=np.uint64)
        fa, fb = ob.field_vector(a), ob.field_vector(b)
        oracle.field_vector_inner_product(ob.ptr(want), C.byref(fa), C.byref(fb))
        assert np.array_equal(got, want)


def test_inner_product_length_mismatch_leaves_result_untouched(capfd):
    """cuda_inner_product.cu:100-103: message on stderr, silent return, *result untouched."""
    import cudabulletproof_b200 as cbp
    a = to_fe([1, 2, 3])
    b = to_fe([1, 2])
    res = np.full(4, 0xDEADBEEF, dtype=np.uint64)
    cbp.cuda_field_vector_inner_product(a, b, result=res)
    assert (res == 0xDEADBEEF).all()
    assert "Vector lengths must match" in capfd.readouterr().err


@pytest.mark.parametrize("count", [1 << 16, (1 << 18) + 12345])
def test_batch_invert_tree_path_matches_single_kernel_and_python(count):
    """large arrays with a workspace take the multi-level Montgomery tree; it must give the same canonical
    bytes as the single-kernel path (which is pinned to the oracle above), zeros included"""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    a = cbp.synth_scalars(count, seed=0x1A7 + count, bits=255)
    a[7] = 0  # inv(0) = 0
    a[count - 1] = 0
    nb = C.c_size_t(0)
    assert lib.bpk_fe_batch_invert_workspace_bytes(count, C.byref(nb)) == 0 and nb.value > 0
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    tree, single = torch.empty_like(a), torch.empty_like(a)
    assert lib.bpk_fe_batch_invert_device(tree.data_ptr(), a.data_ptr(), count, ws.data_ptr(), ws.numel(), None) == 0
    assert lib.bpk_fe_batch_invert_device(single.data_ptr(), a.data_ptr(), count, None, 0, None) == 0
    torch.cuda.synchronize()
    assert torch.equal(tree, single)
    h_a, h_o = a.cpu().numpy().view(np.uint64).reshape(count, 4), tree.cpu().numpy().view(np.uint64).reshape(count, 4)
    for i in [0, 1, 7, 255, 256, 4095, 4096, count // 2, count - 2, count - 1]:
        v = ob.fe_to_int(h_a[i]) % P
        assert ob.fe_to_int(h_o[i]) == (pow(v, P - 2, P) if v else 0)


@pytest.mark.parametrize("n,num_vectors", [(1, 1), (5, 3), (64, 17), (513, 4), (4096, 2), (0, 3), (700, 1)])
def test_batched_inner_product_bit_exact(oracle, n, num_vectors):
    """cuda_batch_field_vector_inner_product (cuda_inner_product.cu:302-348, kernel K18; defined in the reference
    but missing from its header): num_vectors separately allocated vector pairs of one length -> one result each,
    against the oracle's field_vector_inner_product (mod l) per pair and Python big ints."""
    import cudabulletproof_b200 as cbp
    rng = random.Random(7000 + 31 * n + num_vectors)
    va = [rand_fe(rng, n, edge=(v == 0)) for v in range(num_vectors)]
    vb = [rand_fe(rng, n, edge=(v == 1)) for v in range(num_vectors)]
    a, b = [to_fe(x) for x in va], [to_fe(x) for x in vb]
    got = cbp.cuda_batch_field_vector_inner_product(a, b)
    assert got.shape == (num_vectors, 4)
    for v in range(num_vectors):
        assert ob.fe_to_int(got[v]) == sum(x * y for x, y in zip(va[v], vb[v])) % L, v
        if n:
            want = np.zeros(4, dtype=np.uint64)
            fa, fb = ob.field_vector(a[v]), ob.field_vector(b[v])
            oracle.field_vector_inner_product(ob.ptr(want), C.byref(fa), C.byref(fb))
            assert np.array_equal(got[v], want), v


def test_batched_inner_product_ragged_batch_is_refused(capfd):
    """The reference takes n from the first pair and copies n elements of every vector
    (cuda_inner_product.cu:306,312-315): a ragged batch reads out of bounds there.  Here it is refused like the
    single-pair length mismatch (message on stderr, results untouched)."""
    import cudabulletproof_b200 as cbp
    a = [to_fe([1, 2, 3]), to_fe([4, 5])]
    b = [to_fe([1, 2, 3]), to_fe([4, 5])]
    res = np.full((2, 4), 0xDEADBEEF, dtype=np.uint64)
    cbp.cuda_batch_field_vector_inner_product(a, b, results=res)
    assert (res == 0xDEADBEEF).all()
    assert "Vector lengths must match" in capfd.readouterr().err


@pytest.mark.parametrize("n,num_vectors", [(64, 1000), (4096, 37), (3, 5000)])
def test_batched_inner_product_device_api(n, num_vectors):
    """bpk_sc_inner_product_batch_device on contiguous device arrays (one CTA per pair), against the
    single-pair kernel pair (pinned to the oracle above) and Python big ints on a sample."""
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    a = cbp.synth_scalars(n * num_vectors, seed=0xBA7C4 + n, bits=256)
    b = cbp.synth_scalars(n * num_vectors, seed=0xBA7C5 + n, bits=253)
    out = torch.zeros((num_vectors, 32), dtype=torch.uint8, device="cuda")
    assert lib.bpk_sc_inner_product_batch_device(out.data_ptr(), a.data_ptr(), b.data_ptr(), n, num_vectors, None) == 0
    nb = C.c_size_t(0)
    lib.bpk_sc_inner_product_workspace_bytes(n, C.byref(nb))
    ws = torch.zeros(max(nb.value, 16), dtype=torch.uint8, device="cuda")
    one = torch.zeros(32, dtype=torch.uint8, device="cuda")
    h_a = a.cpu().numpy().view(np.uint64).reshape(num_vectors, n, 4)
    h_b = b.cpu().numpy().view(np.uint64).reshape(num_vectors, n, 4)
    h_o = out.cpu().numpy().view(np.uint64).reshape(num_vectors, 4)
    for v in sorted({0, 1, num_vectors // 2, num_vectors - 1}):
        assert lib.bpk_sc_inner_product_device(one.data_ptr(), a[v * n:].data_ptr(), b[v * n:].data_ptr(), n,
                                               ws.data_ptr(), ws.numel(), None) == 0
        assert np.array_equal(one.cpu().numpy().view(np.uint64), h_o[v])
        want = sum(ob.fe_to_int(x) * ob.fe_to_int(y) for x, y in zip(h_a[v], h_b[v])) % L
        assert ob.fe_to_int(h_o[v]) == want
