"""Error behaviour of the device-resident C ABI (include/bpk.h): bad arguments and short workspaces return a
status, touch nothing, never exit the process (the reference's CUDA_CHECK calls exit(), cuda_field_ops.cu:14-22)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

BPK_OK, BPK_ERR_ARG, BPK_ERR_CUDA, BPK_ERR_WORKSPACE = 0, 1, 2, 3


def test_msm_argument_and_workspace_errors():
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    n = 1000
    pts, _ = cbp.synth_points(n, seed=1)
    sc = cbp.synth_scalars(n, seed=2)
    out = torch.full((128,), 0xAB, dtype=torch.uint8, device="cuda")
    nb = C.c_size_t(0)
    assert lib.bpk_msm_workspace_bytes(n, 0, C.byref(nb)) == BPK_OK and nb.value > 0
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    lib.bpk_clear_last_error()
    # workspace one byte short
    assert lib.bpk_msm_device(sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), ws.data_ptr(), nb.value - 1, 0, 1, None) == BPK_ERR_WORKSPACE
    assert lib.bpk_last_error() == BPK_ERR_WORKSPACE
    # unsupported window widths, null pointers
    assert lib.bpk_msm_device(sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), ws.data_ptr(), nb.value, 3, 1, None) == BPK_ERR_ARG
    assert lib.bpk_msm_device(sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), ws.data_ptr(), nb.value, 18, 1, None) == BPK_ERR_ARG
    assert lib.bpk_msm_device(None, pts.data_ptr(), n, out.data_ptr(), ws.data_ptr(), nb.value, 0, 1, None) == BPK_ERR_ARG
    assert lib.bpk_msm_device(sc.data_ptr(), pts.data_ptr(), n, None, ws.data_ptr(), nb.value, 0, 1, None) == BPK_ERR_ARG
    torch.cuda.synchronize()
    assert bool((out == 0xAB).all().item())  # nothing was written by the failed calls
    # and the same call with valid arguments still works afterwards
    assert lib.bpk_msm_device(sc.data_ptr(), pts.data_ptr(), n, out.data_ptr(), ws.data_ptr(), nb.value, 0, 1, None) == BPK_OK
    torch.cuda.synchronize()
    assert not bool((out == 0xAB).all().item())
    lib.bpk_clear_last_error()
    assert lib.bpk_last_error() == BPK_OK


def test_range_api_argument_errors():
    import torch
    import cudabulletproof_b200 as cbp
    lib = cbp.load()
    nb = C.c_size_t(0)
    for bad_n in (0, 3, 48, 128):  # widths must be powers of two <= 64
        assert lib.bpk_gens_workspace_bytes(bad_n, C.byref(nb)) == BPK_ERR_ARG
        assert lib.bpk_range_verify_workspace_bytes(bad_n, 10, C.byref(nb)) == BPK_ERR_ARG
        assert lib.bpk_range_prove_workspace_bytes(bad_n, 10, C.byref(nb)) == BPK_ERR_ARG
    assert lib.bpk_gens_workspace_bytes_ex(64, 12, C.byref(nb)) == BPK_ERR_ARG  # window width 8 or 16 only
    # a verification call with a table this process did not build is refused, the accept mask untouched
    fake = torch.zeros(1 << 20, dtype=torch.uint8, device="cuda")
    proofs = torch.zeros((4, lib.bpk_proof_record_bytes(16)), dtype=torch.uint8, device="cuda")
    acc = torch.full((4,), 7, dtype=torch.uint8, device="cuda")
    assert lib.bpk_range_verify_workspace_bytes(16, 4, C.byref(nb)) == BPK_OK
    ws = torch.empty(nb.value, dtype=torch.uint8, device="cuda")
    assert lib.bpk_range_verify_batch_device(fake.data_ptr(), proofs.data_ptr(), None, 16, 4, acc.data_ptr(), ws.data_ptr(),
                                             nb.value, None) == BPK_ERR_ARG
    torch.cuda.synchronize()
    assert acc.cpu().tolist() == [7, 7, 7, 7]
    assert lib.bpk_ipa_prove_workspace_bytes(48, C.byref(nb)) == BPK_ERR_ARG
    assert lib.bpk_ipa_prove_workspace_bytes(1, C.byref(nb)) == BPK_ERR_ARG


def test_all_zero_proof_records_are_rejected_not_crashing(oracle):
    """garbage in: every record all zeros (points with Z = 0) -> reject, for every proof, through the batch verifier"""
    import torch
    import cudabulletproof_b200 as cbp
    from tests.helpers import Gens
    g = Gens(oracle, 16)
    dg = cbp.Generators(g.G, g.H, g.g, g.h)
    m = 70
    proofs = torch.zeros((m, dg.record_bytes), dtype=torch.uint8, device="cuda")
    acc = cbp.RangeVerifier(dg, m)(proofs).cpu().numpy()
    assert not acc.any()
