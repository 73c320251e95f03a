"""tcgen05 3xTF32 GEMM building block (csrc/tc_gemm.cuh) against float64 matmul."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def pb():
    import uav_wrf_les_ppo_lstm_b200 as m
    return m


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (128, 128, 256), (1000, 128, 256), (384, 256, 128), (4096, 256, 128),
                                   (65536, 128, 256)])
def test_tc_gemm_3xtf32(M, N, K):
    m = pb()
    lib = m._lib.load()
    rng = np.random.default_rng(M + N + K)
    A = (rng.standard_normal((M, K)) * np.exp(rng.uniform(-3, 3, (M, 1)))).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.plume_tc_gemm(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.plume_last_error()
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = c.cpu().numpy().astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T       # condition-aware bound
    err = np.abs(got - want) / scale
    assert np.isfinite(got).all()
    assert err.max() < 2e-6, err.max()
    # fp32 matmul of the same data is not meaningfully better
    ref32 = (torch.from_numpy(A) @ torch.from_numpy(B).T).numpy().astype(np.float64)
    err32 = np.abs(ref32 - want) / scale
    assert err.max() < 20 * max(err32.max(), 1e-8)


@pytest.mark.parametrize("scaled", [1, 0, 2])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (128, 128, 256), (1000, 128, 256), (384, 256, 128), (65536, 128, 256)])
def test_tc_gemm_f16_two_term_split(M, N, K, scaled):
    """kind::f16 with x = hi + lo/s.  Scaled lo (s = 2^11, separate accumulator): fp32-grade for any magnitude that fits
    fp16's hi range; unscaled lo (one accumulator): fp32-grade for O(1) operands, which is what the update kernel feeds it;
    scaled = 2: unscaled lo with the A operand staged MN-major (instruction-descriptor bit 15, leading / stride byte
    offsets swapped roles): the form in which the update kernel's G2 reads the resident dz2 operand."""
    m = pb()
    lib = m._lib.load()
    rng = np.random.default_rng(M + N + K + scaled)
    spread = 3.0 if scaled == 1 else 0.5
    A = (rng.standard_normal((M, K)) * np.exp(rng.uniform(-spread, spread, (M, 1)))).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.plume_tc_gemm_f16(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, scaled,
                               torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.plume_last_error()
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = c.cpu().numpy().astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    err = np.abs(got - want) / scale
    assert np.isfinite(got).all()
    assert err.max() < (2e-6 if scaled == 1 else 4e-6), err.max()


@pytest.mark.parametrize("M,K", [(128, 128), (256, 256), (1024, 384)])
def test_tc_gemm_f16_b_mn_major_bulk_round_trip(M, K):
    """The form of the update kernel's dW2 GEMM: A K-major with the strides of the resident dz2 operand, B (N = 64)
    staged MN-major in the layout of a forward-activation chunk (instruction-descriptor bit 16) after a cp.async.bulk
    round trip shared -> global -> shared (the activation stash)."""
    m = pb()
    lib = m._lib.load()
    N = 64
    rng = np.random.default_rng(M + K)
    A = (rng.standard_normal((M, K)) * np.exp(rng.uniform(-0.5, 0.5, (M, 1)))).astype(np.float32)
    B = rng.standard_normal((N, K)).astype(np.float32)
    a, b = torch.from_numpy(A).cuda(), torch.from_numpy(B).cuda()
    c = torch.full((M, N), float("nan"), dtype=torch.float32, device="cuda")
    rc = lib.plume_tc_gemm_f16(a.data_ptr(), b.data_ptr(), c.data_ptr(), M, N, K, 3, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.plume_last_error()
    torch.cuda.synchronize()
    want = A.astype(np.float64) @ B.astype(np.float64).T
    got = c.cpu().numpy().astype(np.float64)
    scale = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    assert np.isfinite(got).all()
    assert (np.abs(got - want) / scale).max() < 4e-6
