"""Debug helper: accumulation bias of the 3xTF32 tcgen05 GEMM vs float64 as K grows (positive inputs)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import uav_wrf_les_ppo_lstm_b200 as m
lib = m._lib.load()
torch.manual_seed(0)
torch.backends.cuda.matmul.allow_tf32 = False
for K in (32, 256, 1024, 8192, 65536):
    for kind in ("positive", "signed"):
        A = torch.rand(256, K, device="cuda") if kind == "positive" else torch.randn(256, K, device="cuda")
        B = torch.rand(128, K, device="cuda") if kind == "positive" else torch.randn(128, K, device="cuda")
        C = torch.empty(256, 128, device="cuda")
        rc = lib.plume_tc_gemm(A.data_ptr(), B.data_ptr(), C.data_ptr(), 256, 128, K, torch.cuda.current_stream().cuda_stream)
        assert rc == 0
        ref = A.double() @ B.double().T
        c32 = A @ B.T
        scale = (A.double().abs() @ B.double().abs().T)
        e = (C.double() - ref) / scale
        e32 = (c32.double() - ref) / scale
        print(f"K={K:6d} {kind:8s} tc: mean {e.mean():+.2e} max {e.abs().max():.2e} | cublas fp32: mean {e32.mean():+.2e} max {e32.abs().max():.2e}")
