"""Micro-test of the no-swizzle (interleaved) K-major shared-memory descriptor used by the SMPL transform-blend GEMM."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from seeme_b200 import _lib  # noqa: E402

g = torch.Generator().manual_seed(0)
A = torch.randn(128, 32, generator=g).half().cuda()
B = torch.randn(48, 32, generator=g).half().cuda()
D = torch.zeros(128, 48, device="cuda")
rc = _lib.lib().seeme_test_umma_interleave(C.c_void_p(A.data_ptr()), C.c_void_p(B.data_ptr()), C.c_void_p(D.data_ptr()), C.c_void_p(0))
torch.cuda.synchronize()
ref = A.float() @ B.float().t()
err = (D - ref).abs().max().item()
print("rc", rc, "max err", err, "ref scale", ref.abs().max().item())
if err > 1e-3:
    # which (row, col) pattern is wrong?
    bad = (D - ref).abs() > 1e-3
    print("bad rows", bad.any(1).nonzero().flatten()[:16].tolist(), "bad cols", bad.any(0).nonzero().flatten()[:16].tolist())
    print(D[:2, :8], ref[:2, :8])
