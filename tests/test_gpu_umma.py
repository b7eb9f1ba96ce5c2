"""GPU tests of the tcgen05/TMA linear (seeme_test_umma_linear) against fp32 matmul references."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run(A, W, bias, R, act, npass, colmax_group=0, split=False, zsplit=False):
    from seeme_b200 import _lib
    M, K = A.shape
    N = W.shape[0]
    Y = torch.empty(M, N, device=DEV)
    Ys = torch.empty(M, N, device=DEV)
    Zs = torch.empty(M, N, device=DEV)
    cm = None
    if colmax_group:
        cm = torch.zeros((M + colmax_group - 1) // colmax_group, N, device=DEV, dtype=torch.int32)
    _lib.check(_lib.lib().seeme_test_umma_linear(A.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None,
                                                 R.data_ptr() if R is not None else None, Y.data_ptr(), M, N, K, act, npass,
                                                 cm.data_ptr() if cm is not None else None, colmax_group,
                                                 Ys.data_ptr() if split else None, Zs.data_ptr() if zsplit else None,
                                                 torch.cuda.current_stream().cuda_stream), "seeme_test_umma_linear")
    if split or zsplit:
        return Y, cm, Ys, Zs
    return Y, cm


def ord2f(u):
    u = u.to(torch.int64) & 0xFFFFFFFF
    neg = (u & 0x80000000) == 0
    bits = torch.where(neg, (~u) & 0xFFFFFFFF, u & 0x7FFFFFFF)
    return bits.to(torch.int32).view(torch.float32) if False else torch.tensor(
        [__import__("struct").unpack("f", __import__("struct").pack("I", int(b)))[0] for b in bits.flatten().tolist()]).view(u.shape)


ACTS = {0: lambda x: x, 1: torch.relu, 2: torch.nn.functional.gelu, 3: torch.nn.functional.silu}


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 256), (300, 256, 512), (1000, 768, 256), (517, 1024, 256),
                                   (512, 256, 1024), (77, 128, 128), (4096, 512, 256)])
@pytest.mark.parametrize("npass", [1, 3])
def test_umma_linear_matches_fp32(M, N, K, npass):
    g = torch.Generator().manual_seed(M * 7 + N + K)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    R = torch.randn(M, N, generator=g).to(DEV)
    act = (M + N) % 4
    Y, _ = run(A, W, bias, R, act, npass)
    if npass == 1:
        ref = ACTS[act](A.bfloat16().float().double() @ W.bfloat16().float().double().T + bias.double()).float() + R
        tol = 2e-4
    else:
        ref = ACTS[act](A.double() @ W.double().T + bias.double()).float() + R
        tol = 2e-4
    err = float((Y - ref).abs().max())
    assert err < tol, (err, M, N, K, npass)


def test_umma_linear_colmax_two_groups_per_tile():
    g = torch.Generator().manual_seed(3)
    M, N, K, grp = 1000, 256, 256, 200          # groups of 200 rows: tiles straddle group boundaries
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / 16).to(DEV)
    Y, cm = run(A, W, None, None, 0, 3, colmax_group=grp)
    got = ord2f(cm.cpu())
    ref = Y.cpu().view(5, grp, N).max(dim=1)[0]
    assert torch.equal(got, ref)


@pytest.mark.parametrize("M,N,K", [(300, 256, 256), (2500, 128, 128), (4100, 256, 512)])
def test_umma_linear_bf16_outputs(M, N, K):
    """the TMA-stored bf16 (hi, lo) outputs of the result and of relu(result)"""
    g = torch.Generator().manual_seed(M + N)
    A = torch.randn(M, K, generator=g).to(DEV)
    W = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    ref = (A.double() @ W.double().T + bias.double()).float()
    Y, _, Ys, _ = run(A, W, bias, None, 0, 3, split=True)
    assert float((Y - ref).abs().max()) < 2e-4
    assert float((Ys - Y).abs().max()) < 1e-4          # hi + lo reproduces the fp32 value to ~2^-17 relative
    _, _, Ys2, Zs = run(A, W, bias, None, 0, 3, split=True, zsplit=True)
    assert float((Ys2 - ref).abs().max()) < 3e-4
    assert float((Zs - torch.relu(ref)).abs().max()) < 3e-4
