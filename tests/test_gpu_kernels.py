"""Kernel-level parity on the GPU: GEMMs (CUDA-core fp32 and tcgen05 bf16) and the depthwise stage,
called through the C ABI and compared with plain torch on the same inputs."""
import pytest
import torch

from conformer_b200 import _lib
from conformer_b200._lib import check, lib

pytestmark = pytest.mark.gpu


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ref_epilogue(acc, bias, epi, res, scale):
    v = acc + bias
    if epi == _lib.EPI_BIAS_SWISH:
        v = v * torch.sigmoid(v)
    elif epi == _lib.EPI_BIAS_RELU:
        v = torch.relu(v)
    elif epi == _lib.EPI_BIAS_RES:
        v = res + scale * v
    return v


@pytest.mark.parametrize("M,N,K", [(64, 64, 16), (130, 72, 36), (1000, 144, 144), (513, 1024, 256)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_gemm_f32(M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    A = torch.randn(M, K, device="cuda", generator=g)
    W = torch.randn(N, K, device="cuda", generator=g) / K ** 0.5
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda")
    check(lib().cfm_gemm_f32(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, epi,
                             res.data_ptr(), 0.5, _stream()))
    want = _ref_epilogue(A.double() @ W.double().T, bias.double(), epi, res.double(), 0.5)
    torch.testing.assert_close(out.double(), want, rtol=1e-5, atol=1e-5)


TC_SHAPES = [(128, 128, 64), (128, 256, 64), (256, 256, 256), (300, 256, 256), (1000, 144, 144), (777, 1024, 256),
             (640, 256, 1024), (2000, 768, 256), (129, 512, 4864), (23936, 256, 256)]


@pytest.mark.parametrize("M,N,K", TC_SHAPES)
@pytest.mark.parametrize("epi", [0, 1, 2, 3, 4])
def test_gemm_bf16_tcgen05(M, N, K, epi):
    g = torch.Generator(device="cuda").manual_seed(M + N + K + epi)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    res = torch.randn(M, N, device="cuda", generator=g)
    f32_out = epi in (_lib.EPI_BIAS_RES, _lib.EPI_BIAS_F32)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if f32_out else torch.bfloat16)
    check(lib().cfm_gemm_bf16(A.data_ptr(), W.data_ptr(), bias.data_ptr(), out.data_ptr(), M, N, K, epi,
                              res.data_ptr(), 0.5, _stream()))
    torch.cuda.synchronize()
    want = _ref_epilogue(A.double() @ W.double().T, bias.double(), epi, res.double(), 0.5)
    # inputs are exactly representable; the only error is fp32 accumulation (+ bf16 rounding of the output)
    tol = dict(rtol=2e-4, atol=2e-4) if f32_out else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(out.double(), want, **tol)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("B,T,d,k", [(2, 50, 64, 31), (3, 129, 144, 31), (1, 7, 32, 7), (2, 300, 256, 31), (2, 40, 48, 5)])
def test_dwconv_stage(B, T, d, k, dtype):
    g = torch.Generator(device="cuda").manual_seed(B * T + d)
    u = torch.randn(B, T, 2 * d, device="cuda", generator=g).to(dtype)
    w = torch.randn(k, d, device="cuda", generator=g) / k ** 0.5
    bias = torch.randn(d, device="cuda", generator=g)
    y = torch.empty(B, T, d, device="cuda", dtype=dtype)
    check(lib().cfm_dwconv_glu_bn_swish(u.data_ptr(), w.data_ptr(), bias.data_ptr(), y.data_ptr(), B, T, d, k,
                                        0 if dtype == torch.float32 else 1, _stream()))
    uu = u.double()
    gl = uu[..., :d] * torch.sigmoid(uu[..., d:])
    conv = torch.nn.functional.conv1d(gl.transpose(1, 2), w.double().T[:, None, :], bias.double(), padding=(k - 1) // 2,
                                      groups=d).transpose(1, 2)
    want = conv * torch.sigmoid(conv)
    tol = dict(rtol=1e-5, atol=1e-5) if dtype == torch.float32 else dict(rtol=1e-2, atol=1e-2)
    torch.testing.assert_close(y.double(), want, **tol)


@pytest.mark.parametrize("M,d,K", [(300, 64, 128), (1000, 256, 256), (777, 512, 512)])
def test_gemm_bf16_glu_epilogue(M, d, K):
    """EPI 5: out[:, c] = (A W^T + b)[c] * sigmoid((A W^T + b)[c + d]) with the value/gate rows of W packed
    per 64-row group (32 value rows, then their 32 gate rows)."""
    g = torch.Generator(device="cuda").manual_seed(M + d)
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = (torch.randn(2 * d, K, device="cuda", generator=g) / K ** 0.5).bfloat16()
    bias = torch.randn(2 * d, device="cuda", generator=g)
    r = torch.arange(2 * d, device="cuda")
    src = torch.where((r & 63) >= 32, d, 0) + (r >> 6) * 32 + (r & 31)
    Wp, bp = W[src].contiguous(), bias[src].contiguous()
    out = torch.full((M, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    check(lib().cfm_gemm_bf16(A.data_ptr(), Wp.data_ptr(), bp.data_ptr(), out.data_ptr(), M, 2 * d, K, 5, None, 0.0,
                              _stream()))
    acc = A.double() @ W.double().T + bias.double()
    want = acc[:, :d] * torch.sigmoid(acc[:, d:])
    torch.testing.assert_close(out.double(), want, rtol=1e-2, atol=1e-2)


@pytest.mark.parametrize("K,M,N", [(64, 128, 128), (200, 64, 256), (11968, 256, 1024), (747, 256, 256), (1000, 40, 72)])
def test_gemm_bf16_tn_weight_gradient_form(K, M, N):
    """cfm_gemm_bf16_tn: C += A^T B with row-major A [K,M], B [K,N] consumed in place as MN-major tcgen05 operands
    (the dW = dY^T X form of the training step), K / M / N tails, split-K for tall contractions, accumulation into C."""
    from conformer_b200._lib import check, lib
    g = torch.Generator().manual_seed(K + M + N)
    A = torch.randn(K, M, generator=g).bfloat16().to("cuda")
    B = torch.randn(K, N, generator=g).bfloat16().to("cuda")
    C0 = torch.randn(M, N, generator=g).to("cuda")
    C = C0.clone()
    check(lib().cfm_gemm_bf16_tn(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, torch.cuda.current_stream().cuda_stream))
    want = C0.double() + A.double().T @ B.double()
    err = (C.double() - want).abs().max().item()
    assert err < 2e-3 * (K ** 0.5), err
