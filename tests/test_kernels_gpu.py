"""Kernel-level GPU checks: tcgen05 GEMM and flash attention against plain torch fp32 references."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (200, 136, 192), (1500, 3840, 1280), (3000, 384, 1152), (256, 51872, 384)])
@pytest.mark.parametrize("fp32,gelu,bias", [(1, 0, 0), (0, 1, 1)])
def test_gemm_tcgen05(lib, M, N, K, fp32, gelu, bias):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
    torch.cuda.synchronize()
    lib.b200TestGemm(A.data_ptr(), B.data_ptr(), b.data_ptr() if bias else None, C.data_ptr(), M, N, K, fp32, gelu, 0)
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert not torch.isnan(C.float()).any()
    assert _rel(C.float(), ref) < (1e-5 if fp32 else 4e-3)


def test_gemm_gelu_epilogue_accuracy(lib):
    """The GEMM epilogue's erf-GELU is a fitted form (common.cuh: gelu_erf_fit): within 1e-4 absolute of torch's exact GELU."""
    M, N, K = 256, 512, 64
    g = torch.Generator(device="cuda").manual_seed(7)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.6).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.6).bfloat16()
    b = torch.randn(N, device="cuda", generator=g)
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32)
    torch.cuda.synchronize()
    lib.b200TestGemm(A.data_ptr(), B.data_ptr(), b.data_ptr(), C.data_ptr(), M, N, K, 1, 1, 0)
    pre = A.double() @ B.double().t() + b.double()
    ref = torch.nn.functional.gelu(pre)
    assert float(pre.abs().max()) > 8.0                   # the saturated tails are exercised too
    assert float((C.double() - ref).abs().max()) < 1e-4


@pytest.mark.parametrize("M,N,K", [(256, 256, 64), (3000, 1280, 1280), (1500, 3840, 1280), (3000, 5120, 1280), (3000, 1280, 5120), (777, 392, 192)])
@pytest.mark.parametrize("fp32,gelu,bias", [(1, 0, 0), (0, 1, 1)])
def test_gemm_cta_pair(lib, M, N, K, fp32, gelu, bias):
    """The cta_group::2 (256 x 256 tile) kernel, forced for every shape, against torch fp32."""
    g = torch.Generator(device="cuda").manual_seed(M + N + K + 1)
    A = (torch.randn(M, K, device="cuda", generator=g) * 0.5).bfloat16()
    B = (torch.randn(N, K, device="cuda", generator=g) * 0.5).bfloat16()
    b = torch.randn(N, device="cuda", generator=g) if bias else None
    C = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if fp32 else torch.bfloat16)
    torch.cuda.synchronize()
    lib.b200TestGemmTile(3)
    try:
        lib.b200TestGemm(A.data_ptr(), B.data_ptr(), b.data_ptr() if bias else None, C.data_ptr(), M, N, K, fp32, gelu, 0)
    finally:
        lib.b200TestGemmTile(0)
    ref = A.float() @ B.float().t()
    if bias:
        ref = ref + b
    if gelu:
        ref = torch.nn.functional.gelu(ref)
    assert not torch.isnan(C.float()).any()
    assert _rel(C.float(), ref) < (1e-5 if fp32 else 4e-3)


@pytest.mark.parametrize("ramp", [0.0, 6.0])
@pytest.mark.parametrize("n_tok,heads,batch", [(1500, 6, 1), (1500, 20, 2), (128, 2, 1), (77, 1, 3), (200, 4, 1)])
def test_flash_attention_tcgen05(lib, n_tok, heads, batch, ramp):
    """ramp > 0 makes the keys grow along the sequence, so the running row maximum keeps rising and the kernel's lazy
    rescale of the TMEM-resident output fires in most key blocks."""
    d = heads * 64
    g = torch.Generator(device="cuda").manual_seed(n_tok + heads)
    qkv = torch.randn(batch, n_tok, 3 * d, device="cuda", generator=g)
    qkv[..., :d] *= 0.5                                   # scores with a few units of spread, like 0.125-scaled encoder keys
    if ramp:
        qkv[..., d:2 * d] *= (1.0 + ramp * torch.arange(n_tok, device="cuda") / n_tok)[None, :, None]
    qkv = qkv.bfloat16().contiguous()
    out = torch.full((batch, n_tok, d), float("nan"), device="cuda", dtype=torch.bfloat16)
    chk = torch.empty_like(out)
    torch.cuda.synchronize()
    lib.b200TestAttention(qkv.data_ptr(), out.data_ptr(), n_tok, heads, batch, 0)
    lib.b200TestAttention(qkv.data_ptr(), chk.data_ptr(), n_tok, heads, batch, 1)
    q, k, v = [t.float().view(batch, n_tok, heads, 64).transpose(1, 2) for t in qkv.split(d, dim=-1)]
    ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).transpose(1, 2).reshape(batch, n_tok, d)
    assert not torch.isnan(out.float()).any()
    assert _rel(chk.float(), ref) < 1e-2                  # SIMT checker vs torch
    assert _rel(out.float(), ref) < 1e-2                  # tcgen05 kernel vs torch (bf16 P and output)
