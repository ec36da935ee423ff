"""GPU probe: tcgen05 GEMM vs torch and vs the SIMT checker; prints timing."""
import ctypes, sys, os, torch
lib = ctypes.CDLL(os.path.join(os.path.dirname(__file__), "..", "whisper.coreml_b200", "libwhisper_b200.so"))
lib.b200TestGemm.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int] * 6
lib.b200LastError.argtypes = [ctypes.c_char_p, ctypes.c_int]
torch.manual_seed(0)
dev = "cuda"
def run(M, N, K, fp32, gelu, bias, simt=0):
    A = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    B = (torch.randn(N, K, device=dev) * 0.5).bfloat16()
    b = torch.randn(N, device=dev) if bias else None
    C = torch.full((M, N), float("nan"), device=dev, dtype=torch.float32 if fp32 else torch.bfloat16)
    lib.b200TestGemm(A.data_ptr(), B.data_ptr(), b.data_ptr() if bias else None, C.data_ptr(), M, N, K, fp32, gelu, simt)
    torch.cuda.synchronize()
    ref = A.float() @ B.float().t()
    if bias: ref = ref + b
    if gelu: ref = torch.nn.functional.gelu(ref)
    err = (C.float() - ref).abs().max().item()
    rel = ((C.float() - ref).norm() / ref.norm()).item()
    buf = ctypes.create_string_buffer(1024); ne = lib.b200LastError(buf, 1024)
    print(f"M{M} N{N} K{K} fp32={fp32} gelu={gelu} bias={bias} simt={simt}: maxerr {err:.4g} rel {rel:.3g} nan {torch.isnan(C.float()).sum().item()} errs {ne} {buf.value.decode() if ne else ''}", flush=True)
    return rel
run(128, 128, 64, 1, 0, 0, simt=1)
for (M, N, K) in [(128, 128, 64), (128, 128, 256), (256, 256, 128), (128, 256, 64), (200, 136, 192), (1500, 1280, 1280), (1500, 3840, 1280), (3000, 5120, 1280), (12000, 1280, 5120)]:
    run(M, N, K, 1, 0, 0)
run(1500, 5120, 1280, 0, 1, 1)
run(1500, 384, 384, 0, 0, 1)
# timing
for (M, N, K) in [(12000, 5120, 1280), (12000, 1280, 5120), (12000, 3840, 1280), (1500, 5120, 1280), (24000, 5120, 1280)]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(3): lib.b200TestGemm(A.data_ptr(), B.data_ptr(), None, C.data_ptr(), M, N, K, 0, 0, 0)
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record()
    for _ in range(10): lib.b200TestGemm(A.data_ptr(), B.data_ptr(), None, C.data_ptr(), M, N, K, 0, 0, 0)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    s.record()
    for _ in range(10): torch.matmul(A, B.t(), out=C)
    e.record(); torch.cuda.synchronize()
    ms2 = s.elapsed_time(e) / 10
    print(f"time M{M} N{N} K{K}: mine {ms:.3f} ms = {2*M*N*K/ms/1e9:.0f} TF/s ; cublas {ms2:.3f} ms = {2*M*N*K/ms2/1e9:.0f} TF/s", flush=True)
