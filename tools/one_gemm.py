"""One GEMM shape through a forced tile configuration (for ncu):  python tools/one_gemm.py <sel> [M N K mode]"""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import whisper_b200._lib as L
lib = L.load()
sel = int(sys.argv[1]) if len(sys.argv) > 1 else 2
M, N, K = (int(x) for x in sys.argv[2:5]) if len(sys.argv) > 4 else (3000, 3840, 1280)
mode = int(sys.argv[5]) if len(sys.argv) > 5 else 1
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
lib.b200TestGemmTile(sel)
print(sel, lib.b200TestGemmTime(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, mode, 3))
