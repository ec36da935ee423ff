import os, sys, torch
sys.path.insert(0, "/root/repo")
import whisper_b200._lib as L
lib = L.load()
M, N, K = 3000, 3840, 1280
A = torch.randn(M, K, device="cuda").bfloat16(); B = torch.randn(N, K, device="cuda").bfloat16()
C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
lib.b200TestGemmTile(2)
print(lib.b200TestGemmTime(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, 1, 3))
