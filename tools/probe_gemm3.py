"""GPU probe: time of the four GEMMs of an encoder layer (M = 3000) per tile configuration, with their real epilogues
(mode bits: 1 bias, 2 GELU, 4 fp32 output + residual).  B200_GEMM_SKIP=1/2 removes the stores / the whole epilogue."""
import os, sys, torch
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import whisper_b200._lib as L
lib = L.load()
dev = "cuda"
tot = {1: 0.0, 2: 0.0, 3: 0.0}
for name, (M, N, K), mode in [("qkv", (3000, 3840, 1280), 1), ("out", (3000, 1280, 1280), 5), ("mlp1", (3000, 5120, 1280), 3), ("mlp2", (3000, 1280, 5120), 5),
                              ("plain", (3000, 3840, 1280), 0), ("big", (12000, 5120, 1280), 0)]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = []
    for sel in (1, 2, 3):
        lib.b200TestGemmTile(sel)
        ms = lib.b200TestGemmTime(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, mode, 20)
        if name in ("qkv", "out", "mlp1", "mlp2"): tot[sel] += ms
        out.append(f"sel{sel} {ms*1e3:.1f} us = {2*M*N*K/ms/1e9:.0f} TF/s")
    print(f"skip={os.environ.get('B200_GEMM_SKIP','0')} {name} M{M} N{N} K{K} mode {mode}: " + " ; ".join(out), flush=True)
print(f"layer GEMMs: sel1 {tot[1]*1e3:.1f} us, sel2 {tot[2]*1e3:.1f} us, sel3 {tot[3]*1e3:.1f} us (118 GF -> {118.0/tot[1]:.0f} / {118.0/tot[2]:.0f} / {118.0/tot[3]:.0f} TF/s)")
import numpy as np
lib.b200TestGemmTile(2)
for name, (M, N, K), mode in [("qkv", (3000, 3840, 1280), 1), ("mlp1", (3000, 5120, 1280), 3), ("mlp2", (3000, 1280, 5120), 5)]:
    A = torch.randn(M, K, device=dev).bfloat16(); B = torch.randn(N, K, device=dev).bfloat16()
    C = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    buf = np.zeros(8 * 16, dtype=np.uint64)
    n = lib.b200TestGemmTimeline(A.data_ptr(), B.data_ptr(), C.data_ptr(), M, N, K, mode, buf.ctypes.data, 8)
    m = buf.reshape(8, 16).astype(np.int64); t0 = m[0, 0]
    print(f"timeline {name}: CTA 0, {n} tiles; cycles since the epilogue warp entered tile 0")
    print("tile  enter  acc-ready | c0: issued landed done | c1 | c2 | c3 | released | mma-issued")
    for j in range(n):
        print(f"{j:3d} " + " ".join(f"{int(v - t0):7d}" if v else "      -" for v in m[j]))
lib.b200TestGemmTile(0)
L.check_errors("probe")
