"""Launch one GEMM configuration a few times (for ncu captures):  python scripts/gemm_one.py CFG SPLITS M N K [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops, _lib
cfg, sp, M, N, K = map(int, sys.argv[1:6])
iters = int(sys.argv[6]) if len(sys.argv) > 6 else 4
copies = max(2, min(8, int(400e6 // (N * K * 2)) + 1))
Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(copies)]
x = torch.randn(M, K, device="cuda").bfloat16()
out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
_lib.lib().uvx_debug_gemm_override(cfg, sp)
for i in range(iters):
    ops.linear(x, Ws[i % copies], out=out)
torch.cuda.synchronize()
print("done", cfg, sp, M, N, K)
