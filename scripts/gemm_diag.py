"""GPU-side diagnostic for the tcgen05 GEMM (run under gpurun): prints error structure, not just pass/fail."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops

torch.manual_seed(0)
def run(M, N, K, kind="rand"):
    if kind == "rand":
        x = torch.randn(M, K).bfloat16().cuda(); w = (torch.randn(N, K) * 0.1).bfloat16().cuda()
    elif kind == "eye":   # y[m, n] = x[m, n] for n < K : shows row/column permutations directly
        x = (torch.arange(M)[:, None] * 1.0 + torch.arange(K)[None, :] * 0.001).bfloat16().cuda()
        w = torch.eye(N, K).bfloat16().cuda()
    try:
        y = ops.linear(x, w).float()
        torch.cuda.synchronize()
    except Exception as e:
        print(f"[{kind} {M}x{N}x{K}] EXC {e}"); return
    ref = x.float() @ w.float().T
    err = (y - ref).abs()
    print(f"[{kind} {M}x{N}x{K}] rel={((y-ref).norm()/ref.norm()).item():.3e} max={err.max().item():.3e} "
          f"nan={torch.isnan(y).sum().item()} zero_frac={(y==0).float().mean().item():.3f}")
    if ((y - ref).norm() / ref.norm()) > 1e-2:
        rows = err.mean(1); cols = err.mean(0)
        print("  row err (first 16):", [f"{v:.2g}" for v in rows[:16].tolist()])
        print("  row err by 32-row group:", [f"{rows[i:i+32].mean().item():.2g}" for i in range(0, min(M, 256), 32)])
        print("  col err by 16-col group:", [f"{cols[i:i+16].mean().item():.2g}" for i in range(0, min(N, 128), 16)])
        print("  y[0,:8]  ", y[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
        print("  y[1,:8]  ", y[1, :8].tolist()); print("  ref[1,:8]", ref[1, :8].tolist())
        # partial-K hypothesis: does y match a GEMM over only some K slices?
        for k0 in range(0, min(K, 128), 16):
            part = x[:, k0:k0 + 16].float() @ w[:, k0:k0 + 16].float().T
            print(f"  corr with k-slice [{k0},{k0+16}): {torch.nn.functional.cosine_similarity(y.flatten(), part.flatten(), dim=0).item():.3f}")

for kind in ("eye", "rand"):
    run(128, 64, 64, kind)
    run(128, 128, 64, kind)
run(128, 128, 256); run(256, 256, 512); run(201, 4096, 4096); run(1500, 1280, 1280)
print("diag done")
