"""Kernel-level library bar (run under gpurun): the step's GEMM shapes through torch.matmul (cuBLAS / cuBLASLt bf16) and through
uvx_gemm_bf16 in its production configuration, in-graph microseconds per launch over rotating weight copies (> L2)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ultravox_b200 import ops

dev = "cuda"
COPIES, LAUNCHES = 4, 16


def timed(fn_of_i):
    fn_of_i(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(LAUNCHES):
            fn_of_i(i)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / LAUNCHES)
    del g
    return best


shapes = [("llama qkv", 201, 6144, 4096, False, False), ("llama o_proj (+residual)", 201, 4096, 4096, False, True),
          ("llama gate|up", 201, 28672, 4096, False, False), ("llama down (+residual)", 201, 4096, 14336, False, True),
          ("enc qkv (+bias)", 1500, 3840, 1280, True, False), ("enc out (+bias +residual)", 1500, 1280, 1280, True, True),
          ("enc fc1 (+bias, GELU separate for cuBLAS)", 1500, 5120, 1280, True, False), ("enc fc2 (+bias +residual)", 1500, 1280, 5120, True, True)]
for name, M, N, K, bias, resid in shapes:
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.03).bfloat16() for _ in range(COPIES)]
    b = torch.randn(N, device=dev).bfloat16() if bias else None
    r = torch.randn(M, N, device=dev).bfloat16() if resid else None
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)

    def lib(i):
        if bias:
            y = torch.addmm(b, x, Ws[i % COPIES].t())
        else:
            y = torch.mm(x, Ws[i % COPIES].t())
        if resid:
            y = y + r
        return y

    def ours(i):
        ops.linear(x, Ws[i % COPIES], bias=b, residual=r, out=out)

    rec = dict(shape=name, M=M, N=N, K=K, cublas_us=round(timed(lib), 2), uvx_us=round(timed(ours), 2))
    rec["uvx_over_cublas"] = round(rec["uvx_us"] / rec["cublas_us"], 3)
    print(json.dumps(rec), flush=True)
    del Ws
    torch.cuda.empty_cache()
