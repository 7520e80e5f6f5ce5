"""Llama o_proj / down_proj at M = 201 (residual + fused RMSNorm, the production call) under forced tile / split-K configurations of
uvx_gemm_bf16 (run under gpurun): in-graph microseconds incl. the split-K reduce kernel, rotating weights."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ultravox_b200 import _lib, ops

lib = _lib.lib()
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


M = 201
for name, N, K in (("o_proj", 4096, 4096), ("down", 4096, 14336)):
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.03).bfloat16() for _ in range(COPIES)]
    h = torch.randn(M, N, device=dev).bfloat16()
    nw = torch.ones(N, dtype=torch.bfloat16, device=dev)
    xn = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    rec = dict(shape=name)
    for cfg, sp in ((0, 0), (2128, 2), (2128, 3), (2128, 4), (2128, 6), (2128, 8), (2064, 1), (2064, 2), (2064, 3), (2064, 4), (2256, 4), (2256, 8), (2256, 9)):
        lib.uvx_debug_gemm_override(cfg, sp)
        try:
            rec["%d/%d" % (cfg, sp)] = round(timed(lambda i: ops.linear(x, Ws[i % COPIES], residual=h, out=h, norm=(nw, 1e-5, xn))), 2)
        except Exception as e:
            rec["%d/%d" % (cfg, sp)] = "err " + str(e)[:50]
        finally:
            lib.uvx_debug_gemm_override(0, 0)
    print(json.dumps(rec), flush=True)
    del Ws
    torch.cuda.empty_cache()
