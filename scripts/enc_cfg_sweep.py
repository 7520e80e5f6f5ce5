"""Encoder GEMM shapes (T = 1500) with their production epilogues under every tile configuration of uvx_gemm_bf16 (run under gpurun):
heuristic vs forced 1-SM tiles vs the 2-SM pair kernels, optionally split-K; in-graph microseconds over rotating weights."""
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


M = 1500
shapes = [("qkv", 3840, 1280, True, False, False), ("out", 1280, 1280, True, True, False), ("fc1", 5120, 1280, True, False, True),
          ("fc2", 1280, 5120, True, True, False)]
cfgs = [(0, 0), (4128, 1), (9128, 1), (4256, 1), (9256, 1)]
for name, N, K, bias, resid, gelu in shapes:
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.03).bfloat16() for _ in range(COPIES)]
    b = torch.randn(N, device=dev).bfloat16()
    r = torch.randn(M, N, device=dev).bfloat16() if resid else None
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    rec = dict(shape=name, N=N, K=K)
    for cfg, sp in cfgs:
        if sp > 1 and K < 2048:
            continue
        lib.uvx_debug_gemm_override(cfg, sp if cfg else 0)
        try:
            us = timed(lambda i: ops.linear(x, Ws[i % COPIES], bias=b, residual=r, act=ops.ACT_GELU if gelu else ops.ACT_NONE, out=out))
            rec["%d/%d" % (cfg, sp)] = round(us, 2)
        except Exception as e:
            rec["%d/%d" % (cfg, sp)] = "err " + str(e)[:60]
        finally:
            lib.uvx_debug_gemm_override(0, 0)
    print(json.dumps(rec), flush=True)
    del Ws
    torch.cuda.empty_cache()
