"""HBM-side efficiency of the weight stream (run under gpurun): N=28672 x K=1024 (58.7 MB) with the loads-only / full kernel,
(a) one weight copy re-used (L2-resident after the first pass) vs (b) 8 rotating copies (streamed from HBM)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops, _lib
lib = _lib.lib()
M, N = 201, 28672
for K in (1024, 2048):
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(8)]
    x = torch.randn(M, K, device="cuda").bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for cfg in (2208, 1128):
        for mode, label in ((0, "full"), (1, "loads_only")):
            for copies in (1, 8):
                lib.uvx_debug_gemm_override(cfg, 1); lib.uvx_debug_gemm_mode(mode)
                ops.linear(x, Ws[0], out=out); torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    for i in range(24):
                        ops.linear(x, Ws[i % copies], out=out)
                g.replay(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1e3 / 24
                print(f"K={K} cfg={cfg} {label:10s} copies={copies} {us:7.2f} us  W-stream {N * K * 2 / us / 1e3:7.1f} GB/s", flush=True)
                del g
lib.uvx_debug_gemm_mode(0); lib.uvx_debug_gemm_override(0, 0)
