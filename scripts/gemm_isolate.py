"""Which pipeline bounds uvx_gemm_bf16 (run under gpurun): full kernel vs loads-only vs MMAs-only (uvx_debug_gemm_mode),
graph-timed over rotating weight copies."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops, _lib
lib = _lib.lib()
if os.environ.get("CL"):
    lib.uvx_debug_gemm_cluster(*map(int, os.environ["CL"].split(",")))
PLAN = [("llm_gate_up", 201, 28672, 4096, [(2208, 1), (2256, 1), (1128, 1)]), ("llm_qkv", 201, 6144, 4096, [(1128, 1)]),
        ("llm_down", 201, 4096, 14336, [(2128, 4)]), ("enc_fc1", 1500, 5120, 1280, [(1256, 1), (1128, 1)]),
        ("enc_fc2", 1500, 1280, 5120, [(1128, 1)]), ("big", 8192, 8192, 8192, [(1256, 1), (1128, 1), (4256, 1), (5512, 1), (4128, 1), (1064, 1)])]
if len(sys.argv) > 1:
    PLAN = [pl for pl in PLAN if pl[0] in sys.argv[1:]]
for name, M, N, K, cfgs in PLAN:
    copies = max(2, min(8, int(400e6 // (N * K * 2)) + 1))
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(copies)]
    x = torch.randn(M, K, device="cuda").bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    for cfg, sp in cfgs:
        row = {}
        for mode, label in ((0, "full"), (1, "loads_only"), (2, "mma_only")):
            lib.uvx_debug_gemm_override(cfg, sp); lib.uvx_debug_gemm_mode(mode)
            ops.linear(x, Ws[0], out=out); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(20):
                    ops.linear(x, Ws[i % copies], out=out)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            row[label] = round(e0.elapsed_time(e1) * 1e3 / 20, 2)
            del g
        lib.uvx_debug_gemm_mode(0)
        row["tflops_full"] = round(2.0 * M * N * K / row["full"] / 1e6, 1)
        print(name, cfg, sp, row, flush=True)
    del Ws; torch.cuda.empty_cache()
lib.uvx_debug_gemm_override(0, 0)
