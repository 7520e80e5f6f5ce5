"""Epilogue cost of uvx_gemm_bf16 (run under gpurun): encoder shapes with none / bias / bias+GELU / bias+residual
epilogues, graph-timed over rotating weight copies (> L2).  `python scripts/gemm_epi.py [one NAME EPI]` launches a single
variant a few times (for ncu)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops, _lib

SHAPES = {"enc_qkv": (1500, 3840, 1280), "enc_out": (1500, 1280, 1280), "enc_fc1": (1500, 5120, 1280), "enc_fc2": (1500, 1280, 5120)}
EPIS = ["none", "bias", "bias_gelu", "bias_res"]


def make(name):
    M, N, K = SHAPES[name]
    copies = max(2, min(8, int(400e6 // (N * K * 2)) + 1))
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(copies)]
    x = torch.randn(M, K, device="cuda").bfloat16()
    b = torch.randn(N, device="cuda").bfloat16()
    r = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    return M, N, K, Ws, x, b, r, out


def call(epi, x, w, b, r, out):
    if epi == "none":
        ops.linear(x, w, out=out)
    elif epi == "bias":
        ops.linear(x, w, bias=b, out=out)
    elif epi == "bias_gelu":
        ops.linear(x, w, bias=b, act=ops.ACT_GELU, out=out)
    else:
        ops.linear(x, w, bias=b, residual=r, out=out)


if len(sys.argv) > 1 and sys.argv[1] == "one":
    M, N, K, Ws, x, b, r, out = make(sys.argv[2])
    for i in range(4):
        call(sys.argv[3], x, Ws[i % len(Ws)], b, r, out)
    torch.cuda.synchronize()
    print("done")
    sys.exit(0)

res = []
for name in SHAPES:
    M, N, K, Ws, x, b, r, out = make(name)
    for cfg in (0, 1256, 4256, 5512):
        _lib.lib().uvx_debug_gemm_override(cfg, 1 if cfg else 0)
        for epi in EPIS:
            for i in range(3):
                call(epi, x, Ws[i % len(Ws)], b, r, out)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(20):
                    call(epi, x, Ws[i % len(Ws)], b, r, out)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            res.append(dict(shape=name, cfg=cfg, epi=epi, us=round(us, 2), TFs=round(2.0 * M * N * K / us / 1e6, 1)))
            print(res[-1], flush=True)
    _lib.lib().uvx_debug_gemm_override(0, 0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_epi.json", "w"), indent=1)
