"""Round-2 sweep of the Llama-prefill weight-streaming GEMMs (run under gpurun): in-graph microseconds per launch for the four
shapes of one Llama-3.1-8B layer at S = 201, row-major vs pre-tiled weights, L2 prefetch distance 0..24 k-blocks, fused SwiGLU /
RoPE epilogues.  Every graph holds 16 launches over 4 rotating weight copies (> L2 for the big shapes), timed by CUDA events."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ultravox_b200 import _lib, ops

lib = _lib.lib()
M = 201
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


res = []
shapes = [("qkv", 6144, 4096, 128), ("o_proj", 4096, 4096, 128), ("gate_up", 28672, 4096, 208), ("down", 4096, 14336, 128)]
only = sys.argv[1:] or [s[0] for s in shapes]
for name, N, K, R in shapes:
    if name not in only:
        continue
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.03).bfloat16() for _ in range(COPIES)]
    out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    res_in = (torch.randn(M, N, device=dev)).bfloat16()
    nw = torch.ones(N, dtype=torch.bfloat16, device=dev)
    nout = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    resid = name in ("o_proj", "down")
    floor_us = N * K * 2 / 6.33e12 * 1e6
    for layout in ("rowmajor", "tiled"):
        tws = [ops.TiledWeight(w, R) for w in Ws] if layout == "tiled" else None
        for pf in (0, 4, 8, 12, 16, 24):
            lib.uvx_debug_gemm_pf(pf)

            def run(i):
                kw = dict(residual=res_in, norm=(nw, 1e-5, nout)) if resid else {}
                if tws is not None:
                    ops.linear_tiled(x, tws[i % COPIES], out=out, **kw)
                else:
                    ops.linear(x, Ws[i % COPIES], out=out, **kw)
            us = timed(run)
            rec = dict(shape=name, N=N, K=K, layout=layout, pf=pf, us=round(us, 2), w_gbs=round(N * K * 2 / us / 1e3), hbm_floor_us=round(floor_us, 2))
            res.append(rec)
            print(json.dumps(rec), flush=True)
        del tws
    lib.uvx_debug_gemm_pf(-1)
    if name == "gate_up":
        tws = [ops.TiledWeight(w, 208, swiglu=True) for w in Ws]
        act = torch.empty(M, N // 2, dtype=torch.bfloat16, device=dev)
        for pf in (0, 12, 24):
            lib.uvx_debug_gemm_pf(pf)
            us = timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=act, act=ops.ACT_SWIGLU))
            rec = dict(shape=name, layout="tiled+swiglu", pf=pf, us=round(us, 2), w_gbs=round(N * K * 2 / us / 1e3), hbm_floor_us=round(floor_us, 2))
            res.append(rec)
            print(json.dumps(rec), flush=True)
        us = timed(lambda i: ops.swiglu(out, gate_first=True, out=act))
        print(json.dumps(dict(shape="swiglu_kernel_alone", us=round(us, 2))), flush=True)
        del tws
    if name == "qkv":
        inv = ops.llama3_inv_freq(128, 500000.0, None)
        cos, sin = ops.rope_tables(inv, 512, dev)
        rope = (cos, sin, None, M, 0, 40 * 128)
        tws = [ops.TiledWeight(w, 128) for w in Ws]
        for pf in (0, 12):
            lib.uvx_debug_gemm_pf(pf)
            us = timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=out, rope=rope))
            rec = dict(shape=name, layout="tiled+rope", pf=pf, us=round(us, 2), w_gbs=round(N * K * 2 / us / 1e3), hbm_floor_us=round(floor_us, 2))
            res.append(rec)
            print(json.dumps(rec), flush=True)
        us = timed(lambda i: ops.rope_(out, 32, 8, 128, cos, sin, rows_per_seq=M))
        print(json.dumps(dict(shape="rope_kernel_alone", us=round(us, 2))), flush=True)
        del tws
    lib.uvx_debug_gemm_pf(-1)
    del Ws
    torch.cuda.empty_cache()
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r2_ws_sweep.json"), "w"), indent=1)
