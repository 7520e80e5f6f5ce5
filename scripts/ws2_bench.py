"""Round-2 weight-streaming GEMM (csrc/gemm_ws.cu) against gemm_tc_kernel on the four Llama-3.1-8B prefill shapes at S = 201
(run under gpurun): in-graph microseconds per launch, 16 launches over 4 rotating weight copies (> L2), CUDA events; plus the
pipeline-isolation modes of the new kernel (loads only / MMAs only / no epilogue)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ultravox_b200 import _lib, ops

lib = _lib.lib()
M = int(os.environ.get("WS_M", "201"))
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


shapes = [("qkv", 6144, 4096), ("o_proj", 4096, 4096), ("gate_up", 28672, 4096), ("down", 4096, 14336)]
only = sys.argv[1:] or [s[0] for s in shapes]
res = []
for name, N, K in shapes:
    if name not in only:
        continue
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.03).bfloat16() for _ in range(COPIES)]
    resid = name in ("o_proj", "down")
    swiglu = name == "gate_up"
    n_out = N // 2 if swiglu else N
    out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
    res_in = torch.randn(M, n_out, device=dev).bfloat16()
    nw = torch.ones(n_out, dtype=torch.bfloat16, device=dev)
    nout = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
    kw = dict(residual=res_in, norm=(nw, 1e-5, nout)) if resid else {}
    floor_us = N * K * 2 / 6.33e12 * 1e6
    rec = dict(shape=name, M=M, N=N, K=K, hbm_floor_us=round(floor_us, 2))
    # old kernel, production configuration of round 2 (gate|up: 208-row image + fused SwiGLU; others row-major)
    lib.uvx_debug_gemm_ws(0, 0, 0)
    if swiglu:
        old = [ops.TiledWeight(w, 208, swiglu=True) for w in Ws]
        rec["tc_us"] = round(timed(lambda i: ops.linear_tiled(x, old[i % COPIES], out=out, act=ops.ACT_SWIGLU)), 2)
        del old
    else:
        rec["tc_us"] = round(timed(lambda i: ops.linear(x, Ws[i % COPIES], out=out, **kw)), 2)
    # weight-streaming form: row-major weights, then the 128-row image
    lib.uvx_debug_gemm_ws(1, 0, 0)
    if not swiglu:
        rec["ws_rowmajor_us"] = round(timed(lambda i: ops.linear(x, Ws[i % COPIES], out=out, **kw)), 2)
    tws = [ops.TiledWeight(w, 128, swiglu=swiglu) for w in Ws]
    act = ops.ACT_SWIGLU if swiglu else ops.ACT_NONE
    rec["ws_us"] = round(timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=out, act=act, **kw)), 2)
    if resid:
        rec["ws_no_norm_us"] = round(timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=out, residual=res_in)), 2)
    for mode, key in ((1, "ws_loads_only_us"), (2, "ws_mma_only_us"), (3, "ws_no_epilogue_us")):
        lib.uvx_debug_gemm_ws(1, mode, 0)
        rec[key] = round(timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=out, act=act)), 2)
    lib.uvx_debug_gemm_ws(1, 0, 0)
    for grid in (144, 132, 120):
        lib.uvx_debug_gemm_ws(1, 0, grid)
        rec["ws_grid%d_us" % grid] = round(timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=out, act=act)), 2)
    lib.uvx_debug_gemm_ws(-1, 0, 0)
    rec["ws_w_gbs"] = round(N * K * 2 / rec["ws_us"] / 1e3)
    res.append(rec)
    print(json.dumps(rec), flush=True)
    del tws, Ws
    torch.cuda.empty_cache()
json.dump(res, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r2_ws2_bench.json"), "w"), indent=1)
