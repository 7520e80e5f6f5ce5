"""Round-2 diagnosis of the M=201 weight-streaming GEMM (run under gpurun): where do the microseconds go?
Ring depth 2/3/4, 8 vs 4 epilogue warps, pipeline isolation with the production (convergent) loops - loads only, MMAs only,
no epilogue - and the 2-SM pair kernel, for gate|up (28672 x 4096) and down (4096 x 14336) at M = 201."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import _lib, ops
lib = _lib.lib()
M, COPIES, LAUNCHES = 201, 4, 16


def timed(fn):
    fn(0); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(LAUNCHES):
            fn(i)
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / LAUNCHES)
    return best


out = []
for name, N, K, R, resid in (("gate_up", 28672, 4096, 208, False), ("down", 4096, 14336, 128, True)):
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device="cuda") * 0.03).bfloat16() for _ in range(COPIES)]
    tws = [ops.TiledWeight(w, R) for w in Ws]
    o = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16()
    nw = torch.ones(N, dtype=torch.bfloat16, device="cuda"); no = torch.empty_like(o)
    kw = dict(residual=r, norm=(nw, 1e-5, no)) if resid else {}

    def run_rm(i):
        ops.linear(x, Ws[i % COPIES], out=o, **kw)

    def run_t(i):
        ops.linear_tiled(x, tws[i % COPIES], out=o, **kw)

    def rec(label, fn, **extra):
        us = timed(fn)
        d = dict(shape=name, label=label, us=round(us, 2), w_gbs=round(N * K * 2 / us / 1e3), **extra)
        out.append(d); print(json.dumps(d), flush=True)

    for st in (2, 3, 4, 0):
        lib.uvx_debug_gemm_stages(st)
        rec(f"rowmajor stages<={st or 'max'}", run_rm)
        rec(f"tiled    stages<={st or 'max'}", run_t)
    lib.uvx_debug_gemm_stages(0)
    base = 2000 + R
    for cfg, lab in ((6000 + R, "8 epilogue warps"), (7000 + R, "diag twin mode 0")):
        lib.uvx_debug_gemm_override(cfg, 0)
        rec("rowmajor " + lab, run_rm); rec("tiled    " + lab, run_t)
    lib.uvx_debug_gemm_override(7000 + R, 0)
    for mode, lab in ((1, "loads only"), (2, "MMAs only"), (3, "no epilogue")):
        lib.uvx_debug_gemm_mode(mode)
        rec("rowmajor diag " + lab, run_rm); rec("tiled    diag " + lab, run_t)
        for st in (2, 3):
            if mode == 1:
                lib.uvx_debug_gemm_stages(st)
                rec(f"tiled    diag loads only stages<={st}", run_t)
                lib.uvx_debug_gemm_stages(0)
    lib.uvx_debug_gemm_mode(0)
    if name == "gate_up":
        lib.uvx_debug_gemm_override(5416, 1)
        rec("rowmajor 2-SM pair 256x416 (cta_group::2)", run_rm)
        lib.uvx_debug_gemm_override(2256, 1)
        rec("rowmajor (2,256) 112 CTAs", run_rm)
        lib.uvx_debug_gemm_override(0, 0)
        tsw = [ops.TiledWeight(w, 208, swiglu=True) for w in Ws]
        act = torch.empty(M, N // 2, dtype=torch.bfloat16, device="cuda")
        for cfg, lab in ((0, "4 epilogue warps"), (6208, "8 epilogue warps")):
            lib.uvx_debug_gemm_override(cfg, 0)
            rec("tiled + fused swiglu, " + lab, lambda i: ops.linear_tiled(x, tsw[i % COPIES], out=act, act=ops.ACT_SWIGLU))
        del tsw
    lib.uvx_debug_gemm_override(0, 0)
    del Ws, tws
    torch.cuda.empty_cache()
json.dump(out, open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r2_ws_diag.json"), "w"), indent=1)
