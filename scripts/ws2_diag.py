"""Where the weight-streaming GEMM's time goes (run under gpurun): in-graph microseconds with parts of the epilogue switched
off (uvx_debug_gemm_ws mode bits), and per-CTA phase timestamps of one launch."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ultravox_b200 import _lib, ops

lib = _lib.lib()
M, dev = 201, "cuda"
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


shapes = [("qkv", 6144, 4096, 144), ("qkv", 6144, 4096, 0), ("o_proj", 4096, 4096, 0), ("gate_up", 28672, 4096, 0), ("down", 4096, 14336, 0)]
for name, N, K, grid in shapes:
    x = (torch.randn(M, K, device=dev) * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device=dev) * 0.03).bfloat16() for _ in range(COPIES)]
    swiglu = name == "gate_up"
    resid = name in ("o_proj", "down")
    n_out = N // 2 if swiglu else N
    out = torch.empty(M, n_out, dtype=torch.bfloat16, device=dev)
    res_in = torch.randn(M, n_out, device=dev).bfloat16()
    tws = [ops.TiledWeight(w, 128, swiglu=swiglu) for w in Ws]
    act = ops.ACT_SWIGLU if swiglu else ops.ACT_NONE
    kw = dict(residual=res_in) if resid else {}
    rec = dict(shape=name, grid=grid or 148)
    for label, mode in (("full", 0), ("no_slot_stores", 8), ("no_slot_reads", 16), ("no_slot_rw", 24), ("no_flags_no_reads", 32), ("no_slots_no_flags", 40),
                        ("no_out_stores", 64), ("nothing_but_tmem", 8 + 32 + 64), ("no_epilogue", 3)):
        lib.uvx_debug_gemm_ws(1, mode, grid)
        rec[label] = round(timed(lambda i: ops.linear_tiled(x, tws[i % COPIES], out=out, act=act, **kw)), 2)
    print(json.dumps(rec), flush=True)
    # per-CTA timestamps of one eager launch (after a warm one)
    lib.uvx_debug_gemm_ws(1, 0, grid)
    G = grid or 148
    buf = torch.zeros(G, 16, dtype=torch.int64, device=dev)
    ops.linear_tiled(x, tws[0], out=out, act=act, **kw)
    torch.cuda.synchronize()
    lib.uvx_debug_gemm_ws_times(buf.data_ptr())
    ops.linear_tiled(x, tws[1], out=out, act=act, **kw)
    torch.cuda.synchronize()
    lib.uvx_debug_gemm_ws_times(None)
    t = buf.cpu().double()
    ghz = (t[:, 13] - t[:, 0]) / (t[:, 15] - t[:, 14]).clamp_min(1)        # cycles per ns
    us = lambda col: ((t[:, col] - t[:, 0]) / ghz / 1e3)
    g0 = t[:, 14].min()
    print("  SM clock GHz median %.2f; kernel wall (globaltimer) %.1f us" % (float(ghz.median()), float((t[:, 15].max() - g0) / 1e3)))
    cols = [("setup", 1), ("first_stage", 2), ("mma_issued", 3), ("s0_acc", 4), ("s0_flags", 5), ("s0_done", 6), ("s1_acc", 7), ("s1_flags", 8), ("s1_done", 9),
            ("s2_acc", 10), ("s2_flags", 11), ("s2_done", 12), ("exit", 13)]
    for nm, c in cols:
        v = us(c)[t[:, c] > 0]
        if len(v):
            print("  %-12s n=%3d  min %7.2f  med %7.2f  max %7.2f us after entry" % (nm, len(v), float(v.min()), float(v.median()), float(v.max())))
    print("  entry spread (globaltimer) %.2f us; exit spread %.2f us" % (float((t[:, 14].max() - g0) / 1e3), float((t[:, 15].max() - t[:, 15].min()) / 1e3)))
    for c in (0, 1, 2, 47, 48, 147):
        if c < G:
            print("   cta %3d: " % c + " ".join("%s=%.1f" % (nm, float(us(col)[c])) for nm, col in cols if t[c, col] > 0))
    del tws, Ws
    torch.cuda.empty_cache()
lib.uvx_debug_gemm_ws(-1, 0, 0)
