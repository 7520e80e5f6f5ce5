"""Per-CTA phase timeline of the M=201 weight-streaming GEMM from the diagnostic twin (clock64 inside the kernel):
entry -> set-up done -> first stage landed -> last MMA issued -> accumulators complete -> epilogue done."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import _lib, ops
lib = _lib.lib()
M = 201
for name, N, K, R, resid in (("gate_up", 28672, 4096, 208, False), ("down", 4096, 14336, 128, True), ("qkv", 6144, 4096, 128, False)):
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    Ws = [(torch.randn(N, K, device="cuda") * 0.03).bfloat16() for _ in range(3)]
    tws = [ops.TiledWeight(w, R) for w in Ws]
    o = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    r = torch.randn(M, N, device="cuda").bfloat16()
    nw = torch.ones(N, dtype=torch.bfloat16, device="cuda"); no = torch.empty_like(o)
    kw = dict(residual=r, norm=(nw, 1e-5, no)) if resid else {}
    buf = torch.zeros(148 * 8, dtype=torch.int64, device="cuda")
    lib.uvx_debug_gemm_override(7000 + R, 0)
    for mode in (0, 1, 2, 3):
        lib.uvx_debug_gemm_mode(mode)
        lib.uvx_debug_gemm_times(None)
        for i in range(12):                                     # hot GPU: the measured launch follows a burst of the same kernel
            ops.linear_tiled(x, tws[i % 3], out=o, **kw)
        lib.uvx_debug_gemm_times(buf.data_ptr())
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.linear_tiled(x, tws[0], out=o, **kw)
        e1.record()
        torch.cuda.synchronize()
        t = buf.view(148, 8).cpu().double()
        act = t[:, 0] > 0
        t = t[act]
        wall_us = (t[:, 7] - t[:, 6]) / 1e3                    # globaltimer, ns
        mhz = float(((t[:, 5] - t[:, 0]) / wall_us).median())  # SM clock actually seen inside the kernel
        d = lambda a, b: (t[:, a] - t[:, b]) / mhz             # us at the measured clock
        g0 = t[:, 6].min()
        print(f"{name} mode {mode} ({['full','loads only','MMAs only','no epilogue'][mode]}): kernel {e0.elapsed_time(e1)*1e3:.1f} us (eager, incl. launch), "
              f"CTAs {int(act.sum())}, SM clock inside the kernel {mhz:.0f} MHz, CTA wall {float(wall_us.mean()):.1f} us")
        print(f"   start skew (globaltimer) max {float((t[:,6]-g0).max())/1e3:.2f} us | setup {d(1,0).mean():.2f} us | first stage landed +{d(2,1).mean():.2f} "
              f"(max {d(2,1).max():.2f}) | MMA issue span {d(3,2).mean():.2f} (max {d(3,2).max():.2f}) | acc complete - last issue {d(4,3).mean():.2f} "
              f"| epilogue {d(5,4).mean():.2f} (max {d(5,4).max():.2f}) | CTA total {d(5,0).mean():.2f} (min {d(5,0).min():.2f} max {d(5,0).max():.2f})")
    lib.uvx_debug_gemm_mode(0); lib.uvx_debug_gemm_times(None); lib.uvx_debug_gemm_override(0, 0)
    del Ws, tws
