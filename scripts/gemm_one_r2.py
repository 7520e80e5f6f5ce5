"""A few launches of one production-path kernel for ncu captures:  python scripts/gemm_one_r2.py gate_up|qkv|down|o_proj|attn"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops
what = sys.argv[1] if len(sys.argv) > 1 else "gate_up"
M = 201
if what == "attn":
    Hq, Hkv, D = 32, 8, 128
    qkv = (torch.randn(M, (Hq + 2 * Hkv) * D, device="cuda") * 0.5).bfloat16()
    for _ in range(4):
        ops.attention_fused_qkv(qkv, 1, M, Hq, Hkv, D, D ** -0.5, True)
else:
    N, K, R = dict(gate_up=(28672, 4096, 208), qkv=(6144, 4096, 128), down=(4096, 14336, 128), o_proj=(4096, 4096, 128))[what]
    x = (torch.randn(M, K, device="cuda") * 0.5).bfloat16()
    tws = [ops.TiledWeight((torch.randn(N, K, device="cuda") * 0.03).bfloat16(), R, swiglu=(what == "gate_up")) for _ in range(3)]
    for i in range(9):
        if what == "gate_up":
            ops.linear_tiled(x, tws[i % 3], act=ops.ACT_SWIGLU)
        else:
            ops.linear_tiled(x, tws[i % 3])
torch.cuda.synchronize()
print("done", what)
