import sys, os
sys.path.insert(0, "/root/repo")
import torch
from ultravox_b200 import ops, _lib
lib = _lib.lib()
torch.manual_seed(0)
M, F, K = 201, 14336, 4096
x = (torch.randn(M, K, device="cuda")).bfloat16()
w = (torch.randn(2 * F, K, device="cuda") * 0.03).bfloat16()
tw = ops.TiledWeight(w, 208, swiglu=True)
lib.uvx_debug_gemm_tma_store(0)
slow = ops.linear_tiled(x, tw, act=ops.ACT_SWIGLU).clone()
lib.uvx_debug_gemm_tma_store(1)
fast = ops.linear_tiled(x, tw, act=ops.ACT_SWIGLU).clone()
torch.cuda.synchronize()
d = (slow != fast)
print("mismatch count", int(d.sum()), "of", d.numel())
rows = d.any(1).nonzero().flatten().tolist()
cols = d.any(0).nonzero().flatten().tolist()
print("rows", rows[:40], len(rows))
print("cols", cols[:60], len(cols))
print("cols mod 104", sorted(set(c % 104 for c in cols))[:60])
if rows:
    r, c = rows[0], cols[0]
    print(slow[r, c:c+8], fast[r, c:c+8])
