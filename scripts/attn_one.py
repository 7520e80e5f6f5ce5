"""Launch the tcgen05 encoder attention a few times (for ncu): python scripts/attn_one.py [B] [S] [H]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
H = int(sys.argv[3]) if len(sys.argv) > 3 else 20
qkv = torch.randn(B * S, 3 * H * 64, device="cuda").bfloat16()
out = torch.empty(B * S, H * 64, dtype=torch.bfloat16, device="cuda")
for _ in range(4):
    ops.attention_encoder_tc(qkv, B, S, H, 0.125, None, 0, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attention_encoder_tc(qkv, B, S, H, 0.125, None, 0, out=out)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
print(f"attn_tc B={B} S={S} H={H}: {us:.1f} us/launch  {4*B*H*S*S*64/us/1e6:.0f} TFLOP/s")
