"""Decode step of the 8B backbone, 1..8 concurrent streams: matrix-vector kernels vs the tcgen05 GEMM for the linears (run under
gpurun).  ms per CUDA-graphed step, CUDA events over 32 steps."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ultravox_b200.engine as eng_mod
from ultravox_b200.config import preset
from ultravox_b200.engine import DecodeEngine
from ultravox_b200.model import UltravoxModel

cfg = preset("v0_5_8b")
cfg.audio_config.encoder_layers = 1
model = UltravoxModel(cfg, device="cuda").init_random_(seed=42)
S = 201
emb1 = (torch.randn(1, S, cfg.text_config.hidden_size, generator=torch.Generator().manual_seed(0)) * 0.5).to(torch.bfloat16).cuda()
from ultravox_b200 import _lib
for maxb, ws in ((1, 2),):
    eng_mod.GEMV_MAX_B = maxb
    _lib.lib().uvx_debug_gemm_ws(ws, 0, 0)            # ws = 1: the opt-in weight-streaming GEMM (tokens on the UMMA N dimension: N = 16)
    for B in (1, 2, 4, 8):
        de = DecodeEngine(model, B, S + 40)
        de.prefill(emb1.expand(B, -1, -1).contiguous().clone())
        for _ in range(3):
            de.step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(32):
            de.step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 32
        print(json.dumps(dict(gemv_max_b=maxb, gemm_ws=ws, streams=B, ms_per_step=round(ms, 3), tok_per_s=round(B / ms * 1e3, 1), launches=de.launches_per_step)), flush=True)
        del de
        torch.cuda.empty_cache()
_lib.lib().uvx_debug_gemm_ws(-1, 0, 0)
