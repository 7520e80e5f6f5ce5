"""Cluster-shape sweep for uvx_gemm_bf16 (run under gpurun): tile config x split-K x (cm, cn) multicast cluster, graph-timed
over rotating weight copies (> L2), each variant checked against an fp32 reference."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ultravox_b200 import ops, _lib

lib = _lib.lib()
LLM_CL = [(1, 1), (1, 2), (1, 4)]
ENC_CL = [(1, 1), (2, 1), (1, 2), (2, 2), (4, 1), (1, 4)]
PLAN = [
    ("llm_gate_up", 201, 28672, 4096, [(2208, 1), (2256, 1), (4256, 1), (5416, 1), (5512, 1)], LLM_CL),
    ("llm_qkv", 201, 6144, 4096, [(1128, 1), (2128, 1), (2128, 2), (2064, 1), (2208, 1), (4128, 1), (4256, 1)], LLM_CL + [(2, 1), (2, 2)]),
    ("llm_o", 201, 4096, 4096, [(2128, 4), (2128, 2), (2064, 2), (2064, 1), (2256, 4)], LLM_CL),
    ("llm_down", 201, 4096, 14336, [(2128, 4), (2128, 2), (2064, 2), (2256, 4), (2256, 8)], LLM_CL),
    ("enc_qkv", 1500, 3840, 1280, [(1128, 1), (1256, 1), (4256, 1), (5512, 1)], ENC_CL),
    ("enc_out", 1500, 1280, 1280, [(1128, 1), (1064, 1), (4128, 1)], ENC_CL),
    ("enc_fc1", 1500, 5120, 1280, [(1128, 1), (1256, 1), (4256, 1), (5512, 1)], ENC_CL),
    ("enc_fc2", 1500, 1280, 5120, [(1128, 1), (1256, 1), (4128, 1), (4256, 1)], ENC_CL),
]
only = sys.argv[1:] or None
if os.environ.get("BASEONLY"):
    LLM_CL[:] = [(1, 1)]
    ENC_CL[:] = [(1, 1)]
res = []
for name, M, N, K, cfgs, clusters in PLAN:
    if only and not any(o in name for o in only):
        continue
    wbytes = N * K * 2
    copies = int(os.environ.get("COPIES", 0)) or max(2, min(8, int(400e6 // wbytes) + 1))
    Ws = [(torch.randn(N, K, device="cuda") * 0.05).bfloat16() for _ in range(copies)]
    x = torch.randn(M, K, device="cuda").bfloat16()
    ref = (x.float() @ Ws[0].float().T)
    out = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
    rows = []
    for cfg, sp in cfgs:
        for cm, cn in clusters:
            if (cfg // 1000 == 2 and cm > 1) or (cfg >= 4000 and (cm, cn) != (1, 1)):
                continue
            lib.uvx_debug_gemm_override(cfg, sp)
            lib.uvx_debug_gemm_cluster(cm, cn)
            try:
                ops.linear(x, Ws[0], out=out)
                torch.cuda.synchronize()
            except Exception as e:
                print(name, cfg, sp, cm, cn, "EXC", str(e)[:100], flush=True); continue
            err = ((out.float() - ref).norm() / ref.norm()).item()
            if err > 5e-3 or err != err:
                print(f"{name} cfg={cfg} sp={sp} cl={cm}x{cn} WRONG rel={err:.3e}", flush=True); continue
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for i in range(20):
                    ops.linear(x, Ws[i % copies], out=out)
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / 20
            r = dict(shape=name, cfg=cfg, splits=sp, cm=cm, cn=cn, us=round(us, 2), TFs=round(2.0 * M * N * K / us / 1e6, 1),
                     GBs=round((wbytes + M * K * 2 + M * N * 2) / us / 1e3, 1))
            rows.append(r); res.append(r)
            del g
    rows.sort(key=lambda r: r["us"])
    for r in rows[:6]:
        print("TOP ", r, flush=True)
    base = [r for r in rows if r["cm"] == 1 and r["cn"] == 1]
    if base:
        print("BASE", min(base, key=lambda r: r["us"]), flush=True)
    del Ws
    torch.cuda.empty_cache()
lib.uvx_debug_gemm_override(0, 0)
lib.uvx_debug_gemm_cluster(0, 0)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(res, open("gpurun_out/gemm_cluster_sweep.json", "w"), indent=0)
