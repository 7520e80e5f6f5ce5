"""Graph-timed stages of the cfg2 prefill (run under gpurun): log-mel, encoder, projector+splice, Llama stack, lm_head.
Each stage is captured in its own CUDA graph and replayed; between replays a 256 MB buffer is overwritten to flush L2, so
weights come from HBM as they do in the full step.  Prints one JSON line."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ultravox_b200 import ops
from ultravox_b200.config import preset
from ultravox_b200.model import UltravoxModel

cfg = preset("v0_5_8b")
dev = torch.device("cuda", 0)
model = UltravoxModel(cfg, device=dev).init_random_(seed=42)
ac, tc = cfg.audio_config, cfg.text_config
n = 480000
wave = torch.from_numpy(np.random.default_rng(1000).standard_normal(n).astype(np.float32))[None].to(dev)
kv = torch.tensor([1500], dtype=torch.int32, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
S = 201
emb = (torch.randn(1, S, tc.hidden_size, device=dev) * 0.02).bfloat16()


def timeit(fn, iters=20):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    tot = 0.0
    for _ in range(iters):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / iters


res = {}
res["mel_ms"] = timeit(lambda: ops.logmel(wave, ac.num_mel_bins, want_f32=False, want_tm=True))
tm = ops.logmel(wave, ac.num_mel_bins, want_f32=False, want_tm=True)
res["encoder_ms"] = timeit(lambda: model.encode_audio(tm, None, kv_len=kv))
enc = model.encode_audio(tm, None, kv_len=kv)
res["projector_ms"] = timeit(lambda: model.project_audio(enc))
res["llama_ms"] = timeit(lambda: model.llama_hidden(emb.clone()))
hid = model.llama_hidden(emb.clone())
res["lm_head_ms"] = timeit(lambda: ops.argmax(ops.lm_head(hid[:, -1], model.language_model.lm_head.weight)))
res["sum_ms"] = sum(v for v in res.values())
res["fuse_norm"] = os.environ.get("UVX_FUSE_NORM", "1")
print(json.dumps(res))
