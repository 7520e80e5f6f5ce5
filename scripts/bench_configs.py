"""Secondary BASELINE.json configurations, each printing one JSON line on rank 0.

  cfg3  adapter-only training, bf16, per-GPU batch of 30 s clips, data-parallel, ONE gradient all-reduce per step
        python scripts/bench_configs.py train [--batch 32] [--steps 3] [--preset v0_5_8b]
        python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_configs.py train
  cfg4  30 s clip prefill + 128-token greedy decode, one stream per GPU (replicas), 70B-shaped backbone by default
        python scripts/bench_configs.py decode [--preset v0_5_70b] [--new-tokens 128]
  cfg5  log-mel + Whisper-encoder only throughput sweep (clip length x batch) with the tensor / HBM roofline fractions
        python scripts/bench_configs.py encoder [--secs 1,5,30] [--batches 1,8,64]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402


def setup_dist():
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, world, local


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "fallback": True}


def timed(fn, steps, warmup, world):
    import torch.distributed as dist
    for _ in range(warmup):
        fn()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms) / steps


def synth_batch(cfg, B, secs, rank, seed=7):
    n = int(16000 * secs)
    frames = -(-n // 160)
    n_tok = -(-frames // 16)
    g = torch.Generator().manual_seed(seed + rank)
    S = 8 + n_tok + 5
    ids = torch.randint(0, min(cfg.vocab_size, 128000), (B, S), generator=g)
    labels = ids.clone()
    labels[:, :-5] = -100                      # CE on the last 5 tokens
    waves = np.stack([np.random.default_rng(1000 + rank * 1000 + i).standard_normal(n).astype(np.float32) for i in range(B)])
    pad = (-n) % 160
    waves = np.pad(waves, ((0, 0), (0, pad)))
    return dict(waves=torch.from_numpy(waves), n=n, frames=frames, n_tok=n_tok, input_ids=ids, labels=labels,
                start=torch.full((B,), 8), lens=torch.full((B,), frames), tok=torch.full((B,), n_tok, dtype=torch.int32),
                abs=torch.ones(B, dtype=torch.int64))


def cmd_train(a):
    from ultravox_b200 import ops
    from ultravox_b200.config import preset
    from ultravox_b200.model import UltravoxModel
    from ultravox_b200.training import AdapterTrainer
    rank, world, local = setup_dist()
    cfg = preset(a.preset)
    model = UltravoxModel(cfg, device=f"cuda:{local}").init_random_(seed=42)
    tr = AdapterTrainer(model, lr=2e-3)
    sb = synth_batch(cfg, a.batch, a.secs, rank)
    waves = sb["waves"].cuda()

    def step():
        tm = ops.logmel(waves, cfg.audio_config.num_mel_bins, want_f32=False, want_tm=True)
        return tr.train_step(input_ids=sb["input_ids"], audio_values=None, audio_token_start_idx=sb["start"], audio_lens=sb["lens"],
                             audio_token_len=sb["tok"], audio_batch_size=sb["abs"], labels=sb["labels"], audio_tm=tm)
    loss0 = float(step())
    ms = timed(step, a.steps, 1, world)
    loss1 = float(tr.last["loss"])
    if rank == 0:
        clips = a.batch * world
        print(json.dumps({"config": "cfg3 adapter-only training", "preset": a.preset, "n_gpus": world, "per_gpu_batch": a.batch,
                          "clip_seconds": a.secs, "ms_per_step": ms, "clips_per_s": clips / (ms * 1e-3),
                          "audio_sec_per_s": clips * a.secs / (ms * 1e-3), "loss_first": loss0, "loss_last": loss1,
                          "collective": "1 all-reduce of %d fp32 gradient elements per step" % tr.grad.numel(),
                          "dtype": "bf16 (fp32 accumulate, fp32 grads/moments)", "data": "synthetic"}))


def cmd_decode(a):
    from ultravox_b200 import ops
    from ultravox_b200.config import preset
    from ultravox_b200.model import UltravoxModel
    rank, world, local = setup_dist()
    cfg = preset(a.preset)
    t0 = time.perf_counter()
    model = UltravoxModel(cfg, device=f"cuda:{local}").init_random_(seed=42)
    init_s = time.perf_counter() - t0
    sb = synth_batch(cfg, 1, a.secs, rank)
    waves = sb["waves"].cuda()
    lm = model.language_model
    res = {}

    from ultravox_b200.engine import DecodeEngine
    S_prompt = sb["input_ids"].shape[1]
    eng = DecodeEngine(model, 1, S_prompt + a.new_tokens + 2)

    def run():
        torch.cuda.synchronize()
        s0 = time.perf_counter()
        tm = ops.logmel(waves, cfg.audio_config.num_mel_bins, want_f32=False, want_tm=True)
        emb = model._prepare_audio_embeds(sb["input_ids"].cuda(), None, sb["start"], sb["lens"], sb["tok"], sb["abs"], audio_tm=tm)
        eng.prefill(emb)
        torch.cuda.synchronize()
        res["ttft"] = time.perf_counter() - s0
        for _ in range(a.new_tokens - 1):
            eng.step()
        torch.cuda.synchronize()
        res["total"] = time.perf_counter() - s0
    run()
    run()
    pk = peaks()
    wbytes = sum(p.numel() for n, p in lm.named_parameters() if "embed_tokens" not in n) * 2
    dec = (res["total"] - res["ttft"]) / max(1, a.new_tokens - 1)
    if rank == 0:
        print(json.dumps({"config": "cfg4 prefill + greedy decode, 1 stream per GPU (replicas)", "preset": a.preset, "n_gpus": world,
                          "clip_seconds": a.secs, "new_tokens": a.new_tokens, "ttft_ms": res["ttft"] * 1e3,
                          "decode_ms_per_token": dec * 1e3, "decode_tok_per_s_per_stream": 1.0 / dec,
                          "aggregate_audio_sec_per_s": world * a.secs / res["total"],
                          "decode_hbm_frac": wbytes / dec / 1e9 / pk["hbm_gbs"], "weights_gb": wbytes / 1e9, "init_s": init_s,
                          "note": "decode step = one CUDA-graph replay (%d libuvx launches), GEMV linears" % eng.launches_per_step, "data": "synthetic, random-init"}))


def cmd_encoder(a):
    from ultravox_b200 import ops
    from ultravox_b200.config import preset
    from ultravox_b200.model import UltravoxModel
    rank, world, local = setup_dist()
    cfg = preset(a.preset)
    cfg.text_config.num_hidden_layers = 1      # the LLM is not exercised here; keep it tiny
    model = UltravoxModel(cfg, device=f"cuda:{local}").init_random_(seed=42)
    ac = cfg.audio_config
    pk = peaks()
    cells = []
    for secs in [float(x) for x in a.secs_list.split(",")]:
        for B in [int(x) for x in a.batches.split(",")]:
            sb = synth_batch(cfg, B, secs, rank)
            waves = sb["waves"].cuda()
            kv = ((sb["lens"] - 1) // 2 + 1).to(torch.int32).cuda()

            def step():
                tm = ops.logmel(waves, ac.num_mel_bins, want_f32=False, want_tm=True)
                return model.encode_audio(tm, None, kv_len=kv)
            step()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                step()
            ms = timed(g.replay, a.steps, 2, 1)
            T2 = (sb["frames"] + 1) // 2
            d, f, L = ac.d_model, ac.encoder_ffn_dim, ac.encoder_layers
            flops = B * (L * (2 * T2 * d * (4 * d + 2 * f) + 4 * T2 * T2 * d) + 2 * sb["frames"] * 3 * ac.num_mel_bins * d
                         + 2 * T2 * 3 * d * d)
            cells.append({"secs": secs, "batch": B, "ms": ms, "audio_sec_per_s": B * secs / (ms * 1e-3),
                          "tflops": flops / (ms * 1e-3) / 1e12, "tensor_frac": flops / (ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"]})
            del g
    if rank == 0:
        print(json.dumps({"config": "cfg5 log-mel + Whisper encoder sweep", "preset": a.preset, "cells": cells,
                          "peak_tflops": pk["bf16_tflops_sustained"], "data": "synthetic"}))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest="cmd", required=True)
    t = sub.add_parser("train"); t.add_argument("--batch", type=int, default=32); t.add_argument("--steps", type=int, default=3)
    t.add_argument("--preset", default="v0_5_8b"); t.add_argument("--secs", type=float, default=30.0)
    d = sub.add_parser("decode"); d.add_argument("--preset", default="v0_5_70b"); d.add_argument("--new-tokens", type=int, default=128)
    d.add_argument("--secs", type=float, default=30.0)
    e = sub.add_parser("encoder"); e.add_argument("--preset", default="v0_5_8b"); e.add_argument("--steps", type=int, default=10)
    e.add_argument("--secs-list", default="1,5,30"); e.add_argument("--batches", default="1,8,32")
    args = ap.parse_args()
    {"train": cmd_train, "decode": cmd_decode, "encoder": cmd_encoder}[args.cmd](args)
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
        dist.destroy_process_group()
