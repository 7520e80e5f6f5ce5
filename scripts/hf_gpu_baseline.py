"""GPU *library* baseline: the path the reference actually reaches on a GPU - stock
transformers modules in bf16 (cuBLAS/cuBLASLt GEMMs + torch SDPA), eager PyTorch, same shapes / clip / prompt layout as
``bench.py``'s cfg2 workload, random-init weights.  What it executes, stage by stage (ref:ultravox/model/ultravox_model.py):

* ``WhisperFeatureExtractor`` on the host CPU (the reference's processor runs it in the DataLoader / caller process,
  ref:ultravox/model/ultravox_processing.py:295-303), then ``audio_values`` H2D;
* ``transformers.models.whisper.modeling_whisper.WhisperEncoder`` (what ``ModifiedWhisperEncoder`` subclasses, :803-994);
  a full 30 s clip has no padded keys, so no mask is passed (the cheapest case for the library);
* StackAudioFrames + RMSNorm + Linear + SwiGLU + RMSNorm + Linear (``UltravoxProjector``, :745-800) as plain torch ops;
* embedding lookup + the splice slice-assignment (:390-394);
* ``LlamaForCausalLM(inputs_embeds=..., logits_to_keep=1)`` (:328-334) + argmax.

None of this repo's kernels run here.  Reported: eager ms/clip (host-launch bound at batch 1), and the same device work
replayed from a CUDA graph (launch overhead removed - the strongest "library kernels" bar), each with and without the CPU mel.
"""
from __future__ import annotations

import time

import torch
import torch.nn.functional as F


class _Projector(torch.nn.Module):
    def __init__(self, d_in, hidden, d_out, stack, dtype, device):
        super().__init__()
        from transformers.models.llama.modeling_llama import LlamaRMSNorm
        self.stack = stack
        self.ln_pre = LlamaRMSNorm(d_in * stack, eps=1e-6).to(device, dtype)
        self.linear_1 = torch.nn.Linear(d_in * stack, hidden, bias=False, device=device, dtype=dtype)
        self.ln_mid = LlamaRMSNorm(hidden // 2, eps=1e-6).to(device, dtype)
        self.linear_2 = torch.nn.Linear(hidden // 2, d_out, bias=False, device=device, dtype=dtype)

    def forward(self, x):
        B, T, C = x.shape
        Tp = (T + self.stack - 1) // self.stack * self.stack
        x = F.pad(x, (0, 0, 0, Tp - T)).view(B, Tp // self.stack, C * self.stack)
        x = self.linear_1(self.ln_pre(x))
        a, gate = x.chunk(2, dim=-1)
        x = self.ln_mid(F.silu(gate) * a)
        return self.linear_2(x)


def build(cfg, device, attn="sdpa", dtype=torch.bfloat16):
    from transformers import LlamaForCausalLM
    from transformers.models.whisper.modeling_whisper import WhisperEncoder
    ac, tc = cfg.audio_config, cfg.text_config
    ac._attn_implementation = attn
    tc._attn_implementation = attn
    with torch.device(device):
        prev = torch.get_default_dtype()
        torch.set_default_dtype(dtype)
        try:
            enc = WhisperEncoder(ac).eval()
            llm = LlamaForCausalLM(tc).eval()
        finally:
            torch.set_default_dtype(prev)
    enc = enc.to(device, dtype)
    llm = llm.to(device, dtype)
    proj = _Projector(ac.d_model, cfg.hidden_size, tc.hidden_size, cfg.stack_factor, dtype, device).eval()
    return enc, proj, llm


@torch.no_grad()
def device_step(enc, proj, llm, audio_values, input_ids, start, n_tok):
    """audio_values [1, n_mels, 3000] bf16 on device -> next-token id tensor."""
    h = enc(audio_values).last_hidden_state
    aud = proj(h)
    emb = llm.get_input_embeddings()(input_ids)
    emb[0, start:start + n_tok] = aud[0, :n_tok]
    out = llm(inputs_embeds=emb, logits_to_keep=1, use_cache=False)
    return out.logits[:, -1].argmax(-1)


def load_weights(enc, proj, llm, state):
    """Copies an ``UltravoxModel``-named state dict (audio_tower.* / multi_modal_projector.* / language_model.*: the reference's
    names are transformers' own) into the stock modules, so both paths run on IDENTICAL weights."""
    def sub(prefix):
        return {k[len(prefix):]: v for k, v in state.items() if k.startswith(prefix)}
    enc.load_state_dict(sub("audio_tower."), strict=True)
    proj.load_state_dict(sub("multi_modal_projector."), strict=True)
    res = llm.load_state_dict(sub("language_model."), strict=False)
    assert not res.unexpected_keys and all("rotary" in k or "inv_freq" in k for k in res.missing_keys), res


def run(cfg, wl, device="cuda", iters=20, warmup=3, attn="sdpa", try_graph=True, state=None, check_mel=None):
    """Returns a dict for bench.py's ``gpu_library_baseline`` key.  ``state`` (optional): weights to run on (the measured model's
    own); ``check_mel`` [1, n_mels, T] (optional): the result also carries ``_check_logits`` = this path's last-row logits for
    that mel, for the side-by-side error against the fp32 oracle (parity is judged against HF's own bf16 error)."""
    import numpy as np
    from transformers import WhisperFeatureExtractor
    dev = torch.device(device)
    on_gpu = dev.type == "cuda"
    enc, proj, llm = build(cfg, dev, attn)
    if state is not None:
        load_weights(enc, proj, llm, state)
    fe = WhisperFeatureExtractor(feature_size=cfg.audio_config.num_mel_bins)
    wave = np.random.default_rng(1000).standard_normal(wl["n"]).astype(np.float32)
    ids = wl["input_ids"].to(dev)
    start, n_tok = int(wl["start"][0]), int(wl["n_tok"])

    def host_mel():
        f = fe(wave, sampling_rate=16000, padding="longest", pad_to_multiple_of=160, truncation=False, return_tensors="pt")
        return f["input_features"]

    def sync():
        if on_gpu:
            torch.cuda.synchronize(dev)

    mel = host_mel()
    T = mel.shape[-1]
    want = cfg.audio_config.max_source_positions * 2
    if T < want:                                   # stock WhisperEncoder insists on the full 3000-frame window
        mel = F.pad(mel, (0, want - T))
    av = mel.to(dev, torch.bfloat16)
    for _ in range(warmup):
        tok = device_step(enc, proj, llm, av, ids, start, n_tok)
    sync()
    # eager, device part only
    t = []
    for _ in range(iters):
        sync()
        t0 = time.perf_counter()
        tok = device_step(enc, proj, llm, av, ids, start, n_tok)
        sync()
        t.append(time.perf_counter() - t0)
    eager_ms = sorted(t)[len(t) // 2] * 1e3
    # eager, with the reference's CPU mel + H2D in the timed region (what a caller of the reference waits for)
    t = []
    for _ in range(max(3, iters // 4)):
        sync()
        t0 = time.perf_counter()
        m = host_mel()
        if m.shape[-1] < want:
            m = F.pad(m, (0, want - m.shape[-1]))
        tok = device_step(enc, proj, llm, m.to(dev, torch.bfloat16, non_blocking=True), ids, start, n_tok)
        int(tok[0])
        t.append(time.perf_counter() - t0)
    e2e_ms = sorted(t)[len(t) // 2] * 1e3
    out = {"what": "stock transformers WhisperEncoder + LlamaForCausalLM (bf16, attn=%s, cuBLAS + torch SDPA), random-init, "
                   "same clip / prompt layout; none of this repo's kernels" % attn,
           "eager_ms_device_part": eager_ms, "eager_ms_with_cpu_mel": e2e_ms, "iters": iters,
           "token": int(tok[0])}
    if on_gpu and try_graph:
        try:
            g = torch.cuda.CUDAGraph()
            static_av = av.clone()
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                for _ in range(2):
                    device_step(enc, proj, llm, static_av, ids, start, n_tok)
            torch.cuda.current_stream().wait_stream(s)
            with torch.cuda.graph(g):
                gtok = device_step(enc, proj, llm, static_av, ids, start, n_tok)
            for _ in range(3):
                g.replay()
            sync()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                g.replay()
            e1.record()
            sync()
            out["graph_ms_device_part"] = e0.elapsed_time(e1) / iters
            out["graph_token"] = int(gtok[0])
        except Exception as e:  # capture is best effort: the eager numbers stand on their own
            out["graph_error"] = repr(e)[:300]
    if check_mel is not None:
        with torch.no_grad():
            m = check_mel.to(dev, torch.bfloat16)
            if m.shape[-1] < want:
                m = F.pad(m, (0, want - m.shape[-1]))
            h = enc(m).last_hidden_state
            aud = proj(h)
            emb = llm.get_input_embeddings()(ids)
            emb[0, start:start + n_tok] = aud[0, :n_tok]
            out["_check_logits"] = llm(inputs_embeds=emb, logits_to_keep=1, use_cache=False).logits[:, -1].float().cpu().view(-1)
        out["weights"] = "the measured model's own (identical to the B200 path)" if state is not None else "own random init"
    secs = wl["n"] / 16000.0
    best = out.get("graph_ms_device_part", eager_ms)
    out["audio_sec_per_s_best"] = secs / (best * 1e-3)
    out["audio_sec_per_s_eager"] = secs / (eager_ms * 1e-3)
    del enc, proj, llm
    if on_gpu:
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import workload
    from ultravox_b200.config import preset
    name = sys.argv[1] if len(sys.argv) > 1 else "v0_5_8b"
    dev = "cuda" if torch.cuda.is_available() else "cpu"
    cfg = preset(name)
    print(json.dumps(run(cfg, workload(cfg, 30.0), dev, iters=20 if dev == "cuda" else 1, warmup=3 if dev == "cuda" else 1)))
