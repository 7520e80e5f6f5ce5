"""Prefill engine: the whole hot path (waveform -> log-mel -> Whisper encoder -> projector -> splice -> Llama prefill
-> last-position logits -> greedy token) recorded once into a CUDA graph and replayed per request.

This is the serving-side caller of ``UltravoxModel`` for fixed request shapes (what ``LocalInference.infer`` +
``model.generate(max_new_tokens=1)`` does in the reference, ref:ultravox/inference/infer.py:125-153,309-342), with the
host work (processor bookkeeping) hoisted out: per request only the waveform changes.  One engine per process / GPU;
multi-GPU inference is independent replicas - no collective.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import _lib, ops
from .model import UltravoxModel


class PrefillEngine:
    def __init__(self, model: UltravoxModel, clip_samples: int, input_ids: torch.Tensor,
                 audio_token_start_idx: torch.Tensor, audio_token_len: torch.Tensor, audio_batch_size: torch.Tensor,
                 n_clips: Optional[int] = None, use_graph: bool = True):
        """All index tensors follow the processor's output contract; one <=30 s chunk per clip."""
        self.model = model
        dev = model.device
        hop = 160
        L = -(-clip_samples // hop) * hop
        T = L // hop
        if T > model.audio_tower.max_context_length:
            raise ValueError("PrefillEngine handles clips of at most 30 s (one encoder chunk per clip)")
        N = int(n_clips if n_clips is not None else audio_token_start_idx.numel())
        self.N, self.L, self.T = N, L, T
        self.n_mels = model.audio_tower.n_mels
        self.wave = torch.zeros(N, L, dtype=torch.float32, device=dev)          # static input buffer
        self.input_ids = input_ids.to(dev).contiguous()
        self.start = audio_token_start_idx.to(dev, torch.int64).contiguous()
        self.tok_len = audio_token_len.to(dev, torch.int32).contiguous()
        self.abs = audio_batch_size.to(dev, torch.int64).reshape(-1).contiguous()
        frames = torch.full((N,), -(-clip_samples // hop), dtype=torch.int64)
        self.kv_len = ((frames - 1) // 2 + 1).to(torch.int32).to(dev)             # encoder key lengths
        self.audio_lens_host = frames
        self.token = torch.zeros(self.input_ids.shape[0], dtype=torch.int64, device=dev)
        self.logits = None
        self.graph = None
        self.launches_per_step = 0
        self._host_token = torch.zeros(self.input_ids.shape[0], dtype=torch.int64).pin_memory()
        # warm-up outside capture: builds device tables, sets func attributes, sizes the allocator pools
        for _ in range(2):
            self._step()
        torch.cuda.synchronize()
        if use_graph:
            before = _lib.launch_count()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._step()
            self.launches_per_step = _lib.launch_count() - before
            self.graph = g
        else:
            before = _lib.launch_count()
            self._step()
            self.launches_per_step = _lib.launch_count() - before
        torch.cuda.synchronize()

    # the hot path, in order (each call is one libuvx kernel or a short sequence of them)
    def _step(self):
        m = self.model
        tm = ops.logmel(self.wave, self.n_mels, want_f32=False, want_tm=True)
        enc = m.encode_audio(tm, None, kv_len=self.kv_len)
        aud = m.project_audio(enc)
        B, S = self.input_ids.shape
        src = ops.splice_plan(self.start, self.tok_len, self.abs, B, S, aud.shape[1])
        emb = ops.embed_splice(self.input_ids, m.language_model.model.embed_tokens.weight, aud, src)
        hidden = m.llama_hidden(emb)
        self.logits = ops.lm_head(hidden[:, -1, :], m.language_model.lm_head.weight)
        ops.argmax(self.logits, out=self.token)

    def run(self) -> torch.Tensor:
        """One prefill over whatever is in ``self.wave``; returns the device token tensor (no sync)."""
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.token

    def run_e2e(self, wave_host_pinned: torch.Tensor) -> torch.Tensor:
        """Host waveform (pinned fp32 [N, L]) in, host token out: H2D + prefill + D2H + sync."""
        self.wave.copy_(wave_host_pinned, non_blocking=True)
        self.run()
        self._host_token.copy_(self.token, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return self._host_token


class DecodeEngine:
    """Greedy decode after a prefill, one token per stream per step, the whole step (embedding -> 32/80 layers -> lm head
    -> argmax -> position bump) captured ONCE in a CUDA graph and replayed per token: positions, KV lengths and the
    current token live on the device, so nothing in the graph changes between steps.  Linear layers are weight-streaming
    matrix-vector kernels (uvx_gemv_bf16).  This is the serving loop of ``LocalInference._generate`` with
    ``temperature in {None, 0}`` (ref:ultravox/inference/infer.py:309-342)."""

    def __init__(self, model: UltravoxModel, batch: int, max_len: int, use_graph: bool = True):
        if batch > 8:
            raise ValueError("DecodeEngine handles up to 8 streams per GPU (uvx_gemv_bf16)")
        self.model, self.B, self.max_len = model, batch, max_len
        dev = model.device
        tc, lm = model.config.text_config, model.language_model
        self.cache = model.new_cache(batch, max_len)
        self.pos = torch.zeros(batch, dtype=torch.int32, device=dev)       # position of the token being fed
        self.lens = torch.zeros(batch, dtype=torch.int32, device=dev)      # keys visible to it (= pos + 1)
        self.token = torch.zeros(batch, 1, dtype=torch.int64, device=dev)
        self.cos, self.sin = model._rope_tables(max_len)
        self.graph = None
        self.use_graph = use_graph
        self.launches_per_step = 0

    def prefill(self, inputs_embeds: torch.Tensor) -> torch.Tensor:
        """Runs the prompt through the LLM, fills the cache, returns the first generated token [B]."""
        m = self.model
        B, S, _ = inputs_embeds.shape
        self.cache.length = 0
        hid = m.llama_hidden(inputs_embeds, self.cache)
        tok = ops.argmax(ops.lm_head(hid[:, -1, :], m.language_model.lm_head.weight))
        self.token.copy_(tok.view(B, 1))
        self.pos.fill_(S)
        self.lens.fill_(S + 1)
        return tok

    def _step(self):
        m = self.model
        lm, tc = m.language_model, m.config.text_config
        nq, nkv, hd, Dm = tc.num_attention_heads, tc.num_key_value_heads, lm.head_dim, tc.hidden_size
        B = self.B
        h = ops.embed_splice(self.token, lm.model.embed_tokens.weight, None, None).view(B, Dm)
        smax = self.cache.k.shape[2]
        for li, layer in enumerate(lm.model.layers):
            sa, mlp = layer.self_attn, layer.mlp
            x = ops.rmsnorm(h, layer.input_layernorm.weight, tc.rms_norm_eps)
            qkv = ops.gemv(x, sa.qkv_w)
            ops.rope_(qkv, nq, nkv, hd, self.cos, self.sin, rows_per_seq=1, positions=self.pos)
            kc, vc = self.cache.k[li], self.cache.v[li]
            ops.kv_append(qkv, kc, vc, self.pos, nq, nkv, hd)
            att = torch.empty(B, nq * hd, dtype=torch.bfloat16, device=h.device)
            rs = qkv.stride(0)
            ops.attention(qkv.data_ptr(), kc.data_ptr(), vc.data_ptr(), att, B, nq, nkv, 1, smax, hd,
                          (rs, rs, nkv * hd, smax * nkv * hd, nkv * hd, smax * nkv * hd, nq * hd, nq * hd), hd ** -0.5, False,
                          self.lens, 0)
            h = ops.gemv(att, sa.o_proj.weight, residual=h)
            x = ops.rmsnorm(h, layer.post_attention_layernorm.weight, tc.rms_norm_eps)
            act = ops.swiglu(ops.gemv(x, mlp.gate_up_w), gate_first=True)
            h = ops.gemv(act, mlp.down_proj.weight, residual=h)
        hn = ops.rmsnorm(h, lm.model.norm.weight, tc.rms_norm_eps)
        logits = ops.lm_head(hn, lm.lm_head.weight)
        ops.argmax(logits, out=self.token.view(-1))
        ops.add_i32_(self.pos, self.lens, 1)

    def step(self) -> torch.Tensor:
        """Feeds ``self.token`` (the previous output), writes the next token into it; returns the device tensor."""
        if self.use_graph and self.graph is None:
            self._step_warm()
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step()
        return self.token

    def _step_warm(self):
        # warm-up on a scratch copy of the state, then capture; the state is restored so no token is lost
        pos, lens, tok = self.pos.clone(), self.lens.clone(), self.token.clone()
        self._step()
        torch.cuda.synchronize()
        self.pos.copy_(pos), self.lens.copy_(lens), self.token.copy_(tok)
        before = _lib.launch_count()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self._step()
        self.launches_per_step = _lib.launch_count() - before
        self.pos.copy_(pos), self.lens.copy_(lens), self.token.copy_(tok)
        torch.cuda.synchronize()
        self.graph = g
