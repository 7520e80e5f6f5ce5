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
