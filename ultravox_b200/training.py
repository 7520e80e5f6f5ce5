"""Adapter-only training step: encoder and LLM frozen, projector trained.

What the reference does with HF Trainer + autograd + DDP (``ref:ultravox/training/train.py:250-330``:
``model(**batch)`` -> ``loss.backward()`` -> DDP bucket all-reduce -> AdamW), written out explicitly over libuvx kernels:

  forward   encoder (no grad) -> projector (activations kept) -> splice -> Llama layers (per-layer inputs kept) ->
            final norm -> logits ONLY for rows that carry a label (the reference computes all rows) -> fp32 CE
  backward  CE -> lm_head dgrad (row-scattered) -> per layer {down, SwiGLU, gate/up, RMSNorm, o_proj, attention,
            RoPE, qkv, RMSNorm} data gradients against pre-transposed frozen weights (no weight gradients - the LLM
            is frozen, ref apply_lora r=0) -> gather at the audio positions -> projector dgrad + the four weight
            gradients (fp32, written straight into one flat buffer)
  exchange  ONE all-reduce (NCCL over NVLink / NVSwitch via torch.distributed) on the flat projector gradient,
            averaged over ranks - the only collective on the path
  update    one AdamW launch over the flat parameter buffer (fp32 moments, bf16 parameters)
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .model import BF16, UltravoxModel


class AdapterTrainer:
    def __init__(self, model: UltravoxModel, lr: float = 2e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, process_group=None, encoder_lora=None):
        """``encoder_lora`` (``autograd.EncoderLora``): also train LoRA adapters on the encoder's q / k projections - the
        ``audio_model_lora_config: {r: 8}`` of the released recipes (ref:ultravox/training/configs/v0.5_config.yaml:5-6); the
        encoder then runs its training forward (activations kept) and a full data-gradient backward."""
        self.model = model
        self.lora = encoder_lora
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.pg = process_group
        pj = model.multi_modal_projector
        n = pj.flat.numel()
        dev = pj.flat.device
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)     # flat fp32 gradient (all-reduced)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.lora_state = []
        if encoder_lora is not None:
            for prm, _ in encoder_lora.params_and_grads():
                self.lora_state.append((torch.zeros(prm.numel(), dtype=torch.float32, device=dev),
                                        torch.zeros(prm.numel(), dtype=torch.float32, device=dev)))
        self.step_count = 0
        self.last = {}
        self._comm_stream: Optional[torch.cuda.Stream] = None
        self._pending: list = []

    def grad_view(self, name: str) -> torch.Tensor:
        off, n, shape = self.model.multi_modal_projector.slices[name]
        return self.grad[off:off + n].view(*shape)

    # -- forward + backward --------------------------------------------------------------------------
    def forward_backward(self, input_ids, audio_values, audio_token_start_idx, audio_lens, audio_token_len,
                         audio_batch_size, labels, audio_tm: Optional[torch.Tensor] = None, alt_input_ids=None,
                         alt_labels=None, alt_attention_mask=None, attention_mask=None, audio_waveforms=None,
                         audio_num_frames=None, audio_pad_frames=None, **_) -> torch.Tensor:
        """Accumulates d(loss)/d(projector) into ``self.grad`` (zeroed first) and returns the loss (device scalar).  Drives the
        same forward / backward pieces as the autograd path (``autograd.py``) without building a graph; labels follow the HF
        convention; ``attention_mask`` may carry right padding."""
        from . import autograd as ag
        m, cfg = self.model, self.model.config
        lm, pj = m.language_model, m.multi_modal_projector
        dev = m.device
        input_ids = input_ids.to(dev)
        B, S = input_ids.shape
        Dm = m.config.text_config.hidden_size
        kv_start, kv_len = m._pad_bounds(attention_mask.to(dev) if attention_mask is not None else None)
        if kv_start is not None:
            raise NotImplementedError("training batches are right-padded (ref ultravox_processing.py:43-51); left padding is for generation")
        self.grad.zero_()
        with torch.no_grad():
            # ---- forward: audio tower (frozen, nothing kept) + projector (kept)
            if audio_tm is None and audio_waveforms is not None:
                audio_tm = m.mel_chunks_from_waveforms(audio_waveforms, audio_num_frames, audio_pad_frames=audio_pad_frames)
            if audio_tm is None:
                audio_tm = ops.mel_to_timemajor(audio_values.to(dev, torch.float32))
            sv_e = None
            if self.lora is not None:
                self.lora.merge_into(m)                                            # q|k|v weights <- base + s B A (this step's adapters)
                enc, sv_e = ag.encoder_forward_train(m, audio_tm, audio_lens)
            else:
                enc = m.encode_audio(audio_tm, audio_lens).clone()                 # [N, T2, d]
            aud, sv_p = ag.projector_forward(m, enc)
            src = ops.splice_plan(audio_token_start_idx.to(dev, torch.int64).contiguous(),
                                  audio_token_len.to(dev, torch.int32).contiguous(),
                                  audio_batch_size.to(dev, torch.int64).reshape(-1).contiguous(), B, S, sv_p["rows_a"])
            h = ops.embed_splice(input_ids, lm.model.embed_tokens.weight, aud, src).view(B * S, Dm)
            # ---- forward: Llama with per-layer activations kept, loss on the labelled rows
            hn, sv_l = ag.llama_stack_forward(m, h, B, S, kv_len)
            loss, keep = ag.head_loss_forward(m, hn, labels, alt_input_ids, alt_labels)
            # ---- backward: head -> frozen LLM (data gradients only) -> splice rows -> projector (fp32 weight gradients
            #      straight into the flat buffer)
            dh = ag.llama_stack_backward(m, sv_l, ag.head_loss_backward(m, keep))
            d_aud = ops.gather_rows(dh, ops.splice_inverse(src, sv_p["N"] * sv_p["rows_a"]))
            names = ag.projector_param_names(cfg)
            d_enc = ag.projector_backward(m, sv_p, d_aud, {n: self.grad_view(n) for n in names}, want_d_enc=sv_e is not None)
            if sv_e is not None:
                self.lora.zero_grad()
                ag.encoder_backward(m, sv_e, d_enc, self.lora)
        self.last = dict(loss=loss, rows=keep["n_rows"])
        return loss

    # -- exchange + update ---------------------------------------------------------------------------
    def all_reduce(self) -> float:
        """The single data-path collective: SUM of the flat projector gradient over the data-parallel ranks (NCCL over NVLink /
        NVSwitch when launched with one process per GPU).  Returns the scale (1 / world) the optimizer kernel applies - the
        mean is folded into ``uvx_adamw(grad_scale)`` instead of a separate pass over the 201 MB buffer."""
        from .dist_utils import allreduce_sum_
        if self.lora is not None:
            for _, g in self.lora.params_and_grads():
                allreduce_sum_(g, self.pg)
        return allreduce_sum_(self.grad, self.pg)

    def optimizer_step(self, grad_scale: float = 1.0):
        self.step_count += 1
        ops.adamw_(self.model.multi_modal_projector.flat, self.grad, self.m, self.v, self.step_count, self.lr, self.betas,
                   self.eps, self.wd, grad_scale)
        if self.lora is not None:
            for (prm, g), (m1, v1) in zip(self.lora.params_and_grads(), self.lora_state):
                ops.adamw_(prm.data.view(-1), g.view(-1), m1, v1, self.step_count, self.lr, self.betas, self.eps, self.wd, grad_scale)

    def train_step(self, **batch) -> torch.Tensor:
        loss = self.forward_backward(**batch)
        scale = self.all_reduce()
        self.optimizer_step(scale)
        return loss
