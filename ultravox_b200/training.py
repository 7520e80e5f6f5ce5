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
from .losses import causal_lm_loss, causal_lm_loss_bwd
from .config import LossFunction
from .model import BF16, UltravoxModel


class AdapterTrainer:
    def __init__(self, model: UltravoxModel, lr: float = 2e-3, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 0.0, process_group=None):
        self.model = model
        self.lr, self.betas, self.eps, self.wd = lr, betas, eps, weight_decay
        self.pg = process_group
        pj = model.multi_modal_projector
        n = pj.flat.numel()
        dev = pj.flat.device
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)     # flat fp32 gradient (all-reduced)
        self.m = torch.zeros(n, dtype=torch.float32, device=dev)
        self.v = torch.zeros(n, dtype=torch.float32, device=dev)
        self.step_count = 0
        self._wT: Optional[list] = None
        self.last = {}

    # -- one-time: transposed copies of the frozen LLM weights (dgrad operands) ----------------------
    def _transposed_weights(self):
        if self._wT is None:
            lm = self.model.language_model
            self._wT = []
            for layer in lm.model.layers:
                self._wT.append(dict(qkv=ops.transpose(layer.self_attn.qkv_w), o=ops.transpose(layer.self_attn.o_proj.weight),
                                     gate_up=ops.transpose(layer.mlp.gate_up_w), down=ops.transpose(layer.mlp.down_proj.weight)))
            self._lm_head_T = ops.transpose(lm.lm_head.weight)
        return self._wT

    def grad_view(self, name: str) -> torch.Tensor:
        off, n, shape = self.model.multi_modal_projector.slices[name]
        return self.grad[off:off + n].view(*shape)

    # -- forward + backward --------------------------------------------------------------------------
    def forward_backward(self, input_ids, audio_values, audio_token_start_idx, audio_lens, audio_token_len,
                         audio_batch_size, labels, audio_tm: Optional[torch.Tensor] = None, alt_input_ids=None,
                         alt_labels=None, alt_attention_mask=None, attention_mask=None, audio_waveforms=None,
                         audio_num_frames=None, audio_pad_frames=None, **_) -> torch.Tensor:
        """Accumulates d(loss)/d(projector) into ``self.grad`` (zeroed first) and returns the loss (device scalar).
        Unpadded batches of equal length (the cfg3 synthetic workload); labels follow the HF convention."""
        m, cfg = self.model, self.model.config
        lm, tc, pj = m.language_model, m.config.text_config, m.multi_modal_projector
        dev = m.device
        wT = self._transposed_weights()
        input_ids = input_ids.to(dev)
        B, S = input_ids.shape
        nq, nkv, hd, Dm, ffn = tc.num_attention_heads, tc.num_key_value_heads, lm.head_dim, tc.hidden_size, tc.intermediate_size
        eps = tc.rms_norm_eps
        self.grad.zero_()

        # ---- forward: audio tower (frozen, nothing kept) + projector (kept)
        if attention_mask is not None and not bool(torch.as_tensor(attention_mask).to(torch.bool).all()):
            raise NotImplementedError("AdapterTrainer.forward_backward takes unpadded batches of equal length (cfg3 workload); "
                                      "use model(**batch).loss.backward() for padded batches")
        if audio_tm is None and audio_waveforms is not None:
            audio_tm = m.mel_chunks_from_waveforms(audio_waveforms, audio_num_frames, audio_pad_frames=audio_pad_frames)
        if audio_tm is None:
            audio_tm = ops.mel_to_timemajor(audio_values.to(dev, torch.float32))
        enc = m.encode_audio(audio_tm, audio_lens).clone()                 # [N, T2, d]
        N, T2, dE = enc.shape
        rows_a = (T2 + cfg.stack_factor - 1) // cfg.stack_factor
        xs = ops.stack_rmsnorm(enc, pj.ln_pre.weight, cfg.stack_factor, 1e-6)          # [N, rows_a, 8d]
        y1 = ops.linear(xs, pj.linear_1.weight)
        z = ops.swiglu(y1, gate_first=False)
        zn = ops.rmsnorm(z, pj.ln_mid.weight, 1e-6) if cfg.projector_ln_mid else z
        aud = ops.linear(zn, pj.linear_2.weight)                                        # [N, rows_a, D]
        if not cfg.projector_ln_mid:
            aud_pre = aud
            aud = ops.rmsnorm(aud_pre, pj.ln_post.weight, 1e-6)
        src = ops.splice_plan(audio_token_start_idx.to(dev, torch.int64).contiguous(),
                              audio_token_len.to(dev, torch.int32).contiguous(),
                              audio_batch_size.to(dev, torch.int64).reshape(-1).contiguous(), B, S, rows_a)
        h = ops.embed_splice(input_ids, lm.model.embed_tokens.weight, aud, src).view(B * S, Dm)

        # ---- forward: Llama with per-layer activations kept
        cos, sin = m._rope_tables(S)
        saved = []
        x = torch.empty_like(h)
        for layer in lm.model.layers:
            sa, mlp = layer.self_attn, layer.mlp
            h_in = h
            ops.rmsnorm(h_in, layer.input_layernorm.weight, eps, out=x)
            qkv = ops.linear(x, sa.qkv_w)
            ops.rope_(qkv, nq, nkv, hd, cos, sin, rows_per_seq=S)
            att = torch.empty(B * S, nq * hd, dtype=BF16, device=dev)
            lse = torch.empty(B * nq * S, dtype=torch.float32, device=dev)
            ops.attention_fused_qkv_train(qkv, B, S, nq, nkv, hd, hd ** -0.5, True, att, lse)
            h_mid = ops.linear(att, sa.o_proj.weight, residual=h_in)
            ops.rmsnorm(h_mid, layer.post_attention_layernorm.weight, eps, out=x)
            gu = ops.linear(x, mlp.gate_up_w)
            act = ops.swiglu(gu, gate_first=True)
            h = ops.linear(act, mlp.down_proj.weight, residual=h_mid)
            saved.append((h_in, qkv, att, lse, h_mid, gu))
        h_last = h
        hn = ops.rmsnorm(h_last, lm.model.norm.weight, eps)

        # ---- loss on the labelled rows only (row index computed on the host: labels come from the collator)
        lab = labels.to("cpu")
        shifted = torch.full_like(lab, cfg.ignore_index)
        shifted[:, :-1] = lab[:, 1:]
        rows = torch.nonzero(shifted.reshape(-1) != cfg.ignore_index).reshape(-1)
        tgt = shifted.reshape(-1)[rows].to(dev)
        rows_dev = rows.to(dev, torch.int32)
        hn_sel = ops.gather_rows(hn, rows_dev)
        logits = ops.linear(hn_sel, lm.lm_head.weight, out_dtype=torch.float32)        # [R, V]
        keep: dict = {}
        if m.loss_config.loss_function == LossFunction.KL_Divergence:
            # teacher = the same frozen LLM on the text-only twin of the sample (ref ultravox_model.py:202-226), no grad
            if alt_input_ids is None or alt_labels is None:
                raise ValueError("labels must be provided")
            from .losses import kl_distill_loss, kl_distill_loss_bwd, prediction_rows
            t_rows, _ = prediction_rows(alt_labels, cfg.ignore_index)
            _, is_eot = prediction_rows(labels, cfg.ignore_index)
            if t_rows.numel() != rows.numel():
                raise ValueError("student and teacher must predict the same number of tokens for the KL loss")
            alt_ids = alt_input_ids.to(dev)
            t_emb = ops.embed_splice(alt_ids, lm.model.embed_tokens.weight, None, None)
            t_hid = m.llama_hidden(t_emb).view(-1, Dm)
            t_logits = ops.linear(ops.gather_rows(t_hid, t_rows.to(dev, torch.int32)), lm.lm_head.weight, out_dtype=torch.float32)
            loss = kl_distill_loss(logits, t_logits, is_eot, m.loss_config.kl_temperature, m.loss_config.eot_loss_weight, keep=keep)
            dlogits = kl_distill_loss_bwd(keep)
        else:
            loss = causal_lm_loss(logits, tgt, cfg.ignore_index, keep=keep, shift=False)
            dlogits = causal_lm_loss_bwd(keep)                                         # [R, V] bf16

        # ---- backward through the head and the frozen LLM (data gradients only)
        d_hn = torch.zeros(B * S, Dm, dtype=BF16, device=dev)
        ops.linear(dlogits, self._lm_head_T, out=d_hn, row_map=rows_dev)
        dh = ops.rmsnorm_bwd(d_hn, h_last, lm.model.norm.weight, eps)
        for li in range(len(saved) - 1, -1, -1):
            layer = lm.model.layers[li]
            h_in, qkv, att, lse, h_mid, gu = saved[li]
            w = wT[li]
            d_act = ops.linear(dh, w["down"])                                          # [M, ffn]
            d_gu = ops.swiglu_bwd(gu, d_act, gate_first=True)
            dx2 = ops.linear(d_gu, w["gate_up"])                                       # [M, D]
            dh_mid = ops.rmsnorm_bwd(dx2, h_mid, layer.post_attention_layernorm.weight, eps, dres=dh)
            d_att = ops.linear(dh_mid, w["o"])
            dqkv = ops.attention_fused_qkv_bwd(qkv, att, d_att, lse, B, S, nq, nkv, hd, hd ** -0.5, True)
            ops.rope_bwd_(dqkv, nq, nkv, hd, cos, sin, rows_per_seq=S)
            dx1 = ops.linear(dqkv, w["qkv"])
            dh = ops.rmsnorm_bwd(dx1, h_in, layer.input_layernorm.weight, eps, dres=dh_mid)
            saved[li] = None

        # ---- splice backward: rows of d(inputs_embeds) at the audio positions
        inv = ops.splice_inverse(src, N * rows_a)
        d_aud = ops.gather_rows(dh, inv)                                               # [N*rows_a, D]

        # ---- projector backward (weight gradients in fp32 straight into the flat buffer)
        Ma = N * rows_a
        if not cfg.projector_ln_mid:
            d_aud = ops.rmsnorm_bwd(d_aud, aud_pre, pj.ln_post.weight, 1e-6, dw=self.grad_view("ln_post"))
        zn2, xs2 = zn.reshape(Ma, -1), xs.reshape(Ma, -1)
        ops.linear(ops.transpose(d_aud), ops.transpose(zn2), out=self.grad_view("linear_2"))       # dW2 = d_aud^T zn
        d_zn = ops.linear(d_aud, ops.transpose(pj.linear_2.weight))
        if cfg.projector_ln_mid:
            d_z = ops.rmsnorm_bwd(d_zn, z.reshape(Ma, -1), pj.ln_mid.weight, 1e-6, dw=self.grad_view("ln_mid"))
        else:
            d_z = d_zn
        d_y1 = ops.swiglu_bwd(y1.reshape(Ma, -1), d_z, gate_first=False)
        ops.linear(ops.transpose(d_y1), ops.transpose(xs2), out=self.grad_view("linear_1"))        # dW1 = d_y1^T xs
        d_xs = ops.linear(d_y1, ops.transpose(pj.linear_1.weight))
        ops.rmsnorm_bwd(d_xs, enc, pj.ln_pre.weight, 1e-6, want_dx=False, dw=self.grad_view("ln_pre"),
                        stack=(rows_a, T2 * dE))
        self.last = dict(loss=loss, rows=int(rows.numel()))
        return loss

    # -- exchange + update ---------------------------------------------------------------------------
    def all_reduce(self) -> float:
        """The single data-path collective: mean of the flat projector gradient over the data-parallel ranks
        (NCCL over NVLink / NVSwitch when launched with one process per GPU)."""
        from .dist_utils import allreduce_mean_
        allreduce_mean_(self.grad, self.pg)
        return 1.0

    def optimizer_step(self, grad_scale: float = 1.0):
        self.step_count += 1
        ops.adamw_(self.model.multi_modal_projector.flat, self.grad, self.m, self.v, self.step_count, self.lr, self.betas,
                   self.eps, self.wd, grad_scale)

    def train_step(self, **batch) -> torch.Tensor:
        loss = self.forward_backward(**batch)
        scale = self.all_reduce()
        self.optimizer_step(scale)
        return loss
