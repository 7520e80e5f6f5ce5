"""The adapter-training path as ``torch.autograd.Function``s over the libuvx kernels (north_star: "bound
through a thin C-ABI into PyTorch autograd.Functions").

What autograd does for the reference between ``model(**batch)`` and ``loss.backward()``
(``ref:ultravox/training/train.py:250-330``; frozen towers: ``apply_lora`` r=0, ``ref:ultravox/model/ultravox_model.py:690-709``)
is spelled out here as four nodes, each a forward over libuvx kernels that keeps exactly what its hand-written backward needs:

  ProjectorFn   StackAudioFrames + ln_pre + linear_1 + SwiGLU + ln_mid|ln_post + linear_2 (ref :768-800)
                backward: data gradient + the four WEIGHT gradients (fp32 accumulation, split-K-free wgrad GEMMs)
  SpliceFn      embedding gather + audio splice (ref :354-396); backward: gather of the gradient rows at the audio positions
  LlamaStackFn  all decoder layers + final RMSNorm with per-layer activations kept (hf:modeling_llama.py:292-426);
                backward: DATA gradients only, against pre-transposed frozen weights (no weight gradients - the LLM is frozen)
  HeadLossFn    lm_head on the labelled rows only + fp32 cross entropy (hf:loss/loss_utils.py:28-67) or the reference's KL
                distillation loss (ref :202-257); backward: d(loss)/d(hidden)

``UltravoxModel.forward`` composes them when gradients are enabled and a projector parameter requires grad, so
``model(**batch).loss.backward()`` - the door HF ``Trainer`` / DDP use - fills ``multi_modal_projector.*.grad``.
``training.AdapterTrainer`` drives the same ``*_forward`` / ``*_backward`` functions directly (fp32 gradients straight into
one flat buffer, no autograd graph): one implementation, two callers.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from .config import LossFunction

BF16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------ projector
def projector_forward(model, enc: torch.Tensor) -> tuple[torch.Tensor, dict]:
    """enc [N, T2, d] -> audio embeddings [N, rows_a, D]; ``saved`` holds the activations the backward reads."""
    pj, cfg = model.multi_modal_projector, model.config
    if cfg.projector_act != "swiglu":
        raise NotImplementedError(f"projector_act={cfg.projector_act!r}: only 'swiglu' (all released configs) is built")
    N, T2, dE = enc.shape
    rows_a = (T2 + cfg.stack_factor - 1) // cfg.stack_factor
    xs = ops.stack_rmsnorm(enc, pj.ln_pre.weight, cfg.stack_factor, 1e-6)          # [N, rows_a, 8d]
    y1 = ops.linear(xs, pj.linear_1.weight)
    z = ops.swiglu(y1, gate_first=False)
    zn = ops.rmsnorm(z, pj.ln_mid.weight, 1e-6) if cfg.projector_ln_mid else z
    aud = ops.linear(zn, pj.linear_2.weight)                                        # [N, rows_a, D]
    aud_pre = None
    if not cfg.projector_ln_mid:
        aud_pre = aud
        aud = ops.rmsnorm(aud_pre, pj.ln_post.weight, 1e-6)
    return aud, dict(enc=enc, xs=xs, y1=y1, z=z, zn=zn, aud_pre=aud_pre, N=N, T2=T2, dE=dE, rows_a=rows_a)


def projector_backward(model, saved: dict, d_aud: torch.Tensor, grads: dict, want_d_enc: bool = False) -> Optional[torch.Tensor]:
    """d_aud [N*rows_a, D] bf16 -> fp32 weight gradients WRITTEN into ``grads[name]`` (name in ln_pre / linear_1 / ln_mid |
    ln_post / linear_2; tensors of the parameter's shape, overwritten for the linears, accumulated for the norms - zero them
    first).  With a frozen encoder no gradient flows further back; ``want_d_enc`` (encoder LoRA training) also returns
    d(loss)/d(encoder output) [N, T2, dE]."""
    pj, cfg = model.multi_modal_projector, model.config
    N, rows_a, T2, dE = saved["N"], saved["rows_a"], saved["T2"], saved["dE"]
    Ma = N * rows_a
    if not cfg.projector_ln_mid:
        d_aud = ops.rmsnorm_bwd(d_aud, saved["aud_pre"], pj.ln_post.weight, 1e-6, dw=grads["ln_post"])
    zn2, xs2 = saved["zn"].reshape(Ma, -1), saved["xs"].reshape(Ma, -1)
    ops.linear(ops.transpose(d_aud), ops.transpose(zn2), out=grads["linear_2"])       # dW2 = d_aud^T zn   (fp32 out)
    d_zn = ops.linear(d_aud, ops.transpose(pj.linear_2.weight))
    if cfg.projector_ln_mid:
        d_z = ops.rmsnorm_bwd(d_zn, saved["z"].reshape(Ma, -1), pj.ln_mid.weight, 1e-6, dw=grads["ln_mid"])
    else:
        d_z = d_zn
    d_y1 = ops.swiglu_bwd(saved["y1"].reshape(Ma, -1), d_z, gate_first=False)
    ops.linear(ops.transpose(d_y1), ops.transpose(xs2), out=grads["linear_1"])        # dW1 = d_y1^T xs
    d_xs = ops.linear(d_y1, ops.transpose(pj.linear_1.weight))
    d_st = ops.rmsnorm_bwd(d_xs, saved["enc"], pj.ln_pre.weight, 1e-6, want_dx=want_d_enc, dw=grads["ln_pre"], stack=(rows_a, T2 * dE))
    if not want_d_enc:
        return None
    # the stacked rows are the encoder frames in order (StackAudioFrames pads the tail of the last row with zeros)
    return d_st.view(N, rows_a * cfg.stack_factor, dE)[:, :T2].contiguous()


def projector_param_names(cfg) -> list[str]:
    return ["ln_pre", "linear_1", "ln_mid" if cfg.projector_ln_mid else "ln_post", "linear_2"]


# ------------------------------------------------------------------------------------------------ llama stack
def transposed_llm_weights(model) -> list:
    """One-time [K, N] copies of the frozen LLM weights: the B operands of the data-gradient GEMMs (cached on the model)."""
    cached = getattr(model, "_wT", None)
    if cached is None:
        lm = model.language_model
        cached = []
        for layer in lm.model.layers:
            cached.append(dict(qkv=ops.transpose(layer.self_attn.qkv_w), o=ops.transpose(layer.self_attn.o_proj.weight),
                               gate_up=ops.transpose(layer.mlp.gate_up_w), down=ops.transpose(layer.mlp.down_proj.weight)))
        model._wT = cached
        model._lm_head_T = ops.transpose(lm.lm_head.weight)
    return cached


def llama_stack_forward(model, h: torch.Tensor, B: int, S: int, kv_len: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, dict]:
    """h [B*S, D] (residual stream, not modified) -> final-norm output [B*S, D]; keeps (h_in, qkv, att, lse, h_mid, gu) per layer.
    ``kv_len`` [B] int32: right-padded batches (the training collator pads on the right, ref ultravox_processing.py:43-51)."""
    lm, tc = model.language_model, model.config.text_config
    nq, nkv, hd = tc.num_attention_heads, tc.num_key_value_heads, lm.head_dim
    eps = tc.rms_norm_eps
    dev = h.device
    cos, sin = model._rope_tables(S)
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
        ops.attention_fused_qkv_train(qkv, B, S, nq, nkv, hd, hd ** -0.5, True, att, lse, kv_len)
        h_mid = ops.linear(att, sa.o_proj.weight, residual=h_in)
        ops.rmsnorm(h_mid, layer.post_attention_layernorm.weight, eps, out=x)
        gu = ops.linear(x, mlp.gate_up_w)
        act = ops.swiglu(gu, gate_first=True)
        h = ops.linear(act, mlp.down_proj.weight, residual=h_mid)
        saved.append((h_in, qkv, att, lse, h_mid, gu))
    hn = ops.rmsnorm(h, lm.model.norm.weight, eps)
    return hn, dict(layers=saved, h_last=h, B=B, S=S, cos=cos, sin=sin, kv_len=kv_len)


def llama_stack_backward(model, saved: dict, d_hn: torch.Tensor) -> torch.Tensor:
    """d(final-norm output) [B*S, D] bf16 -> d(inputs_embeds) [B*S, D]; frees the per-layer activations as it goes."""
    lm, tc = model.language_model, model.config.text_config
    nq, nkv, hd = tc.num_attention_heads, tc.num_key_value_heads, lm.head_dim
    eps = tc.rms_norm_eps
    B, S, cos, sin = saved["B"], saved["S"], saved["cos"], saved["sin"]
    wT = transposed_llm_weights(model)
    layers = saved["layers"]
    dh = ops.rmsnorm_bwd(d_hn, saved["h_last"], lm.model.norm.weight, eps)
    for li in range(len(layers) - 1, -1, -1):
        layer = lm.model.layers[li]
        h_in, qkv, att, lse, h_mid, gu = layers[li]
        w = wT[li]
        d_act = ops.linear(dh, w["down"])                                          # [M, ffn]
        d_gu = ops.swiglu_bwd(gu, d_act, gate_first=True)
        dx2 = ops.linear(d_gu, w["gate_up"])                                       # [M, D]
        dh_mid = ops.rmsnorm_bwd(dx2, h_mid, layer.post_attention_layernorm.weight, eps, dres=dh)
        d_att = ops.linear(dh_mid, w["o"])
        dqkv = ops.attention_fused_qkv_bwd(qkv, att, d_att, lse, B, S, nq, nkv, hd, hd ** -0.5, True, kv_len=saved["kv_len"])
        ops.rope_bwd_(dqkv, nq, nkv, hd, cos, sin, rows_per_seq=S)
        dx1 = ops.linear(dqkv, w["qkv"])
        dh = ops.rmsnorm_bwd(dx1, h_in, layer.input_layernorm.weight, eps, dres=dh_mid)
        layers[li] = None
    return dh


# ------------------------------------------------------------------------------------------------ head + loss
def head_loss_forward(model, hn: torch.Tensor, labels: torch.Tensor, alt_input_ids=None, alt_labels=None) -> tuple[torch.Tensor, dict]:
    """hn [B*S, D] -> scalar loss (device fp32).  Logits are computed ONLY for rows that carry a label (the reference computes
    every row; the loss is identical).  CE, or - when ``model.loss_config`` says so - the KL distillation loss against this
    same frozen LLM run on the text-only ``alt_*`` twin (ref ultravox_model.py:202-257)."""
    from .losses import causal_lm_loss, kl_distill_loss, prediction_rows
    cfg, lm = model.config, model.language_model
    dev = hn.device
    Dm = hn.shape[-1]
    lab = labels.to("cpu")
    shifted = torch.full_like(lab, cfg.ignore_index)
    shifted[:, :-1] = lab[:, 1:]
    rows = torch.nonzero(shifted.reshape(-1) != cfg.ignore_index).reshape(-1)
    tgt = shifted.reshape(-1)[rows].to(dev)
    rows_dev = rows.to(dev, torch.int32)
    logits = ops.linear(ops.gather_rows(hn, rows_dev), lm.lm_head.weight, out_dtype=torch.float32)        # [R, V]
    keep: dict = {}
    if model.loss_config.loss_function == LossFunction.KL_Divergence:
        if alt_input_ids is None or alt_labels is None:
            raise ValueError("labels must be provided")
        t_rows, _ = prediction_rows(alt_labels, cfg.ignore_index)
        _, is_eot = prediction_rows(labels, cfg.ignore_index)
        if t_rows.numel() != rows.numel():
            raise ValueError("student and teacher must predict the same number of tokens for the KL loss")
        with torch.no_grad():
            t_emb = ops.embed_splice(alt_input_ids.to(dev), lm.model.embed_tokens.weight, None, None)
            t_hid = model.llama_hidden(t_emb).view(-1, Dm)
            t_logits = ops.linear(ops.gather_rows(t_hid, t_rows.to(dev, torch.int32)), lm.lm_head.weight, out_dtype=torch.float32)
        loss = kl_distill_loss(logits, t_logits, is_eot, model.loss_config.kl_temperature, model.loss_config.eot_loss_weight, keep=keep)
        keep["kind"] = "kl"
    else:
        loss = causal_lm_loss(logits, tgt, cfg.ignore_index, keep=keep, shift=False)
        keep["kind"] = "ce"
    keep.update(rows_dev=rows_dev, n_rows=int(rows.numel()), shape=tuple(hn.shape))
    return loss, keep


def head_loss_backward(model, keep: dict, grad_scale: float = 1.0) -> torch.Tensor:
    """-> d(loss)/d(hn) [B*S, D] bf16 (zero on rows without a label)."""
    from .losses import causal_lm_loss_bwd, kl_distill_loss_bwd
    transposed_llm_weights(model)
    dlogits = kl_distill_loss_bwd(keep, grad_scale) if keep["kind"] == "kl" else causal_lm_loss_bwd(keep, grad_scale)
    rows, Dm = keep["shape"]
    d_hn = torch.zeros(rows, Dm, dtype=BF16, device=dlogits.device)
    ops.linear(dlogits, model._lm_head_T, out=d_hn, row_map=keep["rows_dev"])
    return d_hn


# ------------------------------------------------------------------------------------------------ autograd nodes
class ProjectorFn(torch.autograd.Function):
    """aud = UltravoxProjector(enc); the four projector weights are autograd inputs so their ``.grad`` is filled."""

    @staticmethod
    def forward(ctx, model, enc, *weights):
        aud, saved = projector_forward(model, enc)
        ctx.model, ctx.saved = model, saved
        return aud

    @staticmethod
    def backward(ctx, d_aud):
        model, saved = ctx.model, ctx.saved
        names = projector_param_names(model.config)
        pj = model.multi_modal_projector
        grads = {n: torch.zeros(getattr(pj, n).weight.shape, dtype=torch.float32, device=d_aud.device) for n in names}
        d_enc = projector_backward(model, saved, d_aud.reshape(saved["N"] * saved["rows_a"], -1).contiguous(), grads,
                                   want_d_enc=ctx.needs_input_grad[1])      # (encoder LoRA training: the encoder output has a grad_fn)
        ctx.saved = None
        # parameters are bf16: autograd wants gradients in the parameter's dtype (the fp32 accumulation happened in the kernels)
        return (None, d_enc) + tuple(grads[n].to(getattr(pj, n).weight.dtype) for n in names)


class SpliceFn(torch.autograd.Function):
    """inputs_embeds = splice(embed_tokens[input_ids], aud) - bit-exact copy forward, row gather backward."""

    @staticmethod
    def forward(ctx, model, aud, input_ids, src):
        ctx.src, ctx.n_rows, ctx.shape = src, aud.shape[0] * aud.shape[1], tuple(aud.shape)
        return ops.embed_splice(input_ids, model.language_model.model.embed_tokens.weight, aud, src)

    @staticmethod
    def backward(ctx, d_emb):
        inv = ops.splice_inverse(ctx.src, ctx.n_rows)
        d = d_emb.reshape(-1, d_emb.shape[-1])
        if d.dtype != BF16 or not d.is_contiguous():
            d = d.to(BF16).contiguous()
        return None, ops.gather_rows(d, inv).view(ctx.shape), None, None


class LlamaStackFn(torch.autograd.Function):
    """hidden = final_norm(decoder_layers(inputs_embeds)) for the frozen LLM: data gradients only."""

    @staticmethod
    def forward(ctx, model, inputs_embeds, kv_len):
        B, S, Dm = inputs_embeds.shape
        hn, saved = llama_stack_forward(model, inputs_embeds.reshape(B * S, Dm), B, S, kv_len)
        ctx.model, ctx.saved, ctx.shape = model, saved, (B, S, Dm)
        return hn.view(B, S, Dm)

    @staticmethod
    def backward(ctx, d_hn):
        B, S, Dm = ctx.shape
        d = d_hn.reshape(B * S, Dm)
        if d.dtype != BF16 or not d.is_contiguous():
            d = d.to(BF16).contiguous()
        dh = llama_stack_backward(ctx.model, ctx.saved, d)
        ctx.saved = None
        return None, dh.view(B, S, Dm), None


class HeadLossFn(torch.autograd.Function):
    """loss = CE | KL(lm_head(hidden[labelled rows]))."""

    @staticmethod
    def forward(ctx, model, hidden, labels, alt_input_ids, alt_labels):
        B, S, Dm = hidden.shape
        loss, keep = head_loss_forward(model, hidden.reshape(B * S, Dm), labels, alt_input_ids, alt_labels)
        ctx.model, ctx.keep, ctx.shape = model, keep, (B, S, Dm)
        return loss.clone()

    @staticmethod
    def backward(ctx, d_loss):
        # d_loss: 1.0 from loss.backward(), or the Trainer's loss scaling (1 / gradient_accumulation_steps): read once on the
        # host and folded into the loss-gradient kernel's scale
        d_hn = head_loss_backward(ctx.model, ctx.keep, float(d_loss) if d_loss is not None else 1.0)
        ctx.keep = None
        return None, d_hn.view(ctx.shape), None, None, None


def adapter_loss(model, input_ids, enc: torch.Tensor, src_args: tuple, labels, alt_input_ids=None, alt_labels=None,
                 kv_len: Optional[torch.Tensor] = None):
    """Composes the four nodes.  ``enc``: encoder output (no grad - the tower is frozen); ``src_args`` =
    (audio_token_start_idx, audio_token_len, audio_batch_size).  Returns (loss with a grad_fn, hidden [B, S, D] detached)."""
    pj = model.multi_modal_projector
    names = projector_param_names(model.config)
    weights = [getattr(pj, n).weight for n in names]
    aud = ProjectorFn.apply(model, enc, *weights)
    dev = enc.device
    B, S = input_ids.shape
    start, tok_len, abs_ = src_args
    src = ops.splice_plan(start.to(dev, torch.int64).contiguous(), tok_len.to(dev, torch.int32).contiguous(),
                          abs_.to(dev, torch.int64).reshape(-1).contiguous(), B, S, aud.shape[1])
    emb = SpliceFn.apply(model, aud, input_ids, src)
    hidden = LlamaStackFn.apply(model, emb, kv_len)
    loss = HeadLossFn.apply(model, hidden, labels, alt_input_ids, alt_labels)
    return loss, hidden.detach()


# ------------------------------------------------------------------------------------------------ encoder LoRA
class EncoderLora(torch.nn.Module):
    """LoRA adapters on ``q_proj`` / ``k_proj`` of every Whisper encoder layer - what ``audio_model_lora_config: {r: 8}`` of the
    released recipes turns on (ref:ultravox/training/configs/v0.5_config.yaml:5-6; defaults ref:ultravox/model/ultravox_config.py:10-24:
    lora_alpha 8, target modules k_proj / q_proj; ``peft.get_peft_model``, ref:ultravox/model/ultravox_model.py:690-709).

    Per layer the adapters live in GEMM-shaped (zero-padded to 64) bf16 buffers so that every product runs on uvx_gemm_bf16:
    ``A`` [64, d] (rows 0..r-1 = lora_A of q_proj, rows r..2r-1 = lora_A of k_proj), ``Bq`` / ``Bk`` [d, 64] (columns 0..r-1 /
    r..2r-1 = lora_B).  PEFT initialisation: A ~ kaiming-uniform(a = sqrt 5) = U(-1/sqrt d, 1/sqrt d), B = 0.  Gradients are fp32.
    ``merge_into(model)`` writes W + (alpha / r) B A into the fused q|k|v weight - the forward (training and inference) then
    runs the unchanged encoder kernels on the adapted weights, exactly PEFT's ``x W^T + s (x A^T) B^T``."""

    def __init__(self, model, r: int = 8, alpha: float = 8.0, seed: int = 0):
        super().__init__()
        at = model.audio_tower
        d, L = at.d, len(at.layers)
        if not (1 <= r <= 32):
            raise ValueError("EncoderLora: 1 <= r <= 32")
        self.r, self.scaling, self.d, self.L = int(r), float(alpha) / float(r), d, L
        dev = model.device
        g = torch.Generator().manual_seed(seed)
        bound = d ** -0.5
        A = torch.zeros(L, 64, d)
        A[:, :2 * r] = (torch.rand(L, 2 * r, d, generator=g) * 2 - 1) * bound
        self.A = torch.nn.Parameter(A.to(dev, BF16), requires_grad=False)
        self.Bq = torch.nn.Parameter(torch.zeros(L, d, 64, dtype=BF16, device=dev), requires_grad=False)
        self.Bk = torch.nn.Parameter(torch.zeros(L, d, 64, dtype=BF16, device=dev), requires_grad=False)
        # frozen base q / k rows of the fused q|k|v weights (the model's own tensors hold the merged weights)
        self.base_qk = [layer.self_attn.qkv_w[:2 * d].detach().clone() for layer in at.layers]
        self.gA = torch.zeros(L, 64, d, dtype=torch.float32, device=dev)
        self.gBq = torch.zeros(L, d, 64, dtype=torch.float32, device=dev)
        self.gBk = torch.zeros(L, d, 64, dtype=torch.float32, device=dev)

    def params_and_grads(self):
        return [(self.A, self.gA), (self.Bq, self.gBq), (self.Bk, self.gBk)]

    def zero_grad(self):
        for _, g in self.params_and_grads():
            g.zero_()

    @torch.no_grad()
    def merge_into(self, model) -> None:
        """qkv_w[:2d] <- base + s * B A for every layer (two K = 64 GEMMs per layer with the base rows as the residual)."""
        d = self.d
        for li, layer in enumerate(model.audio_tower.layers):
            At = ops.transpose(self.A[li])                                   # [d, 64]
            w = layer.self_attn.qkv_w
            ops.linear(self.Bq[li], At, residual=self.base_qk[li][:d], out=w[:d], alpha=self.scaling)
            ops.linear(self.Bk[li], At, residual=self.base_qk[li][d:], out=w[d:2 * d], alpha=self.scaling)

    @torch.no_grad()
    def unmerge(self, model) -> None:
        """Puts the frozen base q / k rows back (drops the adapters from the model's weights)."""
        for li, layer in enumerate(model.audio_tower.layers):
            layer.self_attn.qkv_w[:2 * self.d].copy_(self.base_qk[li])

    def peft_state_dict(self, prefix: str = "audio_tower.base_model.model.") -> dict:
        """The adapters under PEFT's names (what ``save_pretrained`` of the reference writes for a LoRA-wrapped tower)."""
        out, r = {}, self.r
        for li in range(self.L):
            base = f"{prefix}layers.{li}.self_attn."
            out[base + "q_proj.lora_A.default.weight"] = self.A[li, :r].detach().clone()
            out[base + "k_proj.lora_A.default.weight"] = self.A[li, r:2 * r].detach().clone()
            out[base + "q_proj.lora_B.default.weight"] = self.Bq[li, :, :r].detach().clone()
            out[base + "k_proj.lora_B.default.weight"] = self.Bk[li, :, r:2 * r].detach().clone()
        return out


def transposed_encoder_weights(model) -> list:
    """One-time [K, N] copies of the frozen encoder weights that the data-gradient GEMMs read (q|k|v is re-transposed every step:
    its q / k rows carry the adapters)."""
    cached = getattr(model, "_enc_wT", None)
    if cached is None:
        cached = [dict(o=ops.transpose(layer.self_attn.out_proj.weight), fc1=ops.transpose(layer.fc1.weight),
                       fc2=ops.transpose(layer.fc2.weight)) for layer in model.audio_tower.layers]
        model._enc_wT = cached
    return cached


def encoder_forward_train(model, x_tm: torch.Tensor, audio_lens: Optional[torch.Tensor],
                          kv_len: Optional[torch.Tensor] = None) -> tuple[torch.Tensor, dict]:
    """``UltravoxModel.encode_audio`` with the per-layer activations kept for ``encoder_backward`` (hf:modeling_whisper.py:403-440):
    residual stream before each LayerNorm, LN1 output (the adapters' input), q|k|v, attention output + log-sum-exp, fc1
    pre-activation.  The conv stem runs as in inference (no trainable parameter in front of layer 0)."""
    at = model.audio_tower
    N, Tp, _ = x_tm.shape
    T = Tp - 2
    if T > at.max_context_length:
        raise ValueError(f"Whisper expects the mel input features to be of length {at.max_context_length} or less, but found {T}.")
    d, H = at.d, at.heads
    hd = d // H
    T2 = (T + 1) // 2
    dev = x_tm.device
    h1 = torch.zeros(N, T + 2, d, dtype=BF16, device=dev)
    ops.conv1d_k3(x_tm, model._derived["conv1_w"], at.conv1.bias, 1, h1, out_guard=True)
    h = torch.empty(N, T2, d, dtype=BF16, device=dev)
    ops.conv1d_k3(h1, model._derived["conv2_w"], at.conv2.bias, 2, h, out_guard=False, pos=at.embed_positions.weight[:T2])
    if kv_len is None and audio_lens is not None:
        kv_len = ((audio_lens.to(torch.int64) - 1) // 2 + 1).to(torch.int32).to(dev)
    block = int(model.config.audio_latency_block_size or 0)
    if block:
        raise NotImplementedError("encoder LoRA training with the block-causal streaming mask")
    rows = N * T2
    h = h.view(rows, d)
    layers = []
    for layer in at.layers:
        sa = layer.self_attn
        h_in = h
        ln1 = ops.layernorm(h_in, layer.self_attn_layer_norm.weight, layer.self_attn_layer_norm.bias, 1e-5)
        qkv = ops.linear(ln1, sa.qkv_w, sa.qkv_b)
        att = torch.empty(rows, d, dtype=BF16, device=dev)
        lse = torch.empty(N, H, T2, dtype=torch.float32, device=dev)
        ops.attention_fused_qkv_train(qkv, N, T2, H, H, hd, hd ** -0.5, False, att, lse, kv_len)
        h_mid = ops.linear(att, sa.out_proj.weight, sa.out_proj.bias, residual=h_in)
        ln2 = ops.layernorm(h_mid, layer.final_layer_norm.weight, layer.final_layer_norm.bias, 1e-5)
        pre = ops.linear(ln2, layer.fc1.weight, layer.fc1.bias)
        h = ops.linear(ops.gelu(pre), layer.fc2.weight, layer.fc2.bias, residual=h_mid)
        layers.append(dict(h_in=h_in, ln1=ln1, qkv=qkv, att=att, lse=lse, h_mid=h_mid, pre=pre))
    out = ops.layernorm(h, at.layer_norm.weight, at.layer_norm.bias, 1e-5)
    return out.view(N, T2, d), dict(layers=layers, h_last=h, N=N, T2=T2, kv_len=kv_len)


def encoder_backward(model, saved: dict, d_enc: torch.Tensor, lora: EncoderLora) -> None:
    """d(loss)/d(encoder output) [N, T2, d] -> fp32 adapter gradients ACCUMULATED into ``lora.gA / gBq / gBk``.  Data gradients run
    back through every layer (final LayerNorm, fc2 / GELU / fc1, LayerNorm, out_proj, attention, q|k|v, LayerNorm, residuals)
    against pre-transposed weights; the base weights get no gradient (frozen, ref ``apply_lora``).  Per layer and projection:
    dB = s dq^T (ln1 A^T), dA = s (dq B)^T ln1 - low-rank first, so no [d, d] weight gradient is ever formed."""
    at = model.audio_tower
    d, H = at.d, at.heads
    hd = d // H
    N, T2 = saved["N"], saved["T2"]
    rows = N * T2
    r, s = lora.r, lora.scaling
    wT = transposed_encoder_weights(model)
    dh = ops.layernorm_bwd(d_enc.reshape(rows, d).contiguous(), saved["h_last"], at.layer_norm.weight, 1e-5)
    for li in range(len(at.layers) - 1, -1, -1):
        layer, sv = at.layers[li], saved["layers"][li]
        dg = ops.linear(dh, wT[li]["fc2"])                                                 # [rows, ffn]
        dpre = ops.gelu_bwd(sv["pre"], dg)
        dln2 = ops.linear(dpre, wT[li]["fc1"])
        dh_mid = ops.layernorm_bwd(dln2, sv["h_mid"], layer.final_layer_norm.weight, 1e-5, dres=dh)
        datt = ops.linear(dh_mid, wT[li]["o"])
        dqkv = ops.attention_fused_qkv_bwd(sv["qkv"], sv["att"], datt, sv["lse"], N, T2, H, H, hd, hd ** -0.5, False, kv_len=saved["kv_len"])
        # ---- adapter gradients (zero-padded rank dimension of 64: every product is a uvx_gemm_bf16 call)
        dq, dk = dqkv[:, :d], dqkv[:, d:2 * d]
        u = ops.linear(sv["ln1"], lora.A[li])                                              # [rows, 64]: (ln1 Aq^T | ln1 Ak^T | 0)
        uT = ops.transpose(u)                                                              # [64, rows']
        gq = torch.empty(d, 64, dtype=torch.float32, device=dh.device)
        gk = torch.empty(d, 64, dtype=torch.float32, device=dh.device)
        ops.linear(ops.transpose(dq), uT, out=gq)                                          # dq^T u
        ops.linear(ops.transpose(dk), uT, out=gk)
        lora.gBq[li, :, :r].add_(gq[:, :r], alpha=s)
        lora.gBk[li, :, r:2 * r].add_(gk[:, r:2 * r], alpha=s)
        t = ops.linear(dq, ops.transpose(lora.Bq[li]))                                     # [rows, 64]: dq Bq in columns 0..r-1
        t = ops.linear(dk, ops.transpose(lora.Bk[li]), residual=t)                         # + dk Bk in columns r..2r-1
        ga = torch.empty(64, d, dtype=torch.float32, device=dh.device)
        ops.linear(ops.transpose(t), ops.transpose(sv["ln1"]), out=ga)                     # t^T ln1
        lora.gA[li, :2 * r].add_(ga[:2 * r], alpha=s)
        # ---- back through the (adapted) q|k|v projection and the first LayerNorm
        dln1 = ops.linear(dqkv, ops.transpose(layer.self_attn.qkv_w))
        dh = ops.layernorm_bwd(dln1, sv["h_in"], layer.self_attn_layer_norm.weight, 1e-5, dres=dh_mid)


class EncoderLoraFn(torch.autograd.Function):
    """enc = WhisperEncoder(mel) with LoRA adapters on q / k: the adapters are autograd inputs, so ``loss.backward()`` fills their
    ``.grad`` (zero-padded rank dimension: only the first r / 2r rows / columns are ever non-zero)."""

    @staticmethod
    def forward(ctx, model, x_tm, audio_lens, A, Bq, Bk):
        lora = model.encoder_lora
        lora.merge_into(model)
        enc, saved = encoder_forward_train(model, x_tm, audio_lens)
        ctx.model, ctx.saved = model, saved
        return enc.clone()

    @staticmethod
    def backward(ctx, d_enc):
        model, lora = ctx.model, ctx.model.encoder_lora
        lora.zero_grad()
        d = d_enc if d_enc.dtype == BF16 else d_enc.to(BF16)
        encoder_backward(model, ctx.saved, d.contiguous(), lora)
        ctx.saved = None
        return None, None, None, lora.gA.to(lora.A.dtype), lora.gBq.to(lora.Bq.dtype), lora.gBk.to(lora.Bk.dtype)
