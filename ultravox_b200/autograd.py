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


def projector_backward(model, saved: dict, d_aud: torch.Tensor, grads: dict) -> None:
    """d_aud [N*rows_a, D] bf16 -> fp32 weight gradients WRITTEN into ``grads[name]`` (name in ln_pre / linear_1 / ln_mid |
    ln_post / linear_2; tensors of the parameter's shape, overwritten for the linears, accumulated for the norms - zero them
    first).  The encoder is frozen, so no gradient flows further back."""
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
    ops.rmsnorm_bwd(d_xs, saved["enc"], pj.ln_pre.weight, 1e-6, want_dx=False, dw=grads["ln_pre"], stack=(rows_a, T2 * dE))


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
        projector_backward(model, saved, d_aud.reshape(saved["N"] * saved["rows_a"], -1).contiguous(), grads)
        ctx.saved = None
        # parameters are bf16: autograd wants gradients in the parameter's dtype (the fp32 accumulation happened in the kernels)
        return (None, None) + tuple(grads[n].to(getattr(pj, n).weight.dtype) for n in names)


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
