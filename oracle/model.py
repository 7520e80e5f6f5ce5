"""Oracle (test infrastructure): fp32 CPU restatement of the Ultravox forward.

Functional torch (CPU, float32 unless told otherwise) over a flat state-dict that uses the
reference's parameter names.  Each function cites what it follows:

* ``whisper_encoder``  ref:ultravox/model/ultravox_model.py:865-994 (``ModifiedWhisperEncoder.forward``)
  + layer math of hf:models/whisper/modeling_whisper.py:241-414 (third-party)
* ``stack_frames``     ref:ultravox/model/ultravox_model.py:722-730
* ``rms_norm``         ref:ultravox/model/ultravox_model.py:733-736 / hf:models/llama/modeling_llama.py:53-67
* ``swiglu``           ref:ultravox/model/ultravox_model.py:739-742
* ``projector``        ref:ultravox/model/ultravox_model.py:768-800
* ``splice``           ref:ultravox/model/ultravox_model.py:259-275,390-394
* ``llama_forward``    hf:models/llama/modeling_llama.py:225-499, rope hf:modeling_rope_utils.py:550-626
* ``causal_lm_loss``   hf:loss/loss_utils.py:28-67
* ``forward``          ref:ultravox/model/ultravox_model.py:277-352

Float parity of the reference is unpinned by its own tests; this file is pinned in
``tests/test_oracle_cpu.py`` against the transformers modules themselves (WhisperEncoderLayer,
LlamaForCausalLM) and against fixtures produced by the reference's own projector / stack / splice
code (``tests/golden/projector_*.npz``, ``scripts/make_golden.py``).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Optional

import torch
import torch.nn.functional as F


@dataclasses.dataclass
class Shapes:
    # audio tower
    n_mels: int = 128
    enc_d: int = 1280
    enc_layers: int = 32
    enc_heads: int = 20
    enc_ffn: int = 5120
    enc_max_pos: int = 1500
    # projector
    stack: int = 8
    proj_hidden: int = 4096
    proj_ln_mid: bool = True
    proj_act: str = "swiglu"
    # llm
    d: int = 4096
    layers: int = 32
    heads: int = 32
    kv_heads: int = 8
    head_dim: int = 128
    ffn: int = 14336
    vocab: int = 128256
    rms_eps: float = 1e-5
    rope_theta: float = 500000.0
    rope_llama3: bool = True
    rope_factor: float = 8.0
    rope_low: float = 1.0
    rope_high: float = 4.0
    rope_orig_ctx: int = 8192
    tie_embeddings: bool = False
    latency_block: Optional[int] = None


# --------------------------------------------------------------------------- audio tower
def _finfo_min(dtype):
    return torch.finfo(dtype).min


def encoder_masks(audio_len: torch.Tensor, T2: int, dtype, latency_block: Optional[int], max_ctx: int = 3000):
    """Additive masks of ref :915-936.  Returns [N,1,1|T2,T2]."""
    feat_len = (audio_len.to(torch.int64) - 1) // 2 + 1            # hf _get_feat_extract_output_lengths
    keep = torch.arange(T2)[None, :] < feat_len.view(-1, 1)
    m = (1.0 - keep[:, None, None, :].to(dtype)) * _finfo_min(dtype)
    if latency_block is not None:
        assert max_ctx % latency_block == 0
        nb = max_ctx // latency_block
        blk = torch.tril(torch.ones(nb, nb)).repeat_interleave(latency_block, 0).repeat_interleave(latency_block, 1)
        blk = ((1.0 - blk) * _finfo_min(dtype))[None, None, :T2, :T2].to(dtype)
        m = torch.minimum(blk, m)
    return m


def whisper_layer(sd: dict, p: str, h: torch.Tensor, mask, n_heads: int) -> torch.Tensor:
    N, T, D = h.shape
    hd = D // n_heads
    x = F.layer_norm(h, (D,), sd[p + "self_attn_layer_norm.weight"], sd[p + "self_attn_layer_norm.bias"], 1e-5)
    q = F.linear(x, sd[p + "self_attn.q_proj.weight"], sd[p + "self_attn.q_proj.bias"]) * (hd ** -0.5)
    k = F.linear(x, sd[p + "self_attn.k_proj.weight"])
    v = F.linear(x, sd[p + "self_attn.v_proj.weight"], sd[p + "self_attn.v_proj.bias"])
    q = q.view(N, T, n_heads, hd).transpose(1, 2)
    k = k.view(N, T, n_heads, hd).transpose(1, 2)
    v = v.view(N, T, n_heads, hd).transpose(1, 2)
    s = q @ k.transpose(-1, -2)
    if mask is not None:
        s = s + mask
    a = torch.softmax(s, dim=-1) @ v
    a = a.transpose(1, 2).reshape(N, T, D)
    h = h + F.linear(a, sd[p + "self_attn.out_proj.weight"], sd[p + "self_attn.out_proj.bias"])
    x = F.layer_norm(h, (D,), sd[p + "final_layer_norm.weight"], sd[p + "final_layer_norm.bias"], 1e-5)
    x = F.gelu(F.linear(x, sd[p + "fc1.weight"], sd[p + "fc1.bias"]))
    x = F.linear(x, sd[p + "fc2.weight"], sd[p + "fc2.bias"])
    return h + x


def whisper_encoder(sd: dict, sh: Shapes, mel: torch.Tensor, audio_len: Optional[torch.Tensor],
                    prefix: str = "audio_tower.", collect: Optional[list] = None) -> torch.Tensor:
    """mel [N, n_mels, T<=3000] -> [N, ceil(T/2), enc_d]."""
    if mel.shape[-1] > sh.enc_max_pos * 2:
        raise ValueError("Whisper expects the mel input features to be of length "
                         f"{sh.enc_max_pos * 2} or less, but found {mel.shape[-1]}.")
    p = prefix
    h = F.gelu(F.conv1d(mel, sd[p + "conv1.weight"], sd[p + "conv1.bias"], padding=1))
    h = F.gelu(F.conv1d(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], stride=2, padding=1))
    h = h.permute(0, 2, 1)
    T2 = h.shape[1]
    h = h + sd[p + "embed_positions.weight"][:T2]
    if collect is not None:
        collect.append(h)
    mask = None
    if audio_len is not None:
        mask = encoder_masks(audio_len, T2, h.dtype, sh.latency_block, sh.enc_max_pos * 2)
    elif sh.latency_block is not None:
        mask = encoder_masks(torch.full((h.shape[0],), 2 * T2), T2, h.dtype, sh.latency_block, sh.enc_max_pos * 2)
    for i in range(sh.enc_layers):
        h = whisper_layer(sd, f"{p}layers.{i}.", h, mask, sh.enc_heads)
        if collect is not None:
            collect.append(h)
    return F.layer_norm(h, (sh.enc_d,), sd[p + "layer_norm.weight"], sd[p + "layer_norm.bias"], 1e-5)


# --------------------------------------------------------------------------- projector
def stack_frames(x: torch.Tensor, k: int) -> torch.Tensor:
    B, T, C = x.shape
    Tp = (T + k - 1) // k * k
    x = F.pad(x, (0, 0, 0, Tp - T))
    return x.reshape(B, Tp // k, C * k)


def rms_norm(x: torch.Tensor, w: torch.Tensor, eps: float) -> torch.Tensor:
    dt = x.dtype
    x32 = x.to(torch.float32)
    x32 = x32 * torch.rsqrt(x32.pow(2).mean(-1, keepdim=True) + eps)
    return w * x32.to(dt)


def swiglu(x: torch.Tensor) -> torch.Tensor:
    a, gate = x.chunk(2, dim=-1)
    return F.silu(gate) * a


def projector(sd: dict, sh: Shapes, enc_out: torch.Tensor, prefix: str = "multi_modal_projector.",
              collect: Optional[dict] = None) -> torch.Tensor:
    p = prefix
    x = stack_frames(enc_out, sh.stack)
    x = rms_norm(x, sd[p + "ln_pre.weight"], 1e-6)
    x = F.linear(x, sd[p + "linear_1.weight"])
    x = swiglu(x) if sh.proj_act == "swiglu" else F.gelu(x)
    if sh.proj_ln_mid:
        x = rms_norm(x, sd[p + "ln_mid.weight"], 1e-6)
    if collect is not None:
        collect["mid"] = x
    x = F.linear(x, sd[p + "linear_2.weight"])
    if not sh.proj_ln_mid:
        x = rms_norm(x, sd[p + "ln_post.weight"], 1e-6)
    return x


def splice(inputs_embeds: torch.Tensor, audio_embeds: torch.Tensor, start_idx, tok_len, audio_batch_size):
    """In-place, chunk after chunk in batch order (ref :259-275, :390-394)."""
    a = 0
    for b, cnt in enumerate(audio_batch_size.reshape(-1).tolist()):
        for _ in range(int(cnt)):
            s, n = int(start_idx[a]), int(tok_len[a])
            inputs_embeds[b][s: s + n] = audio_embeds[a][:n]
            a += 1
    return inputs_embeds


# --------------------------------------------------------------------------- llama
def rope_inv_freq(sh: Shapes) -> torch.Tensor:
    dim = sh.head_dim
    inv = 1.0 / (sh.rope_theta ** (torch.arange(0, dim, 2, dtype=torch.int64).to(torch.float32) / dim))
    if not sh.rope_llama3:
        return inv
    low_wl = sh.rope_orig_ctx / sh.rope_low
    high_wl = sh.rope_orig_ctx / sh.rope_high
    wl = 2 * math.pi / inv
    inv_l = torch.where(wl > low_wl, inv / sh.rope_factor, inv)
    smooth = (sh.rope_orig_ctx / wl - sh.rope_low) / (sh.rope_high - sh.rope_low)
    sm = (1 - smooth) * inv_l / sh.rope_factor + smooth * inv_l
    mid = ~(wl < high_wl) * ~(wl > low_wl)
    return torch.where(mid, sm, inv_l)


def rope_cos_sin(sh: Shapes, position_ids: torch.Tensor):
    freqs = position_ids[..., None].to(torch.float32) * rope_inv_freq(sh)[None, None, :]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def _rot_half(x):
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def llama_layer(sd, p, sh: Shapes, h, cos, sin, mask, kv_out: Optional[list] = None):
    B, S, D = h.shape
    x = rms_norm(h, sd[p + "input_layernorm.weight"], sh.rms_eps)
    q = F.linear(x, sd[p + "self_attn.q_proj.weight"]).view(B, S, sh.heads, sh.head_dim).transpose(1, 2)
    k = F.linear(x, sd[p + "self_attn.k_proj.weight"]).view(B, S, sh.kv_heads, sh.head_dim).transpose(1, 2)
    v = F.linear(x, sd[p + "self_attn.v_proj.weight"]).view(B, S, sh.kv_heads, sh.head_dim).transpose(1, 2)
    c, s_ = cos[:, None], sin[:, None]
    q = q * c + _rot_half(q) * s_
    k = k * c + _rot_half(k) * s_
    if kv_out is not None:
        kv_out.append((k, v))
    rep = sh.heads // sh.kv_heads
    k = k.repeat_interleave(rep, dim=1)
    v = v.repeat_interleave(rep, dim=1)
    sc = (q @ k.transpose(-1, -2)) * (sh.head_dim ** -0.5) + mask
    a = torch.softmax(sc.to(torch.float32), dim=-1).to(q.dtype) @ v
    a = a.transpose(1, 2).reshape(B, S, sh.heads * sh.head_dim)
    h = h + F.linear(a, sd[p + "self_attn.o_proj.weight"])
    x = rms_norm(h, sd[p + "post_attention_layernorm.weight"], sh.rms_eps)
    x = F.linear(F.silu(F.linear(x, sd[p + "mlp.gate_proj.weight"])) * F.linear(x, sd[p + "mlp.up_proj.weight"]),
                 sd[p + "mlp.down_proj.weight"])
    return h + x


def llama_forward(sd, sh: Shapes, inputs_embeds, attention_mask=None, position_ids=None,
                  prefix="language_model.", collect: Optional[list] = None, last_only: bool = False,
                  n_layers: Optional[int] = None):
    B, S, D = inputs_embeds.shape
    if position_ids is None:
        position_ids = torch.arange(S)[None, :].expand(B, S)
    cos, sin = rope_cos_sin(sh, position_ids)
    neg = _finfo_min(inputs_embeds.dtype)
    causal = torch.triu(torch.full((S, S), neg, dtype=inputs_embeds.dtype), diagonal=1)[None, None]
    mask = causal
    if attention_mask is not None:
        pad = (1.0 - attention_mask[:, None, None, :].to(inputs_embeds.dtype)) * neg
        mask = torch.minimum(causal.expand(B, 1, S, S), pad)
    h = inputs_embeds
    p = prefix + "model."
    for i in range(sh.layers if n_layers is None else n_layers):
        h = llama_layer(sd, f"{p}layers.{i}.", sh, h, cos, sin, mask)
        if collect is not None:
            collect.append(h)
    h = rms_norm(h, sd[p + "norm.weight"], sh.rms_eps)
    head = sd[p + "embed_tokens.weight"] if sh.tie_embeddings else sd[prefix + "lm_head.weight"]
    if last_only:
        h = h[:, -1:, :]
    return F.linear(h, head)


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100) -> torch.Tensor:
    lg = logits.to(torch.float32)
    lab = F.pad(labels, (0, 1), value=ignore_index)[..., 1:].contiguous()
    return F.cross_entropy(lg.view(-1, lg.shape[-1]), lab.view(-1), ignore_index=ignore_index, reduction="mean")


# --------------------------------------------------------------------------- end to end
def forward(sd, sh: Shapes, input_ids, audio_values=None, audio_token_start_idx=None, audio_lens=None,
            audio_token_len=None, audio_batch_size=None, attention_mask=None, labels=None,
            last_only: bool = False, stages: Optional[dict] = None):
    emb = sd["language_model.model.embed_tokens.weight"][input_ids].clone()
    if audio_values is not None and len(audio_values) > 0:
        enc = whisper_encoder(sd, sh, audio_values.to(emb.dtype), audio_lens)
        aud = projector(sd, sh, enc)
        if stages is not None:
            stages["encoder"], stages["projector"] = enc, aud
        emb = splice(emb, aud, audio_token_start_idx, audio_token_len, audio_batch_size)
    if stages is not None:
        stages["inputs_embeds"] = emb.clone()
    logits = llama_forward(sd, sh, emb, attention_mask, last_only=last_only)
    loss = causal_lm_loss(logits, labels) if labels is not None else None
    return logits, loss


# --------------------------------------------------------------------------- adapters (duck-typed; no product import)
def shapes_from_config(cfg) -> Shapes:
    """Build ``Shapes`` from an UltravoxConfig-like object (reference or ultravox_b200 - same field names)."""
    ac, tc = cfg.audio_config, cfg.text_config
    rp = getattr(tc, "rope_parameters", None) or {}
    sc = getattr(tc, "rope_scaling", None) or rp
    llama3 = bool(sc) and sc.get("rope_type", sc.get("type", "default")) == "llama3"
    theta = rp.get("rope_theta", None) or getattr(tc, "rope_theta", 10000.0)
    return Shapes(
        n_mels=ac.num_mel_bins, enc_d=ac.d_model, enc_layers=ac.encoder_layers, enc_heads=ac.encoder_attention_heads,
        enc_ffn=ac.encoder_ffn_dim, enc_max_pos=ac.max_source_positions, stack=cfg.stack_factor,
        proj_hidden=cfg.hidden_size, proj_ln_mid=cfg.projector_ln_mid, proj_act=cfg.projector_act, d=tc.hidden_size,
        layers=tc.num_hidden_layers, heads=tc.num_attention_heads, kv_heads=tc.num_key_value_heads,
        head_dim=getattr(tc, "head_dim", None) or tc.hidden_size // tc.num_attention_heads, ffn=tc.intermediate_size,
        vocab=tc.vocab_size, rms_eps=tc.rms_norm_eps, rope_theta=float(theta), rope_llama3=llama3,
        rope_factor=float(sc.get("factor", 8.0)) if llama3 else 8.0,
        rope_low=float(sc.get("low_freq_factor", 1.0)) if llama3 else 1.0,
        rope_high=float(sc.get("high_freq_factor", 4.0)) if llama3 else 4.0,
        rope_orig_ctx=int(sc.get("original_max_position_embeddings", 8192)) if llama3 else 8192,
        tie_embeddings=bool(getattr(tc, "tie_word_embeddings", False)),
        latency_block=getattr(cfg, "audio_latency_block_size", None))


def state_dict_fp32(module) -> dict:
    """fp32 CPU copy of a torch module's state dict (bit-identical values: bf16 -> fp32 is exact)."""
    return {k: v.detach().to("cpu", torch.float32) for k, v in module.state_dict().items()}
