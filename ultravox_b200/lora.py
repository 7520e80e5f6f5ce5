"""LoRA checkpoints on the inference path (state-dict row of 8b).

The released >= v0.5 recipes train rank-8 LoRA adapters on the Whisper ``q_proj`` / ``k_proj`` (ref
ultravox_config.py:10-24 ``LoraConfigSimplified``; ``apply_lora`` -> ``peft.get_peft_model``, ref ultravox_model.py:690-709).
A checkpoint saved from such a model names the wrapped weights PEFT-style (ref training/model_types.py:300-333):

    audio_tower.base_model.model.layers.3.self_attn.q_proj.base_layer.weight     (frozen W  [out, in])
    audio_tower.base_model.model.layers.3.self_attn.q_proj.lora_A.default.weight (A [r, in])
    audio_tower.base_model.model.layers.3.self_attn.q_proj.lora_B.default.weight (B [out, r])

For inference the adapter is a rank-r update of W, so the B200 path folds it in at load time -
W' = W + (lora_alpha / r) * B @ A  (PEFT's default scaling; ``merge_and_unload`` in the reference's ``push_to_hub``,
ref ultravox_model.py:560-562) - and the fused q|k|v GEMM runs unchanged.  Training THROUGH the encoder adapters (encoder
backward) is not built.
"""
from __future__ import annotations

import re
from typing import Dict, Iterable, Mapping, Optional

import torch

_LORA_A = re.compile(r"^(?P<stem>.+)\.lora_A\.(?P<adapter>[^.]+)\.weight$")
_PEFT_ROOT = ".base_model.model."


def has_lora_keys(state_dict: Mapping[str, torch.Tensor]) -> bool:
    return any(".lora_A." in k or ".base_layer." in k or _PEFT_ROOT in k for k in state_dict)


def lora_scaling(lora_config: Optional[Mapping]) -> float:
    """``lora_alpha / r`` (0 when the component has no adapter)."""
    r = int((lora_config or {}).get("r", 0) or 0)
    return float((lora_config or {}).get("lora_alpha", 8)) / r if r > 0 else 0.0


def plain_name(key: str) -> str:
    """PEFT-wrapped parameter name -> the name the un-wrapped module uses."""
    return key.replace(_PEFT_ROOT, ".", 1).replace(".base_layer.", ".")


def merge_lora_state_dict(state_dict: Mapping[str, torch.Tensor], scaling: Mapping[str, float],
                          adapter: str = "default") -> Dict[str, torch.Tensor]:
    """Fold every ``lora_A/lora_B`` pair of ``adapter`` into its ``base_layer`` weight and strip the PEFT infixes.

    ``scaling`` maps a top-level component prefix (``"audio_tower"``, ``"language_model"``) to its ``lora_alpha / r``.
    The merge is done in fp32 and rounded once to the base weight's dtype."""
    out: Dict[str, torch.Tensor] = {}
    used = set()
    for key, a in state_dict.items():
        m = _LORA_A.match(key)
        if not m:
            continue
        if m["adapter"] != adapter:
            used.add(key)
            continue
        stem = m["stem"]
        kb, kw = f"{stem}.lora_B.{adapter}.weight", f"{stem}.base_layer.weight"
        if kb not in state_dict or kw not in state_dict:
            raise KeyError(f"incomplete LoRA triple for {stem}: need {kb} and {kw}")
        comp = stem.split(".", 1)[0]
        if comp not in scaling or scaling[comp] <= 0:
            raise ValueError(f"checkpoint has LoRA weights under '{comp}' but its lora config has r = 0")
        w, b = state_dict[kw], state_dict[kb]
        if a.shape[0] != b.shape[1] or b.shape[0] != w.shape[0] or a.shape[1] != w.shape[1]:
            raise ValueError(f"LoRA shapes do not fit {stem}: W {tuple(w.shape)}, A {tuple(a.shape)}, B {tuple(b.shape)}")
        merged = w.to(torch.float32) + scaling[comp] * (b.to(torch.float32) @ a.to(torch.float32))
        out[plain_name(kw)] = merged.to(w.dtype)
        used.update((key, kb, kw))
    for key, t in state_dict.items():
        if key in used or ".lora_A." in key or ".lora_B." in key:
            continue
        out[plain_name(key)] = t
    return out


def to_lora_names(state_dict: Mapping[str, torch.Tensor], prefix: str, target_modules: Iterable[str]) -> Dict[str, torch.Tensor]:
    """Plain names -> the names a PEFT-wrapped ``prefix`` component expects (what the reference does before loading a plain
    checkpoint into a LoRA model, ref training/model_types.py:300-333); keys that already carry the infix pass through."""
    targets = tuple(target_modules)
    out: Dict[str, torch.Tensor] = {}
    for key, t in state_dict.items():
        if not key.startswith(prefix + ".") or _PEFT_ROOT in key:
            out[key] = t
            continue
        new = prefix + _PEFT_ROOT + key[len(prefix) + 1:]
        for mod in targets:
            if f".{mod}." in new:
                new = new.replace(f".{mod}.", f".{mod}.base_layer.", 1)
                break
        out[new] = t
    return out
