"""Drop-in ``UltravoxConfig`` surface for the B200-native path.

Mirrors the public fields, defaults and (de)serialisation behaviour of
``ref:ultravox/model/ultravox_config.py:8-53`` (LoRA / loss dataclasses + enums) and ``:56-203``
(``UltravoxConfig``).  Only the fields that change the math matter to the CUDA path:
``stack_factor``, ``hidden_size``, ``projector_act``, ``projector_ln_mid``, ``norm_init``,
``audio_latency_block_size`` plus the two sub-configs; the rest are carried so that configs written
by the reference load unchanged.
"""
from __future__ import annotations

import dataclasses
import enum
from typing import Any, Optional

import transformers


@dataclasses.dataclass
class LoraConfigSimplified:
    """r == 0 means "freeze" (ref ultravox_model.py:690-709)."""
    r: int = 0
    lora_alpha: float = 8
    target_modules: Optional[list[str]] = dataclasses.field(
        default_factory=lambda: ["k_proj", "q_proj", "linear_k", "linear_q"])
    unfreeze_layers: Optional[list[str]] = None


class LossMaskType(str, enum.Enum):
    LAST_ASSISTANT = "last_assistant"
    ALL = "all"
    AFTER_AUDIO = "after_audio"


class LossFunction(str, enum.Enum):
    CrossEntropy = "ce"
    KL_Divergence = "kl"


@dataclasses.dataclass
class LossConfig:
    loss_function: LossFunction = LossFunction.CrossEntropy
    kl_temperature: float = 2.0
    initial_tokens_to_ignore: int = 0
    eot_loss_weight: float = 1.0

    @property
    def requires_alt_fields(self) -> bool:
        return self.loss_function == LossFunction.KL_Divergence


def _sub_config(cfg, model_id, default_type):
    if model_id is not None:
        return transformers.AutoConfig.from_pretrained(model_id)
    cfg = cfg or {}
    if isinstance(cfg, dict):
        cfg = transformers.CONFIG_MAPPING[cfg.get("model_type", default_type)](**cfg)
    return cfg


def _lora_dict(c):
    return c if isinstance(c, dict) else dataclasses.asdict(c or LoraConfigSimplified())


class UltravoxConfig(transformers.PretrainedConfig):
    model_type = "ultravox"
    is_composition = False

    def __init__(self, audio_config=None, text_config=None, audio_model_id: str | None = None,
                 text_model_id: str | None = None, llm_only_training: bool = False, ignore_index: int = -100,
                 audio_token_index: int | None = None, hidden_size: int = 4096, stack_factor: int = 8,
                 norm_init: float = 0.4, projector_act: str = "swiglu", projector_ln_mid: bool = False,
                 text_model_lora_config=None, audio_model_lora_config=None,
                 audio_latency_block_size: int | None = None, **kwargs):
        self.ignore_index = ignore_index
        self.audio_model_id, self.text_model_id = audio_model_id, text_model_id
        self.audio_token_index = audio_token_index
        self.hidden_size, self.stack_factor, self.norm_init = hidden_size, stack_factor, norm_init
        self.projector_act, self.projector_ln_mid = projector_act, projector_ln_mid
        self.text_config = _sub_config(text_config, text_model_id, "llama")
        self.audio_config = _sub_config(audio_config, audio_model_id, "whisper")
        self.llm_only_training = llm_only_training
        self.text_model_lora_config = _lora_dict(text_model_lora_config)
        self.audio_model_lora_config = _lora_dict(audio_model_lora_config)
        self.audio_latency_block_size = audio_latency_block_size
        tc = self.text_config
        if hasattr(tc, "text_config"):
            tc.vocab_size, tc.hidden_size = tc.text_config.vocab_size, tc.text_config.hidden_size
        self.vocab_size = tc.vocab_size
        self.initializer_range = tc.initializer_range
        super().__init__(**kwargs)

    def to_diff_dict(self) -> dict[str, Any]:
        d = super().to_diff_dict()
        for key, mid in (("text_config", self.text_model_id), ("audio_config", self.audio_model_id)):
            if mid is not None:
                d.pop(key, None)
            elif key in d:
                d[key].pop("_attn_implementation_autoset", None)
        return d


# ---------------------------------------------------------------------------------------------
# Named shape presets used by tests / bench (cfg1, cfg2, cfg4).
def _whisper(d, layers, heads, ffn, mels, name):
    return dict(model_type="whisper", d_model=d, encoder_layers=layers, encoder_attention_heads=heads,
                encoder_ffn_dim=ffn, num_mel_bins=mels, max_source_positions=1500, decoder_layers=0,
                _name_or_path=name)


def _llama(h, layers, heads, kv, ffn, hd, vocab=128256, tie=False, theta=500000.0, llama3=True, eps=1e-5):
    d = dict(model_type="llama", hidden_size=h, num_hidden_layers=layers, num_attention_heads=heads,
             num_key_value_heads=kv, intermediate_size=ffn, head_dim=hd, vocab_size=vocab,
             tie_word_embeddings=tie, rms_norm_eps=eps, max_position_embeddings=131072, rope_theta=theta)
    if llama3:
        d["rope_scaling"] = dict(rope_type="llama3", factor=8.0, low_freq_factor=1.0, high_freq_factor=4.0,
                                 original_max_position_embeddings=8192)
    return d


PRESETS = {
    # cfg1: Whisper-tiny + Llama-3.2-1B
    "tiny_1b": dict(audio_config=_whisper(384, 4, 6, 1536, 80, "openai/whisper-tiny"),
                    text_config=_llama(2048, 16, 32, 8, 8192, 64, tie=True), hidden_size=4096),
    # cfg2: Whisper-large-v3(-turbo) encoder + Llama-3.1-8B (Ultravox v0.5)
    "v0_5_8b": dict(audio_config=_whisper(1280, 32, 20, 5120, 128, "openai/whisper-large-v3-turbo"),
                    text_config=_llama(4096, 32, 32, 8, 14336, 128), hidden_size=4096),
    # cfg4: same encoder + Llama-3.3-70B
    "v0_5_70b": dict(audio_config=_whisper(1280, 32, 20, 5120, 128, "openai/whisper-large-v3-turbo"),
                     text_config=_llama(8192, 80, 64, 8, 28672, 128), hidden_size=4096),
    # unit-test size: every dimension small but structurally identical (GQA 4:1, llama3 rope, ln_mid)
    "micro": dict(audio_config=_whisper(128, 2, 2, 256, 80, "openai/whisper-micro"),
                  text_config=_llama(256, 2, 4, 2, 512, 64, vocab=1024), hidden_size=512),
}


def preset(name: str, **overrides) -> UltravoxConfig:
    kw = dict(PRESETS[name])
    kw.update(stack_factor=8, projector_act="swiglu", projector_ln_mid=True, norm_init=0.4)
    kw.update(overrides)
    return UltravoxConfig(**kw)
