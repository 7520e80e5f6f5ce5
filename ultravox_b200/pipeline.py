"""``UltravoxPipeline`` - the reference's one-call wrapper (ref ultravox/model/ultravox_pipeline.py:15-133) over the B200 model.

Same call contract: ``pipe({"audio": ndarray, "turns": [...], "sampling_rate": 16000, "prompt": ...}, max_new_tokens=..,
temperature=.., repetition_penalty=..) -> str`` with the same four stages (``_sanitize_parameters``, ``preprocess``,
``_forward``, ``postprocess``), the same audio dtype normalisation, prompt / ``<|audio|>`` handling, terminators and default
repetition penalty (1.1).  It is a plain class rather than a ``transformers.Pipeline`` subclass: that base class loads and
moves ``PreTrainedModel`` instances, and the B200 model owns its device placement.  Tokenizer and processor are passed in
(the hub download the reference falls back to is out of scope here - no network).  Registered as ``"ultravox-pipeline"``.
"""
from __future__ import annotations

import logging
from typing import Any, Dict, List, Optional

import numpy as np
import torch

AUDIO_PLACEHOLDER = "<|audio|>"


class UltravoxPipeline:
    def __init__(self, model, tokenizer=None, audio_processor=None, chat_template: Optional[str] = None, processor=None,
                 **kwargs):
        if tokenizer is None:
            raise ValueError("pass the text tokenizer: the hub lookup of the reference pipeline needs network access")
        if chat_template:
            tokenizer.chat_template = chat_template
        self.model = model
        self.tokenizer = tokenizer
        if processor is None:
            from .processing import MelSpec, UltravoxProcessor
            n_mels = model.config.audio_config.num_mel_bins
            processor = UltravoxProcessor(audio_processor if audio_processor is not None else MelSpec(feature_size=n_mels), tokenizer,
                                          stack_factor=model.config.stack_factor,
                                          audio_context_size=model.audio_tower.max_context_length, mel_device=str(model.device))
        self.processor = processor

    def _sanitize_parameters(self, **kwargs):
        generation_keys = ["temperature", "max_new_tokens", "repetition_penalty"]
        return {}, {k: kwargs[k] for k in kwargs if k in generation_keys}, {}

    @staticmethod
    def _as_float_pcm(audio):
        """float64 -> float32, int16 / int32 -> float32 in [-1, 1); anything else is left to the processor to judge."""
        if not isinstance(audio, np.ndarray):
            return audio
        full_scale = {np.dtype(np.int16): 2.0 ** 15, np.dtype(np.int32): 2.0 ** 31}.get(audio.dtype)
        if full_scale is not None:
            return audio.astype(np.float32) / np.float32(full_scale)
        return audio.astype(np.float32) if audio.dtype == np.float64 else audio

    def preprocess(self, inputs: Dict[str, Any]):
        """``{"audio", "turns", "prompt", "sampling_rate"}`` -> processor features.  A clip without a trailing user turn gets one
        made from ``prompt`` (default: just the placeholder; the placeholder is appended when the prompt lacks it)."""
        turns: list = inputs.get("turns", [])
        audio = self._as_float_pcm(inputs.get("audio"))
        needs_user_turn = audio is not None and not (turns and turns[-1]["role"] == "user")
        if needs_user_turn:
            prompt = inputs.get("prompt", AUDIO_PLACEHOLDER)
            if AUDIO_PLACEHOLDER not in prompt:
                logging.warning("Prompt does not contain '<|audio|>', appending '<|audio|>' to the end of the prompt.")
                prompt = f"{prompt} {AUDIO_PLACEHOLDER}"
            turns.append({"role": "user", "content": prompt})
        if audio is not None and "sampling_rate" not in inputs:
            logging.warning("No sampling rate provided, using default of 16kHz. We highly recommend providing the correct "
                            "sampling rate.")
        rendered = self.processor.tokenizer.apply_chat_template(turns, add_generation_prompt=True, tokenize=False)
        return self.processor(text=rendered, audio=audio, sampling_rate=inputs.get("sampling_rate", 16000))

    def _forward(self, model_inputs: Dict[str, Any], temperature: Optional[float] = None,
                 max_new_tokens: Optional[int] = None, repetition_penalty: float = 1.1) -> List[int]:
        temperature = temperature or None
        terminators = [self.tokenizer.eos_token_id]
        if "<|eot_id|>" in getattr(self.tokenizer, "added_tokens_encoder", {}):
            terminators.append(self.tokenizer.convert_tokens_to_ids("<|eot_id|>"))
        input_len = model_inputs["input_ids"].shape[1]
        dev = self.model.device
        tensors = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in model_inputs.items()}
        outputs = self.model.generate(**tensors, do_sample=temperature is not None, temperature=temperature,
                                      max_new_tokens=max_new_tokens if max_new_tokens is not None else 20,
                                      repetition_penalty=repetition_penalty, eos_token_id=terminators)
        return outputs[0][input_len:]

    def postprocess(self, model_outputs) -> str:
        return self.tokenizer.decode(model_outputs, skip_special_tokens=True)

    def __call__(self, inputs: Dict[str, Any], **kwargs) -> str:
        _, forward_kwargs, _ = self._sanitize_parameters(**kwargs)
        with torch.no_grad():
            return self.postprocess(self._forward(self.preprocess(inputs), **forward_kwargs))


def _register_with_transformers() -> None:
    """ref ultravox_pipeline.py:128-133."""
    import transformers
    try:
        transformers.pipelines.PIPELINE_REGISTRY.register_pipeline("ultravox-pipeline", pipeline_class=UltravoxPipeline,
                                                                   pt_model=transformers.AutoModel, type="multimodal")
    except Exception as e:  # registry API drift must not break the import of the hot path
        logging.getLogger(__name__).warning("could not register 'ultravox-pipeline': %s", e)


_register_with_transformers()
