"""Host-side feed of the hot path: sample container + dataset -> model-features preprocessing.

Mirrors the reference interface so a training / eval script can switch imports:
  * ``VoiceSample``      ref:ultravox/data/data_sample.py:60-130  (messages + mono PCM; int16 / int32 / float64 -> float32)
  * ``Dataproc``         ref:ultravox/data/datasets.py:592-616    (iterable wrapper applying ``_process``)
  * ``UltravoxDataproc`` ref:ultravox/model/ultravox_data_proc.py:10-154 (chat template -> processor -> labels with the loss
                         mask -> optional text-only ``alt_*`` twin for the KL loss -> response truncation)
Only integer / string work happens here; with ``UltravoxProcessor(defer_mel=True)`` the log-mel stays on the GPU.
Parity: ``tests/test_host_cpu.py::test_dataproc_matches_reference_fixtures`` replays fixtures produced by the reference's
own ``UltravoxDataproc._process`` (``scripts/make_golden.py:dataproc_goldens``).
"""
from __future__ import annotations

import dataclasses
from typing import Any, Dict, Iterable, List, Optional

import numpy as np

from .config import LossMaskType

SAMPLE_RATE = 16000
AUDIO_PLACEHOLDER = "<|audio|>"


def messages_from_prompt(prompt: str) -> List[Dict[str, str]]:
    """A bare prompt is one user turn (ref data_sample.py:45-47)."""
    return [{"role": "user", "content": prompt}]


@dataclasses.dataclass
class VoiceSample:
    """One conversation with at most one audio clip.  ``audio`` is mono PCM at ``sample_rate``; integer and float64 inputs
    are normalised to float32 in [-1, 1) exactly like the reference (int16 / 2^15, int32 / 2^31)."""
    messages: List[Dict[str, str]]
    audio: Optional[np.ndarray] = None
    sample_rate: int = SAMPLE_RATE
    audio_transcript: Optional[str] = None
    label: Optional[str] = None
    extra_kwargs: Optional[Dict[str, Any]] = None

    def __post_init__(self):
        a = self.audio
        if a is None:
            return
        scale = {np.dtype(np.int16): 32768.0, np.dtype(np.int32): 2147483648.0}.get(a.dtype)
        if scale is not None:
            a = a.astype(np.float32) / np.float32(scale)
        elif a.dtype == np.float64:
            a = a.astype(np.float32)
        assert a.dtype == np.float32, f"Unexpected audio dtype: {a.dtype}"
        assert a.ndim == 1, f"Unexpected audio shape: {a.shape}"
        self.audio = a

    @staticmethod
    def from_prompt(prompt: str) -> "VoiceSample":
        return VoiceSample(messages_from_prompt(prompt), None)

    @staticmethod
    def from_prompt_and_raw(prompt: str, buf: np.ndarray, sample_rate: int) -> "VoiceSample":
        return VoiceSample(messages_from_prompt(prompt), buf, sample_rate)

    def add_past_messages(self, past_messages: List[Dict[str, str]]) -> None:
        self.messages = past_messages + self.messages


class Dataproc:
    """Iterable view of a dataset of ``VoiceSample`` with ``_process`` applied to every element."""

    def __init__(self, dataset: Iterable[VoiceSample]) -> None:
        self._dataset = dataset

    def _process(self, sample: VoiceSample) -> Dict[str, Any]:
        raise NotImplementedError

    def __iter__(self):
        for sample in self._dataset:
            yield self._process(sample)

    def __len__(self):
        return len(self._dataset)

    def __str__(self):
        return f"Dataproc({self._dataset})"

    @property
    def name(self):
        return getattr(self._dataset, "name", type(self._dataset).__name__)


class UltravoxDataproc(Dataproc):
    """Same constructor arguments and output keys as the reference class.

    Output of ``_process``: the processor's features (``input_ids``, ``attention_mask``, ``audio_values`` or deferred
    waveform, ``audio_lens``, ``audio_token_len``, ``audio_token_start_idx``, ...) with the batch dimension squeezed from
    ``input_ids`` / ``attention_mask``, plus ``labels`` (list, ``-100`` on the masked prefix) and, if
    ``include_alt_fields``, ``alt_input_ids`` / ``alt_attention_mask`` / ``alt_labels`` for the text-only teacher pass.
    """

    def __init__(self, dataset, processor, loss_mask_type: LossMaskType, augmentation=None, inference_mode: bool = False,
                 include_alt_fields: bool = False, max_response_tokens: Optional[int] = None,
                 chat_template: Optional[str] = None) -> None:
        super().__init__(dataset)
        self.processor = processor
        self.loss_mask_type = loss_mask_type
        self.augmentation = augmentation
        self.inference_mode = inference_mode
        self.include_alt_fields = include_alt_fields
        self.max_response_tokens = max_response_tokens
        self.chat_template = chat_template

    # -- helpers ------------------------------------------------------------------------------------
    def _render(self, messages) -> str:
        return self.processor.tokenizer.apply_chat_template(messages, tokenize=False, chat_template=self.chat_template)

    def _prompt_token_count(self, text: str, audio, sample_rate: int) -> int:
        return int(self.processor(text=text, audios=audio, sampling_rate=sample_rate)["input_ids"].shape[-1])

    def _compute_loss_mask_len(self, sample: VoiceSample, audio) -> int:
        """Number of leading tokens that carry no loss (ref :45-75): everything up to and including the audio
        placeholder (AFTER_AUDIO), everything before the last assistant turn (LAST_ASSISTANT), or nothing (ALL)."""
        kind = self.loss_mask_type
        if kind == LossMaskType.ALL:
            return 0
        if kind == LossMaskType.AFTER_AUDIO:
            head = self._render(sample.messages).split(AUDIO_PLACEHOLDER)[0] + AUDIO_PLACEHOLDER
            return self._prompt_token_count(head, audio, sample.sample_rate)
        if kind == LossMaskType.LAST_ASSISTANT:
            return self._prompt_token_count(self._render(sample.messages[:-1]), audio, sample.sample_rate)
        raise ValueError(f"Unsupported loss mask type: {kind}")

    # -- the per-sample transform -------------------------------------------------------------------
    def _process(self, sample: VoiceSample) -> Dict[str, Any]:
        if self.augmentation:
            sample = self.augmentation.apply_sample(sample)
        if self.inference_mode:
            sample.messages = sample.messages[:-1]          # the model generates the assistant turn itself
        text = self._render(sample.messages)
        audio = None if sample.audio is None else np.expand_dims(sample.audio, axis=0)   # [C=1, samples]
        inputs = self.processor(text=text, audios=audio, return_tensors="pt", sampling_rate=sample.sample_rate)
        input_ids = inputs["input_ids"].squeeze_(0)
        inputs["attention_mask"].squeeze_(0)

        labels = input_ids.clone()                          # the model shifts internally
        n_masked = self._compute_loss_mask_len(sample, audio)
        labels[:n_masked] = -100

        alt_masked = None
        if self.include_alt_fields:
            alt_text = text.replace(AUDIO_PLACEHOLDER, sample.audio_transcript or "")
            alt = self.processor(text=alt_text, audio=None, return_tensors="pt")
            alt_ids = alt["input_ids"].squeeze_(0)
            alt["attention_mask"].squeeze_(0)
            alt_masked = n_masked + len(alt_ids) - len(input_ids)     # the transcript replaces the audio tokens
            alt_labels = alt_ids.clone()
            alt_labels[:alt_masked] = -100
            inputs["alt_input_ids"] = alt_ids
            inputs["alt_attention_mask"] = alt["attention_mask"]
            inputs["alt_labels"] = alt_labels.tolist()

        cap = self.max_response_tokens
        if cap and n_masked + cap < len(input_ids):
            keep = n_masked + cap
            inputs["input_ids"] = inputs["input_ids"][:keep]
            inputs["attention_mask"] = inputs["attention_mask"][:keep]
            labels = labels[:keep]
            if self.include_alt_fields:
                alt_keep = alt_masked + cap
                for key in ("alt_input_ids", "alt_attention_mask", "alt_labels"):
                    inputs[key] = inputs[key][:alt_keep]
        return {**inputs, "labels": labels.tolist()}
