"""Host side of the hot path: text + waveforms -> model inputs (drop-in for the reference processor).

Mirrors the call surface, output keys/dtypes and error behaviour of
``ref:ultravox/model/ultravox_processing.py``:

* ``UltravoxProcessor.__call__``             (:217-370)
* ``UltravoxProcessor._chunk_and_pad_audio`` (:153-215)
* ``DataCollatorForSeq2SeqWithAudio``        (:12-64)

What is different, B200-first: the log-mel front end (third-party ``WhisperFeatureExtractor`` on the
host CPU in the reference, ``:295-303``) runs on the GPU through ``libuvx`` (``uvx_logmel_*``), either
right here (``audio_values`` comes back as a CUDA tensor with the reference's exact layout
``[N, n_mels, T]`` fp32) or - with ``defer_mel=True`` - inside the model, in which case the processor
hands over the zero-padded waveforms (``audio_waveforms`` ``[B, L]``) and never touches CUDA (safe in
forked DataLoader workers).  All integer bookkeeping (frame counts, 30 s chunking, ``audio_token_len``,
placeholder expansion, start indices) is computed arithmetically on the host and is bit-exact with the
reference.  There is no CPU mel fallback.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Any, Optional, Sequence, Union

import numpy as np
import torch
import torch.nn.functional as F
import transformers

from .config import UltravoxConfig


@dataclasses.dataclass
class MelSpec:
    """The numbers of ``WhisperFeatureExtractor`` the path depends on."""
    feature_size: int = 80
    hop_length: int = 160
    n_fft: int = 400
    sampling_rate: int = 16000

    @classmethod
    def of(cls, audio_processor) -> "MelSpec":
        if audio_processor is None:
            return cls()
        if isinstance(audio_processor, MelSpec):
            return audio_processor
        fe = getattr(audio_processor, "feature_extractor", audio_processor)
        return cls(int(fe.feature_size), int(fe.hop_length), int(fe.n_fft), int(fe.sampling_rate))


def frame_chunks(frame_lens: Sequence[int], context: int):
    """Chunk plan of ref :172-199: per chunk (clip, frame offset, valid frames, is_continuation)."""
    plan, per_clip = [], []
    for i, n in enumerate(frame_lens):
        n = int(n)
        per_clip.append(int(math.ceil(n / context)))
        for off in range(0, n, context):
            plan.append((i, off, min(n - off, context), off > 0))
    return plan, per_clip


@dataclasses.dataclass
class DataCollatorForSeq2SeqWithAudio(transformers.DataCollatorForSeq2Seq):
    """ref :12-64.  Audio lists are flattened across samples, the mel is right-padded on time to the
    batch maximum, and start indices move right by the amount of left padding each sample received."""
    include_alt_fields: bool = False

    def __call__(self, features, *args, **kwargs):
        def flat(key):
            return [x for f in features for x in f.pop(key, [])]
        # deferred-mel samples (``UltravoxProcessor(defer_mel=True)``, the mode that is safe in forked DataLoader workers) carry
        # zero-padded waveforms instead of a mel: clips are flattened across samples and padded (as waveforms, to a multiple
        # of the hop) to the batch-longest; ``audio_pad_frames`` remembers each clip's own padded width so the model can
        # reproduce the reference's padding content frame for frame (see UltravoxModel.mel_chunks_from_waveforms)
        waves = [w for f in features for w in f.pop("audio_waveforms", [])]
        n_frames = [n for f in features for n in f.pop("audio_num_frames", [])]
        vals, lens = flat("audio_values"), flat("audio_lens")
        tok_len, starts = flat("audio_token_len"), flat("audio_token_start_idx")
        alt = None
        if self.include_alt_fields:
            alt = [{"input_ids": f.pop("alt_input_ids"), "attention_mask": f.pop("alt_attention_mask"),
                    "labels": f.pop("alt_labels")} for f in features]
        batch = super().__call__(features, *args, **kwargs)
        if alt is not None:
            ab = super().__call__(alt, *args, **kwargs)
            for k in ("input_ids", "attention_mask", "labels"):
                batch["alt_" + k] = ab[k]
        has_mel = bool(vals) and len(vals[0]) > 0
        if has_mel or waves:
            batch["audio_token_start_idx"] = torch.stack(starts)
            batch["audio_lens"] = torch.stack(lens)
            batch["audio_token_len"] = torch.stack(tok_len)
            if has_mel:
                width = max(v.shape[-1] for v in vals)
                batch["audio_values"] = torch.stack([F.pad(v, (0, width - v.shape[-1])) for v in vals])
            else:
                waves = [torch.as_tensor(w) for w in waves]
                width = max(w.shape[-1] for w in waves)
                batch["audio_waveforms"] = torch.stack([F.pad(w, (0, width - w.shape[-1])) for w in waves])
                batch["audio_num_frames"] = torch.stack([torch.as_tensor(n) for n in n_frames]).to(torch.int64)
                batch["audio_pad_frames"] = torch.tensor([w.shape[-1] // 160 for w in waves], dtype=torch.int64)
            if self.tokenizer.padding_side == "left":
                own = torch.LongTensor([f["input_ids"].shape[-1] for f in features])
                shift = (batch["input_ids"].shape[-1] - own).repeat_interleave(
                    batch["audio_batch_size"].squeeze(-1))
                batch["audio_token_start_idx"] += shift.to(batch["audio_token_start_idx"].device)
        return batch


class UltravoxProcessor:
    """Same constructor and ``__call__`` contract as the reference class (:67-382)."""

    attributes = ["audio_processor", "tokenizer"]

    def __init__(self, audio_processor=None, tokenizer=None, audio_padding: str = "longest",
                 encoder_ds_factor: int = 2, stack_factor: int = 8, audio_placeholder: str = "<|audio|>",
                 audio_context_size: Optional[int] = 3000, mel_device: Union[str, torch.device] = "cuda",
                 defer_mel: bool = False):
        self.audio_padding = audio_padding
        self.encoder_ds_factor = encoder_ds_factor
        self.stack_factor = stack_factor
        self.audio_placeholder = audio_placeholder
        self.audio_context_size = audio_context_size
        assert tokenizer.eos_token is not None, "The tokenizer has no EOS token. Cannot recover."
        self.vocab = tokenizer.get_vocab()
        self.audio_token_replacement = tokenizer.eos_token
        if tokenizer.pad_token_id is None:
            tokenizer.pad_token_id = tokenizer.eos_token_id
        self.audio_processor = audio_processor
        self.tokenizer = tokenizer
        self.mel_spec = MelSpec.of(audio_processor)
        self.mel_device = mel_device
        self.defer_mel = defer_mel

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, **kwargs):
        config: UltravoxConfig = transformers.AutoConfig.from_pretrained(pretrained_model_name_or_path, **kwargs)
        audio_processor = transformers.AutoProcessor.from_pretrained(
            config.audio_model_id or config.audio_config._name_or_path or "openai/whisper-tiny")
        tokenizer = transformers.AutoTokenizer.from_pretrained(pretrained_model_name_or_path, **kwargs)
        tokenizer.padding_side = "left"
        tokenizer.pad_token = tokenizer.eos_token
        return cls(audio_processor=audio_processor, tokenizer=tokenizer, stack_factor=config.stack_factor)

    # -- audio ---------------------------------------------------------------------------------
    def pad_waveforms(self, audios: list[np.ndarray]) -> tuple[torch.Tensor, list[int]]:
        """Zero-pad like the feature extractor (``padding="longest"``, ``pad_to_multiple_of=hop``,
        ref :295-303 / hf feature_extraction_whisper.py:296-303).  Returns pinned-able fp32 ``[B, L]``
        and the per-clip valid frame counts (= ones in ``sample_mask[:, ::hop]``)."""
        hop = self.mel_spec.hop_length
        audios = [np.pad(a, (0, 2 * hop - len(a))) if len(a) < 2 * hop else a for a in audios]
        longest = max(len(a) for a in audios)
        width = -(-longest // hop) * hop
        out = torch.zeros(len(audios), width, dtype=torch.float32)
        for i, a in enumerate(audios):
            out[i, : len(a)] = torch.as_tensor(np.asarray(a, dtype=np.float32))
        return out, [-(-len(a) // hop) for a in audios]

    def _chunk_and_pad_audio(self, audio_values: torch.Tensor, audio_lens: torch.Tensor,
                             include_audio_num_chunks: bool = False) -> dict[str, Any]:
        """ref :153-215 on a ``[B, n_mels, T]`` mel (any device)."""
        context = self.audio_context_size or audio_values.shape[-1]
        plan, per_clip = frame_chunks(audio_lens.tolist(), context)
        pieces = []
        for clip, off, _, cont in plan:
            piece = audio_values[clip, :, off: off + context]
            if cont and piece.shape[-1] < context:
                piece = F.pad(piece, (0, context - piece.shape[-1]))
            pieces.append(piece)
        data = {"audio_values": torch.stack(pieces, 0),
                "audio_lens": torch.tensor([p[2] for p in plan], dtype=torch.int64),
                "audio_is_continuation": torch.tensor([p[3] for p in plan], dtype=torch.bool),
                "audio_batch_size": torch.tensor([len(plan)])}
        if include_audio_num_chunks:
            data["audio_num_chunks"] = torch.tensor(per_clip, dtype=torch.int64)
        return data

    # -- main entry ----------------------------------------------------------------------------
    def __call__(self, text: Optional[str] = None, audio=None, audios=None, sampling_rate: Optional[int] = None,
                 return_tensors="pt", include_audio_num_chunks: bool = False, **kwargs) -> transformers.BatchFeature:
        if audio is not None and audios is not None:
            raise ValueError("Only one of `audio` or `audios` should be provided.")
        if audio is not None:
            audios = audio if isinstance(audio, list) or audio.ndim == 2 else [audio]
        elif audios is None:
            audios = []
        if sampling_rate is not None and len(audios) > 0 and sampling_rate != self.mel_spec.sampling_rate:
            raise ValueError(f"The model corresponding to this feature extractor was trained using a sampling rate of "
                             f"{self.mel_spec.sampling_rate}; got {sampling_rate}.")
        data: dict[str, Any] = {}
        continuation: list[bool] = []
        if len(audios) > 0:
            audios = [x.numpy() if isinstance(x, torch.Tensor) else np.asarray(x) for x in audios]
            waves, frame_lens = self.pad_waveforms(audios)
            if self.defer_mel:
                context = self.audio_context_size or (waves.shape[-1] // self.mel_spec.hop_length)
                plan, per_clip = frame_chunks(frame_lens, context)
                data["audio_waveforms"] = waves
                data["audio_num_frames"] = torch.tensor(frame_lens, dtype=torch.int64)
                data["audio_lens"] = torch.tensor([p[2] for p in plan], dtype=torch.int64)
                data["audio_batch_size"] = torch.tensor([len(plan)])
                if include_audio_num_chunks:
                    data["audio_num_chunks"] = torch.tensor(per_clip, dtype=torch.int64)
                continuation = [p[3] for p in plan]
            else:
                from . import ops  # CUDA only; raises if libuvx is missing
                mel = ops.logmel(waves.to(self.mel_device, non_blocking=True), self.mel_spec.feature_size)
                chunked = self._chunk_and_pad_audio(mel, torch.tensor(frame_lens), include_audio_num_chunks)
                continuation = chunked.pop("audio_is_continuation").tolist()
                data.update(chunked)
            data["audio_token_len"] = torch.ceil(
                data["audio_lens"] / (self.encoder_ds_factor * self.stack_factor)).to(dtype=torch.int)

        if text is not None:
            if not isinstance(text, str):
                raise ValueError("Text must be a string. Batch mode not supported yet.")
            parts = self.tokenizer(text.split("<|audio|>"), add_special_tokens=False, **kwargs)["input_ids"]
            placeholder_id = self.vocab[self.audio_token_replacement]
            ids: list[int] = []
            starts: list[int] = []
            slot = -1
            for i, n in enumerate(data.get("audio_token_len", [])):
                if not continuation[i]:
                    slot += 1
                    if slot >= len(parts):
                        raise ValueError(f"Text contains too few audio placeholders. (Expected {len(audios)} placeholders)")
                    ids.extend(parts[slot])
                starts.append(len(ids))
                ids.extend([placeholder_id] * int(n))
            slot += 1
            if slot != len(parts) - 1:
                raise ValueError(f"Text contains too many audio placeholders. (Expected {len(audios)} placeholders)")
            ids.extend(parts[slot])
            if "audio_token_len" in data:
                data["audio_token_start_idx"] = torch.as_tensor(starts)
            data["input_ids"] = [ids]
            data["attention_mask"] = [[1] * len(ids)]
        return transformers.BatchFeature(data=data, tensor_type=return_tensors)

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)

    @property
    def model_input_names(self):
        names = list(getattr(self.tokenizer, "model_input_names", ["input_ids", "attention_mask"]))
        return list(set(names + ["input_features"]))


def _register_with_transformers() -> None:
    """``AutoProcessor`` mapping, as the reference registers it at import time (ref ultravox_processing.py:385-387)."""
    import transformers
    from .config import UltravoxConfig
    transformers.AutoProcessor.register(UltravoxConfig, UltravoxProcessor, exist_ok=True)


_register_with_transformers()
