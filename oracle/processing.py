"""Oracle (test infrastructure): the integer / token-index path of the Ultravox processor.

Plain-python / numpy restatement of
* ``UltravoxProcessor.__call__``            ref:ultravox/model/ultravox_processing.py:217-370
* ``UltravoxProcessor._chunk_and_pad_audio`` ref:ultravox/model/ultravox_processing.py:153-215
* ``DataCollatorForSeq2SeqWithAudio.__call__`` ref:ultravox/model/ultravox_processing.py:17-64

Bit-exact contract.  Pinned against the tokenizer-independent literals of
``ref:ultravox/model/ultravox_processing_test.py:46-137,177-186`` and
``ref:ultravox/inference/infer_test.py:72-109`` and against fixtures produced by the reference
class itself (``tests/golden/processor_cases.json``, made by ``scripts/make_golden.py``).
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np

from . import logmel

HOP = logmel.HOP


def chunk_plan(frame_lens: Sequence[int], context: int = 3000):
    """ref :153-215 - split every clip's frames into <= ``context`` pieces.

    Returns (chunk_lens, is_continuation, num_chunks, (clip, offset) per chunk)."""
    chunk_lens, cont, num_chunks, src = [], [], [], []
    for i, n in enumerate(frame_lens):
        n = int(n)
        num_chunks.append(int(math.ceil(n / context)))
        for off in range(0, n, context):
            chunk_lens.append(min(n - off, context))
            cont.append(off > 0)
            src.append((i, off))
    return chunk_lens, cont, num_chunks, src


def chunk_and_pad(mel: np.ndarray, frame_lens: Sequence[int], context: int = 3000):
    """mel [B, n_mels, T] -> stacked chunks [N, n_mels, T'] exactly as ref :175-199.

    Only *continuation* chunks are right-padded with literal 0.0 to ``context``; the first chunk of
    every clip keeps the batch width (clipped at ``context``)."""
    chunk_lens, cont, num_chunks, src = chunk_plan(frame_lens, context)
    pieces = []
    for (i, off), c in zip(src, cont):
        piece = mel[i, :, off: off + context]
        if c and piece.shape[-1] < context:
            piece = np.pad(piece, ((0, 0), (0, context - piece.shape[-1])))
        pieces.append(piece)
    return np.stack(pieces, 0), chunk_lens, cont, num_chunks


def audio_token_len(chunk_lens: Sequence[int], ds: int = 2, stack: int = 8) -> list[int]:
    """ref :316-318 - ceil(audio_lens / (encoder_ds_factor * stack_factor)), int32."""
    return [int(math.ceil(n / (ds * stack))) for n in chunk_lens]


def build_input_ids(split_ids: list[list[int]], tok_lens: Sequence[int], cont: Sequence[bool],
                    placeholder_id: int, n_audios: int):
    """ref :325-366 - interleave tokenised text parts with placeholder runs."""
    ids: list[int] = []
    starts: list[int] = []
    ph = -1
    for n, c in zip(tok_lens, cont):
        if not c:
            ph += 1
            if ph >= len(split_ids):
                raise ValueError(f"Text contains too few audio placeholders. (Expected {n_audios} placeholders)")
            ids.extend(split_ids[ph])
        starts.append(len(ids))
        ids.extend([placeholder_id] * int(n))
    ph += 1
    if ph != len(split_ids) - 1:
        raise ValueError(f"Text contains too many audio placeholders. (Expected {n_audios} placeholders)")
    ids.extend(split_ids[ph])
    return ids, starts


def process(text: str | None, audios: list[np.ndarray], tokenize: Callable[[list[str]], list[list[int]]],
            placeholder_id: int, n_mels: int = 128, context: int = 3000, ds: int = 2, stack: int = 8,
            include_audio_num_chunks: bool = False, with_mel: bool = True) -> dict:
    """End-to-end restatement of ``UltravoxProcessor.__call__`` (numpy outputs)."""
    data: dict = {}
    cont: list[bool] = []
    if len(audios) > 0:
        audios = [np.pad(np.asarray(a), (0, 2 * HOP - len(a))) if len(a) < 2 * HOP else np.asarray(a)
                  for a in audios]                                              # ref :283-292
        padded, frame_lens = logmel.pad_batch(audios)
        if with_mel:
            mel = logmel.log_mel(padded, n_mels)
            vals, chunk_lens, cont, num_chunks = chunk_and_pad(mel, frame_lens, context)
            data["audio_values"] = vals
        else:
            chunk_lens, cont, num_chunks, _ = chunk_plan(frame_lens, context)
        data["audio_lens"] = np.asarray(chunk_lens, dtype=np.int64)
        data["audio_batch_size"] = np.asarray([len(chunk_lens)], dtype=np.int64)
        if include_audio_num_chunks:
            data["audio_num_chunks"] = np.asarray(num_chunks, dtype=np.int64)
        data["audio_token_len"] = np.asarray(audio_token_len(chunk_lens, ds, stack), dtype=np.int32)
    if text is not None:
        if not isinstance(text, str):
            raise ValueError("Text must be a string. Batch mode not supported yet.")
        split_ids = tokenize(text.split("<|audio|>"))
        ids, starts = build_input_ids(split_ids, data.get("audio_token_len", []), cont, placeholder_id,
                                      len(audios))
        if "audio_token_len" in data:
            data["audio_token_start_idx"] = np.asarray(starts, dtype=np.int64)
        data["input_ids"] = np.asarray([ids], dtype=np.int64)
        data["attention_mask"] = np.ones((1, len(ids)), dtype=np.int64)
    return data


def collate(features: list[dict], pad_id: int, padding_side: str = "right", label_pad: int = -100) -> dict:
    """ref :17-64 (+ hf DataCollatorForSeq2Seq padding): flatten audio lists, pad ids/mask/labels,
    right-pad mel on time, displace start indices under left padding."""
    vals = [x for f in features for x in f.get("audio_values", [])]
    lens = [x for f in features for x in f.get("audio_lens", [])]
    tlen = [x for f in features for x in f.get("audio_token_len", [])]
    start = [x for f in features for x in f.get("audio_token_start_idx", [])]
    L = max(len(f["input_ids"]) for f in features)
    B = len(features)
    ids = np.full((B, L), pad_id, dtype=np.int64)
    mask = np.zeros((B, L), dtype=np.int64)
    has_labels = all("labels" in f and f["labels"] is not None for f in features)
    labels = np.full((B, L), label_pad, dtype=np.int64) if has_labels else None
    for i, f in enumerate(features):
        n = len(f["input_ids"])
        sl = slice(L - n, L) if padding_side == "left" else slice(0, n)
        ids[i, sl] = f["input_ids"]
        mask[i, sl] = 1
        if has_labels:
            labels[i, sl] = f["labels"]
    batch = {"input_ids": ids, "attention_mask": mask, "labels": labels}
    if "audio_batch_size" in features[0]:
        batch["audio_batch_size"] = np.stack([np.asarray(f["audio_batch_size"]) for f in features])
    if len(vals) > 0 and len(vals[0]) > 0:
        batch["audio_token_start_idx"] = np.asarray(start, dtype=np.int64)
        batch["audio_lens"] = np.asarray(lens, dtype=np.int64)
        batch["audio_token_len"] = np.asarray(tlen, dtype=np.int32)
        T = max(v.shape[-1] for v in vals)
        batch["audio_values"] = np.stack([np.pad(v, ((0, 0), (0, T - v.shape[-1]))) for v in vals])
        if padding_side == "left":
            disp = np.asarray([L - len(f["input_ids"]) for f in features], dtype=np.int64)
            disp = np.repeat(disp, batch["audio_batch_size"].reshape(-1))
            batch["audio_token_start_idx"] = batch["audio_token_start_idx"] + disp
    return batch
