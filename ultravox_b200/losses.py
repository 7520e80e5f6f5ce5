"""Losses of the hot path (CUDA, through libuvx)."""
from __future__ import annotations

import torch

from ._lib import check, lib
from .ops import _cuda, _stream


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100, keep: dict | None = None):
    """Shifted next-token cross entropy, fp32 (hf:loss/loss_utils.py:28-67).  logits [B, S, V] fp32, labels [B, S]."""
    _cuda(logits, torch.float32, "logits"), _cuda(labels, torch.int64, "labels")
    B, S, V = logits.shape
    lg = logits.reshape(B * S, V)
    labels = labels.contiguous()
    row_loss = torch.empty(B * S, dtype=torch.float32, device=logits.device)
    row_lse = torch.empty_like(row_loss)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    check(lib().uvx_ce_loss(lg.data_ptr(), lg.stride(0), labels.data_ptr(), B, S, V, ignore_index, row_loss.data_ptr(),
                            row_lse.data_ptr(), out.data_ptr(), _stream()), "uvx_ce_loss")
    if keep is not None:
        keep.update(row_lse=row_lse, count=out[1])
    return out[0]
