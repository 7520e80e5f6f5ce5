"""Losses of the hot path (CUDA, through libuvx)."""
from __future__ import annotations

import torch

from ._lib import check, lib
from .ops import _cuda, _stream


def causal_lm_loss(logits: torch.Tensor, labels: torch.Tensor, ignore_index: int = -100, keep: dict | None = None,
                   shift: bool = True):
    """Next-token cross entropy, fp32 (hf:loss/loss_utils.py:28-67).  logits [B, S, V] fp32, labels [B, S]; with
    ``shift=False`` logits are [R, V] pre-gathered rows and labels [R] their targets."""
    _cuda(logits, torch.float32, "logits"), _cuda(labels, torch.int64, "labels")
    if logits.dim() == 2:
        logits = logits[None]
        labels = labels.reshape(1, -1)
    B, S, V = logits.shape
    lg = logits.reshape(B * S, V)
    labels = labels.contiguous()
    row_loss = torch.empty(B * S, dtype=torch.float32, device=logits.device)
    row_lse = torch.empty_like(row_loss)
    out = torch.empty(2, dtype=torch.float32, device=logits.device)
    check(lib().uvx_ce_loss(lg.data_ptr(), lg.stride(0), labels.data_ptr(), B, S, V, ignore_index, int(shift),
                            row_loss.data_ptr(), row_lse.data_ptr(), out.data_ptr(), _stream()), "uvx_ce_loss")
    if keep is not None:
        keep.update(row_lse=row_lse, loss2=out, logits2d=lg, labels=labels, B=B, S=S, V=V, shift=int(shift),
                    ignore_index=ignore_index)
    return out[0]


def causal_lm_loss_bwd(keep: dict, grad_scale: float = 1.0) -> torch.Tensor:
    """d(loss)/d(logits) as bf16 [B*S, V] for the rows/labels recorded by ``causal_lm_loss(keep=...)``."""
    lg = keep["logits2d"]
    d = torch.empty(lg.shape, dtype=torch.bfloat16, device=lg.device)
    check(lib().uvx_ce_bwd(lg.data_ptr(), lg.stride(0), keep["labels"].data_ptr(), keep["B"], keep["S"], keep["V"],
                           keep["ignore_index"], keep["shift"], keep["row_lse"].data_ptr(), keep["loss2"].data_ptr(),
                           grad_scale, d.data_ptr(), _stream()), "uvx_ce_bwd")
    return d
