"""Mirror of tf_seq2seq_losses/classic_ctc_loss.py: the repeat-collapsing CTC loss of Graves et al. (2006)."""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import _lib
from .base_loss import BaseCtcLossData, ctc_loss


class ClassicCtcLossData(BaseCtcLossData):
    """Mirror of ClassicCtcLossData (classic_ctc_loss.py:73-669).

    alpha / beta have shape [B, T+1, U, 2]; the last axis is the reference's state index (0 closed, 1 open).
    """

    _variant = _lib.CLASSIC


def classic_ctc_loss(labels: torch.Tensor, logits: torch.Tensor, label_length: torch.Tensor,
                     logit_length: torch.Tensor, blank_index: Union[int, torch.Tensor] = 0,
                     max_label_length: Optional[int] = None, logits_time_major: bool = False) -> torch.Tensor:
    """Drop-in for classic_ctc_loss (classic_ctc_loss.py:33-70), same argument order and meaning as
    ``tf.nn.ctc_loss(..., logits_time_major=False)``.

    Args:
        labels:        int32   [batch, max_label_length]
        logits:        float32 [batch, max_length, num_tokens] on a CUDA device
        label_length:  int32   [batch]
        logit_length:  int32   [batch]
        blank_index:   python int or scalar tensor
        max_label_length: optional keyword extension: max(label_length), to avoid reading it back from the device
        logits_time_major: optional keyword extension (as in tf.nn.ctc_loss): logits are [max_length, batch, num_tokens]

    Returns: float32 [batch] per-sample loss (+inf where the label cannot be emitted); differentiable twice
    w.r.t. ``logits``.
    """
    return ctc_loss(labels=labels, logits=logits, label_length=label_length, logit_length=logit_length,
                    blank_index=blank_index, ctc_loss_data_cls=ClassicCtcLossData, max_label_length=max_label_length,
                    logits_time_major=logits_time_major)
