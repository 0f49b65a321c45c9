"""Mirror of tf_seq2seq_losses/simplified_ctc_loss.py: CTC with trivial decoding (blank removal only)."""
from __future__ import annotations

from typing import Optional, Union

import torch

from . import _lib
from .base_loss import BaseCtcLossData, ctc_loss


class SimplifiedCtcLossData(BaseCtcLossData):
    """Mirror of SimplifiedCtcLossData (simplified_ctc_loss.py:70-534).  alpha / beta have shape [B, T+1, U]."""

    _variant = _lib.SIMPLIFIED


def simplified_ctc_loss(labels: torch.Tensor, logits: torch.Tensor, label_length: torch.Tensor,
                        logit_length: torch.Tensor, blank_index: Union[int, torch.Tensor] = 0,
                        max_label_length: Optional[int] = None, logits_time_major: bool = False) -> torch.Tensor:
    """Drop-in for simplified_ctc_loss (simplified_ctc_loss.py:32-67); see classic_ctc_loss for the arguments."""
    return ctc_loss(labels=labels, logits=logits, label_length=label_length, logit_length=logit_length,
                    blank_index=blank_index, ctc_loss_data_cls=SimplifiedCtcLossData,
                    max_label_length=max_label_length,
                    logits_time_major=logits_time_major)


# README.md:22,32 and tests/benchmark.py:72 of the reference call it simple_ctc_loss
simple_ctc_loss = simplified_ctc_loss
