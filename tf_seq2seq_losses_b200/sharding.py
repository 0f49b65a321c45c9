"""Batch sharding across the GPUs of one box (one process per GPU).

Utterances are independent (every tensor of the path carries a leading batch axis and the loss is per-sample,
tf_seq2seq_losses/base_loss.py:140-155), so the batch is cut into contiguous slices, one per rank, with no exchange on
the data path.  The only collective is the optional reduction of the scalar summed loss that a caller performing
``tf.reduce_sum/mean(loss)`` (tests/benchmark.py:199 of the reference) needs.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(batch: int, world_size: int, rank: int) -> Tuple[int, int]:
    """Contiguous [begin, end) slice of the batch owned by ``rank``; sizes differ by at most one."""
    assert world_size >= 1 and 0 <= rank < world_size
    base, rem = divmod(batch, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def sharded_loss_and_grad(loss_grad_fn: Callable, labels, logits, label_length, logit_length, blank_index=0,
                          reduce_loss: bool = True, group: Optional[dist.ProcessGroup] = None):
    """Runs ``loss_grad_fn(labels, logits, label_length, logit_length, blank_index) -> (loss[b], grad[b,T,V])`` on this
    rank's slice of a replicated batch and (optionally) all-reduces the summed loss.

    Returns (local_loss, local_grad, (begin, end), total_loss_or_None).
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    begin, end = shard_bounds(logits.shape[0], world, rank)
    loss, grad = loss_grad_fn(labels[begin:end], logits[begin:end], label_length[begin:end],
                              logit_length[begin:end], blank_index)
    total = None
    if reduce_loss:
        finite = torch.where(torch.isinf(loss), torch.zeros_like(loss), loss)
        total = finite.sum().reshape(1).clone()
        if world > 1:
            dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    return loss, grad, (begin, end), total
