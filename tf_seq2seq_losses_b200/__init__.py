"""tf_seq2seq_losses_b200: B200-native (sm_100a) CTC loss / gradient / Hessian hot path behind the
tf_seq2seq_losses call signatures (tf_seq2seq_losses/__init__.py:22-28)."""
from .base_loss import BaseCtcLossData, ctc_loss, ctc_loss_from_logproba
from .classic_ctc_loss import ClassicCtcLossData, classic_ctc_loss
from .simplified_ctc_loss import SimplifiedCtcLossData, simple_ctc_loss, simplified_ctc_loss
from .sharding import shard_bounds, sharded_loss_and_grad
from ._lib import greedy_decode as ctc_greedy_decode

__version__ = "0.1.0"
__all__ = [
    "classic_ctc_loss", "simplified_ctc_loss", "simple_ctc_loss", "ctc_loss", "ctc_loss_from_logproba",
    "BaseCtcLossData", "ClassicCtcLossData", "SimplifiedCtcLossData", "shard_bounds", "sharded_loss_and_grad", "ctc_greedy_decode",
]
