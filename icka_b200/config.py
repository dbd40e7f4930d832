"""Duck-typed stand-in for the ``config`` object the reference modules read.

The reference passes an HF ``RobertaConfig`` (My_cross_attention.py:671) or its own ``BertConfig``
(CMIM:45-139); the cross-modal modules only read the attributes below.
"""
from dataclasses import dataclass


@dataclass
class FusionConfig:
    hidden_size: int = 768
    num_attention_heads: int = 12
    intermediate_size: int = 3072
    hidden_act: str = 'gelu'
    hidden_dropout_prob: float = 0.1
    attention_probs_dropout_prob: float = 0.1
    layer_norm_eps: float = 1e-12
    vocab_size: int = 30522
