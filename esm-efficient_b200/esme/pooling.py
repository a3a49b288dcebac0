"""Pooling of packed representations (drop-in for esme/pooling.py): the per-sequence mean
pool and the class-token attention pool with the heads built on it.  The reference's attention pool calls
flash_attn_varlen_func with one query per (class token, sequence) (pooling.py:126-134); here it is one
`esmk_attn_pool` launch -- no `repeat`ed copies of k and v per class token."""
from typing import Tuple

import torch
import torch.nn as nn

from . import ops


def partition_mean_pool(embed: torch.Tensor, cu_lens: torch.Tensor) -> torch.Tensor:
    """Mean of the rows of every `cu_lens` partition: [T, D] -> [B, D].  The reference accumulates with
    bf16 `index_add_` (atomic, order-dependent); this kernel accumulates in fp32 and rounds once."""
    return ops.mean_pool(embed, cu_lens)


class PartitionMeanPool(nn.Module):
    def forward(self, embed, cu_lens):
        return partition_mean_pool(embed, cu_lens)

    @staticmethod
    def _indices(cu_lens):
        lens = (cu_lens[1:] - cu_lens[:-1]).long()
        return torch.repeat_interleave(torch.arange(lens.numel(), device=cu_lens.device), lens)


class AttentionPool(nn.Module):
    """esme/pooling.py:72-136: class tokens `cls` [C, D] attend to the tokens of every sequence; keys are
    `self.k(embed)`, values the embeddings themselves.  -> [B, C, D]."""

    def __init__(self, attention_heads: int, embed_dim: int, dropout_p=0.0, dtype=torch.bfloat16):
        super().__init__()
        if dropout_p != 0.0:
            raise NotImplementedError('attention dropout is a training feature; the inference kernels use p=0')
        self.attention_heads = attention_heads
        self.dropout_p = dropout_p
        self.k = nn.Linear(embed_dim, embed_dim, dtype=dtype)

    def forward(self, cls: torch.Tensor, embed: torch.Tensor, pad_args: Tuple[torch.Tensor, int]):
        cu_lens, _max_len = pad_args
        k = ops.linear(embed, self.k.weight, self.k.bias)
        return ops.attn_pool(cls, k, embed, cu_lens.to(torch.int32), self.attention_heads)


class LearnedAttentionPool(AttentionPool):
    """esme/pooling.py:139-179: the class tokens are parameters (initialised to ones, as the reference)."""

    def __init__(self, num_cls, attention_heads, embed_dim: int, dropout_p=0.0, dtype=torch.bfloat16):
        super().__init__(attention_heads, embed_dim, dropout_p, dtype)
        self.cls = nn.Parameter(torch.ones(num_cls, embed_dim, dtype=dtype))

    def forward(self, embed: torch.Tensor, pad_args: Tuple[torch.Tensor, int]):
        return super().forward(self.cls, embed, pad_args)


class LearnedAggregation(nn.Module):
    """esme/pooling.py:182-215: attention pool -> linear -> ReLU -> linear(1); [B, C, 1] squeezed on dim 1."""

    def __init__(self, num_cls, attention_heads: int, embed_dim: int, dropout_p=.0, dtype=torch.bfloat16):
        super().__init__()
        self.attn = LearnedAttentionPool(num_cls, attention_heads, embed_dim, dropout_p=dropout_p, dtype=dtype)
        self.linear = nn.Linear(embed_dim, embed_dim, dtype=dtype)
        self.relu = nn.ReLU()
        self.final = nn.Linear(embed_dim, 1, dtype=dtype)

    def forward(self, embed: torch.Tensor, pad_args: Tuple[torch.Tensor, int]):
        x = self.attn(embed, pad_args)                                   # [B, C, D]
        h = self.relu(ops.linear(x, self.linear.weight, self.linear.bias))
        return ops.linear(h, self.final.weight, self.final.bias).squeeze(1)


class BinaryLearnedAggregation(LearnedAggregation):
    """esme/pooling.py:218-228."""

    def __init__(self, attention_heads: int, embed_dim: int, dropout_p=0.0, dtype=torch.bfloat16):
        super().__init__(1, attention_heads, embed_dim, dropout_p, dtype)

    def forward(self, embed: torch.Tensor, pad_args: Tuple[torch.Tensor, int]):
        return super().forward(embed, pad_args).squeeze(-1)
