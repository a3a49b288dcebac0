"""Learned positional embedding of ESM-1b / ESM-1v (drop-in for esme/embedding.py:7-107).  Positions count from
`padding_idx + 1` inside each sequence; the table has `num_embeddings + 2` rows."""
import torch
import torch.nn as nn

from . import ops
from .alphabet import Alphabet


class LearnedPositionalEmbedding(nn.Embedding):
    def __init__(self, num_embeddings: int, embedding_dim: int, dtype=torch.bfloat16):
        super().__init__(num_embeddings + 2, embedding_dim, Alphabet.padding_idx, dtype=dtype)
        self.max_positions = num_embeddings

    def positions(self, input: torch.Tensor) -> torch.Tensor:
        """esme/embedding.py:36-52: [B,S] tokens -> position ids (padding stays at padding_idx)."""
        if input.size(1) > self.max_positions:
            raise ValueError(f'Sequence length {input.size(1)} above maximum  sequence length of {self.max_positions}')
        pad = input.ne(self.padding_idx).int()
        return (torch.cumsum(pad, dim=1).type_as(pad) * pad).long() + self.padding_idx

    def position_unpad(self, input: torch.Tensor, pad_args) -> torch.Tensor:
        """esme/embedding.py:54-79: packed tokens -> position ids 2, 3, ... restarting per sequence."""
        assert input.ndim == 1
        cu_lens, max_len = pad_args
        if max_len > self.max_positions:
            raise ValueError(f'Sequence length {max_len} above maximum  sequence length of {self.max_positions}')
        pos, _ = ops.batch_meta(cu_lens.to(torch.int32).contiguous(), input.numel())
        return pos.long() + 1 + self.padding_idx

    def forward(self, input: torch.Tensor, pad_args=None) -> torch.Tensor:
        ids = self.positions(input) if pad_args is None else self.position_unpad(input, pad_args)
        rows = ops.embed(ids.reshape(-1), self.weight, zero_token=self.padding_idx)   # padding row -> zeros
        return rows.reshape(*ids.shape, -1)
