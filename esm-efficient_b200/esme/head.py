"""LM head (drop-in for esme/head.py:8-27): final(LN(gelu(dense(x))))."""
import torch
import torch.nn as nn

from . import _lib as L
from . import ops


class RobertaLMHead(nn.Module):
    """Parameters use the reference's names (`dense`, `layer_norm`, `final`) so
    checkpoints load unchanged; `forward` runs dense+bias+GELU as one tcgen05 GEMM
    epilogue, the LayerNorm kernel, and the vocabulary projection GEMM."""

    def __init__(self, embed_dim, vocab_size, dtype=torch.bfloat16):
        super().__init__()
        self.dense = nn.Linear(embed_dim, embed_dim, dtype=dtype)
        self.layer_norm = nn.LayerNorm(embed_dim, dtype=dtype)
        self.final = nn.Linear(embed_dim, vocab_size, dtype=dtype)

    def forward(self, features: torch.Tensor) -> torch.Tensor:
        x = ops.linear(features, self.dense.weight, self.dense.bias, epilogue=L.EPI_BIAS_GELU)
        x = ops.layernorm(x, self.layer_norm.weight, self.layer_norm.bias, self.layer_norm.eps)
        return ops.linear(x, self.final.weight, self.final.bias)
