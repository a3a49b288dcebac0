"""Per-sequence mean pooling of packed representations (drop-in for the inference part of
esme/pooling.py:8-69).  The attention-pool / learned-aggregation heads of the
reference are fine-tuning modules and stay out of scope."""
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
