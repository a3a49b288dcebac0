"""Rotary position embedding for packed sequences (drop-in for esme/rotary.py).

The reference builds the per-token position index with ~17 small PyTorch kernels
and three host syncs per call; here positions come from one `esmk_batch_meta`
launch per batch and the rotation is one `esmk_qk_norm_rope` launch (or, on the
model path, the epilogue of the QKV GEMM)."""
from typing import Tuple

import torch

from . import ops


class RotaryEmbedding(torch.nn.Module):
    """Same constructor and call signature as the reference module
    (esme/rotary.py:81-165): ``forward(q, k, cu_lens, max_len) -> (q, k)`` for
    q, k of shape [T, H, dim].  cos/sin follow the reference exactly: fp32 angles,
    tables rounded to bf16, bf16 arithmetic with three rounding points."""

    def __init__(self, dim: int, base=10000.0, pos_idx_in_fp32=True, device=None):
        super().__init__()
        if float(base) != 10000.0 or not pos_idx_in_fp32:
            raise NotImplementedError('only base=10000, pos_idx_in_fp32=True are supported (the values esme uses)')
        self.dim = dim
        self.base = float(base)
        self.pos_idx_in_fp32 = pos_idx_in_fp32
        inv_freq = 1.0 / (self.base ** (torch.arange(0, dim, 2, device=device, dtype=torch.float32) / dim))
        self.register_buffer('inv_freq', inv_freq, persistent=False)
        self._seq_len_cached = 0
        self._cos_cached = None
        self._sin_cached = None

    def _update_cos_sin_cache(self, seqlen, device=None, dtype=None):
        if dtype not in (None, torch.bfloat16):
            raise NotImplementedError('the B200 kernels are bf16-only')
        if (seqlen > self._seq_len_cached or self._cos_cached is None
                or self._cos_cached.device != torch.device(device)):
            self._seq_len_cached = seqlen
            self._cos_cached, self._sin_cached = ops.rope_tables(seqlen, self.dim, device)

    def forward(self, q: torch.Tensor, k: torch.Tensor, cu_lens: torch.Tensor,
                max_len: int) -> Tuple[torch.Tensor, torch.Tensor]:
        ops._need_cuda(q, k, cu_lens)
        self._update_cos_sin_cache(max_len, device=q.device, dtype=q.dtype)
        T, H, hd = q.shape
        pos, _ = ops.batch_meta(cu_lens.to(torch.int32), T)
        # out of place like the reference: rotate copies laid out as one [T, 2, H*hd] block
        qk = torch.stack((q.reshape(T, H * hd), k.reshape(T, H * hd)), dim=1)
        ops.qk_norm_rope_(qk[:, 0], qk[:, 1], H, hd, cos=self._cos_cached, sin=self._sin_cached, pos=pos)
        return qk[:, 0].reshape(T, H, hd), qk[:, 1].reshape(T, H, hd)
