"""Multi-adapter LoRA around the attention projections (drop-in for esme/lora.py and the LoRA plumbing of
esme/esm.py:495-607): same module layout and state-dict keys (`...self_attn.q.layer.weight`,
`...self_attn.q.lora_A.<name>`, `...lora_B.<name>`), same `add_lora / load_lora / save_lora / lora_state_dict /
mark_only_lora_as_trainable` surface, `lora_names` selecting adapters per call (None = all, as the reference).

Inference only.  The adapters stay UNMERGED, as in the reference's forward (lora.py:73-91): merging `B A * scaling`
into a bf16 weight would round away deltas smaller than the weight's bf16 ulp.  Each adapter is two small tcgen05
GEMMs: u = x A^T, then y += bf(u B^T * scaling) through the residual epilogue (in place, also on a column block of
the packed QKV output, before RoPE).  A model carrying adapters runs the per-layer operator path
(`FlashTransformerLayer.forward`) instead of the whole-model `esmk_forward` call."""
import math
from typing import Dict, Iterable, Optional

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .quantization import dense_weight


class LoRA(nn.Module):
    """esme/lora.py:8-95."""

    def __init__(self, layer: nn.Module, rank: int = 16, alpha: float = 1, dropout_p: float = 0., names: list = None,
                 dtype=None):
        super().__init__()
        assert getattr(layer, 'in_features', None) is not None, 'The layer must have an attribute in_features'
        assert getattr(layer, 'out_features', None) is not None, 'The layer must have an attribute out_features'
        assert rank >= 0, 'The rank must be a non-negative integer'
        if dropout_p > 0.:
            raise NotImplementedError('LoRA dropout is a training feature; this build is inference-only')
        self.layer = layer
        self.rank = rank
        self.alpha = alpha
        self.dropout_p = dropout_p
        self.scaling = self.alpha / self.rank
        self.in_features = layer.in_features
        self.out_features = layer.out_features
        names = names or ['default']
        self.names = set(names)
        p0 = next(layer.parameters())
        device = p0.device
        dtype = dtype or p0.dtype
        if dtype in (torch.uint8, torch.int8):
            dtype = torch.bfloat16
        self.lora_A = nn.ParameterDict({n: nn.Parameter(torch.zeros((rank, self.in_features), device=device, dtype=dtype))
                                        for n in names})
        self.lora_B = nn.ParameterDict({n: nn.Parameter(torch.zeros((self.out_features, rank), device=device, dtype=dtype))
                                        for n in names})
        self.reset_parameters()

    def reset_parameters(self):
        for name in self.names:
            if self.lora_A[name].device.type != 'meta':
                nn.init.kaiming_uniform_(self.lora_A[name], a=math.sqrt(5))
                nn.init.zeros_(self.lora_B[name])

    @property
    def bias(self):
        return self.layer.bias

    def _padded(self, name):
        """A [r8, in], B [out, r8] with the rank zero-padded to a multiple of 8 (16-byte TMA pitch)."""
        A, B = self.lora_A[name], self.lora_B[name]
        r8 = (self.rank + 7) // 8 * 8
        if r8 != self.rank:
            A = torch.cat((A, A.new_zeros(r8 - self.rank, A.shape[1])), 0)
            B = torch.cat((B, B.new_zeros(B.shape[0], r8 - self.rank)), 1)
        return A.contiguous(), B.contiguous()

    def add_adapters_(self, x: torch.Tensor, y: torch.Tensor, names: Optional[Iterable[str]] = None) -> torch.Tensor:
        """y += sum over adapters of bf(bf(x A^T) B^T * scaling), in place; y may be a column block of a wider
        tensor (the packed QKV output)."""
        for name in (names or self.names):
            if name not in self.lora_A:
                raise KeyError(name)
            A, B = self._padded(name)
            u = ops.linear(x, A)
            ops.linear(u, B, epilogue=L.EPI_RESIDUAL, residual=y, residue_scaling=1.0 / self.scaling, out=y)
        return y

    def forward(self, x: torch.Tensor, names=None) -> torch.Tensor:
        y = ops.linear(x, dense_weight(self.layer), self.layer.bias)
        return self.add_adapters_(x, y, names)

    def extra_repr(self):
        return f'in_features={self.in_features}, out_features={self.out_features}, rank={self.rank}, ' \
               f'alpha={self.alpha}, dropout_p={self.dropout_p}'


def mark_only_lora_as_trainable(model: nn.Module, names=None) -> None:
    """esme/lora.py:97-110."""
    names = set(names or [])
    for n, p in model.named_parameters():
        trainable = ('.lora_A.' in n or '.lora_B.' in n) and (not names or n.split('.')[-1] in names)
        if p.dtype.is_floating_point:
            p.requires_grad = trainable


def lora_state_dict(model: nn.Module, names=None) -> Dict[str, torch.Tensor]:
    """esme/lora.py:113-124."""
    names = set(names or [])
    return {k: v for k, v in model.state_dict().items()
            if ('.lora_A.' in k or '.lora_B.' in k) and (not names or k.split('.')[-1] in names)}


def has_lora(model: nn.Module) -> bool:
    return any(isinstance(m, LoRA) for m in model.modules())
