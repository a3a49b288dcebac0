"""Weight-only quantised linears (drop-in for the bitsandbytes modules the reference's quantised loaders
install: esme/esm.py:414-472 `_load_linear8bit` / `_load_linear4bit`, applied to q, k, v, out and the two FFN
linears by `_load_quantize`, esme/esm.py:449-472 and 916-946).

bitsandbytes is a third-party dependency that is absent from /root/reference, so the storage formats are this
library's own (declared in include/esmk.h; PARITY UNPINNED against bitsandbytes):

  Linear4bit   `weight` uint8 [N*K/2, 1] (bitsandbytes' Params4bit shape): fp4 codebook, blocks of 64 along K,
               fp32 `absmax` per block, no double quantisation
  Linear8bit   `weight` int8 [N, K], fp32 `SCB`-style per-row scale (`scale` = row absmax / 127); the reference's
               LLM.int8 activation quantisation / outlier split is not reproduced (weights only)

Biases stay exact bf16, as the reference's tests require (tests/test_esm.py:123-154).  Like bitsandbytes' batched
path (dequantize, then F.linear) the GEMM runs in bf16 on the dequantised weight; inside `model(...)` the engine
expands each weight into workspace scratch right before the GEMM that uses it.
"""
from typing import Optional

import torch
import torch.nn as nn

from . import ops


class _QuantLinear(nn.Module):
    bits = 0

    def __init__(self, in_features: int, out_features: int, data: torch.Tensor, scale: torch.Tensor,
                 bias: Optional[torch.Tensor]):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.weight = nn.Parameter(data, requires_grad=False)
        self.register_buffer('scale', scale)
        self.bias = None if bias is None else nn.Parameter(bias.detach().clone(), requires_grad=False)

    def _apply(self, fn, *a, **kw):
        """Device moves apply; dtype casts (model.bfloat16(), .to(torch.bfloat16)) must not touch the integer payload
        (torch leaves integer tensors alone) NOR the fp32 block scales, which the kernels read as fp32."""
        scale = self.scale
        out = super()._apply(fn, *a, **kw)
        if self.scale.dtype != torch.float32:
            self.scale = scale.to(self.scale.device)
        return out

    @classmethod
    def from_linear(cls, linear: nn.Linear) -> '_QuantLinear':
        """Quantise an nn.Linear whose bf16 weight already lives on a CUDA device."""
        data, scale = ops.quantize(linear.weight.detach(), cls.bits)
        return cls(linear.in_features, linear.out_features, data, scale, linear.bias)

    def dequantize(self) -> torch.Tensor:
        return ops.dequantize(self.weight.data, self.scale, self.out_features, self.in_features, self.bits)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return ops.linear(x, self.dequantize(), self.bias)

    def extra_repr(self) -> str:
        return f'in_features={self.in_features}, out_features={self.out_features}, bits={self.bits}, ' \
               f'bias={self.bias is not None}'


class Linear4bit(_QuantLinear):
    bits = 4


class Linear8bit(_QuantLinear):
    bits = 8


def dense_weight(linear) -> torch.Tensor:
    """bf16 [N,K] weight of an nn.Linear or of a quantised linear (dequantised)."""
    return linear.dequantize() if isinstance(linear, _QuantLinear) else linear.weight


def quantize_model_(model, bits: int, which=('q', 'k', 'v', 'out', 'ffn')):
    """Replace linears of every layer of an ESM2 / ESMC model (already on a CUDA device) in place.  `which` selects
    among q, k, v, out and ffn (= final.1 / final.3 for ESM2, final.1.activation / .fc and final.2 for ESMC);
    the reference's `quantization=` loaders convert all of them."""
    cls = {4: Linear4bit, 8: Linear8bit}[bits]
    for layer in model.layers:
        sa = layer.self_attn
        for name in ('q', 'k', 'v', 'out'):
            if name in which and isinstance(getattr(sa, name), nn.Linear):
                setattr(sa, name, cls.from_linear(getattr(sa, name)))
        if 'ffn' in which:
            if layer.final_activation == 'gelu':
                for idx in (1, 3):
                    if isinstance(layer.final[idx], nn.Linear):
                        layer.final[idx] = cls.from_linear(layer.final[idx])
            else:
                glu = layer.final[1]
                for name in ('activation', 'fc'):
                    if isinstance(getattr(glu, name), nn.Linear):
                        setattr(glu, name, cls.from_linear(getattr(glu, name)))
                if isinstance(layer.final[2], nn.Linear):
                    layer.final[2] = cls.from_linear(layer.final[2])
    model._engine = None
    return model
