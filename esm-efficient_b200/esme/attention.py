"""Transformer block operators (drop-in for esme/attention.py).

Module / parameter names mirror the reference so its checkpoints load unchanged
(`self_attn.{norm,q,k,v,out,layernorm_q,layernorm_k}`, `final.{0,1,2|3}`); the
`forward` methods call the sm_100a kernels through the C ABI:

  FlashMultiheadAttention.forward   LN -> one QKV GEMM (RoPE fused in the epilogue
                                    for ESM2; QK-LayerNorm+RoPE kernel for ESMC)
                                    -> varlen attention -> out projection
  FlashTransformerLayer.forward     the two residual branches with the residual
                                    add (and 1/residue_scaling) fused into the
                                    out-projection / FFN-down GEMM epilogues.
"""
import torch
import torch.nn as nn
from torch import Tensor

from . import _lib as L
from . import ops
from .quantization import dense_weight
from .rotary import RotaryEmbedding


def _base(linear):
    """The wrapped linear of a LoRA module, or the module itself."""
    from .lora import LoRA
    return linear.layer if isinstance(linear, LoRA) else linear


def _adapters(linear, x, y, lora_names):
    """y += the LoRA adapters of `linear` (if it is wrapped) applied to x, in place (esme/attention.py:94-101,
    136-139: lora_names = None selects every adapter, as LoRA.forward does)."""
    from .lora import LoRA
    if isinstance(linear, LoRA):
        linear.add_adapters_(x, y, lora_names)
    return y


class SwiGLU(nn.Module):
    """silu(activation(x)) * fc(x)  (reference: esme/attention.py:258-281)."""

    def __init__(self, in_features, out_features, bias=False, dtype=torch.bfloat16):
        super().__init__()
        if bias:
            raise NotImplementedError('SwiGLU with bias is not used by any ESM model')
        self.activation = nn.Linear(in_features, out_features, bias=bias, dtype=dtype)
        self.fc = nn.Linear(in_features, out_features, bias=bias, dtype=dtype)

    @staticmethod
    def interleave(act_w: Tensor, fc_w: Tensor) -> Tensor:
        """[F,D],[F,D] -> [2F,D] in alternating 32-row blocks (the layout the SWIGLU GEMM epilogue expects)."""
        F_, D = act_w.shape
        assert F_ % 32 == 0
        return torch.stack((act_w.reshape(F_ // 32, 32, D), fc_w.reshape(F_ // 32, 32, D)), dim=1) \
            .reshape(2 * F_, D).contiguous()

    def forward(self, x: Tensor) -> Tensor:
        w = self.interleave(dense_weight(self.activation), dense_weight(self.fc))
        return ops.linear(x, w, None, epilogue=L.EPI_SWIGLU)


class FlashMultiheadAttention(nn.Module):
    """Variable-length multi-head self attention over packed sequences
    (reference: esme/attention.py:10-139)."""

    def __init__(self, embed_dim: int, num_heads: int, dropout=0.0, pre_layernorm=True,
                 rotary_embedding=True, bias=False, dtype=torch.bfloat16):
        super().__init__()
        if dropout != 0.0:
            raise NotImplementedError('attention dropout is a training feature; the inference kernels use p=0')
        if dtype != torch.bfloat16:
            raise NotImplementedError('the B200 kernels are bf16-only')
        self.embed_dim = embed_dim
        self.num_heads = num_heads
        self.dropout = dropout
        self.head_dim = embed_dim // num_heads
        assert self.head_dim * num_heads == embed_dim, 'embed_dim must be divisible by num_heads'

        self.norm = nn.LayerNorm(embed_dim, dtype=dtype)
        self.q = nn.Linear(embed_dim, embed_dim, bias=bias, dtype=dtype)
        self.k = nn.Linear(embed_dim, embed_dim, bias=bias, dtype=dtype)
        self.v = nn.Linear(embed_dim, embed_dim, bias=bias, dtype=dtype)
        self.out = nn.Linear(embed_dim, embed_dim, bias=bias, dtype=dtype)
        self.rot_emb = RotaryEmbedding(dim=self.head_dim) if rotary_embedding else None
        self.pre_layernorm = pre_layernorm
        if pre_layernorm:
            self.layernorm_q = nn.LayerNorm(embed_dim, bias=bias, dtype=dtype)
            self.layernorm_k = nn.LayerNorm(embed_dim, bias=bias, dtype=dtype)

    def packed_qkv(self):
        """Concatenated [3D,D] weight and [3D] bias (or None) for the single QKV GEMM."""
        q, k, v = _base(self.q), _base(self.k), _base(self.v)
        w = torch.cat((dense_weight(q), dense_weight(k), dense_weight(v)), dim=0).contiguous()
        b = None
        if q.bias is not None:
            b = torch.cat((q.bias, k.bias, v.bias), dim=0).contiguous()
        return w, b

    def _has_qkv_adapters(self) -> bool:
        from .lora import LoRA
        return any(isinstance(m, LoRA) for m in (self.q, self.k, self.v))

    def _qkv_rot(self, x: Tensor, cu_lens: Tensor, max_len: int, lora_names=None):
        """-> qkv [T,3D] with q,k already QK-normalised (ESMC) and rotated."""
        D, H, hd = self.embed_dim, self.num_heads, self.head_dim
        T = x.shape[0]
        h = ops.layernorm(x, self.norm.weight, self.norm.bias, self.norm.eps)
        w, b = self.packed_qkv()
        pos, tile_info = ops.batch_meta(cu_lens, T)
        cos = sin = None
        if self.rot_emb is not None:
            self.rot_emb._update_cos_sin_cache(max_len, device=x.device, dtype=x.dtype)
            cos, sin = self.rot_emb._cos_cached, self.rot_emb._sin_cached
        adapters = self._has_qkv_adapters()      # LoRA deltas are added BEFORE the rotation: no fused RoPE epilogue
        fuse = (cos is not None) and (not self.pre_layernorm) and hd in (16, 32, 64) and (2 * D) % 64 == 0 \
            and not adapters
        if fuse:
            qkv = ops.linear(h, w, b, epilogue=L.EPI_QKV_ROPE, rope=(cos, sin, pos, hd, 2 * D))
        else:
            qkv = ops.linear(h, w, b)
            if adapters:
                for i, proj in enumerate((self.q, self.k, self.v)):
                    _adapters(proj, h, qkv[:, i * D:(i + 1) * D], lora_names)
            if self.pre_layernorm or cos is not None:
                ops.qk_norm_rope_(qkv[:, :D], qkv[:, D:2 * D], H, hd,
                                  self.layernorm_q.weight if self.pre_layernorm else None,
                                  self.layernorm_k.weight if self.pre_layernorm else None, cos, sin, pos)
        return qkv, tile_info

    def _attn(self, qkv: Tensor, cu_lens: Tensor, max_len: int, tile_info=None) -> Tensor:
        T = qkv.shape[0]
        D, H, hd = self.embed_dim, self.num_heads, self.head_dim
        q, k, v = (qkv[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
        return ops.attn_varlen(q, k, v, cu_lens, max_len, tile_info)

    def forward(self, x: Tensor, cu_lens, max_len, lora_names=None) -> Tensor:
        qkv, tile_info = self._qkv_rot(x, cu_lens, max_len, lora_names)
        a = self._attn(qkv, cu_lens, max_len, tile_info)
        out = _base(self.out)
        return _adapters(self.out, a, ops.linear(a, dense_weight(out), out.bias), lora_names)


class FlashTransformerLayer(nn.Module):
    """Pre-LN transformer block (reference: esme/attention.py:142-255)."""

    def __init__(self, embed_dim, expand_dim, attention_heads, rotary_embedding=True, pre_layernorm=False,
                 bias=False, residue_scaling=1., final_activation='swiglu', dropout=0.0, dtype=torch.bfloat16):
        super().__init__()
        self.embed_dim = embed_dim
        self.expand_dim = expand_dim
        self.attention_heads = attention_heads
        self.residue_scaling = residue_scaling
        self.final_activation = final_activation
        self.self_attn = FlashMultiheadAttention(embed_dim, attention_heads, pre_layernorm=pre_layernorm, bias=bias,
                                                 dropout=dropout, rotary_embedding=rotary_embedding, dtype=dtype)
        if final_activation == 'swiglu':
            hidden = int(((expand_dim * embed_dim) + 255) // 256 * 256)
            self.final = nn.Sequential(
                nn.LayerNorm(embed_dim, dtype=dtype),
                SwiGLU(embed_dim, hidden, bias=bias, dtype=dtype),
                nn.Linear(hidden, embed_dim, bias=bias, dtype=dtype))
        elif final_activation == 'gelu':
            self.final = nn.Sequential(
                nn.LayerNorm(embed_dim, dtype=dtype),
                nn.Linear(embed_dim, embed_dim * expand_dim, bias=bias, dtype=dtype),
                nn.GELU(),
                nn.Linear(embed_dim * expand_dim, embed_dim, bias=bias, dtype=dtype))
        else:
            raise ValueError('Invalid final activation function. Must be "swiglu" or "gelu".')

    @property
    def ffn_dim(self) -> int:
        return self.final[-1].in_features

    def forward(self, x: Tensor, cu_lens, max_len, lora_names=None) -> Tensor:
        sa, s = self.self_attn, float(self.residue_scaling)
        qkv, tile_info = sa._qkv_rot(x, cu_lens, max_len, lora_names)
        a = sa._attn(qkv, cu_lens, max_len, tile_info)
        out = _base(sa.out)
        if out is sa.out:
            # x + out(a) / s, fused into the out-projection epilogue
            x = ops.linear(a, dense_weight(out), out.bias, epilogue=L.EPI_RESIDUAL, residual=x, residue_scaling=s)
        else:
            # with adapters on the output projection the reference's order applies: o = out(a) + adapters, x + o / s
            o = _adapters(sa.out, a, ops.linear(a, dense_weight(out), out.bias), lora_names)
            x = ops.residual_add(x, o, s)
        ln = self.final[0]
        h = ops.layernorm(x, ln.weight, ln.bias, ln.eps)
        if self.final_activation == 'gelu':
            u = ops.linear(h, dense_weight(self.final[1]), self.final[1].bias, epilogue=L.EPI_BIAS_GELU)
            down = self.final[3]
        else:
            u = self.final[1](h)
            down = self.final[2]
        return ops.linear(u, dense_weight(down), down.bias, epilogue=L.EPI_RESIDUAL, residual=x, residue_scaling=s)
