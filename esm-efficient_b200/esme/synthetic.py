"""Seeded synthetic checkpoints and packed batches for benchmarks and full-size
tests (no network: the 650M / 3B / ESMC weights cannot be downloaded).  The
recipes are those of SURVEY.md §8d; `tests/test_synthetic.py` pins them against
the test infrastructure's own copy.

The batch rule is the reference's token-budget sampler (esme/data.py:42-51):
sequences are appended in draw order until the next one would exceed the budget.
"""
import math
from typing import Dict, List, Sequence

import torch

CLS, EOS = 0, 2


def ffn_dim(family: str, embed_dim: int) -> int:
    if family == 'esm2':
        return 4 * embed_dim
    return int(((8 / 3 * embed_dim) + 255) // 256 * 256)


def synthetic_state_dict(family: str, num_layers: int, embed_dim: int, seed: int = 1,
                         qk_gain: float = None) -> Dict[str, torch.Tensor]:
    """bf16 tensors keyed like the reference's checkpoints.  Linear ~ N(0, 0.02^2) (q, k scaled by
    qk_gain, default sqrt(2 / (D * 0.02^2)) so attention logits have std ~2: a non-trivial but
    well-conditioned softmax), LayerNorm weight 1 + N(0, 0.02^2), biases N(0, 0.02^2)."""
    g = torch.Generator().manual_seed(seed)
    D, F = embed_dim, ffn_dim(family, embed_dim)
    V = 33 if family == 'esm2' else 64
    bias = family == 'esm2'
    if qk_gain is None:
        qk_gain = math.sqrt(2.0 / (D * 0.02 ** 2))
    W: Dict[str, torch.Tensor] = {}

    def lin(name, n_out, n_in, gain=1.0, with_bias=bias):
        W[f'{name}.weight'] = (torch.randn(n_out, n_in, generator=g) * 0.02 * gain).to(torch.bfloat16)
        if with_bias:
            W[f'{name}.bias'] = (torch.randn(n_out, generator=g) * 0.02).to(torch.bfloat16)

    def ln(name, with_bias=True):
        W[f'{name}.weight'] = (1 + torch.randn(D, generator=g) * 0.02).to(torch.bfloat16)
        if with_bias:
            W[f'{name}.bias'] = (torch.randn(D, generator=g) * 0.02).to(torch.bfloat16)

    W['embed_tokens.weight'] = (torch.randn(V, D, generator=g) * 0.5).to(torch.bfloat16)
    for i in range(num_layers):
        pre = f'layers.{i}'
        ln(f'{pre}.self_attn.norm')
        lin(f'{pre}.self_attn.q', D, D, qk_gain)
        lin(f'{pre}.self_attn.k', D, D, qk_gain)
        lin(f'{pre}.self_attn.v', D, D)
        lin(f'{pre}.self_attn.out', D, D)
        ln(f'{pre}.final.0')
        if family == 'esm2':
            lin(f'{pre}.final.1', F, D)
            lin(f'{pre}.final.3', D, F)
        else:
            ln(f'{pre}.self_attn.layernorm_q', with_bias=False)
            ln(f'{pre}.self_attn.layernorm_k', with_bias=False)
            lin(f'{pre}.final.1.activation', F, D)
            lin(f'{pre}.final.1.fc', F, D)
            lin(f'{pre}.final.2', D, F)
    ln('emb_layer_norm_after', with_bias=(family == 'esm2'))
    lin('lm_head.dense', D, D, with_bias=True)
    ln('lm_head.layer_norm')
    lin('lm_head.final', V, D, gain=5.0, with_bias=True)
    return W


def synthetic_lengths(budget: int, seed: int, dist: str = 'lognormal') -> List[int]:
    """Sequence lengths INCLUDING <cls>/<eos>, greedily packed up to `budget` tokens."""
    g = torch.Generator().manual_seed(seed)
    out, tot = [], 0
    while True:
        if dist == 'lognormal':       # residues ~ round(LogNormal(ln 400, 0.75)) clipped to [30, 3500]
            n = int(torch.exp(torch.randn(1, generator=g) * 0.75 + math.log(400.0)).round().clamp(30, 3500))
        elif dist == 'loguniform':    # residues log-uniform in [128, 2048]
            u = float(torch.rand(1, generator=g))
            n = int(round(math.exp(math.log(128) + u * (math.log(2048) - math.log(128)))))
        else:
            raise ValueError(dist)
        if tot + n + 2 > budget:
            break
        out.append(n + 2)
        tot += n + 2
    return out


def synthetic_batch(lens: Sequence[int], seed: int):
    """-> (tokens int64[T], cu_lens int32[B+1], max_len): uniform residues over the 20 standard amino acids."""
    g = torch.Generator().manual_seed(seed)
    toks = []
    for l in lens:
        body = torch.randint(4, 24, (l - 2,), generator=g, dtype=torch.int64)
        toks.append(torch.cat([torch.tensor([CLS]), body, torch.tensor([EOS])]))
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int32), 0)
    return torch.cat(toks), cu, int(max(lens))


def forward_flops(family: str, num_layers: int, embed_dim: int, lens: Sequence[int]) -> Dict[str, float]:
    """Algorithmic FLOPs of one forward (SURVEY.md §8d, 2 FLOP per MAC)."""
    D, F = embed_dim, ffn_dim(family, embed_dim)
    V = 33 if family == 'esm2' else 64
    T = float(sum(lens))
    ffn = 16 * D * D if family == 'esm2' else 6 * D * F
    gemm = T * num_layers * (8 * D * D + ffn)
    head = T * (2 * D * D + 2 * D * V)
    attn = num_layers * 4.0 * D * float(sum(l * l for l in lens))
    return dict(gemm=gemm, head=head, attention=attn, total=gemm + head + attn)
