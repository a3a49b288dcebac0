"""CPU oracle for the ESM forward hot path.  TEST INFRASTRUCTURE ONLY.

This file is a from-scratch restatement (torch CPU tensors, explicit rounding
points) of the arithmetic the reference `esme` package performs on its
inference path.  It is *not* part of the product: only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product (`esm-efficient_b200/esme`) never does and
fails loudly when its CUDA library is missing.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the real
reference from /root/reference (with an `accelerate` stub and
`flash_attn_varlen_func` replaced by a per-sequence SDPA, because the reference
has no CPU attention path) and commits its logits for the fixtures under
`tests/golden/`; `tests/test_oracle.py` checks this restatement against them.

Reference citations (all paths relative to /root/reference):
  embedding ................ esme/esm.py:176-199 (ESM2), esme/esm.py:876 (ESMC)
  layer loop / final LN .... esme/esm.py:243-252, 890-899
  pad / unpad .............. esme/esm.py:235-239, 254-261
  LN -> q,k,v -> QK-LN ...... esme/attention.py:91-110
  rotary tables ............ esme/rotary.py:110-149
  rotary apply ............. esme/rotary.py:5-43
  attention ................ esme/attention.py:112-124 (flash_attn_varlen_func)
  out proj + residual ...... esme/attention.py:136-139, 253-254
  FFN (GELU / SwiGLU) ...... esme/attention.py:217-236, 258-281, 255
  LM head .................. esme/head.py:25-27
  log-softmax .............. esme/esm.py:297
  tokenizer ................ esme/alphabet.py:93, 117-183

Precision modes
  'bf16' : every tensor the reference materialises in bf16 is rounded to bf16
           here (round-to-nearest-even); reductions / GEMMs run in fp32 (fp64
           accumulate optional) exactly as SURVEY.md Appendix A states.
  'fp32' / 'fp64' : no intermediate rounding, used to measure the bf16 noise
           floor (weights are still the bf16 checkpoint values).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------
# vocabulary (esme/alphabet.py:9-56)
# --------------------------------------------------------------------------
_SPECIAL_HEAD = ['<cls>', '<pad>', '<eos>', '<unk>']
_RESIDUES = list('LAGVSERTIDPKQNFYMHWCXBUZO')
ALPHABET_ESM2 = _SPECIAL_HEAD + _RESIDUES + ['.', '-', '<null_1>', '<mask>']
ALPHABET_ESMC = _SPECIAL_HEAD + _RESIDUES + ['.', '-', '|', '<mask>']
CLS, PAD, EOS, UNK, MASK = 0, 1, 2, 3, 32
AMINO_ACIDS = ALPHABET_ESM2[4:24]

_TOKEN_RE = re.compile(r"<[^>]+>|.")  # esme/alphabet.py:93


def _encode(seq: str, table: Sequence[str]) -> List[int]:
    lut = {t: i for i, t in enumerate(table)}
    return [CLS] + [lut.get(t, UNK) for t in _TOKEN_RE.findall(seq)] + [EOS]


def tokenize(sequences, table: Sequence[str] = ALPHABET_ESMC) -> torch.Tensor:
    """Padded tokenizer, esme/alphabet.py:117-145."""
    if isinstance(sequences, str):
        sequences = [sequences]
    rows = [_encode(s, table) for s in sequences]
    width = max(len(r) for r in rows)
    out = torch.full((len(rows), width), PAD, dtype=torch.int64)
    for i, r in enumerate(rows):
        out[i, :len(r)] = torch.tensor(r, dtype=torch.int64)
    return out


def tokenize_unpad(sequences, table: Sequence[str] = ALPHABET_ESMC):
    """Packed tokenizer, esme/alphabet.py:148-183 -> (tokens, indices, cu_lens, max_len)."""
    if isinstance(sequences, str):
        sequences = [sequences]
    rows = [_encode(s, table) for s in sequences]
    lens = [len(r) for r in rows]
    max_len = max(lens)
    cu = np.zeros(len(rows) + 1, dtype=np.int32)
    cu[1:] = np.cumsum(lens)
    tokens = torch.tensor([t for r in rows for t in r], dtype=torch.int64)
    indices = torch.tensor(
        [i * max_len + j for i, l in enumerate(lens) for j in range(l)], dtype=torch.int64)
    return tokens, indices, torch.from_numpy(cu), max_len


# --------------------------------------------------------------------------
# configuration
# --------------------------------------------------------------------------
@dataclass
class OracleConfig:
    family: str            # 'esm2' | 'esmc'
    num_layers: int
    embed_dim: int
    attention_heads: int

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.attention_heads

    @property
    def gelu_family(self) -> bool:
        """ESM2 block (bias, GELU FFN, mask-row zeroing): esm2 and the rotary-free esm1b / esm1v (esm.py:618-735)."""
        return self.family in ('esm2', 'esm1b', 'esm1v')

    @property
    def rotary(self) -> bool:
        return self.family in ('esm2', 'esmc')

    @property
    def ffn_dim(self) -> int:
        if self.gelu_family:
            return 4 * self.embed_dim                       # esme/esm.py:159
        # esme/attention.py:218-219 with expand 8/3 (esme/esm.py:833)
        return int(((8 / 3 * self.embed_dim) + 255) // 256 * 256)

    @property
    def residue_scaling(self) -> float:
        if self.gelu_family:
            return 1.0
        return math.sqrt(self.num_layers / 36)               # esme/esm.py:839

    @property
    def vocab(self) -> int:
        return 33 if self.gelu_family else 64                # esme/esm.py:154,174,828,850


def config_from_metadata(meta: Dict[str, str]) -> OracleConfig:
    """esme/esm.py:328-339."""
    return OracleConfig(meta['name'].split('_')[0], int(meta['num_layers']),
                        int(meta['embed_dim']), int(meta['attention_heads']))


def load_checkpoint(path: str) -> Tuple[OracleConfig, Dict[str, torch.Tensor]]:
    from safetensors import safe_open
    with safe_open(path, framework='pt', device='cpu') as f:
        cfg = config_from_metadata(f.metadata())
        weights = {k: f.get_tensor(k) for k in f.keys()}
    return cfg, weights


# --------------------------------------------------------------------------
# numerics helpers
# --------------------------------------------------------------------------
_WEIGHT_CACHE: Dict[tuple, tuple] = {}


class _Prec:
    def __init__(self, mode: str):
        assert mode in ('bf16', 'fp32', 'fp64')
        self.mode = mode
        self.dt = torch.float64 if mode == 'fp64' else torch.float32

    def r(self, x: torch.Tensor) -> torch.Tensor:
        """Materialise as the reference would: bf16 RNE in 'bf16' mode."""
        if self.mode == 'bf16':
            return x.to(torch.bfloat16).to(self.dt)
        return x

    def w(self, t: torch.Tensor) -> torch.Tensor:
        """Weights widened to the compute dtype (cached: repeated forwards, e.g. the CPU
        baseline of bench.py, must not pay the bf16 -> fp32 conversion every call)."""
        if t.dtype == self.dt:
            return t
        key = (t.data_ptr(), tuple(t.shape), self.dt)
        hit = _WEIGHT_CACHE.get(key)
        if hit is None or hit[0] is not t:
            hit = (t, t.to(self.dt))
            _WEIGHT_CACHE[key] = hit
        return hit[1]


def _layer_norm(x, w, b, eps=1e-5):
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)          # biased variance
    y = (x - mean) * torch.rsqrt(var + eps) * w
    return y + b if b is not None else y


def _gelu(x):                                               # exact erf GELU
    return x * 0.5 * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def rotary_tables(max_len: int, head_dim: int, p: _Prec):
    """esme/rotary.py:110-149: fp32 inv_freq / outer / cos,sin, cast to model dtype."""
    inv_freq = 1.0 / (10000.0 ** (torch.arange(0, head_dim, 2, dtype=torch.float32) / head_dim))
    t = torch.arange(max_len, dtype=torch.float32)
    freqs = torch.outer(t, inv_freq)
    emb = torch.cat((freqs, freqs), dim=-1)
    return p.r(emb.cos().to(p.dt)), p.r(emb.sin().to(p.dt))


def positions_from_cu_lens(cu_lens: torch.Tensor) -> torch.Tensor:
    """esme/rotary.py:5-14: index of each packed token inside its own sequence."""
    cu = cu_lens.to(torch.int64)
    lens = cu[1:] - cu[:-1]
    starts = torch.repeat_interleave(cu[:-1], lens)
    return torch.arange(int(cu[-1]), dtype=torch.int64) - starts


def apply_rotary(x, cos, sin, pos, p: _Prec):
    """esme/rotary.py:17-43.  x [T,H,hd]; three bf16 roundings in 'bf16' mode."""
    half = x.shape[-1] // 2
    rot = torch.cat((-x[..., half:], x[..., :half]), dim=-1)
    c = cos[pos].unsqueeze(1)
    s = sin[pos].unsqueeze(1)
    return p.r(p.r(x * c) + p.r(rot * s))


def varlen_attention(q, k, v, cu_lens, p: _Prec):
    """esme/attention.py:112-124.  flash-attn semantics: fp32 scores, softmax
    scale hd^-0.5, non-causal, un-normalised P cast to bf16 before P@V, fp32 row
    sum taken from the un-rounded P, output rounded once."""
    T, H, hd = q.shape
    out = torch.empty_like(q)
    scale = hd ** -0.5
    cu = cu_lens.tolist()
    for a, b in zip(cu[:-1], cu[1:]):
        qs, ks, vs = (t[a:b].transpose(0, 1) for t in (q, k, v))    # [H,L,hd]
        s = torch.matmul(qs, ks.transpose(1, 2)) * scale
        m = s.max(-1, keepdim=True).values
        e = torch.exp(s - m)
        o = torch.matmul(p.r(e), vs) / e.sum(-1, keepdim=True)
        out[a:b] = o.transpose(0, 1)
    return p.r(out)


def attention_pool(cls, embed, k_weight, k_bias, cu_lens, num_heads: int, p: _Prec):
    """esme/pooling.py:72-136 `AttentionPool.forward`: class tokens [C,D] attend to the tokens of every sequence;
    keys = k(embed), values = embed; flash-attn semantics as in `varlen_attention`.  -> [B, C, D]."""
    C, D = cls.shape
    hd = D // num_heads
    k = _linear(embed, k_weight, k_bias, p)
    cu = cu_lens.tolist()
    out = torch.empty(len(cu) - 1, C, D, dtype=p.dt)
    q = cls.to(p.dt).reshape(C, num_heads, hd).transpose(0, 1)                   # [H,C,hd]
    for s, (a, b) in enumerate(zip(cu[:-1], cu[1:])):
        ks = k[a:b].reshape(b - a, num_heads, hd).transpose(0, 1)                # [H,L,hd]
        vs = embed[a:b].to(p.dt).reshape(b - a, num_heads, hd).transpose(0, 1)
        sc = torch.matmul(q, ks.transpose(1, 2)) * hd ** -0.5
        e = torch.exp(sc - sc.max(-1, keepdim=True).values)
        o = torch.matmul(p.r(e), vs) / e.sum(-1, keepdim=True)                   # [H,C,hd]
        out[s] = o.transpose(0, 1).reshape(C, D)
    return p.r(out)


# --------------------------------------------------------------------------
# the forward pass
# --------------------------------------------------------------------------
def _linear(x, w, b, p: _Prec):
    y = x @ p.w(w).t()
    if b is not None:
        y = y + p.w(b)
    return p.r(y)


def layer_forward(x, W: Dict[str, torch.Tensor], i: int, cfg: OracleConfig,
                  cu_lens, pos, cos, sin, p: _Prec, taps: Optional[dict] = None):
    """One FlashTransformerLayer, esme/attention.py:241-255."""
    g = lambda k: W.get(f'layers.{i}.{k}')
    H, hd, s = cfg.attention_heads, cfg.head_dim, cfg.residue_scaling
    h = p.r(_layer_norm(x, p.w(g('self_attn.norm.weight')), p.w(g('self_attn.norm.bias'))))
    q = _linear(h, g('self_attn.q.weight'), g('self_attn.q.bias'), p)
    k = _linear(h, g('self_attn.k.weight'), g('self_attn.k.bias'), p)
    v = _linear(h, g('self_attn.v.weight'), g('self_attn.v.bias'), p)
    if cfg.family == 'esmc':                                  # attention.py:104-105
        q = p.r(_layer_norm(q, p.w(g('self_attn.layernorm_q.weight')), None))
        k = p.r(_layer_norm(k, p.w(g('self_attn.layernorm_k.weight')), None))
    T = x.shape[0]
    q, k, v = (t.reshape(T, H, hd) for t in (q, k, v))
    if cfg.rotary:                                            # ESM-1b / 1v: rotary_embedding=False (esm.py:628-629)
        q = apply_rotary(q, cos, sin, pos, p)
        k = apply_rotary(k, cos, sin, pos, p)
    a = varlen_attention(q, k, v, cu_lens, p).reshape(T, H * hd)
    o = _linear(a, g('self_attn.out.weight'), g('self_attn.out.bias'), p)
    x = p.r(x + p.r(o / s))
    if taps is not None:
        taps[f'layer{i}.q_rot'] = q
        taps[f'layer{i}.k_rot'] = k
        taps[f'layer{i}.attn'] = a
        taps[f'layer{i}.x_mid'] = x
    gln = p.r(_layer_norm(x, p.w(g('final.0.weight')), p.w(g('final.0.bias'))))
    if cfg.gelu_family:                                       # attention.py:228-236
        u = _linear(gln, g('final.1.weight'), g('final.1.bias'), p)
        u = p.r(_gelu(u))
        y = _linear(u, g('final.3.weight'), g('final.3.bias'), p)
    else:                                                     # attention.py:221-227, 281
        act = _linear(gln, g('final.1.activation.weight'), None, p)
        fc = _linear(gln, g('final.1.fc.weight'), None, p)
        u = p.r(p.r(act * torch.sigmoid(act)) * fc)
        y = _linear(u, g('final.2.weight'), None, p)
    return p.r(x + p.r(y / s))


def lm_head(z, W, p: _Prec):
    """esme/head.py:25-27."""
    d = p.r(_gelu(_linear(z, W['lm_head.dense.weight'], W['lm_head.dense.bias'], p)))
    d = p.r(_layer_norm(d, p.w(W['lm_head.layer_norm.weight']), p.w(W['lm_head.layer_norm.bias'])))
    return _linear(d, W['lm_head.final.weight'], W['lm_head.final.bias'], p)


def forward_packed(cfg: OracleConfig, W: Dict[str, torch.Tensor], tokens: torch.Tensor,
                   cu_lens: torch.Tensor, max_len: int, mode: str = 'bf16',
                   return_repr: bool = False, taps: Optional[dict] = None,
                   zero_rows: Optional[torch.Tensor] = None):
    """model(tokens_1d, (cu_lens, max_len)) -> logits [T,V] (esme/esm.py:268-282)."""
    p = _Prec(mode)
    assert tokens.ndim == 1
    x = p.w(W['embed_tokens.weight'])[tokens]
    if cfg.gelu_family:                                       # esm.py:189 (ESMC: esm.py:876, no zeroing)
        x = x.masked_fill((tokens == MASK).unsqueeze(-1), 0.0)
    pos = positions_from_cu_lens(cu_lens)
    if not cfg.rotary:
        # ESM-1b / ESM-1v (esm.py:634-656, 696-714; embedding.py:54-92): learned positions count from
        # padding_idx + 1 = 2 inside each sequence; ESM-1b then applies emb_layer_norm_before
        x = p.r(x + p.w(W['embed_positions.weight'])[pos + 2])
        if 'emb_layer_norm_before.weight' in W:
            x = p.r(_layer_norm(x, p.w(W['emb_layer_norm_before.weight']), p.w(W['emb_layer_norm_before.bias'])))
    if zero_rows is not None:
        x = x.masked_fill(zero_rows.unsqueeze(-1), 0.0)
    cos, sin = rotary_tables(max_len, cfg.head_dim, p)
    for i in range(cfg.num_layers):
        x = layer_forward(x, W, i, cfg, cu_lens, pos, cos, sin, p, taps)
        if taps is not None:
            taps[f'layer{i}.x_out'] = x
    z = p.r(_layer_norm(x, p.w(W['emb_layer_norm_after.weight']),
                        p.w(W['emb_layer_norm_after.bias']) if 'emb_layer_norm_after.bias' in W else None))
    if return_repr:
        return z
    return lm_head(z, W, p)


def forward_padded(cfg, W, tokens2d: torch.Tensor, mode: str = 'bf16'):
    """model(tokens_2d) -> logits [B,S,V]: unpad, run, re-pad with zero rows, then
    the LM head runs on every row incl. pads (esme/esm.py:191-193,235-239,254-255,281)."""
    p = _Prec(mode)
    B, S = tokens2d.shape
    keep = tokens2d != PAD
    lens = keep.sum(1).to(torch.int32)
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    flat_idx = torch.nonzero(keep.flatten(), as_tuple=False).flatten()
    max_len = int(lens.max())
    z = forward_packed(cfg, W, tokens2d.flatten()[flat_idx], cu, max_len, mode, return_repr=True)
    full = torch.zeros(B * max_len, cfg.embed_dim, dtype=z.dtype)
    rows = torch.cat([torch.arange(int(l)) + b * max_len for b, l in enumerate(lens)])
    full[rows] = z
    return lm_head(full, W, p).reshape(B, max_len, cfg.vocab)


def log_softmax(logits: torch.Tensor, mode: str = 'bf16'):
    """esme/esm.py:297: log_softmax on bf16 logits, bf16 out."""
    return _Prec(mode).r(torch.log_softmax(logits, dim=-1))


# --------------------------------------------------------------------------
# synthetic checkpoints / batches -- shared by tests and bench
# --------------------------------------------------------------------------
def synthetic_weights(cfg: OracleConfig, seed: int = 1, qk_gain: Optional[float] = None) -> Dict[str, torch.Tensor]:
    """Seeded random-init bf16 weights with the reference's key schema
    (SURVEY.md §3.2).  Linear ~ N(0, 0.02^2) (q,k scaled by qk_gain, default
    sqrt(2/(D*0.02^2)) -> attention logits of std ~2: non-uniform but well
    conditioned), LN weight 1+N(0,0.02^2), biases N(0,0.02^2)."""
    g = torch.Generator().manual_seed(seed)
    if qk_gain is None:
        qk_gain = math.sqrt(2.0 / (cfg.embed_dim * 0.02 ** 2))
    D, F, V = cfg.embed_dim, cfg.ffn_dim, cfg.vocab
    bias = cfg.family == 'esm2'
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
    for i in range(cfg.num_layers):
        pre = f'layers.{i}'
        ln(f'{pre}.self_attn.norm')
        lin(f'{pre}.self_attn.q', D, D, qk_gain)
        lin(f'{pre}.self_attn.k', D, D, qk_gain)
        lin(f'{pre}.self_attn.v', D, D)
        lin(f'{pre}.self_attn.out', D, D)
        ln(f'{pre}.final.0')
        if cfg.family == 'esm2':
            lin(f'{pre}.final.1', F, D)
            lin(f'{pre}.final.3', D, F)
        else:
            ln(f'{pre}.self_attn.layernorm_q', with_bias=False)
            ln(f'{pre}.self_attn.layernorm_k', with_bias=False)
            lin(f'{pre}.final.1.activation', F, D)
            lin(f'{pre}.final.1.fc', F, D)
            lin(f'{pre}.final.2', D, F)
    ln('emb_layer_norm_after', with_bias=(cfg.family == 'esm2'))
    lin('lm_head.dense', D, D, with_bias=True)
    ln('lm_head.layer_norm')
    lin('lm_head.final', V, D, gain=5.0, with_bias=True)
    return W


def save_checkpoint(path: str, cfg: OracleConfig, W: Dict[str, torch.Tensor], tag: str = 'synthetic'):
    from safetensors.torch import save_file
    meta = {'name': f'{cfg.family}_{tag}', 'num_layers': str(cfg.num_layers),
            'embed_dim': str(cfg.embed_dim), 'attention_heads': str(cfg.attention_heads), 'format': 'pt'}
    save_file({k: v.contiguous() for k, v in W.items()}, path, metadata=meta)


def synthetic_lengths(budget: int, seed: int, dist: str = 'lognormal') -> List[int]:
    """sequence lengths (incl. cls/eos) greedily packed in draw order
    until the next one would exceed `budget` tokens (TokenSizeBatchSampler rule,
    esme/data.py:42-51)."""
    g = torch.Generator().manual_seed(seed)
    out, tot = [], 0
    while True:
        if dist == 'lognormal':       # residues ~ round(LogNormal(ln 400, 0.75)) clipped [30,3500]
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
    """Random residues over the 20 standard amino acids with <cls>/<eos> framing."""
    g = torch.Generator().manual_seed(seed)
    toks = []
    for l in lens:
        body = torch.randint(4, 24, (l - 2,), generator=g, dtype=torch.int64)
        toks.append(torch.cat([torch.tensor([CLS]), body, torch.tensor([EOS])]))
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int32), 0)
    return torch.cat(toks), cu, int(max(lens))
