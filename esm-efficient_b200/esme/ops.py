"""Operator-level Python wrappers over the C ABI: torch tensors in, torch tensors
out, work enqueued on the current CUDA stream.  torch is used for device memory
and streams only; every arithmetic kernel is in libesmk.so."""
from typing import Optional, Tuple

import torch

from . import _lib as L

bf16 = torch.bfloat16


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                'esme (B200 build) runs on CUDA tensors only: the sm_100a kernels in libesmk.so are the '
                'only backend and there is no CPU fallback. Move the model and inputs to a CUDA device.')


def _on_tensor_device(fn):
    """Run `fn` with the CUDA device of its tensor arguments current: the launch then goes to that device's
    context and to ITS current stream (a process may hold models on several GPUs).  All CUDA tensor arguments
    must live on one device."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in list(args) + list(kwargs.values()):
            for t in (a if isinstance(a, (tuple, list)) else (a,)):
                if isinstance(t, torch.Tensor) and t.is_cuda:
                    if dev is None:
                        dev = t.device
                    elif t.device != dev:
                        raise RuntimeError(f'{fn.__name__}: tensor arguments live on different devices ({dev} and {t.device})')
        if dev is None and isinstance(kwargs.get('device', None), (torch.device, str, int)):
            dev = torch.device(kwargs['device']) if not isinstance(kwargs['device'], int) else torch.device('cuda', kwargs['device'])
        if dev is None or dev.type != 'cuda' or dev == torch.device('cuda', torch.cuda.current_device()):
            return fn(*args, **kwargs)
        with torch.cuda.device(dev):
            return fn(*args, **kwargs)
    return wrapper


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: torch.Tensor) -> Tuple[int, int]:
    """(row count, row pitch) of a tensor whose last dim is contiguous and whose
    leading dims form one uniformly strided row index."""
    assert t.stride(-1) == 1, 'last dim must be contiguous'
    if t.ndim == 1:
        return 1, t.shape[0]
    t2 = t if t.ndim == 2 else t.flatten(0, -2)
    return t2.shape[0], (t2.stride(0) if t2.shape[0] > 1 else t2.shape[1])


@_on_tensor_device
def batch_meta(cu_lens: torch.Tensor, T: int):
    """-> (pos int32[T], tile_info int32[capacity, 4]): per-token positions (replaces esme/rotary.py:5-14
    culen_indices) and the attention work list, one {seq start, seq length, first query row, seq id}
    record per 128-query tile, longest sequences first."""
    _need_cuda(cu_lens)
    assert cu_lens.dtype == torch.int32 and cu_lens.is_contiguous()
    B = cu_lens.numel() - 1
    pos = torch.empty(T, dtype=torch.int32, device=cu_lens.device)
    tile_info = torch.empty(L.lib.esmk_tile_capacity(T, B), 4, dtype=torch.int32, device=cu_lens.device)
    L.check(L.lib.esmk_batch_meta(cu_lens.data_ptr(), B, T, pos.data_ptr(), tile_info.data_ptr(), _stream()),
            'esmk_batch_meta')
    return pos, tile_info


def rope_tables(max_len: int, head_dim: int, device) -> Tuple[torch.Tensor, torch.Tensor]:
    cos = torch.empty(max_len, head_dim, dtype=bf16, device=device)
    sin = torch.empty_like(cos)
    _need_cuda(cos)
    with torch.cuda.device(cos.device):
        L.check(L.lib.esmk_rope_tables(cos.data_ptr(), sin.data_ptr(), max_len, head_dim, _stream()), 'esmk_rope_tables')
    return cos, sin


@_on_tensor_device
def embed(tokens: torch.Tensor, table: torch.Tensor, zero_token: int = -1,
          zero_rows: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(tokens, table)
    assert tokens.dtype == torch.int64 and table.dtype == bf16 and table.is_contiguous()
    flat = tokens.reshape(-1).contiguous()
    out = torch.empty(flat.numel(), table.shape[1], dtype=bf16, device=table.device)
    zr = None if zero_rows is None else zero_rows.reshape(-1).to(torch.uint8).contiguous()
    L.check(L.lib.esmk_embed(flat.data_ptr(), table.data_ptr(), out.data_ptr(), flat.numel(), table.shape[1],
                             table.shape[0], zero_token, _ptr(zr), _stream()), 'esmk_embed')
    return out.reshape(*tokens.shape, table.shape[1])


@_on_tensor_device
def layernorm(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], eps: float = 1e-5) -> torch.Tensor:
    _need_cuda(x, weight)
    assert x.dtype == bf16 and weight.dtype == bf16
    D = x.shape[-1]
    xc = x if x.stride(-1) == 1 else x.contiguous()
    T, ldx = _rows(xc)
    y = torch.empty(x.shape, dtype=bf16, device=x.device)
    L.check(L.lib.esmk_layernorm(xc.data_ptr(), ldx, weight.data_ptr(), _ptr(bias), y.data_ptr(), D, T, D,
                                 eps, _stream()), 'esmk_layernorm')
    return y


@_on_tensor_device
def qk_norm_rope_(q: torch.Tensor, k: torch.Tensor, H: int, head_dim: int,
                  ln_q_weight=None, ln_k_weight=None, cos=None, sin=None, pos=None):
    """In place on q and k ([T, H*hd] views sharing one row pitch)."""
    _need_cuda(q, k)
    assert q.dtype == bf16 and k.dtype == bf16 and q.shape == k.shape
    q2, k2 = q.reshape(q.shape[0], -1), k.reshape(k.shape[0], -1)
    assert q2.data_ptr() == q.data_ptr() and k2.data_ptr() == k.data_ptr(), 'q/k must be viewable as [T, D]'
    T, ld = _rows(q2)
    assert _rows(k2)[1] == ld
    L.check(L.lib.esmk_qk_norm_rope(q2.data_ptr(), k2.data_ptr(), ld, T, H, head_dim, _ptr(ln_q_weight),
                                    _ptr(ln_k_weight), _ptr(cos), _ptr(sin), _ptr(pos), _stream()),
            'esmk_qk_norm_rope')
    return q, k


@_on_tensor_device
def mean_pool(x: torch.Tensor, cu_lens: torch.Tensor) -> torch.Tensor:
    """[T, D] packed rows + cu_lens int32[B+1] -> [B, D] per-sequence means (esme/pooling.py:44)."""
    _need_cuda(x, cu_lens)
    assert x.dtype == bf16 and x.ndim == 2 and x.stride(1) == 1
    B, D = cu_lens.numel() - 1, x.shape[1]
    cu = cu_lens.to(torch.int32).contiguous()
    out = torch.empty(B, D, dtype=bf16, device=x.device)
    L.check(L.lib.esmk_mean_pool(x.data_ptr(), x.stride(0), cu.data_ptr(), B, D, out.data_ptr(), D, _stream()),
            'esmk_mean_pool')
    return out


@_on_tensor_device
def softmax(logits: torch.Tensor, log: bool) -> torch.Tensor:
    _need_cuda(logits)
    assert logits.dtype == bf16
    xc = logits.contiguous()
    V = xc.shape[-1]
    T = xc.numel() // V
    out = torch.empty_like(xc)
    L.check(L.lib.esmk_softmax(xc.data_ptr(), V, out.data_ptr(), V, T, V, int(log), _stream()), 'esmk_softmax')
    return out


@_on_tensor_device
def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
           epilogue: int = L.EPI_BIAS, residual: Optional[torch.Tensor] = None, residue_scaling: float = 1.0,
           rope: Optional[tuple] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = epilogue(x @ weight.T + bias) through esmk_gemm (tcgen05).  `rope` =
    (cos, sin, pos, head_dim, rope_cols) for EPI_QKV_ROPE."""
    _need_cuda(x, weight)
    assert x.dtype == bf16 and weight.dtype == bf16 and weight.is_contiguous()
    K = x.shape[-1]
    N = weight.shape[0]
    assert weight.shape[1] == K
    xc = x if x.stride(-1) == 1 else x.contiguous()
    x2 = xc.reshape(-1, K)
    M, lda = _rows(x2)
    n_out = N // 2 if epilogue == L.EPI_SWIGLU else N
    if out is None:
        out = torch.empty(*x.shape[:-1], n_out, dtype=bf16, device=x.device)
    o2 = out.reshape(-1, n_out)
    assert o2.data_ptr() == out.data_ptr()
    a = L.GemmArgs()
    a.A, a.lda, a.W, a.bias = x2.data_ptr(), lda, weight.data_ptr(), _ptr(bias)
    a.C, a.ldc = o2.data_ptr(), _rows(o2)[1]
    a.M, a.N, a.K, a.epilogue = M, N, K, epilogue
    a.residue_scaling = float(residue_scaling)
    if epilogue == L.EPI_RESIDUAL:
        assert residual is not None and residual.dtype == bf16
        r2 = residual.reshape(-1, N)
        a.R, a.ldr = r2.data_ptr(), _rows(r2)[1]
    if epilogue == L.EPI_QKV_ROPE:
        cos, sin, pos, hd, rope_cols = rope
        a.rope_cos, a.rope_sin, a.pos, a.head_dim, a.rope_cols = cos.data_ptr(), sin.data_ptr(), pos.data_ptr(), hd, rope_cols
    L.check(L.lib.esmk_gemm(a, _stream()), 'esmk_gemm')
    return out


@_on_tensor_device
def attn_varlen(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_lens: torch.Tensor, max_len: int,
                tile_info: Optional[torch.Tensor] = None, impl: int = 0) -> torch.Tensor:
    """q,k,v: [T,H,hd] views with a common row pitch (e.g. column blocks of the QKV
    GEMM output) -> [T, H*hd].  Replaces flash_attn_varlen_func (esme/attention.py:115)."""
    _need_cuda(q, k, v, cu_lens)
    T, H, hd = q.shape
    for t in (q, k, v):
        assert t.dtype == bf16 and t.shape == (T, H, hd) and t.stride(2) == 1 and t.stride(1) == hd
    ld = q.stride(0)
    assert k.stride(0) == ld and v.stride(0) == ld
    assert cu_lens.dtype == torch.int32
    B = cu_lens.numel() - 1
    if tile_info is None:
        _, tile_info = batch_meta(cu_lens, T)
    out = torch.empty(T, H * hd, dtype=bf16, device=q.device)
    L.check(L.lib.esmk_attn_varlen(q.data_ptr(), k.data_ptr(), v.data_ptr(), ld, out.data_ptr(), H * hd,
                                   cu_lens.data_ptr(), tile_info.data_ptr(), B, T, H, hd, int(max_len), impl,
                                   _stream()), 'esmk_attn_varlen')
    return out


@_on_tensor_device
def quantize(weight: torch.Tensor, bits: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """bf16 [N,K] weight -> (data, scale) in the library's weight-only formats (include/esmk.h):
    bits=4: uint8 [N*K/2, 1] + fp32 absmax [N*K/64];  bits=8: int8 [N,K] + fp32 [N] (row absmax / 127)."""
    _need_cuda(weight)
    assert weight.dtype == bf16 and weight.ndim == 2 and bits in (4, 8)
    w = weight.contiguous()
    N, K = w.shape
    if bits == 4:
        data = torch.empty(N * K // 2, 1, dtype=torch.uint8, device=w.device)
        scale = torch.empty(N * K // 64, dtype=torch.float32, device=w.device)
    else:
        data = torch.empty(N, K, dtype=torch.int8, device=w.device)
        scale = torch.empty(N, dtype=torch.float32, device=w.device)
    L.check(L.lib.esmk_quantize(w.data_ptr(), N, K, bits, data.data_ptr(), scale.data_ptr(), _stream()), 'esmk_quantize')
    return data, scale


@_on_tensor_device
def dequantize(data: torch.Tensor, scale: torch.Tensor, N: int, K: int, bits: int) -> torch.Tensor:
    """Inverse of `quantize`: bf16 [N,K]."""
    _need_cuda(data, scale)
    assert data.is_contiguous() and scale.is_contiguous() and scale.dtype == torch.float32
    w = torch.empty(N, K, dtype=bf16, device=data.device)
    L.check(L.lib.esmk_dequantize(data.data_ptr(), scale.data_ptr(), N, K, bits, w.data_ptr(), _stream()),
            'esmk_dequantize')
    return w


@_on_tensor_device
def attn_pool(cls: torch.Tensor, k: torch.Tensor, v: torch.Tensor, cu_lens: torch.Tensor, num_heads: int) -> torch.Tensor:
    """Class-token attention pooling (esme/pooling.py:72-136): cls [C, D], k / v [T, D] -> [B, C, D]."""
    _need_cuda(cls, k, v, cu_lens)
    assert cls.dtype == bf16 and k.dtype == bf16 and v.dtype == bf16 and cu_lens.dtype == torch.int32
    assert k.shape == v.shape and k.stride(-1) == 1 and v.stride(-1) == 1 and k.stride(0) == v.stride(0)
    cls = cls.contiguous()
    C_, D = cls.shape
    B = cu_lens.numel() - 1
    out = torch.empty(B, C_, D, dtype=bf16, device=k.device)
    L.check(L.lib.esmk_attn_pool(cls.data_ptr(), D, k.data_ptr(), v.data_ptr(), k.stride(0), out.data_ptr(),
                                 cu_lens.contiguous().data_ptr(), B, C_, num_heads, D // num_heads, _stream()),
            'esmk_attn_pool')
    return out


@_on_tensor_device
def residual_add(x: torch.Tensor, y: torch.Tensor, residue_scaling: float = 1.0) -> torch.Tensor:
    """bf(x + bf(y / residue_scaling)) (esme/attention.py:253-255 as stand-alone ops)."""
    _need_cuda(x, y)
    assert x.dtype == bf16 and y.dtype == bf16 and x.shape == y.shape and x.numel() % 8 == 0
    xc, yc = x.contiguous(), y.contiguous()
    out = torch.empty_like(xc)
    L.check(L.lib.esmk_residual_add(xc.data_ptr(), yc.data_ptr(), out.data_ptr(), xc.numel(), float(residue_scaling),
                                    _stream()), 'esmk_residual_add')
    return out


@_on_tensor_device
def add_positions_(x: torch.Tensor, table: torch.Tensor, pos: torch.Tensor, offset: int = 2) -> torch.Tensor:
    """x[t] += table[pos[t] + offset] in place, one bf16 rounding (ESM-1b / ESM-1v learned positions)."""
    _need_cuda(x, table, pos)
    assert x.dtype == bf16 and table.dtype == bf16 and pos.dtype == torch.int32 and x.is_contiguous() and table.is_contiguous()
    L.check(L.lib.esmk_add_positions(x.data_ptr(), table.data_ptr(), pos.data_ptr(), x.shape[0], x.shape[1],
                                     table.shape[0], offset, _stream()), 'esmk_add_positions')
    return x


@_on_tensor_device
def unpad_tokens(tokens2d: torch.Tensor, pad_token: int):
    """[B,S] int64 token grid -> (packed tokens int64[T], flat grid indices int64[T], cu_lens int32[B+1], max_len):
    the job of flash_attn.bert_padding.unpad_input at esme/esm.py:238, on the device; ONE 8-byte read-back of
    {T, max_len} is the entry's only synchronisation (the reference's path has three)."""
    _need_cuda(tokens2d)
    assert tokens2d.ndim == 2 and tokens2d.dtype == torch.int64
    t2 = tokens2d.contiguous()
    B, S = t2.shape
    dev = t2.device
    packed = torch.empty(B * S, dtype=torch.int64, device=dev)
    indices = torch.empty(B * S, dtype=torch.int64, device=dev)
    cu_lens = torch.empty(B + 1, dtype=torch.int32, device=dev)
    lens = torch.empty(B, dtype=torch.int32, device=dev)
    meta = torch.empty(2, dtype=torch.int32, device=dev)
    L.check(L.lib.esmk_unpad_tokens(t2.data_ptr(), B, S, int(pad_token), packed.data_ptr(), indices.data_ptr(),
                                    cu_lens.data_ptr(), lens.data_ptr(), meta.data_ptr(), _stream()), 'esmk_unpad_tokens')
    T, max_len = (int(v) for v in meta.tolist())
    return packed[:T], indices[:T], cu_lens, max_len


@_on_tensor_device
def pad_rows(x: torch.Tensor, indices: torch.Tensor, rows: int) -> torch.Tensor:
    """[T,D] packed rows -> [rows, D] with x[t] at row indices[t] and zeros elsewhere (pad_input, esme/esm.py:255-261)."""
    _need_cuda(x, indices)
    assert x.dtype == bf16 and x.ndim == 2 and x.stride(1) == 1 and indices.dtype == torch.int64
    out = torch.empty(rows, x.shape[1], dtype=bf16, device=x.device)
    inverse = torch.empty(max(rows, 1), dtype=torch.int32, device=x.device)
    L.check(L.lib.esmk_pad_rows(x.data_ptr(), x.stride(0) if x.shape[0] > 1 else x.shape[1], indices.contiguous().data_ptr(),
                                x.shape[0], out.data_ptr(), rows, x.shape[1], inverse.data_ptr(), _stream()), 'esmk_pad_rows')
    return out
