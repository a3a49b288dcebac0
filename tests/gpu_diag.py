"""GPU bring-up diagnostics: runs every kernel family against the oracle / torch
fp32, one SUBPROCESS per case (a trap in one kernel cannot poison the others),
each under a timeout.  Writes gpurun_out/diag.json and prints a table.

    python tests/gpu_diag.py            # all cases
    python tests/gpu_diag.py --case gemm_bias_small
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))
sys.path.insert(0, os.path.join(ROOT, 'tests'))

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _imports():
    import torch
    from esme import ops, _lib as L
    from oracle import esm_oracle as O
    return torch, ops, L, O


def _cmp(name, got, want, out):
    import torch
    got, want = got.float().cpu(), want.float().cpu()
    diff = (got - want).abs()
    denom = want.abs().max().item() + 1e-12
    out[name] = dict(max_abs=diff.max().item(), max_rel_to_peak=diff.max().item() / denom,
                     rms_rel=(diff.pow(2).mean().sqrt() / (want.pow(2).mean().sqrt() + 1e-12)).item(),
                     mismatch_frac=(diff > 0).float().mean().item(), finite=bool(torch.isfinite(got).all()))
    return out[name]


def _where_bad(got, want, tol):
    """Rows / cols pattern of mismatches, for descriptor debugging."""
    bad = ((got.float().cpu() - want.float().cpu()).abs() > tol)
    rows = bad.any(1).nonzero().flatten().tolist()
    cols = bad.any(0).nonzero().flatten().tolist()
    return dict(n_bad=int(bad.sum()), rows=rows[:16], n_rows=len(rows), cols=cols[:16], n_cols=len(cols))


# --------------------------------------------------------------------------
@case
def rowops():
    torch, ops, L, O = _imports()
    dev = 'cuda'
    out = {}
    g = torch.Generator().manual_seed(0)
    p = O._Prec('bf16')
    for D in (320, 960, 1280, 2560):
        x = (torch.randn(777, D, generator=g) * 2 + 0.3).bfloat16()
        w = (1 + 0.1 * torch.randn(D, generator=g)).bfloat16()
        b = (0.1 * torch.randn(D, generator=g)).bfloat16()
        want = p.r(O._layer_norm(x.float(), w.float(), b.float()))
        _cmp(f'layernorm_D{D}', ops.layernorm(x.to(dev), w.to(dev), b.to(dev)), want, out)
        want = p.r(O._layer_norm(x.float(), w.float(), None))
        _cmp(f'layernorm_nobias_D{D}', ops.layernorm(x.to(dev), w.to(dev), None), want, out)
    # embedding
    table = torch.randn(33, 320, generator=g).bfloat16()
    tok = torch.randint(0, 33, (500,), generator=g)
    want = table[tok].float().masked_fill((tok == 32)[:, None], 0)
    _cmp('embed', ops.embed(tok.to(dev), table.to(dev), zero_token=32), want, out)
    # softmax
    lg = (torch.randn(1000, 33, generator=g) * 5).bfloat16()
    _cmp('log_softmax', ops.softmax(lg.to(dev), True), p.r(torch.log_softmax(lg.float(), -1)), out)
    _cmp('softmax', ops.softmax(lg.to(dev), False), p.r(torch.softmax(lg.float(), -1)), out)
    lg = (torch.randn(1000, 64, generator=g) * 5).bfloat16()
    _cmp('log_softmax64', ops.softmax(lg.to(dev), True), p.r(torch.log_softmax(lg.float(), -1)), out)
    # rope tables + apply
    for hd in (16, 32, 64, 128):
        cos, sin = ops.rope_tables(700, hd, dev)
        wc, ws = O.rotary_tables(700, hd, p)
        _cmp(f'rope_cos_hd{hd}', cos, wc, out)
        _cmp(f'rope_sin_hd{hd}', sin, ws, out)
        H = 256 // hd if hd < 128 else 3
        cu = torch.tensor([0, 60, 100, 700, 702], dtype=torch.int32)
        T = 702
        q = torch.randn(T, H, hd, generator=g).bfloat16()
        k = torch.randn(T, H, hd, generator=g).bfloat16()
        pos = O.positions_from_cu_lens(cu)
        wq = O.apply_rotary(q.float(), wc, ws, pos, p)
        wk = O.apply_rotary(k.float(), wc, ws, pos, p)
        dpos, _ = ops.batch_meta(cu.to(dev), T)
        out[f'pos_ok_hd{hd}'] = bool((dpos.cpu() == pos.int()).all())
        qd, kd = q.to(dev).reshape(T, -1).clone(), k.to(dev).reshape(T, -1).clone()
        ops.qk_norm_rope_(qd, kd, H, hd, cos=wc.bfloat16().to(dev), sin=ws.bfloat16().to(dev), pos=dpos)
        _cmp(f'rope_q_hd{hd}', qd.reshape(T, H, hd), wq, out)
        _cmp(f'rope_k_hd{hd}', kd.reshape(T, H, hd), wk, out)
    # ESMC qk-layernorm + rope
    H, hd, T = 15, 64, 333
    D = H * hd
    q = torch.randn(T, D, generator=g).bfloat16()
    k = torch.randn(T, D, generator=g).bfloat16()
    wq_ = (1 + 0.1 * torch.randn(D, generator=g)).bfloat16()
    wk_ = (1 + 0.1 * torch.randn(D, generator=g)).bfloat16()
    cu = torch.tensor([0, 33, 333], dtype=torch.int32)
    pos = O.positions_from_cu_lens(cu)
    wc, ws = O.rotary_tables(300, hd, p)
    eq = O.apply_rotary(p.r(O._layer_norm(q.float(), wq_.float(), None)).reshape(T, H, hd), wc, ws, pos, p)
    ek = O.apply_rotary(p.r(O._layer_norm(k.float(), wk_.float(), None)).reshape(T, H, hd), wc, ws, pos, p)
    qkv = torch.cat([q, k, k], 1).to(dev)
    dpos, _ = ops.batch_meta(cu.to(dev), T)
    ops.qk_norm_rope_(qkv[:, :D], qkv[:, D:2 * D], H, hd, wq_.to(dev), wk_.to(dev), wc.bfloat16().to(dev),
                      ws.bfloat16().to(dev), dpos)
    _cmp('esmc_qkln_rope_q', qkv[:, :D].reshape(T, H, hd), eq, out)
    _cmp('esmc_qkln_rope_k', qkv[:, D:2 * D].reshape(T, H, hd), ek, out)
    return out


def _gemm_case(M, N, K, epi, seed=0, hd=64, scale=1.0, bias=True):
    torch, ops, L, O = _imports()
    dev = 'cuda'
    g = torch.Generator().manual_seed(seed)
    p = O._Prec('bf16')
    x = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / K ** 0.5).bfloat16()
    b = torch.randn(N, generator=g).bfloat16() if bias else None
    acc = x.double() @ w.double().t()
    if b is not None:
        acc = acc + b.double()
    y = p.r(acc.float())
    kw = {}
    if epi == L.EPI_BIAS_GELU:
        want = p.r(O._gelu(y))
    elif epi == L.EPI_RESIDUAL:
        r = torch.randn(M, N, generator=g).bfloat16()
        want = p.r(r.float() + p.r(y / scale))
        kw = dict(residual=r.to(dev), residue_scaling=scale)
    elif epi == L.EPI_QKV_ROPE:
        D = N // 3
        H = D // hd
        lens = []
        left = M
        while left > 0:
            l = min(left, 37 + 61 * len(lens))
            lens.append(l)
            left -= l
        cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(torch.tensor(lens), 0)
        pos = O.positions_from_cu_lens(cu)
        wc, ws = O.rotary_tables(max(lens), hd, p)
        q = O.apply_rotary(y[:, :D].reshape(M, H, hd), wc, ws, pos, p).reshape(M, D)
        k = O.apply_rotary(y[:, D:2 * D].reshape(M, H, hd), wc, ws, pos, p).reshape(M, D)
        want = torch.cat([q, k, y[:, 2 * D:]], 1)
        kw = dict(rope=(wc.bfloat16().to(dev), ws.bfloat16().to(dev), pos.int().to(dev), hd, 2 * D))
    elif epi == L.EPI_SWIGLU:
        F_ = N // 2
        a, f = y[:, :F_], y[:, F_:]
        want = p.r(p.r(a * torch.sigmoid(a)) * f)
        from esme.attention import SwiGLU
        w = SwiGLU.interleave(w[:F_], w[F_:])
    else:
        want = y
    got = ops.linear(x.to(dev), w.to(dev), None if b is None else b.to(dev), epilogue=epi, **kw)
    torch.cuda.synchronize()
    return got, want


@case
def gemm_bias_small():
    torch, ops, L, O = _imports()
    out = {}
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 512, 320), (100, 40, 64)]:
        got, want = _gemm_case(M, N, K, L.EPI_BIAS)
        s = _cmp(f'gemm_{M}x{N}x{K}', got, want, out)
        if s['max_rel_to_peak'] > 0.02:
            out[f'gemm_{M}x{N}x{K}_where'] = _where_bad(got, want, 0.05 * want.abs().max().item())
    return out


@case
def gemm_shapes():
    torch, ops, L, O = _imports()
    out = {}
    for (M, N, K) in [(300, 320, 320), (1000, 3840, 1280), (777, 1280, 5120), (4660, 960, 960), (132, 33, 320),
                      (500, 64, 960), (129, 2560, 2560), (50, 1280, 1288)]:
        got, want = _gemm_case(M, N, K, L.EPI_BIAS, seed=M)
        _cmp(f'gemm_{M}x{N}x{K}', got, want, out)
    return out


@case
def gemm_epilogues():
    torch, ops, L, O = _imports()
    out = {}
    got, want = _gemm_case(700, 1280, 320, L.EPI_BIAS_GELU)
    _cmp('gelu', got, want, out)
    got, want = _gemm_case(700, 320, 1280, L.EPI_RESIDUAL)
    _cmp('residual', got, want, out)
    got, want = _gemm_case(700, 960, 2560, L.EPI_RESIDUAL, scale=0.9128709291752769, bias=False)
    _cmp('residual_scaled_nobias', got, want, out)
    got, want = _gemm_case(900, 2 * 2560, 960, L.EPI_SWIGLU, bias=False)
    _cmp('swiglu', got, want, out)
    for hd, D in ((64, 1280), (16, 320), (32, 640), (64, 960)):
        got, want = _gemm_case(650, 3 * D, D, L.EPI_QKV_ROPE, hd=hd)
        _cmp(f'qkv_rope_hd{hd}_D{D}', got, want, out)
    return out


def _attn_case(lens, H, hd, impl, seed=0, qk_scale=1.0):
    torch, ops, L, O = _imports()
    dev = 'cuda'
    g = torch.Generator().manual_seed(seed)
    p = O._Prec('bf16')
    T = sum(lens)
    D = H * hd
    qkv = torch.randn(T, 3 * D, generator=g)
    qkv[:, :2 * D] *= qk_scale
    qkv = qkv.bfloat16()
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    q, k, v = (qkv[:, i * D:(i + 1) * D].float().reshape(T, H, hd) for i in range(3))
    want = O.varlen_attention(q, k, v, cu, p).reshape(T, D)
    qd = qkv.to(dev)
    qq, kk, vv = (qd[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
    got = ops.attn_varlen(qq, kk, vv, cu.to(dev), max(lens), impl=impl)
    torch.cuda.synchronize()
    return got, want


@case
def attn_generic():
    out = {}
    for hd, H in ((16, 20), (64, 4), (32, 5), (128, 2)):
        got, want = _attn_case([66, 66, 1, 2, 300, 129], H, hd, impl=1, qk_scale=1.5)
        _cmp(f'generic_hd{hd}', got, want, out)
    return out


@case
def attn64_small():
    out = {}
    for name, lens in (('one_tile', [128]), ('short', [64]), ('two_blocks', [256]), ('ragged', [300, 131, 66, 2, 129, 257, 1])):
        got, want = _attn_case(lens, 2, 64, impl=0, qk_scale=1.5)
        s = _cmp(f'attn64_{name}', got, want, out)
        if s['max_rel_to_peak'] > 0.05:
            out[f'attn64_{name}_where'] = _where_bad(got, want, 0.05 * want.abs().max().item())
    return out


@case
def attn64_rescale():
    """Scores whose row maximum keeps growing by > 2^8 block after block: exercises the lazy O rescale."""
    torch, ops, L, O = _imports()
    out = {}
    dev = 'cuda'
    g = torch.Generator().manual_seed(1)
    p = O._Prec('bf16')
    H, hd, Ls = 2, 64, [640, 300]
    T, D = sum(Ls), H * hd
    q = torch.randn(T, D, generator=g)
    k = torch.randn(T, D, generator=g)
    v = torch.randn(T, D, generator=g)
    ramp = torch.cat([torch.arange(l) for l in Ls]).float() / 128.0       # later keys score much higher
    k = k * (1.0 + 3.0 * ramp[:, None])
    qkv = torch.cat([q * 2.0, k, v], 1).bfloat16()
    cu = torch.tensor([0, 640, 940], dtype=torch.int32)
    qq, kk, vv = (qkv[:, i * D:(i + 1) * D].float().reshape(T, H, hd) for i in range(3))
    want = O.varlen_attention(qq, kk, vv, cu, p).reshape(T, D)
    qd = qkv.to(dev)
    a, b, c = (qd[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
    got = ops.attn_varlen(a, b, c, cu.to(dev), 640, impl=0)
    _cmp('attn64_rescale', got, want, out)
    got1 = ops.attn_varlen(a, b, c, cu.to(dev), 640, impl=1)
    _cmp('generic_rescale', got1, want, out)
    return out


@case
def attn_accuracy():
    """v2 / v1 / generic / flash-attn (the reference's kernel) against an exact fp64 attention on the same
    bf16 inputs: the CUDA kernels must not be further from the exact result than the reference's kernel."""
    torch, ops, L, O = _imports()
    dev = 'cuda'
    out = {}
    g = torch.Generator().manual_seed(11)
    H, hd, lens = 4, 64, [700, 1300, 90, 257]
    T, D = sum(lens), H * hd
    qkv = torch.randn(T, 3 * D, generator=g)
    qkv[:, :2 * D] *= 1.4
    qkv = qkv.bfloat16()
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    p64 = O._Prec('fp64')
    q, k, v = (qkv[:, i * D:(i + 1) * D].double().reshape(T, H, hd) for i in range(3))
    exact = O.varlen_attention(q, k, v, cu, p64).reshape(T, D)
    qd = qkv.to(dev)
    a, b, c = (qd[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
    for name, impl in (('tcgen05', 0), ('generic', 1)):
        _cmp(name, ops.attn_varlen(a, b, c, cu.to(dev), max(lens), impl=impl), exact, out)
    _cmp('oracle_bf16', O.varlen_attention(q.float(), k.float(), v.float(), cu, O._Prec('bf16')).reshape(T, D), exact, out)
    _cmp('exact_rounded_to_bf16', exact.bfloat16(), exact, out)
    try:
        from flash_attn import flash_attn_varlen_func
        fa = flash_attn_varlen_func(a.contiguous(), b.contiguous(), c.contiguous(), cu.to(dev), cu.to(dev),
                                    max(lens), max(lens))
        _cmp('flash_attn_library', fa.reshape(T, D), exact, out)
    except Exception as e:
        out['flash_attn_error'] = repr(e)[:200]
    return out


@case
def attn_probe():
    """Where does the tcgen05 kernel's extra noise over flash-attn come from?  (a) q = 0: P = 1 exactly, so any
    error is the P.V accumulation; (b) v = 1: output = sum(bf16 P) / sum(P), isolates the P rounding;
    (c) the plain case, against the exact fp64 result, next to the generic kernel and flash-attn."""
    torch, ops, L, O = _imports()
    dev = 'cuda'
    out = {}
    g = torch.Generator().manual_seed(11)
    H, hd, lens = 4, 64, [700, 1300, 90, 257]
    T, D = sum(lens), H * hd
    base = torch.randn(T, 3 * D, generator=g)
    base[:, :2 * D] *= 1.4
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    p64 = O._Prec('fp64')
    for tag in ('plain', 'intqk', 'onehot_v'):
        qkv = base.clone()
        lens_t, cu_t = lens, cu
        if tag == 'intqk':                      # integer q, k: every score is exact in any accumulator
            qkv[:, :2 * D] = torch.randint(-2, 3, (T, 2 * D), generator=g).float()
        if tag == 'onehot_v':                   # output column d = share of the softmax mass on keys = d mod 64
            oh = torch.zeros(T, hd)
            oh[torch.arange(T), torch.arange(T) % hd] = 1.0
            qkv[:, 2 * D:] = oh.repeat(1, H)
        if tag == 'v1':
            qkv[:, 2 * D:] = 1.0
        if tag == 'short':
            lens_t = [100, 128, 60, 30] * 7
            qkv = qkv[:sum(lens_t)]
            cu_t = torch.zeros(len(lens_t) + 1, dtype=torch.int32)
            cu_t[1:] = torch.cumsum(torch.tensor(lens_t), 0)
        qkv = qkv.bfloat16()
        Tt = qkv.shape[0]
        q, k, v = (qkv[:, i * D:(i + 1) * D].double().reshape(Tt, H, hd) for i in range(3))
        exact = O.varlen_attention(q, k, v, cu_t, p64).reshape(Tt, D)
        qd = qkv.to(dev)
        a, b, c = (qd[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
        for name, impl in (('tcgen05', 0), ('generic', 1)):
            r = _cmp(f'{tag}_{name}', ops.attn_varlen(a, b, c, cu_t.to(dev), max(lens_t), impl=impl), exact, out)
            out[f'{tag}_{name}'] = {k_: r[k_] for k_ in ('max_abs', 'rms_rel')}
        r = _cmp(f'{tag}_oracle_bf16', O.varlen_attention(q.float(), k.float(), v.float(), cu_t, O._Prec('bf16')).reshape(Tt, D), exact, out)
        out[f'{tag}_oracle_bf16'] = {k_: r[k_] for k_ in ('max_abs', 'rms_rel')}
        r = _cmp(f'{tag}_exact_rounded', exact.bfloat16(), exact, out)
        out[f'{tag}_exact_rounded'] = {k_: r[k_] for k_ in ('max_abs', 'rms_rel')}
        try:
            from flash_attn import flash_attn_varlen_func
            fa = flash_attn_varlen_func(a.contiguous(), b.contiguous(), c.contiguous(), cu_t.to(dev), cu_t.to(dev),
                                        max(lens_t), max(lens_t))
            r = _cmp(f'{tag}_flash', fa.reshape(Tt, D), exact, out)
            out[f'{tag}_flash'] = {k_: r[k_] for k_ in ('max_abs', 'rms_rel')}
        except Exception as e:
            out['flash_attn_error'] = repr(e)[:200]
    return out


@case
def perf_attn_cold():
    """Attention kernel timed per launch with CUDA events, (a) back to back, (b) after streaming 1 GB through L2
    (as inside a forward, where the QKV GEMM has just pushed ~0.5 GB through the 126 MB L2)."""
    import subprocess
    torch, ops, L, O = _imports()
    dev = 'cuda'
    out = {}
    lens = O.synthetic_lengths(50000, seed=2)
    T, H, hd = sum(lens), 20, 64
    D = H * hd
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    cu = cu.to(dev)
    qkv = torch.randn(T, 3 * D, device=dev).bfloat16()
    q, k, v = (qkv[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
    flops = 4.0 * D * sum(l * l for l in lens)
    _, tile_info = ops.batch_meta(cu, T)
    junk = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    for cold in (False, True):
        ts = []
        for i in range(13):
            if cold:
                junk.add_(1)
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            ops.attn_varlen(q, k, v, cu, max(lens), tile_info, impl=0)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[len(ts) // 2]
        out['cold_l2' if cold else 'back_to_back'] = dict(ms=ms, tflops=flops / ms / 1e9)
    try:
        out['sm_clock_mhz_after'] = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm', '--format=csv,noheader,nounits'],
                                                   capture_output=True, text=True).stdout.strip()
    except Exception:
        pass
    return out


@case
def perf_attn():
    torch, ops, L, O = _imports()
    dev = 'cuda'
    out = {}
    lens = O.synthetic_lengths(50000, seed=2)
    T, H, hd = sum(lens), 20, 64
    D = H * hd
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(torch.tensor(lens), 0)
    cu = cu.to(dev)
    qkv = torch.randn(T, 3 * D, device=dev).bfloat16()
    q, k, v = (qkv[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
    flops = 4.0 * D * sum(l * l for l in lens)
    _, tile_info = ops.batch_meta(cu, T)

    def timeit(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for name, impl in (('tcgen05', 0),):
        ms = timeit(lambda: ops.attn_varlen(q, k, v, cu, max(lens), tile_info, impl=impl))
        out[name] = dict(ms=ms, tflops=flops / ms / 1e9)
    try:
        from flash_attn import flash_attn_varlen_func
        qc, kc, vc = q.contiguous(), k.contiguous(), v.contiguous()
        ms = timeit(lambda: flash_attn_varlen_func(qc, kc, vc, cu, cu, max(lens), max(lens)))
        out['flash_attn_2.8.3_library'] = dict(ms=ms, tflops=flops / ms / 1e9)
        a = ops.attn_varlen(q, k, v, cu, max(lens), tile_info, impl=0).float()
        b = flash_attn_varlen_func(qc, kc, vc, cu, cu, max(lens), max(lens)).reshape(T, D).float()
        out['v3_vs_flash_attn_max_abs'] = (a - b).abs().max().item()
    except Exception as e:                                  # library baseline is optional
        out['flash_attn_error'] = repr(e)[:200]
    return out


@case
def attn64_big():
    out = {}
    got, want = _attn_case([1026, 700, 3, 2050, 513], 20, 64, impl=0, qk_scale=2.0, seed=3)
    _cmp('attn64_long', got, want, out)
    return out


def _model_case(ckpt, fixture, out, padded=None):
    torch, ops, L, O = _imports()
    import esme
    from conftest import load_golden, err_stats, GOLDEN
    dev = 'cuda'
    g = load_golden(fixture)
    model = esme.ESM.from_pretrained(f'{GOLDEN}/{ckpt}', device=dev)
    cfg, W = O.load_checkpoint(f'{GOLDEN}/{ckpt}')
    tokens, cu, max_len = g['tokens'], g['cu_lens'], g['max_len']
    got = model(tokens.to(dev), (cu.to(dev), max_len)).float().cpu()
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    tag = fixture.split('.')[0]
    out[f'{tag}_new_vs_ref'] = err_stats(got, g['logits'])
    out[f'{tag}_new_vs_fp64'] = err_stats(got, exact)
    out[f'{tag}_ref_vs_fp64'] = err_stats(g['logits'], exact)
    lp = model.predict_log_prob(tokens.to(dev), (cu.to(dev), max_len)).float().cpu()
    out[f'{tag}_logp_new_vs_ref'] = err_stats(lp, g['log_prob'])
    rep = model.forward_representation(tokens.to(dev), (cu.to(dev), max_len)).float().cpu()
    out[f'{tag}_repr_new_vs_ref'] = err_stats(rep, g['representation'])
    # module-level layer 0 vs reference taps
    if 'layer0.x_out' in g:
        x0 = model.embedding(tokens.to(dev))
        y0 = model.layers[0](x0, cu.to(dev), max_len).float().cpu()
        out[f'{tag}_layer0_module_vs_ref'] = err_stats(y0, g['layer0.x_out'])
    return model, g


@case
def model_8m():
    out = {}
    torch, ops, L, O = _imports()
    from conftest import load_golden, err_stats
    model, g = _model_case('esm2_8m.safetensors', 'esm2_8m_cfg1.npz', out)
    toks2d = __import__('esme').tokenize(g['seqs'], __import__('esme').alphabet.Alphabet)
    pl = model(toks2d.cuda()).float().cpu()
    out['cfg1_padded_vs_ref'] = err_stats(pl, g['padded_logits'])
    _model_case('esm2_8m.safetensors', 'esm2_8m_testfa.npz', out)
    gp = load_golden('esm2_8m_padded.npz')
    pl = model(gp['tokens'].cuda()).float().cpu()
    out['padded_mixed_vs_ref'] = err_stats(pl, gp['logits'])
    plp = model.predict_log_prob(gp['tokens'].cuda()).float().cpu()
    out['padded_mixed_logp_vs_ref'] = err_stats(plp, gp['log_prob'])
    return out


@case
def model_tiny_esm2():
    out = {}
    _model_case('esm2_tiny.safetensors', 'esm2_tiny.npz', out)
    return out


@case
def model_tiny_esmc():
    out = {}
    _model_case('esmc_tiny.safetensors', 'esmc_tiny.npz', out)
    return out


@case
def perf_gemm():
    torch, ops, L, O = _imports()
    dev = 'cuda'
    out = {}
    for (M, N, K, epi) in [(50000, 3840, 1280, L.EPI_BIAS), (50000, 5120, 1280, L.EPI_BIAS_GELU),
                           (50000, 1280, 5120, L.EPI_BIAS), (50000, 1280, 1280, L.EPI_BIAS),
                           (8192, 8192, 8192, L.EPI_BIAS)]:
        x = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=dev).bfloat16()
        y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.linear(x, w, b, epilogue=epi, out=y)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            ops.linear(x, w, b, epilogue=epi, out=y)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        e0.record()
        for _ in range(10):
            torch.nn.functional.linear(x, w, b)
        e1.record()
        torch.cuda.synchronize()
        ms_t = e0.elapsed_time(e1) / 10
        out[f'gemm_{M}x{N}x{K}_epi{epi}'] = dict(ms=ms, tflops=2 * M * N * K / ms / 1e9, cublas_ms=ms_t,
                                                 cublas_tflops=2 * M * N * K / ms_t / 1e9)
    return out


@case
def perf_model():
    torch, ops, L, O = _imports()
    import esme
    dev = 'cuda'
    out = {}
    torch.manual_seed(0)
    model = esme.ESM2(33, 1280, 20).to(dev)
    for p_ in model.parameters():
        torch.nn.init.normal_(p_, std=0.02) if p_.ndim > 1 else None
    for m in model.modules():
        if isinstance(m, torch.nn.LayerNorm):
            torch.nn.init.ones_(m.weight)
            if m.bias is not None:
                torch.nn.init.zeros_(m.bias)
        elif isinstance(m, torch.nn.Linear) and m.bias is not None:
            torch.nn.init.zeros_(m.bias)
    lens = O.synthetic_lengths(50000, seed=2)
    tokens, cu, max_len = O.synthetic_batch(lens, seed=3)
    tokens, cu = tokens.to(dev), cu.to(dev)
    T = tokens.numel()
    for _ in range(2):
        y = model(tokens, (cu, max_len))
    torch.cuda.synchronize()
    out['finite'] = bool(torch.isfinite(y.float()).all())
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(5):
        y = model(tokens, (cu, max_len))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    out['esm2_650m'] = dict(T=T, B=len(lens), max_len=max_len, ms=ms, residues_per_s=T / ms * 1e3)
    return out


from oracle.restated_gpu import reference_gpu_forward as _reference_gpu_forward  # noqa: E402


@case
def perf_reference_gpu():
    """Same batch, same weights: esmk engine vs the reference's op sequence on torch/cuBLAS/flash-attn."""
    torch, ops, L, O = _imports()
    import esme
    from conftest import err_stats
    dev = 'cuda'
    out = {}
    cfg = O.OracleConfig('esm2', 33, 1280, 20)
    W = O.synthetic_weights(cfg, seed=1)
    model = esme.ESM2(33, 1280, 20)
    model.load_state_dict(W, strict=True)
    model = model.to(dev).eval()
    Wd = {k: v.to(dev) for k, v in W.items()}
    lens = O.synthetic_lengths(50000, seed=2)
    tokens, cu, max_len = O.synthetic_batch(lens, seed=3)
    tokens, cu = tokens.to(dev), cu.to(dev)
    T = tokens.numel()

    def timeit(fn, n=5):
        for _ in range(2):
            y = fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            y = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, y
    with torch.no_grad():
        ms_ref, y_ref = timeit(lambda: _reference_gpu_forward(Wd, 33, 1280, 20, tokens, cu, max_len))
        ms_new, y_new = timeit(lambda: model(tokens, (cu, max_len)))
    out['reference_restated_gpu'] = dict(ms=ms_ref, residues_per_s=T / ms_ref * 1e3)
    out['esmk'] = dict(ms=ms_new, residues_per_s=T / ms_new * 1e3)
    out['speedup'] = ms_ref / ms_new
    out['new_vs_reference_gpu'] = err_stats(y_new.float().cpu(), y_ref.float().cpu())
    n = int(cu[2])
    exact = O.forward_packed(cfg, W, tokens[:n].cpu(), cu[:3].cpu(), int((cu[1:3] - cu[:2]).max()), 'fp32')
    out['prefix_new_vs_exact'] = err_stats(y_new[:n].float().cpu(), exact)
    out['prefix_reference_gpu_vs_exact'] = err_stats(y_ref[:n].float().cpu(), exact)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--case')
    ap.add_argument('--only', nargs='*')
    ap.add_argument('--timeout', type=int, default=300)
    a = ap.parse_args()
    if a.case:
        res = CASES[a.case]()
        print('@@RESULT@@' + json.dumps(res))
        return
    results = {}
    names = a.only or list(CASES)
    for name in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, '--case', name], capture_output=True, text=True,
                               timeout=a.timeout)
            payload = [l for l in r.stdout.splitlines() if l.startswith('@@RESULT@@')]
            if r.returncode == 0 and payload:
                results[name] = json.loads(payload[-1][len('@@RESULT@@'):])
            else:
                results[name] = dict(error=f'rc={r.returncode}', stdout=r.stdout[-1500:], stderr=r.stderr[-2500:])
        except subprocess.TimeoutExpired as e:
            results[name] = dict(error='timeout', stdout=(e.stdout or b'')[-1500:].decode('utf8', 'replace')
                                 if isinstance(e.stdout, bytes) else str(e.stdout)[-1500:])
        results[name + '__seconds'] = round(time.time() - t0, 1)
        print(f'== {name} ({results[name + "__seconds"]}s)')
        print(json.dumps(results[name], indent=1)[:6000])
        sys.stdout.flush()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'diag.json'), 'w') as f:
        json.dump(results, f, indent=1)


if __name__ == '__main__':
    main()
