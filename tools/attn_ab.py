"""A/B harness for the attention kernel variants (measurement tool, GPU box only).

    python tools/attn_ab.py                      # every configuration below, one child process each
    python tools/attn_ab.py --one '{"ESMK_ATTN_IMPL": "v1"}'

Per configuration (environment variables read once by libesmk): accuracy against an exact fp64 attention on the
same bf16 inputs (rms-relative, next to flash-attn 2.8.3 and the oracle's bf16 restatement), the parity tests'
ragged cases against the CUDA-core kernel, and the time on the BASELINE config-2 batch (ESM2-650M geometry, 49,677
tokens, 20 heads x 64): 20 back-to-back launches (L2-warm) and 20 launches with a 256 MB L2 flush in between.
Writes gpurun_out/attn_ab.json.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))

CONFIGS = [{'ESMK_ATTN_RESCALE_THRESHOLD': t} for t in ('0', '1', '2', '3', '4', '8')] + \
          [{'ESMK_ATTN_RESCALE_THRESHOLD': '0', 'ESMK_ATTN_POLY': '1'}]


def rel(a, b):
    a, b = a.double(), b.double()
    return ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()


def one():
    import torch
    from esme import ops, synthetic
    from oracle import esm_oracle as O
    dev = 'cuda'
    out = {}
    # ---- accuracy: exact fp64 attention on the same bf16 inputs
    for tag, gain, lens in (('gain1.4', 1.4, [700, 1300, 90, 257]), ('gain3', 3.0, [130, 64, 1, 513, 2049])):
        g = torch.Generator().manual_seed(11)
        H, hd = 4, 64
        T, D = sum(lens), H * hd
        qkv = torch.randn(T, 3 * D, generator=g)
        qkv[:, :2 * D] *= gain
        qkv = qkv.bfloat16()
        cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
        cu[1:] = torch.cumsum(torch.tensor(lens), 0)
        q, k, v = (qkv[:, i * D:(i + 1) * D].double().reshape(T, H, hd) for i in range(3))
        exact = O.varlen_attention(q, k, v, cu, O._Prec('fp64')).reshape(T, D)
        qd = qkv.to(dev)
        a, b, c = (qd[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
        got = ops.attn_varlen(a, b, c, cu.to(dev), max(lens)).float().cpu()
        gen = ops.attn_varlen(a, b, c, cu.to(dev), max(lens), impl=1).float().cpu()
        orc = O.varlen_attention(q.float(), k.float(), v.float(), cu, O._Prec('bf16')).reshape(T, D)
        from flash_attn import flash_attn_varlen_func
        fa = flash_attn_varlen_func(a.contiguous(), b.contiguous(), c.contiguous(), cu.to(dev), cu.to(dev),
                                    max(lens), max(lens)).reshape(T, D).float().cpu()
        out[tag] = dict(tcgen05=rel(got, exact), cuda_core=rel(gen, exact), oracle_bf16=rel(orc, exact),
                        flash_attn=rel(fa, exact), finite=bool(torch.isfinite(got).all()),
                        max_abs_vs_cuda_core=(got - gen).abs().max().item())
    # ---- time on the config-2 batch
    lens = synthetic.synthetic_lengths(50000, seed=2)
    H, hd, D = 20, 64, 1280
    T = sum(lens)
    g = torch.Generator(device=dev).manual_seed(5)
    qkv = torch.randn(T, 3 * D, generator=g, device=dev).to(torch.bfloat16)
    q, k, v = (qkv[:, i * D:(i + 1) * D].view(T, H, hd) for i in range(3))
    cu = torch.zeros(len(lens) + 1, dtype=torch.int32, device=dev)
    cu[1:] = torch.cumsum(torch.tensor(lens, dtype=torch.int32, device=dev), 0)
    _, tile_info = ops.batch_meta(cu, T)
    flops = 4.0 * D * sum(l * l for l in lens)
    for _ in range(3):
        y = ops.attn_varlen(q, k, v, cu, max(lens), tile_info)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        y = ops.attn_varlen(q, k, v, cu, max(lens), tile_info)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cold = []
    for _ in range(20):
        flush.zero_()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        ops.attn_varlen(q, k, v, cu, max(lens), tile_info)
        a1.record()
        torch.cuda.synchronize()
        cold.append(a0.elapsed_time(a1))
    cold.sort()
    gen = ops.attn_varlen(q, k, v, cu, max(lens), impl=1)
    out['config2'] = dict(ms_warm=ms, tflops_warm=flops / ms / 1e9, ms_cold_median=cold[len(cold) // 2],
                          tflops_cold=flops / cold[len(cold) // 2] / 1e9,
                          max_abs_vs_cuda_core=(y.float() - gen.float()).abs().max().item(),
                          finite=bool(torch.isfinite(y.float()).all()))
    print('@@AB@@' + json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--one')
    ap.add_argument('--only', type=int, nargs='*')
    a = ap.parse_args()
    if a.one is not None:
        return one()
    results = []
    for i, cfg in enumerate(CONFIGS):
        if a.only and i not in a.only:
            continue
        env = dict(os.environ)
        env.update(cfg)
        try:
            r = subprocess.run([sys.executable, __file__, '--one', json.dumps(cfg)], capture_output=True, text=True,
                               timeout=600, env=env)
            lines = [l for l in r.stdout.splitlines() if l.startswith('@@AB@@')]
            res = json.loads(lines[-1][6:]) if lines else dict(error=f'rc={r.returncode}', stderr=r.stderr[-2000:],
                                                                 stdout=r.stdout[-1000:])
        except subprocess.TimeoutExpired:
            res = dict(error='timeout')
        results.append(dict(config=cfg, result=res))
        print(json.dumps(results[-1]))
        sys.stdout.flush()
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    json.dump(results, open(os.path.join(ROOT, 'gpurun_out', 'attn_ab.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
