"""A/B timing of the residual-epilogue GEMMs (measurement tool, GPU box only): out-projection and FFN-down shapes of
ESM2-650M / 3B / ESMC-300M, in place on x as esmk_forward runs them, with ESMK_GEMM_TMA_RESID on / off, plus a
bit-exactness check of the two paths against each other.  One child process per setting (read once by libesmk)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))
SHAPES = [(49677, 1280, 1280, 1.0), (49677, 1280, 5120, 1.0), (49677, 960, 960, 0.9129), (49677, 960, 2560, 0.9129),
          (24400, 2560, 2560, 1.0), (24400, 2560, 10240, 1.0), (1000, 1280, 1280, 1.0)]


def one():
    import torch
    from esme import ops, _lib as L
    dev = 'cuda'
    out = {}
    for (M, N, K, sc) in SHAPES:
        g = torch.Generator(device=dev).manual_seed(1)
        a = torch.randn(M, K, device=dev, generator=g).bfloat16()
        w = (torch.randn(N, K, device=dev, generator=g) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=dev, generator=g).bfloat16() if sc == 1.0 else None
        x0 = torch.randn(M, N, device=dev, generator=g).bfloat16()
        x = x0.clone()
        ops.linear(a, w, b, epilogue=L.EPI_RESIDUAL, residual=x, residue_scaling=sc, out=x)     # in place, as the engine does
        ref = x0.float() + ((a.float() @ w.float().T + (b.float() if b is not None else 0)).bfloat16().float() / sc).bfloat16().float()
        err = (x.float() - ref).abs().max().item()
        chk = float(x.float().sum().item())
        for _ in range(3):
            ops.linear(a, w, b, epilogue=L.EPI_RESIDUAL, residual=x, residue_scaling=sc, out=x)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.linear(a, w, b, epilogue=L.EPI_RESIDUAL, residual=x, residue_scaling=sc, out=x)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        out[f'{M}x{N}x{K}'] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9, max_abs_err_vs_torch=err, checksum=chk)
    print('@@AB@@' + json.dumps(out))


if __name__ == '__main__':
    if len(sys.argv) > 1:
        one()
    else:
        res = {}
        for flag in ('1', '0'):
            env = dict(os.environ, ESMK_GEMM_TMA_RESID=flag)
            r = subprocess.run([sys.executable, __file__, 'one'], capture_output=True, text=True, env=env, timeout=600)
            lines = [l for l in r.stdout.splitlines() if l.startswith('@@AB@@')]
            res[flag] = json.loads(lines[-1][6:]) if lines else dict(error=r.stderr[-1500:] + r.stdout[-500:])
        for k in res['1']:
            if k == 'error' or 'error' in res['0']:
                print(res)
                break
            a, b = res['1'][k], res['0'][k]
            print(f'{k:22s} tma_resid {a["ms"]:.4f} ms {a["tflops"]:7.1f} TF | per-thread loads {b["ms"]:.4f} ms {b["tflops"]:7.1f} TF | '
                  f'speedup {b["ms"] / a["ms"]:.3f} | err {a["max_abs_err_vs_torch"]:.4f}/{b["max_abs_err_vs_torch"]:.4f} '
                  f'| identical {a["checksum"] == b["checksum"]}')
        json.dump(res, open(os.path.join(ROOT, 'gpurun_out', 'gemm_ab.json'), 'w'), indent=1)
