"""Profiling driver (run under ncu): ESM2-650M-shaped synthetic model, config-2 batch."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))
import torch
import esme
from esme import ops, _lib as L
from oracle import esm_oracle as O

mode = sys.argv[1] if len(sys.argv) > 1 else 'model'
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = 'cuda'
torch.manual_seed(0)
if mode == 'model':
    layers = int(os.environ.get('LAYERS', '33'))
    model = esme.ESM2(layers, 1280, 20).to(dev)
    for p in model.parameters():
        if p.ndim > 1:
            torch.nn.init.normal_(p, std=0.02)
    for m in model.modules():
        if isinstance(m, torch.nn.LayerNorm):
            torch.nn.init.ones_(m.weight)
            torch.nn.init.zeros_(m.bias)
        elif isinstance(m, torch.nn.Linear) and m.bias is not None:
            torch.nn.init.zeros_(m.bias)
    lens = O.synthetic_lengths(50000, seed=2)
    tokens, cu, max_len = O.synthetic_batch(lens, seed=3)
    tokens, cu = tokens.to(dev), cu.to(dev)
    for _ in range(iters):
        y = model(tokens, (cu, max_len))
    torch.cuda.synchronize()
    print('ok', y.shape)
else:
    M, N, K, epi = {'qkv': (50000, 3840, 1280, L.EPI_BIAS), 'gelu': (50000, 5120, 1280, L.EPI_BIAS_GELU),
                    'down': (50000, 1280, 5120, L.EPI_BIAS), 'big': (8192, 8192, 8192, L.EPI_BIAS)}[mode]
    x = torch.randn(M, K, device=dev).bfloat16()
    w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
    b = torch.randn(N, device=dev).bfloat16()
    y = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    for _ in range(iters):
        ops.linear(x, w, b, epilogue=epi, out=y)
    torch.cuda.synchronize()
    print('ok')
