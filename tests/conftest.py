import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
PKG = os.path.join(ROOT, 'esm-efficient_b200')
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    """Load a fixture written by tests/golden/make_golden.py.  uint16 arrays are
    raw bf16 bit patterns and come back as float32 tensors holding exact bf16 values."""
    out = {}
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        for k in z.files:
            a = z[k]
            if a.dtype == np.uint16:
                out[k] = torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).float()
            elif a.dtype.kind in 'US':
                out[k] = [str(s) for s in a.tolist()] if a.ndim else str(a)
            elif a.ndim == 0:
                out[k] = a.item()
            else:
                out[k] = torch.from_numpy(a.copy())
    return out


def err_stats(a: torch.Tensor, b: torch.Tensor):
    """(max-abs, rms-relative, min row cosine, argmax agreement) of a vs b (last dim = vocab/features)."""
    a = a.double().reshape(-1, a.shape[-1])
    b = b.double().reshape(-1, b.shape[-1])
    max_abs = (a - b).abs().max().item()
    rms_rel = ((a - b).pow(2).mean().sqrt() / b.pow(2).mean().sqrt()).item()
    cos = torch.nn.functional.cosine_similarity(a, b, dim=-1).min().item()
    agree = (a.argmax(-1) == b.argmax(-1)).double().mean().item()
    return max_abs, rms_rel, cos, agree


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN
