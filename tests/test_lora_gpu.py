"""LoRA adapters at inference (reference: esme/lora.py, esme/esm.py:495-616, esme/attention.py:91-139): the adapter
checkpoint written by the REAL reference's `save_lora` loads through `load_lora`, and the logits with all adapters
(lora_names=None), with one adapter, and of the base model match the reference's (tests/golden/make_golden_lora.py)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'esm-efficient_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

from conftest import GOLDEN, err_stats, load_golden   # noqa: E402

pytestmark = pytest.mark.gpu


def test_lora_checkpoint_of_the_reference_loads_and_matches():
    import esme
    g = load_golden('lora_tiny.npz')
    tokens, cu, max_len = g['tokens'].cuda(), g['cu_lens'].cuda(), g['max_len']
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device='cuda')
    base = model(tokens, (cu, max_len))
    _, rms, cos, _ = err_stats(base.float().cpu(), g['logits_base'])
    assert rms < 6e-3 and cos > 0.9999
    with pytest.raises(ValueError):
        model(tokens, (cu, max_len), lora_names=['a'])                      # no adapters attached yet
    model.load_lora(f'{GOLDEN}/lora_tiny.safetensors')
    keys = set(model.lora_state_dict())
    assert 'layers.0.self_attn.q.lora_A.a' in keys and 'layers.1.self_attn.out.lora_B.b' in keys
    assert not any('.k.lora' in k for k in keys)                            # layers = query, value, output
    assert all(p.requires_grad == ('.lora_' in n) for n, p in model.named_parameters())
    both = model(tokens, (cu, max_len))                                     # None = every adapter, as LoRA.forward
    only_a = model(tokens, (cu, max_len), lora_names=['a'])
    for got, want in ((both, g['logits_all']), (only_a, g['logits_a'])):
        _, rms, cos, agree = err_stats(got.float().cpu(), want)
        assert rms < 8e-3 and cos > 0.9999 and agree > 0.97, (rms, cos, agree)
    # the adapters matter (the fixture moves the logits by ~3) and are distinguishable
    assert (both.float() - base.float()).abs().max() > 1.0
    assert (only_a.float() - both.float()).abs().max() > 1.0
    # padded entry and log-probs go through the same path
    lp = model.predict_log_prob(tokens, (cu, max_len), lora_names=['a'])
    assert torch.allclose(lp.float().exp().sum(-1), torch.ones(tokens.numel(), device='cuda'), atol=3e-2)
    rep = model.forward_representation(tokens, (cu, max_len), lora_names=['a'], layers=[0])
    assert rep.shape == (tokens.numel(), 2 * model.embed_dim)


def test_lora_module_and_save_load_roundtrip(tmp_path):
    """esme/lora.py unit behaviour (reference tests/test_lora.py:8-36, 118-185): fresh adapters are a no-op
    (B = 0), a rank that is not a multiple of 8 works, save_lora / load_lora round-trips."""
    import esme
    from esme.lora import LoRA
    from safetensors import safe_open
    lin = torch.nn.Linear(64, 128, dtype=torch.bfloat16).cuda()
    lora = LoRA(lin, rank=4, alpha=1, names=['x'])
    assert 'x' in lora.lora_A and lora.lora_A['x'].shape == (4, 64) and lora.lora_B['x'].shape == (128, 4)
    x = torch.randn(40, 64, device='cuda').bfloat16()
    y0 = lora(x)
    assert torch.equal(y0, esme.ops.linear(x, lin.weight, lin.bias))        # B initialised to zero
    with torch.no_grad():
        lora.lora_B['x'].copy_((torch.randn(128, 4) * 0.1).bfloat16())
    want = torch.nn.functional.linear(x.float(), lin.weight.float(), lin.bias.float()) + \
        (x.float() @ lora.lora_A['x'].float().t()) @ lora.lora_B['x'].float().t() * lora.scaling
    assert (lora(x).float() - want).abs().max() < 0.05 * want.abs().max()
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device='cuda')
    model.add_lora(16, 0.5, adapter_names=['test_a', 'test_b'], layers=['query', 'key'])
    path = str(tmp_path / 'lora.safetensors')
    model.save_lora(path)
    with safe_open(path, 'pt') as f:
        assert set(f.keys()) == {f'layers.{i}.self_attn.{j}.lora_{a}.test_{n}'
                                 for i in range(2) for j in 'qk' for a in 'AB' for n in 'ab'}
    other = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device='cuda').load_lora(path)
    for k, v in model.lora_state_dict().items():
        assert torch.equal(other.state_dict()[k], v)
