"""Host-side model logic on CPU: loader, parameter naming, packing, error behaviour."""
import pytest
import torch
from safetensors import safe_open

import esme
from conftest import GOLDEN
from esme.attention import SwiGLU


def test_from_pretrained_round_trip_and_tied_head():
    """Reference tests/test_esm.py:108-120: every parameter equals the file tensor; the
    LM-head projection equals the embedding table for ESM2."""
    path = f'{GOLDEN}/esm2_8m.safetensors'
    model = esme.ESM.from_pretrained(path)
    assert isinstance(model, esme.ESM2) and (model.num_layers, model.embed_dim, model.attention_heads) == (6, 320, 20)
    params = dict(model.named_parameters())
    with safe_open(path, framework='pt') as f:
        assert set(f.keys()) == set(params)
        for k in f.keys():
            assert torch.equal(f.get_tensor(k), params[k]), k
    assert torch.equal(model.lm_head.final.weight, model.embed_tokens.weight)
    assert all(p.dtype == torch.bfloat16 and not p.requires_grad for p in model.parameters())


def test_esmc_architecture():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esmc_tiny.safetensors')
    assert isinstance(model, esme.ESMC)
    layer = model.layers[0]
    assert layer.ffn_dim == 512 and abs(layer.residue_scaling - (2 / 36) ** 0.5) < 1e-12
    assert model.embed_tokens.num_embeddings == 64 and model.lm_head.final.out_features == 64
    assert model.emb_layer_norm_after.bias is None
    assert layer.self_attn.layernorm_q.bias is None and layer.self_attn.q.bias is None
    m300 = esme.ESMC(num_layers=1)
    assert m300.layers[0].ffn_dim == 2560


def test_wrong_family_and_bad_arguments():
    with pytest.raises(AssertionError):
        esme.ESMC.create_model(f'{GOLDEN}/esm2_8m.safetensors')
    with pytest.raises(ValueError):
        esme.ESM.from_pretrained('no_such_model')
    with pytest.raises(AssertionError):
        esme.ESM2.from_pretrained(f'{GOLDEN}/esm2_8m.safetensors', quantization='2bit')
    with pytest.raises(AssertionError):
        esme.ESM2.from_pretrained(f'{GOLDEN}/esm2_8m.safetensors', quantization='4bit', device='cpu')


def test_forward_requires_cuda():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_8m.safetensors')
    tok = torch.tensor([0, 5, 6, 2])
    with pytest.raises(RuntimeError, match='no CPU path'):
        model(tok, (torch.tensor([0, 4], dtype=torch.int32), 4))
    with pytest.raises(AssertionError):
        model.forward_representation(tok, (torch.tensor([0, 4], dtype=torch.int32), 4), layers=[6])


def test_weight_packing_layouts():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esmc_tiny.safetensors')
    sa = model.layers[0].self_attn
    w, b = sa.packed_qkv()
    D = model.embed_dim
    assert b is None and w.shape == (3 * D, D)
    assert torch.equal(w[:D], sa.q.weight) and torch.equal(w[D:2 * D], sa.k.weight) and torch.equal(w[2 * D:], sa.v.weight)
    glu = model.layers[0].final[1]
    inter = SwiGLU.interleave(glu.activation.weight, glu.fc.weight)
    assert inter.shape == (2 * 512, D)
    assert torch.equal(inter[0:32], glu.activation.weight[0:32]) and torch.equal(inter[32:64], glu.fc.weight[0:32])
    assert torch.equal(inter[64:96], glu.activation.weight[32:64])
