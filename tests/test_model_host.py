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


def test_lora_host_plumbing(tmp_path):
    """Module layout / state-dict keys / save-load of the LoRA drop-in, no kernels involved (reference
    tests/test_lora.py:39-185)."""
    from safetensors import safe_open
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors')
    model.add_lora(16, 0.5, adapter_names=['x', 'y'], layers=['query', 'output'])
    keys = set(model.lora_state_dict())
    assert keys == {f'layers.{i}.self_attn.{j}.lora_{a}.{n}' for i in range(model.num_layers) for j in ('q', 'out')
                    for a in 'AB' for n in 'xy'}
    assert set(model.lora_state_dict(['x'])) == {k for k in keys if k.endswith('.x')}
    assert 'layers.0.self_attn.q.layer.weight' in model.state_dict()          # the wrapped linear keeps its weights
    model.mark_only_lora_as_trainable(['x'])
    for n, p in model.named_parameters():
        assert p.requires_grad == (('.lora_A.' in n or '.lora_B.' in n) and n.endswith('.x')), n
    assert len(model.trainable_parameters()) == 2 * 2 * model.num_layers
    path = str(tmp_path / 'lora.safetensors')
    with torch.no_grad():
        model.layers[0].self_attn.q.lora_B['x'].fill_(0.25)
    model.save_lora(path)
    with safe_open(path, 'pt') as f:
        assert f.metadata()['rank'] == '16' and set(f.metadata()['names'].split(',')) == {'x', 'y'}
    other = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors').load_lora(path)
    assert torch.equal(other.layers[0].self_attn.q.lora_B['x'], model.layers[0].self_attn.q.lora_B['x'])
    with pytest.raises(AssertionError):
        model.add_lora(layers=['ffn'])


def test_learned_positions_known_answers():
    """esme/embedding.py docstring example and the packed variant (positions restart at padding_idx + 1 = 2)."""
    from esme.embedding import LearnedPositionalEmbedding
    emb = LearnedPositionalEmbedding(33, 8)
    assert emb.weight.shape == (35, 8)
    x = torch.tensor([[20, 29, 28], [8, 13, 9]])
    assert emb.positions(x).tolist() == [[2, 3, 4], [2, 3, 4]]
    assert emb.positions(torch.tensor([[0, 5, 2, 1, 1]])).tolist() == [[2, 3, 4, 1, 1]]    # padding stays at padding_idx
    with pytest.raises(ValueError):
        emb.positions(torch.zeros(1, 40, dtype=torch.long))
    for cls in (esme.ESM1b, esme.ESM1v):
        m = cls(num_layers=1, embed_dim=64, attention_heads=1, max_seq_len=32)
        assert hasattr(m, 'emb_layer_norm_before') == (cls is esme.ESM1b)
        assert m.layers[0].self_attn.rot_emb is None and m.embed_positions.weight.shape == (34, 64)
