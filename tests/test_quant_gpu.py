"""Weight-only quantised storage (`quantization='4bit' | '8bit'`, reference: esme/esm.py:414-472, 916-946) on the
GPU: the device kernels against the CPU restatement of the formats (bit-exact: codes are integer work), the
loader's contract as the reference's own tests state it (tests/test_esm.py:123-154: integer weight dtypes, exact
biases, tied LM-head weight), and the engine path against a bf16 model carrying the dequantised weights (bit-exact:
the GEMMs are the same, only the weight storage differs)."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'esm-efficient_b200')):
    if p not in sys.path:
        sys.path.insert(0, p)

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def _weights(seed, N, K):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(N, K, generator=g) * 0.05
    w[0, :64] = 0                       # an all-zero block (absmax 0)
    w[1, 5] = 3.0                       # an outlier that sets its block's / row's scale
    return w.bfloat16()


@pytest.mark.parametrize('N,K', [(320, 320), (1280, 320), (96, 2560)])
def test_q4_kernels_match_the_format_oracle(N, K):
    from esme import ops
    from oracle import quant_oracle as Q
    w = _weights(N + K, N, K)
    data, scale = ops.quantize(w.cuda(), 4)
    want_data, want_scale = Q.q4_quantize(w)
    assert data.dtype == torch.uint8 and data.shape == (N * K // 2, 1) and scale.shape == (N * K // 64,)
    assert torch.equal(scale.cpu(), want_scale)
    assert torch.equal(data.cpu(), want_data)
    back = ops.dequantize(data, scale, N, K, 4)
    assert torch.equal(back.cpu(), Q.q4_dequantize(want_data, want_scale, N, K))
    # quantisation error of the fp4 codebook: the largest gap between neighbouring codes is 1/3 of the block
    # absmax (2/3 -> 1), plus the bf16 rounding of the dequantised value
    err = (back.float().cpu() - w.float()).reshape(-1, 64).abs().max(1).values
    assert (err <= want_scale * (1 / 6 + 1 / 128) + 1e-6).all()


@pytest.mark.parametrize('N,K', [(320, 320), (33, 1280)])
def test_q8_kernels_match_the_format_oracle(N, K):
    from esme import ops
    from oracle import quant_oracle as Q
    w = _weights(N * 3 + K, N, K)
    data, scale = ops.quantize(w.cuda(), 8)
    want_data, want_scale = Q.q8_quantize(w)
    assert data.dtype == torch.int8 and torch.equal(scale.cpu(), want_scale)
    assert torch.equal(data.cpu(), want_data)
    back = ops.dequantize(data, scale, N, K, 8)
    assert torch.equal(back.cpu(), Q.q8_dequantize(want_data, want_scale))


def _dense_twin(model, ckpt):
    """bf16 model whose quantised linears carry the dequantised weights of `model`."""
    import esme
    from esme.quantization import _QuantLinear
    twin = esme.ESM.from_pretrained(ckpt, device='cuda')
    mods = dict(model.named_modules())
    with torch.no_grad():
        for name, m in twin.named_modules():
            if isinstance(mods.get(name), _QuantLinear):
                m.weight.copy_(mods[name].dequantize())
    return twin


@pytest.mark.parametrize('ckpt,fixture', [('esm2_8m.safetensors', 'esm2_8m_testfa.npz'),
                                          ('esmc_tiny.safetensors', 'esmc_tiny.npz')])
@pytest.mark.parametrize('mode', ['4bit', '8bit'])
def test_quantised_loader_contract_and_engine_path(ckpt, fixture, mode):
    import esme
    from safetensors import safe_open
    path = os.path.join(GOLDEN, ckpt)
    model = esme.ESM.from_pretrained(path, quantization=mode, device='cuda')
    params = model.state_dict()
    want_dtype = torch.uint8 if mode == '4bit' else torch.int8
    n_quant = 0
    for k, v in params.items():                       # tests/test_esm.py:123-154
        if any(k.endswith(f'.{l}.weight') for l in ('q', 'k', 'v', 'out', 'final.1', 'final.3', 'final.2',
                                                    'activation', 'fc')) and 'lm_head' not in k:
            assert v.dtype == want_dtype, k
            n_quant += 1
    assert n_quant == (6 if 'esm2' in ckpt else 7) * model.num_layers
    with safe_open(path, framework='pt', device='cuda') as f:
        for k in f.keys():
            if 'bias' in k:
                assert torch.equal(params[k], f.get_tensor(k)), k
        if 'esm2' in ckpt:
            assert torch.equal(model.lm_head.final.weight, model.embed_tokens.weight)
    z = np.load(os.path.join(GOLDEN, fixture))
    tokens, cu, max_len = torch.from_numpy(z['tokens']).cuda(), torch.from_numpy(z['cu_lens']).cuda(), int(z['max_len'])
    got = model(tokens, (cu, max_len))
    twin = _dense_twin(model, path)
    want = twin(tokens, (cu, max_len))
    assert torch.equal(got, want)                      # same kernels, same (dequantised) weights
    # operator-level path (FlashTransformerLayer.forward on quantised modules) == engine path
    x = model.embedding(tokens)
    for a, b in zip(model.layers[:2], twin.layers[:2]):
        assert torch.equal(a(x, cu, max_len), b(x, cu, max_len))
    # and the quantised model stays close to the bf16 one (loose: this is the format's error, not the kernels')
    ref = torch.from_numpy(z['logits'].view(np.int16)).view(torch.bfloat16).float()
    cos = torch.nn.functional.cosine_similarity(got.float().cpu(), ref, dim=-1)
    assert cos.min() > (0.90 if mode == '4bit' else 0.995), cos.min()
    # mixed selection: FFN only (BASELINE config 5)
    from esme.quantization import quantize_model_
    ffn_only = quantize_model_(esme.ESM.from_pretrained(path, device='cuda'), 4, which=('ffn',))
    assert torch.equal(ffn_only(tokens, (cu, max_len)), _dense_twin(ffn_only, path)(tokens, (cu, max_len)))
