"""Regression tests for the round-1 review findings: device-detected data errors, LoRA + mask-margin,
`layers=` ordering, dtype guard of the engine, learned-position range check, second-GPU operation."""
import pytest
import torch

import esme
from conftest import GOLDEN, load_golden
from esme import _lib, ops
from esme.variant import predict_mask_margin

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def test_out_of_range_token_is_reported_at_the_next_call():
    """The embedding kernel cannot return a status: an id outside the table sets a sticky error word (mapped
    pinned memory) and the NEXT entry point fails, naming the cause (the reference device-asserts)."""
    table = torch.randn(33, 64, device=DEV).bfloat16()
    ok = ops.embed(torch.tensor([0, 5, 32], device=DEV), table)
    torch.cuda.synchronize()
    assert _lib.lib.esmk_async_error() == 0 and torch.equal(ok[1], table[5])
    ops.embed(torch.tensor([0, 33, 2], device=DEV), table)
    torch.cuda.synchronize()
    with pytest.raises(_lib.EsmkError, match='token id outside'):
        ops.embed(torch.tensor([0, 1], device=DEV), table)
    ops.embed(torch.tensor([0, 1], device=DEV), table)              # flag cleared by the report
    torch.cuda.synchronize()
    ops.embed(torch.tensor([-1], device=DEV), table)
    torch.cuda.synchronize()
    assert _lib.lib.esmk_async_error() == 1 and _lib.lib.esmk_async_error() == 0


def test_mask_margin_on_a_lora_model():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device=DEV)
    seq = 'MKTAYIAKQRQISFVKSHFSRQ'
    base = predict_mask_margin(model, seq, batch_size=8)
    model.add_lora(rank=8, alpha=8, adapter_names=['a'])             # fresh adapters are a no-op (B = 0)
    with_lora = predict_mask_margin(model, seq, batch_size=8)
    assert with_lora.shape == base.shape
    assert (torch.tensor(with_lora['score'].to_numpy()) - torch.tensor(base['score'].to_numpy())).abs().max() < 0.25


def test_intermediate_layers_come_back_in_ascending_order_once():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device=DEV)
    g = load_golden('esm2_tiny.npz')
    tokens, cu, max_len = g['tokens'].to(DEV), g['cu_lens'].to(DEV), g['max_len']
    D = model.embed_dim
    asc = model.forward_representation(tokens, (cu, max_len), layers=[0, 1])
    assert asc.shape[1] == 3 * D
    assert torch.equal(model.forward_representation(tokens, (cu, max_len), layers=[1, 0, 1]), asc)
    model.add_lora(rank=8, alpha=8, adapter_names=['a'])             # operator path: same layout
    lo = model.forward_representation(tokens, (cu, max_len), layers=[1, 0])
    assert lo.shape == asc.shape and (lo.float() - asc.float()).abs().max() < 0.1


def test_engine_rejects_non_bf16_storage():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device=DEV)
    g = load_golden('esm2_tiny.npz')
    tokens, cu, max_len = g['tokens'].to(DEV), g['cu_lens'].to(DEV), g['max_len']
    want = model(tokens, (cu, max_len))
    import copy
    with pytest.raises(RuntimeError, match='bf16 weights only'):
        copy.deepcopy(model).half()(tokens, (cu, max_len))
    with pytest.raises(RuntimeError, match='bf16 weights only'):
        copy.deepcopy(model).float()(tokens, (cu, max_len))
    assert torch.equal(model.bfloat16()(tokens, (cu, max_len)), want)
    q = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', quantization='4bit', device=DEV)
    y = q(tokens, (cu, max_len))
    assert torch.equal(q.to(torch.bfloat16)(tokens, (cu, max_len)), y)          # fp32 block scales survive the cast
    assert q.layers[0].self_attn.q.scale.dtype == torch.float32


def test_learned_positions_reject_too_long_sequences():
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm1v_tiny.safetensors', device=DEV)
    n = model.embed_positions.max_positions + 1
    tokens = torch.full((n,), 5, dtype=torch.int64, device=DEV)
    cu = torch.tensor([0, n], dtype=torch.int32, device=DEV)
    with pytest.raises(ValueError, match='above maximum'):
        model(tokens, (cu, n))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason='needs two GPUs')
def test_second_gpu_in_the_same_process():
    """Kernel attributes (dynamic shared memory opt-in) and the SM count are per device; the operator wrappers
    launch on the device of their tensors, not on the current one."""
    g = load_golden('esm2_tiny.npz')
    outs = []
    for dev in ('cuda:0', 'cuda:1'):
        model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', device=dev)
        outs.append(model(g['tokens'].to(dev), (g['cu_lens'].to(dev), g['max_len'])).cpu())
        x = torch.randn(64, 256, device=dev).bfloat16()
        w = torch.randn(128, 256, device=dev).bfloat16()
        y = ops.linear(x, w)                                           # current device stays cuda:0
        assert (y.float() - x.float() @ w.float().T).abs().max() < 0.5
        q4 = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', quantization='4bit', device=dev)
        assert torch.isfinite(q4(g['tokens'].to(dev), (g['cu_lens'].to(dev), g['max_len'])).float()).all()
    assert torch.equal(outs[0], outs[1])
