"""Parity against the reference's ACTUAL GPU path: the unmodified `esme` package (oracle/_ref, placed there by
oracle/build_ref.py) with the real `flash_attn_varlen_func`, run on the same B200 in a child interpreter
(oracle/ref_runner.py), versus this package on identical tokens and weights.

Bars (same inputs, `exact` = the oracle in fp64):
    rms_rel(new, exact) <= 1.5 x rms_rel(reference, exact)   -- no further from the truth than the reference's
    rms_rel(new, reference) <= 2.5 x rms_rel(reference, exact)   own bf16 noise floor allows
    min row cosine(new, reference) > 0.99                     -- the reference's own bar (its tests/test_esm.py:45-81)
    argmax agreement with the reference, judged next to the reference's own agreement with `exact`
at BASELINE config 1 (real ESM2-8M weights), the 16-protein test.fa batch, and 2-layer models at the dimensions
of BASELINE configs 2, 3 and 4 (ESM2-650M 1280/20, ESMC-300M 960/15 F 2560 V 64, ESM2-3B 2560/40 F 10240).
The second half of the file repeats the dimension cases against the oracle alone, so they also run where
oracle/_ref is absent."""
import pytest
import torch

import esme
from conftest import GOLDEN, err_stats, load_golden
from esme import synthetic
from oracle import esm_oracle as O
from oracle import ref_client as RC

pytestmark = pytest.mark.gpu
DEV = 'cuda'
needs_ref = pytest.mark.skipif(not RC.ref_available(), reason='oracle/_ref absent (python oracle/build_ref.py)')

# (family, layers, D, H, sequence lengths incl. cls/eos)
DIM_CASES = {
    'esm2_650m_dims': ('esm2', 2, 1280, 20, [130, 700, 66, 257, 31]),
    'esmc_300m_dims': ('esmc', 2, 960, 15, [130, 2050, 418, 1021, 257]),      # config 3: mixed 128-2048 residues
    'esm2_3b_dims': ('esm2', 2, 2560, 40, [402, 129, 64, 300]),
}


def _new_model(family, layers, D, H, state):
    cls = esme.ESMC if family == 'esmc' else esme.ESM2
    model = cls(layers, D, H)
    model.load_state_dict(state, strict=True)
    return model.to(DEV).eval().requires_grad_(False)


def _check(new, ref, exact, what):
    _, rms_new, cos_new, agree_new = err_stats(new, exact)
    _, rms_ref, cos_ref, agree_ref = err_stats(ref, exact)
    _, rms_pair, cos_pair, agree_pair = err_stats(new, ref)
    print(f'{what}: rms-rel vs fp64 new={rms_new:.3e} reference={rms_ref:.3e}; new-vs-reference={rms_pair:.3e} '
          f'cos={cos_pair:.6f} argmax new/ref vs exact={agree_new:.4f}/{agree_ref:.4f} new-vs-ref={agree_pair:.4f}')
    assert rms_new <= 1.5 * rms_ref + 1e-4
    assert rms_pair <= 2.5 * rms_ref + 1e-4
    assert cos_pair > 0.99 and cos_new >= min(0.9999, cos_ref - 2e-4)
    assert agree_new >= min(0.985, agree_ref - 0.02)


@needs_ref
@pytest.mark.parametrize('fixture', ['esm2_8m_cfg1.npz', 'esm2_8m_testfa.npz'])
def test_real_weights_against_reference_flash_attn(fixture):
    """BASELINE config 1 (ESM2-8M real weights, 2 x 64 residues) and the reference's test.fa batch."""
    g = load_golden(fixture)
    ckpt = f'{GOLDEN}/esm2_8m.safetensors'
    batch = (g['tokens'], g['cu_lens'], g['max_len'])
    info, ref = RC.run_reference(dict(device='cuda', family='esm2', num_layers=6, embed_dim=320, attention_heads=20,
                                      weights={'safetensors': ckpt}, mode='forward'), batch=batch)
    assert 'flash_attn' in info['attention'] and '/oracle/_ref/esme_ref.zip/esme/' in info['reference_file']
    model = esme.ESM.from_pretrained(ckpt, device=DEV)
    cfg, W = O.load_checkpoint(ckpt)
    exact = O.forward_packed(cfg, W, *batch, 'fp64')
    tok, cu = g['tokens'].to(DEV), g['cu_lens'].to(DEV)
    _check(model(tok, (cu, g['max_len'])).float().cpu(), ref['logits'], exact.float(), f'{fixture} logits')
    _check(model.predict_log_prob(tok, (cu, g['max_len'])).float().cpu(), ref['log_prob'],
           O.log_softmax(exact, 'fp64').float(), f'{fixture} log_prob')
    rep = model.forward_representation(tok, (cu, g['max_len'])).float().cpu()
    assert err_stats(rep, ref['representation'])[1] < 1.5e-2
    # the GPU reference and the committed CPU-substituted goldens are the same model at bf16 noise level
    assert err_stats(ref['logits'], g['logits'])[1] < 6e-3


@needs_ref
@pytest.mark.parametrize('case', sorted(DIM_CASES))
def test_config_dims_against_reference_flash_attn(case):
    family, layers, D, H, lens = DIM_CASES[case]
    state = synthetic.synthetic_state_dict(family, layers, D, seed=11)
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=13)
    info, ref = RC.run_reference(dict(device='cuda', family=family, num_layers=layers, embed_dim=D, attention_heads=H,
                                      weights={'synthetic_seed': 11}, mode='forward'), batch=(tokens, cu, max_len))
    assert 'flash_attn' in info['attention']
    model = _new_model(family, layers, D, H, state)
    cfg = O.OracleConfig(family, layers, D, H)
    W = {k: v.clone() for k, v in state.items()}
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64')
    _check(model(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu(), ref['logits'], exact.float(), f'{case} logits')
    _check(model.predict_log_prob(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu(), ref['log_prob'],
           O.log_softmax(exact, 'fp64').float(), f'{case} log_prob')


@needs_ref
def test_padded_entry_against_reference_flash_attn():
    g = load_golden('esm2_8m_padded.npz')
    ckpt = f'{GOLDEN}/esm2_8m.safetensors'
    cu = torch.tensor([0, 3], dtype=torch.int32)
    _, ref = RC.run_reference(dict(device='cuda', family='esm2', num_layers=6, embed_dim=320, attention_heads=20,
                                   weights={'safetensors': ckpt}, mode='forward'),
                              batch=(torch.tensor([0, 5, 2]), cu, 3), tokens2d=g['tokens'])
    model = esme.ESM.from_pretrained(ckpt, device=DEV)
    got = model(g['tokens'].to(DEV)).float().cpu()
    assert got.shape == ref['logits_padded'].shape
    _, rms, cos, agree = err_stats(got, ref['logits_padded'])
    assert rms < 6e-3 and cos > 0.9999 and agree > 0.98


@pytest.mark.parametrize('case', sorted(DIM_CASES))
def test_config_dims_against_oracle(case):
    """Model-level oracle parity at the BASELINE dimensions (runs with or without oracle/_ref)."""
    family, layers, D, H, lens = DIM_CASES[case]
    lens = [min(l, 600) for l in lens]                       # keeps the fp64 oracle to seconds on the host
    state = synthetic.synthetic_state_dict(family, layers, D, seed=11)
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=13)
    model = _new_model(family, layers, D, H, state)
    cfg = O.OracleConfig(family, layers, D, H)
    W = {k: v.clone() for k, v in state.items()}
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64')
    want = O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16')
    got = model(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    _check(got, want.float(), exact.float(), f'{case} logits vs oracle')
    lp = model.predict_log_prob(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    _check(lp, O.log_softmax(want, 'bf16').float(), O.log_softmax(exact, 'fp64').float(), f'{case} log_prob vs oracle')
