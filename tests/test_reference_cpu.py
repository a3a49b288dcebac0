"""oracle/_ref (the unmodified reference package, oracle/build_ref.py) run out of process on the CPU:
it must reproduce the committed golden logits bit for bit (same code, same SDPA substitution as
tests/golden/make_golden.py), and the oracle restatement must sit at its bf16 noise level.  Skipped where
oracle/_ref is absent."""
import pytest
import torch

from conftest import GOLDEN, err_stats, load_golden
from oracle import esm_oracle as O
from oracle import ref_client as RC

needs_ref = pytest.mark.skipif(not RC.ref_available(), reason='oracle/_ref absent (python oracle/build_ref.py)')


@needs_ref
def test_ref_runner_reproduces_golden_and_pins_oracle():
    g = load_golden('esm2_8m_cfg1.npz')
    ckpt = f'{GOLDEN}/esm2_8m.safetensors'
    batch = (g['tokens'], g['cu_lens'], g['max_len'])
    info, ref = RC.run_reference(dict(device='cpu', family='esm2', num_layers=6, embed_dim=320, attention_heads=20,
                                      weights={'safetensors': ckpt}, mode='forward'), batch=batch)
    assert '/oracle/_ref/esme_ref.zip/esme/' in info['reference_file'] and 'SDPA' in info['attention']
    assert torch.equal(ref['logits'], g['logits']) and torch.equal(ref['log_prob'], g['log_prob'])
    cfg, W = O.load_checkpoint(ckpt)
    exact = O.forward_packed(cfg, W, *batch, 'fp64').float()
    want = O.forward_packed(cfg, W, *batch, 'bf16').float()
    assert err_stats(want, exact)[1] <= 1.5 * err_stats(ref['logits'], exact)[1]


@needs_ref
def test_ref_runner_synthetic_esmc_matches_oracle():
    """Synthetic weights are generated inside the runner from esme/synthetic.py (loaded by path): same tensors
    as in this process, so the oracle and the real reference see identical ESMC models."""
    from esme import synthetic
    family, layers, D, H = 'esmc', 2, 192, 3
    tokens, cu, max_len = synthetic.synthetic_batch([40, 129, 7], seed=9)
    _, ref = RC.run_reference(dict(device='cpu', family=family, num_layers=layers, embed_dim=D, attention_heads=H,
                                   weights={'synthetic_seed': 5}, mode='forward'), batch=(tokens, cu, max_len))
    W = synthetic.synthetic_state_dict(family, layers, D, seed=5)
    cfg = O.OracleConfig(family, layers, D, H)
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    want = O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16').float()
    _, rms_ref, _, _ = err_stats(ref['logits'], exact)
    _, rms_orc, cos, _ = err_stats(want, exact)
    assert rms_orc <= 1.5 * rms_ref + 1e-4 and cos > 0.9999
    assert err_stats(want, ref['logits'])[1] <= 2.5 * rms_ref + 1e-4
