"""Model-level parity on the GPU: the drop-in esme API (C ABI underneath) against
  (1) the committed logits of the real reference (tests/golden, CPU bf16 path),
  (2) the oracle in exact (fp64) mode -- the tolerance statement:

      rms_rel(new, exact) <= 1.5 x rms_rel(reference, exact)      (same inputs)
      min row cosine >= 0.9999, argmax agreement >= 98.5 %

i.e. the CUDA path may be no further from the exact result than the reference's own
bf16 noise floor (2.3e-3 RMS-relative on ESM2-8M); elementwise 1e-3
agreement between two bf16 pipelines is below that floor and is not claimed."""
import pytest
import torch

import esme
from conftest import GOLDEN, err_stats, load_golden
from esme.alphabet import Alphabet
from oracle import esm_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'

CASES = [('esm2_8m.safetensors', 'esm2_8m_cfg1.npz'), ('esm2_8m.safetensors', 'esm2_8m_testfa.npz'),
         ('esm2_tiny.safetensors', 'esm2_tiny.npz'), ('esmc_tiny.safetensors', 'esmc_tiny.npz')]


def _load(ckpt):
    return esme.ESM.from_pretrained(f'{GOLDEN}/{ckpt}', device=DEV)


@pytest.mark.parametrize('ckpt,fixture', CASES)
def test_packed_forward_matches_reference_and_oracle(ckpt, fixture):
    g = load_golden(fixture)
    model = _load(ckpt)
    cfg, W = O.load_checkpoint(f'{GOLDEN}/{ckpt}')
    tokens, cu, max_len = g['tokens'], g['cu_lens'], g['max_len']
    got = model(tokens.to(DEV), (cu.to(DEV), max_len))
    assert got.dtype == torch.bfloat16 and got.shape == g['logits'].shape
    got = got.float().cpu()
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    _, rms_new, cos_new, agree = err_stats(got, exact)
    _, rms_ref, _, _ = err_stats(g['logits'], exact)
    _, rms_pair, cos_pair, _ = err_stats(got, g['logits'])
    assert rms_new <= 1.5 * rms_ref + 1e-4
    assert rms_pair <= 2.5 * rms_ref + 1e-4
    assert cos_new >= 0.9999 and cos_pair >= 0.9999 and agree >= 0.985
    lp = model.predict_log_prob(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    assert err_stats(lp, g['log_prob'])[1] <= 2.5 * rms_ref + 1e-3
    pr = model.predict_prob(tokens.to(DEV), pad_args=(cu.to(DEV), max_len)).float().cpu()
    assert torch.allclose(pr.sum(-1), torch.ones(pr.shape[0]), atol=2e-2)
    rep = model.forward_representation(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    rs = err_stats(rep, g['representation'])
    assert rs[1] < 1.5e-2 and rs[2] > 0.998


@pytest.mark.parametrize('ckpt,fixture', [CASES[0], CASES[2], CASES[3]])
def test_layer_module_matches_reference_taps(ckpt, fixture):
    """FlashTransformerLayer / RotaryEmbedding operator boundaries vs the reference's
    layer-0 outputs (hooks in make_golden.py)."""
    g = load_golden(fixture)
    model = _load(ckpt)
    tokens, cu = g['tokens'].to(DEV), g['cu_lens'].to(DEV)
    x0 = model.embedding(tokens)
    y0 = model.layers[0](x0, cu, g['max_len']).float().cpu()
    _, rms, cos, _ = err_stats(y0, g['layer0.x_out'])
    assert rms < 3e-3 and cos > 0.9999
    sa = model.layers[0].self_attn
    qkv, _ = sa._qkv_rot(x0, cu, g['max_len'])
    D = model.embed_dim
    q = qkv[:, :D].float().cpu().reshape(g['layer0.q_rot'].shape)
    k = qkv[:, D:2 * D].float().cpu().reshape(g['layer0.k_rot'].shape)
    assert err_stats(q.flatten(1), g['layer0.q_rot'].flatten(1))[1] < 1e-3
    assert err_stats(k.flatten(1), g['layer0.k_rot'].flatten(1))[1] < 1e-3


def test_rotary_module_matches_reference_bf16():
    g = load_golden('rope.npz')
    rot = esme.rotary.RotaryEmbedding(dim=64)
    q, k = g['q'].bfloat16().to(DEV), g['k'].bfloat16().to(DEV)
    qr, kr = rot(q, k, g['cu_lens'].to(DEV), g['max_len'])
    assert torch.equal(rot._cos_cached.float().cpu(), g['cos_bf16'])
    assert (rot._sin_cached.float().cpu() - g['sin_bf16']).abs().max() < 1e-5
    mism = (qr.float().cpu() != g['q_rot_bf16']).float().mean().item()
    assert mism < 1e-3 and (kr.float().cpu() - g['k_rot_bf16']).abs().max() < 1e-2
    assert torch.equal(q.float().cpu(), g['q'].bfloat16().float())      # inputs untouched


def test_padded_entry_matches_reference():
    model = _load('esm2_8m.safetensors')
    g = load_golden('esm2_8m_padded.npz')
    got = model(g['tokens'].to(DEV))
    assert got.shape == g['logits'].shape
    _, rms, cos, agree = err_stats(got.float().cpu(), g['logits'])
    assert rms < 6e-3 and cos > 0.9999 and agree > 0.98
    pad = g['tokens'] == Alphabet.padding_idx
    rows = got.float().cpu()[pad]
    assert (rows - rows[0]).abs().max() == 0                            # constant lm_head(0) rows
    lp = model.predict_log_prob(g['tokens'].to(DEV)).float().cpu()
    assert err_stats(lp, g['log_prob'])[1] < 8e-3


def test_packed_equals_padded_and_batch_invariance():
    """Reference tests/test_esm.py:70-81 (packed vs padded) -- here bit-exact, plus batch
    composition invariance (a sequence's logits do not depend on its neighbours)."""
    model = _load('esm2_8m.safetensors')
    g = load_golden('esm2_8m_cfg1.npz')
    tokens, cu, max_len = g['tokens'].to(DEV), g['cu_lens'].to(DEV), g['max_len']
    packed = model(tokens, (cu, max_len))
    padded = model(esme.tokenize(g['seqs'], Alphabet).to(DEV))
    assert torch.equal(padded.reshape(-1, 33), packed)
    first = model(tokens[:66].contiguous(), (cu[:2].contiguous(), 66))
    assert torch.equal(first, packed[:66])
    # pad_output=True with explicit indices
    from esme.alphabet import tokenize_unpad
    t, idx, c, ml = tokenize_unpad(g['seqs'], Alphabet)
    po = model(t.to(DEV), (c.to(DEV), ml), pad_output=True, pad_indices=idx.to(DEV))
    assert torch.equal(po, padded)


def test_intermediate_layers_and_errors():
    model = _load('esm2_tiny.safetensors')
    g = load_golden('esm2_tiny.npz')
    tokens, cu, max_len = g['tokens'].to(DEV), g['cu_lens'].to(DEV), g['max_len']
    rep = model.forward_representation(tokens, (cu, max_len), layers=[0])
    assert rep.shape == (tokens.numel(), 2 * model.embed_dim)
    _, rms, cos, _ = err_stats(rep[:, model.embed_dim:].float().cpu(), g['layer0.x_out'])
    assert rms < 3e-3 and cos > 0.9999
    with pytest.raises(AssertionError):
        model(tokens.reshape(1, -1), (cu, max_len))                     # 2-D tokens with pad_args
    with pytest.raises(AssertionError):
        model(tokens)                                                   # 1-D tokens without pad_args


def test_full_size_650m_properties():
    """BASELINE config 2 shape (ESM2-650M, ~50k packed tokens, synthetic weights): properties
    that do not need an oracle run at full size."""
    cfg = O.OracleConfig('esm2', 33, 1280, 20)
    W = O.synthetic_weights(cfg, seed=1)
    model = esme.ESM2(33, 1280, 20)
    model.load_state_dict(W, strict=True)
    model = model.to(DEV).eval()
    lens = O.synthetic_lengths(50000, seed=2)
    tokens, cu, max_len = O.synthetic_batch(lens, seed=3)
    logp = model.predict_log_prob(tokens.to(DEV), (cu.to(DEV), max_len))
    assert logp.shape == (tokens.numel(), 33) and torch.isfinite(logp.float()).all()
    assert torch.allclose(logp.float().exp().sum(-1), torch.ones(tokens.numel(), device=DEV), atol=3e-2)
    logits = model(tokens.to(DEV), (cu.to(DEV), max_len))
    assert torch.equal(logits, model(tokens.to(DEV), (cu.to(DEV), max_len)))      # deterministic
    # batch-composition invariance at full size: sequences 3..5 alone == inside the 50k batch
    a, b = int(cu[3]), int(cu[6])
    sub_cu = (cu[3:7] - cu[3]).to(DEV)
    alone = model(tokens[a:b].to(DEV), (sub_cu, int((cu[4:7] - cu[3:6]).max())))
    assert torch.equal(alone, logits[a:b])
    # a 2-sequence prefix against the oracle (finishes in seconds on CPU)
    n = int(cu[2])
    got = logits[:n].float().cpu()
    exact = O.forward_packed(cfg, W, tokens[:n], cu[:3], int((cu[1:3] - cu[:2]).max()), 'fp32')
    want = O.forward_packed(cfg, W, tokens[:n], cu[:3], int((cu[1:3] - cu[:2]).max()), 'bf16')
    _, rms_new, cos_new, agree = err_stats(got, exact)
    _, rms_orc, cos_orc, agree_orc = err_stats(want, exact)
    # 33 layers of random weights amplify bf16 rounding noise to ~1.5 % RMS: the CUDA path must sit at
    # the same distance from the exact result as the bf16-faithful oracle does (no worse, same inputs)
    assert rms_new <= 1.5 * rms_orc + 1e-4
    assert cos_new >= min(0.9999, cos_orc - 2e-4) and agree >= agree_orc - 0.02


def test_head_dim_24_model_matches_oracle():
    """ESM2-35M geometry (embed_dim 480, 20 heads -> head_dim 24; the reference dispatches it to flash-attn's
    hd<=32 kernel): stand-alone pair-wise rotary kernel + the tcgen05 attention kernel on heads zero-padded to 32
    columns, against the oracle on seeded synthetic weights (no 35M checkpoint exists offline)."""
    from esme import synthetic
    layers, D, H = 2, 480, 20
    sd = synthetic.synthetic_state_dict('esm2', layers, D, seed=7)
    model = esme.ESM2(layers, D, H)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    lens = [40, 129, 7]
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=9)
    got = model(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    cfg = O.OracleConfig('esm2', layers, D, H)
    W = {k: v.clone() for k, v in sd.items()}
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    want = O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16').float()
    _, rms_new, cos_new, agree = err_stats(got, exact)
    _, rms_orc, _, agree_orc = err_stats(want, exact)
    # random synthetic weights leave near-ties in the logits: argmax agreement is judged next to the oracle's own
    assert rms_new <= 1.5 * rms_orc + 1e-4 and cos_new >= 0.9999 and agree >= min(0.985, agree_orc - 0.02)
    # the rotation itself is bit-exact against the oracle's rotary restatement
    T = tokens.numel()
    g = torch.Generator().manual_seed(3)
    q = torch.randn(T, H, 24, generator=g).bfloat16()
    k = torch.randn(T, H, 24, generator=g).bfloat16()
    rot = esme.rotary.RotaryEmbedding(dim=24)
    qr, kr = rot(q.to(DEV), k.to(DEV), cu.to(DEV), max_len)
    pos = O.positions_from_cu_lens(cu)
    cos, sin = O.rotary_tables(max_len, 24, O._Prec('bf16'))
    p = O._Prec('bf16')
    assert torch.equal(qr.float().cpu(), O.apply_rotary(q.float(), cos, sin, pos, p))
    assert torch.equal(kr.float().cpu(), O.apply_rotary(k.float(), cos, sin, pos, p))


@pytest.mark.parametrize('name', ['esm1b', 'esm1v'])
def test_esm1_models_match_reference(name):
    """ESM-1b / ESM-1v drop-ins (reference esme/esm.py:618-735, esme/embedding.py): engine path, operator-level
    embedding, padded entry -- against the real reference's outputs on tiny models and the oracle."""
    g = load_golden(f'{name}_tiny.npz')
    model = esme.ESM.from_pretrained(f'{GOLDEN}/{name}_tiny.safetensors', device=DEV)
    assert type(model).__name__ == ('ESM1b' if name == 'esm1b' else 'ESM1v')
    assert all(layer.self_attn.rot_emb is None for layer in model.layers)
    tokens, cu, max_len = g['tokens'].to(DEV), g['cu_lens'].to(DEV), g['max_len']
    cfg, W = O.load_checkpoint(f'{GOLDEN}/{name}_tiny.safetensors')
    exact = O.forward_packed(cfg, W, g['tokens'], g['cu_lens'], max_len, 'fp64').float()
    got = model(tokens, (cu, max_len)).float().cpu()
    _, rms_new, cos_new, agree = err_stats(got, exact)
    _, rms_ref, cos_ref, agree_ref = err_stats(g['logits'], exact)       # (~1e-2: random weights with a large gain)
    assert rms_new <= 1.5 * rms_ref + 1e-4 and cos_new >= cos_ref - 2e-4 and agree >= agree_ref - 0.02
    emb = model.embedding(tokens, (cu, max_len)).float().cpu()
    assert (emb != g['embedding']).float().mean() < 2e-3                 # gather + add (+ LayerNorm for 1b)
    rep = model.forward_representation(tokens, (cu, max_len)).float().cpu()
    assert err_stats(rep, g["representation"])[1] < 3e-2
    padded = model(esme.tokenize(g['seqs'], Alphabet).to(DEV)).float().cpu()
    assert padded.shape == g['padded_logits'].shape
    _, rms, cos, _ = err_stats(padded, g['padded_logits'])
    assert rms < 1.5 * rms_ref and cos > 0.999
    with pytest.raises(AssertionError):
        model.embedding(tokens)                                           # packed tokens need pad_args
    # mask-margin scoring runs on these models too (ESM-1v is the variant-effect model of the family)
    from esme.variant import predict_mask_margin
    df = predict_mask_margin(model, 'MKTAYIAKQRQISFVKSHFSRQLEERLGL', batch_size=8)
    assert df.shape[0] == 29 * 20 and torch.isfinite(torch.tensor(df['score'].to_numpy())).all()
