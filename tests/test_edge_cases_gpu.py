"""Edge cases of the packed forward on the GPU (the reference's tests cover ragged and padded inputs, single
sequences and long proteins: tests/test_esm.py, tests/test_attention.py:228-250; its benchmark goes to 3,500
residues, workflow/inference/inference_on_human.py:12): shapes the tile scheduler, the trimmed last key block and
the TMA-store epilogue must survive, each checked against the bf16 oracle or an invariance the domain offers."""
import pytest
import torch

import esme
from conftest import GOLDEN, err_stats
from esme import synthetic
from oracle import esm_oracle as O

pytestmark = pytest.mark.gpu
DEV = 'cuda'


def _tiny(family='esm2', layers=2, D=256, H=4, seed=5):
    sd = synthetic.synthetic_state_dict(family, layers, D, seed=seed)
    cls = esme.ESMC if family == 'esmc' else esme.ESM2
    model = cls(layers, D, H)
    model.load_state_dict(sd, strict=True)
    return model.to(DEV).eval(), O.OracleConfig(family, layers, D, H), {k: v.clone() for k, v in sd.items()}


@pytest.mark.parametrize('lens', [[3], [129], [128, 128], [2, 2, 2, 2, 2], [127, 1 + 128, 3, 257, 64, 65]])
def test_ragged_shapes_against_oracle(lens):
    model, cfg, W = _tiny()
    tokens, cu, max_len = synthetic.synthetic_batch([max(l, 3) for l in lens], seed=sum(lens))
    got = model(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    want = O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16').float()
    _, rms_new, cos_new, _ = err_stats(got, exact)
    _, rms_orc, _, _ = err_stats(want, exact)
    assert torch.isfinite(got).all()
    assert rms_new <= 1.5 * rms_orc + 1e-4 and cos_new >= 0.9999


def test_longest_benchmark_protein_is_batch_invariant():
    """One 3,500-residue protein (28 key blocks per query tile) packed between short ones: same logits as alone."""
    model, cfg, W = _tiny(H=4, D=256)
    lens = [40, 3502, 77]
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=1)
    full = model(tokens.to(DEV), (cu.to(DEV), max_len))
    alone = model(tokens[40:3542].contiguous().to(DEV), (torch.tensor([0, 3502], dtype=torch.int32, device=DEV), 3502))
    assert torch.isfinite(full.float()).all()
    assert torch.equal(full[40:3542], alone)


def test_many_short_sequences():
    """1,000 sequences of 5-40 tokens: more work-list records than 128-row tiles of packed tokens."""
    model, cfg, W = _tiny()
    g = torch.Generator().manual_seed(2)
    lens = torch.randint(5, 41, (1000,), generator=g).tolist()
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=4)
    got = model(tokens.to(DEV), (cu.to(DEV), max_len))
    assert torch.isfinite(got.float()).all()
    pick = [0, 1, 500, 999]
    for s in pick:
        a, b = int(cu[s]), int(cu[s + 1])
        alone = model(tokens[a:b].contiguous().to(DEV), (torch.tensor([0, b - a], dtype=torch.int32, device=DEV), b - a))
        assert torch.equal(got[a:b], alone), s


@pytest.mark.parametrize('quantization', [None, '4bit'])
def test_forward_is_cuda_graph_capturable_and_stream_ordered(quantization):
    """esmk_forward is a fixed launch sequence with no host synchronisation: it can be captured into a CUDA graph
    (the reference's forward cannot: rotary.py:5-14 synchronises twice per layer) and replayed on new inputs.
    Quantised models fork onto a library-owned side stream (weight expansion two GEMMs ahead) and join back inside
    the forward, which capture must follow."""
    model = esme.ESM.from_pretrained(f'{GOLDEN}/esm2_tiny.safetensors', quantization=quantization, device=DEV)
    lens = [200, 131, 515]
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=8)
    tokens, cu = tokens.to(DEV), cu.to(DEV)
    eager = model(tokens, (cu, max_len))
    static_tokens = tokens.clone()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        model(static_tokens, (cu, max_len))                        # warm-up on the capture stream (workspace allocation)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        static_out = model(static_tokens, (cu, max_len))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, eager)
    other, _, _ = synthetic.synthetic_batch(lens, seed=9)
    static_tokens.copy_(other.to(DEV))
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(static_out, model(other.to(DEV), (cu, max_len)))
    assert not torch.equal(static_out, eager)


def test_esmc_ragged_against_oracle():
    model, cfg, W = _tiny('esmc', layers=2, D=192, H=3, seed=6)
    tokens, cu, max_len = synthetic.synthetic_batch([130, 3, 64, 300], seed=12)
    got = model.predict_log_prob(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    exact = O.log_softmax(O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64'), 'fp64').float()
    want = O.log_softmax(O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16'), 'bf16').float()
    _, rms_new, cos_new, _ = err_stats(got, exact)
    _, rms_orc, _, _ = err_stats(want, exact)
    assert rms_new <= 1.5 * rms_orc + 1e-3 and cos_new >= 0.9999


def test_head_dim_128_model_matches_oracle():
    """ESM2-15B geometry in miniature (embed_dim 256, 2 heads -> head_dim 128): stand-alone rotary kernel + the
    tcgen05 attention kernel's head_dim-128 instantiation, against the oracle."""
    model, cfg, W = _tiny('esm2', layers=2, D=256, H=2, seed=8)
    tokens, cu, max_len = synthetic.synthetic_batch([130, 5, 64, 300], seed=14)
    got = model(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    want = O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16').float()
    _, rms_new, cos_new, _ = err_stats(got, exact)
    _, rms_orc, _, _ = err_stats(want, exact)
    assert rms_new <= 1.5 * rms_orc + 1e-4 and cos_new >= 0.9999


def test_head_dim_32_model_matches_oracle():
    """ESM2-150M geometry in miniature (embed_dim 128, 4 heads -> head_dim 32): fused QKV + RoPE epilogue and the
    tcgen05 attention kernel's native head_dim-32 instantiation (64-byte TMA rows), against the oracle."""
    model, cfg, W = _tiny('esm2', layers=2, D=128, H=4, seed=9)
    tokens, cu, max_len = synthetic.synthetic_batch([130, 5, 64, 300], seed=15)
    got = model(tokens.to(DEV), (cu.to(DEV), max_len)).float().cpu()
    exact = O.forward_packed(cfg, W, tokens, cu, max_len, 'fp64').float()
    want = O.forward_packed(cfg, W, tokens, cu, max_len, 'bf16').float()
    _, rms_new, cos_new, _ = err_stats(got, exact)
    _, rms_orc, _, _ = err_stats(want, exact)
    assert rms_new <= 1.5 * rms_orc + 1e-4 and cos_new >= 0.9999


def test_attention_work_list_longer_than_one_launch():
    """More than 65,535 query tiles (here 70,000 two- and three-token sequences): the work list is walked in several
    launches; result against the CUDA-core kernel."""
    from esme import ops
    g = torch.Generator().manual_seed(23)
    B = 70000
    lens = torch.randint(2, 4, (B,), generator=g)
    T = int(lens.sum())
    cu = torch.zeros(B + 1, dtype=torch.int32)
    cu[1:] = torch.cumsum(lens, 0)
    H, hd = 2, 64
    qkv = torch.randn(T, 3 * H * hd, generator=g).bfloat16().to(DEV)
    q, k, v = (qkv[:, i * H * hd:(i + 1) * H * hd].unflatten(1, (H, hd)) for i in range(3))
    got = ops.attn_varlen(q, k, v, cu.to(DEV), 3)
    want = ops.attn_varlen(q, k, v, cu.to(DEV), 3, impl=1)
    assert torch.isfinite(got.float()).all()
    assert (got.float() - want.float()).abs().max() <= 0.04


def test_batch_with_more_than_2_31_activation_elements():
    """600k packed tokens at ESM2-650M width: T x 3D = 2.3e9 and T x F = 3.1e9 elements, beyond 32-bit indexing
    (15 GB of activations).  First, middle and last sequence equal their stand-alone forward bit for bit."""
    layers, D, H = 2, 1280, 20
    sd = synthetic.synthetic_state_dict('esm2', layers, D, seed=3)
    model = esme.ESM2(layers, D, H)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    lens = synthetic.synthetic_lengths(600000, seed=7)
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=8)
    assert tokens.numel() * 3 * D > 2 ** 31
    out = model(tokens.to(DEV), (cu.to(DEV), max_len))
    assert torch.isfinite(out.float()).all()
    for s in (0, len(lens) // 2, len(lens) - 1):
        a, b = int(cu[s]), int(cu[s + 1])
        alone = model(tokens[a:b].contiguous().to(DEV), (torch.tensor([0, b - a], dtype=torch.int32, device=DEV), b - a))
        assert torch.equal(out[a:b], alone), s
    del out
    torch.cuda.empty_cache()


def test_repeated_forwards_are_bit_identical_under_interference():
    """30 forwards of a 650M-width model (TMA-store and TMA-residual epilogues, two CTAs per SM in attention) while a
    second stream keeps HBM and L2 busy: every result equals the first bit for bit (a staging-slot or barrier race
    would show as run-to-run differences)."""
    layers, D, H = 2, 1280, 20
    sd = synthetic.synthetic_state_dict('esm2', layers, D, seed=3)
    model = esme.ESM2(layers, D, H)
    model.load_state_dict(sd, strict=True)
    model = model.to(DEV).eval()
    lens = synthetic.synthetic_lengths(20000, seed=12)
    tokens, cu, max_len = synthetic.synthetic_batch(lens, seed=13)
    tokens, cu = tokens.to(DEV), cu.to(DEV)
    first = model(tokens, (cu, max_len)).clone()
    noise = torch.empty(64 << 20, dtype=torch.float32, device=DEV)
    side = torch.cuda.Stream()
    for i in range(30):
        if i % 2:
            with torch.cuda.stream(side):
                noise.normal_()
                noise.mul_(1.0001)
        assert torch.equal(model(tokens, (cu, max_len)), first), i
    torch.cuda.synchronize()
