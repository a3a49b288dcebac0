"""The product-side synthetic generators (bench / full-size tests) equal the oracle's."""
import torch

from esme import synthetic as S
from oracle import esm_oracle as O


def test_generators_agree_with_oracle():
    for fam, dims in (('esm2', (2, 64, 4)), ('esmc', (2, 128, 2))):
        cfg = O.OracleConfig(fam, *dims)
        a = O.synthetic_weights(cfg, seed=5)
        b = S.synthetic_state_dict(fam, dims[0], dims[1], seed=5)
        assert a.keys() == b.keys() and all(torch.equal(a[k], b[k]) for k in a)
        assert S.ffn_dim(fam, dims[1]) == cfg.ffn_dim
    assert O.synthetic_lengths(20000, 4, 'loguniform') == S.synthetic_lengths(20000, 4, 'loguniform')
    la = S.synthetic_lengths(50000, 2)
    assert la == O.synthetic_lengths(50000, 2)
    ta, tb = O.synthetic_batch(la[:5], 3), S.synthetic_batch(la[:5], 3)
    assert torch.equal(ta[0], tb[0]) and torch.equal(ta[1], tb[1]) and ta[2] == tb[2]


def test_flop_model_matches_survey_constants():
    f = S.forward_flops('esm2', 33, 1280, [1000])
    assert abs(f['gemm'] / 1000 - 1297.6e6) / 1297.6e6 < 1e-3
    assert abs(f['attention'] / 1000 - 168960 * 1000) < 1
    f = S.forward_flops('esmc', 30, 960, [1000])
    assert abs(f['gemm'] / 1000 - 663.6e6) / 663.6e6 < 1e-3
