"""Host-side tokenisation of the drop-in package vs the reference's known answers
(tests/golden/tokenizer.json generated from the reference; p53 vector quoted in the
reference's tests/test_alphabet.py:11-34)."""
import json

import pytest
import torch

from conftest import GOLDEN
from esme.alphabet import (Alphabet, Alphabet3, mask_tokens, pad_tokens, padding_mask, split_alphabet, token_to_str,
                           tokenize, tokenize_unpad)
from oracle import esm_oracle as O

CASES = json.load(open(f'{GOLDEN}/tokenizer.json'))


@pytest.mark.parametrize('name', sorted(CASES))
def test_tokenizers_match_reference(name):
    c = CASES[name]
    alph = Alphabet if c['alphabet'] == 'esm2' else Alphabet3
    padded = tokenize(c['seqs'], alph)
    assert padded.dtype == torch.int64 and padded.tolist() == c['padded']
    tokens, indices, cu_lens, max_len = tokenize_unpad(c['seqs'], alph)
    assert tokens.dtype == torch.int64 and indices.dtype == torch.int64 and cu_lens.dtype == torch.int32
    assert tokens.tolist() == c['tokens'] and indices.tolist() == c['indices']
    assert cu_lens.tolist() == c['cu_lens'] and max_len == c['max_len'] and isinstance(max_len, int)
    assert padding_mask(cu_lens, max_len).int().tolist() == c['padding_mask']
    # the oracle's own restatement agrees too
    table = O.ALPHABET_ESM2 if c['alphabet'] == 'esm2' else O.ALPHABET_ESMC
    assert O.tokenize(c['seqs'], table).tolist() == c['padded']
    ot = O.tokenize_unpad(c['seqs'], table)
    assert ot[0].tolist() == c['tokens'] and ot[1].tolist() == c['indices'] and ot[2].tolist() == c['cu_lens']


def test_p53_known_answer_head_and_tail():
    t = tokenize(CASES['p53']['seqs'])
    assert t.shape == (1, 395)
    assert t[0, :12].tolist() == [0, 20, 9, 9, 14, 16, 8, 13, 14, 8, 7, 9]
    assert t[0, -6:].tolist() == [6, 14, 13, 8, 13, 2]


def test_unpad_is_padded_minus_pads():
    seqs = CASES['p53_list']['seqs']
    padded = tokenize(seqs)
    tokens, indices, cu_lens, max_len = tokenize_unpad(seqs)
    assert torch.equal(padded.flatten()[indices], tokens)
    assert torch.equal(padded.flatten()[indices], padded[padded != Alphabet3.padding_idx])


def test_vocabulary_constants():
    for a in (Alphabet, Alphabet3):
        assert (a.cls_idx, a.padding_idx, a.eos_idx, a.unk_idx, a.mask_idx) == (0, 1, 2, 3, 32)
        assert len(a.alphabet) == 33 and a.amino_acids == list('LAGVSERTIDPKQNFYMHWC')
    assert Alphabet.alphabet[31] == '<null_1>' and Alphabet3.alphabet[31] == '|'
    assert split_alphabet('MK<mask>A') == ['M', 'K', '<mask>', 'A']
    assert token_to_str(tokenize('MKA'))[0] == '<cls>MKA<eos>'


def test_pad_and_mask_tokens():
    rows = [torch.tensor([0, 5, 2]), torch.tensor([0, 5, 6, 7, 2])]
    p = pad_tokens(rows)
    assert p.tolist() == [[0, 5, 2, 1, 1], [0, 5, 6, 7, 2]]
    torch.manual_seed(0)
    tok = tokenize(['ACDEFGHIKLMNPQRSTVWY' * 5, 'MKT'])
    masked, mask = mask_tokens(tok, freq=0.15)
    assert mask.any(dim=1).all()                                   # at least one per row
    special = (tok == 0) | (tok == 1) | (tok == 2)
    assert not (mask & special).any() and torch.equal(masked[~mask], tok[~mask])


def test_fast_tokenizer_path_is_identical_to_the_regex_path():
    """Plain residue strings take a byte lookup table (33 M residues/s on one core); strings with '<...>' tokens,
    newlines or non-ASCII characters take the reference's regex split (esme/alphabet.py:79-98).  Same indices."""
    import random
    from esme import alphabet as A
    random.seed(7)
    chars = 'ACDEFGHIKLMNPQRSTVWYXBUZO.-|*?acd <>\n1é'
    for alpha in (A.Alphabet, A.Alphabet3):
        for _ in range(2000):
            s = ''.join(random.choice(chars) for _ in range(random.randint(1, 40)))
            if random.random() < 0.3:
                s = s.replace('<', '').replace('>', '')
            if random.random() < 0.2:
                s += '<mask>K'
            assert A._encode_rows([s], alpha)[0].tolist() == A._encode(A._TOKEN.findall(s), alpha), s


def test_masked_lm_losses():
    """esme.loss (reference esme/loss.py, tests/test_loss.py): scalar, positive, equal to the mean of the picked
    negative log-probabilities, padding targets ignored."""
    import torch
    from esme.alphabet import Alphabet3
    from esme.loss import cross_entropy, nll_loss
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(12, 33, generator=g)
    logp = torch.log_softmax(logits, -1)
    tokens = torch.randint(4, 24, (12,), generator=g)
    tokens[3] = Alphabet3.padding_idx
    mask = torch.zeros(12, dtype=torch.bool)
    mask[[1, 3, 5, 8]] = True
    want = -(logp[1, tokens[1]] + logp[5, tokens[5]] + logp[8, tokens[8]]) / 3
    assert torch.allclose(nll_loss(logp, tokens, mask), want) and nll_loss(logp, tokens, mask).shape == ()
    assert torch.allclose(cross_entropy(logits, tokens, mask), want)
    assert torch.allclose(nll_loss(logp.reshape(3, 4, 33), tokens.reshape(3, 4), mask.reshape(3, 4)), want)
