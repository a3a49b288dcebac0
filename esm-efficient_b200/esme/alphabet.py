"""Vocabularies and host-side tokenisation (drop-in for esme/alphabet.py of the
reference).  Pure host integer work: outputs are byte-identical to the
reference's (pinned by tests/golden/tokenizer.json)."""
import re
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

_HEAD = ['<cls>', '<pad>', '<eos>', '<unk>']
_BODY = list('LAGVSERTIDPKQNFYMHWCXBUZO')


class _Vocabulary:
    """Shared machinery; subclasses only provide `alphabet` (reference: esme/alphabet.py:9-56)."""
    alphabet: List[str] = []

    def __init_subclass__(cls, **kw):
        super().__init_subclass__(**kw)
        cls.amino_acids = cls.alphabet[4:24]
        cls.amino_acids_idx = list(range(4, 24))
        cls.idx_to_token = dict(enumerate(cls.alphabet))
        cls.token_to_idx = {tok: i for i, tok in enumerate(cls.alphabet)}
        cls.cls_idx = cls.token_to_idx[cls._names[0]]
        cls.padding_idx = cls.token_to_idx[cls._names[1]]
        cls.eos_idx = cls.token_to_idx[cls._names[2]]
        cls.unk_idx = cls.token_to_idx[cls._names[3]]
        cls.mask_idx = cls.token_to_idx[cls._names[4]]

    _names = ('<cls>', '<pad>', '<eos>', '<unk>', '<mask>')


class Alphabet(_Vocabulary):
    """ESM-1b / ESM-1v / ESM2 vocabulary (33 tokens)."""
    alphabet = _HEAD + _BODY + ['.', '-', '<null_1>', '<mask>']


class Alphabet3(_Vocabulary):
    """ESMC vocabulary: identical indices except slot 31 ('|')."""
    alphabet = _HEAD + _BODY + ['.', '-', '|', '<mask>']


_TOKEN = re.compile(r'<[^>]+>|.')


def split_alphabet(seq: Union[str, Sequence[str]]):
    """'MK<mask>A' -> ['M', 'K', '<mask>', 'A'] (a list of such lists for a list input)."""
    if isinstance(seq, str):
        return _TOKEN.findall(seq)
    return [_TOKEN.findall(s) for s in seq]


def token_to_str(tokens: Tensor, alphabet=Alphabet3) -> List[str]:
    return [''.join(alphabet.idx_to_token[i] for i in row) for row in tokens.tolist()]


def _encode(pieces: Sequence[str], alphabet) -> List[int]:
    lut, unk = alphabet.token_to_idx, alphabet.unk_idx
    return [alphabet.cls_idx, *(lut.get(p, unk) for p in pieces), alphabet.eos_idx]


_BYTE_LUT = {}


def _byte_lut(alphabet) -> np.ndarray:
    """256-entry table ASCII byte -> token index (unknown -> <unk>) for the single-character tokens."""
    t = _BYTE_LUT.get(alphabet)
    if t is None:
        t = np.full(256, alphabet.unk_idx, dtype=np.int64)
        for tok, i in alphabet.token_to_idx.items():
            if len(tok) == 1 and ord(tok) < 128:
                t[ord(tok)] = i
        _BYTE_LUT[alphabet] = t
    return t


def _encode_rows(sequences: Sequence[str], alphabet) -> List[np.ndarray]:
    """[cls] + tokens + [eos] per sequence as int64 arrays.  Plain residue strings (ASCII, no '<...>' token, no
    newline -- '.' in the reference's regex does not match one) go through a byte lookup table, ~25x faster than
    the regex split; anything else takes the regex path.  Both paths give identical indices."""
    lut = _byte_lut(alphabet)
    rows = []
    for s in sequences:
        if '<' in s or '\n' in s or not s.isascii():
            rows.append(np.asarray(_encode(_TOKEN.findall(s), alphabet), dtype=np.int64))
        else:
            r = np.empty(len(s) + 2, dtype=np.int64)
            r[0], r[-1] = alphabet.cls_idx, alphabet.eos_idx
            r[1:-1] = lut[np.frombuffer(s.encode('ascii'), dtype=np.uint8)]
            rows.append(r)
    return rows


def tokenize(sequences: Union[List[str], str], alphabet=Alphabet3) -> Tensor:
    """Padded int64 [B, max_len] (reference: esme/alphabet.py:117)."""
    if isinstance(sequences, str):
        sequences = [sequences]
    rows = _encode_rows(sequences, alphabet)
    width = max(map(len, rows))
    arr = np.full((len(rows), width), alphabet.padding_idx, dtype=np.int64)
    for i, r in enumerate(rows):
        arr[i, :len(r)] = r
    return torch.from_numpy(arr)


def tokenize_unpad(sequences: Union[List[str], str], alphabet=Alphabet3) -> Tuple[Tensor, Tensor, Tensor, int]:
    """Packed tokens int64[T], indices int64[T] into the virtual [B,max_len] grid,
    cu_lens int32[B+1], max_len int (reference: esme/alphabet.py:148)."""
    if isinstance(sequences, str):
        sequences = [sequences]
    rows = _encode_rows(sequences, alphabet)
    lens = np.fromiter(map(len, rows), dtype=np.int64, count=len(rows))
    max_len = int(lens.max())
    cu = np.zeros(len(rows) + 1, dtype=np.int32)
    np.cumsum(lens, out=cu[1:])
    tokens = np.concatenate(rows)
    indices = np.concatenate([np.arange(l, dtype=np.int64) + i * max_len for i, l in enumerate(lens)])
    return torch.from_numpy(tokens), torch.from_numpy(indices), torch.from_numpy(cu), max_len


def pad_tokens(tokens: List[Tensor], alphabet=Alphabet3) -> Tensor:
    """Stack 1-D token rows (or concatenate [n, len] blocks) padded to a common length."""
    if tokens[0].ndim == 1:
        width = max(t.size(0) for t in tokens)
        return torch.stack([F.pad(t, (0, width - t.size(0)), value=alphabet.padding_idx) for t in tokens])
    width = max(t.size(1) for t in tokens)
    return torch.cat([F.pad(t, (0, width - t.size(1)), value=1) for t in tokens], dim=0)


def mask_tokens(token: Tensor, freq: float = 0.15, alter: float = 0.1, alphabet=Alphabet3):
    """BERT-style corruption (reference: esme/alphabet.py:215): returns (tokens, mask).
    Of the selected positions ~80 % become <mask>, ~10 % a random residue, ~10 % stay."""
    original = token
    token = token.clone()
    special = (token == alphabet.cls_idx) | (token == alphabet.eos_idx) | (token == alphabet.padding_idx)
    valid = ~special
    mask = (torch.rand_like(token, dtype=torch.float32) < freq) & valid
    empty = mask.sum(dim=-1) == 0
    if empty.any():                      # guarantee at least one masked position per row
        pick = torch.multinomial(valid[empty].float(), 1).squeeze(1)
        if token.ndim == 1:
            mask[pick] = True
        elif token.ndim == 2:
            mask[empty, pick] = True
        else:
            raise ValueError('tokens must be 1D or 2D')
    token[mask] = alphabet.mask_idx
    lo, hi = alphabet.amino_acids_idx[0], alphabet.amino_acids_idx[-1] + 1
    swap = (torch.rand_like(token, dtype=torch.float32) < alter) & mask
    token = torch.where(swap, torch.randint_like(token, lo, hi), token)
    keep = (torch.rand_like(token, dtype=torch.float32) < alter) & mask
    token = torch.where(keep, original, token)
    return token, mask


def padding_mask(cu_lens: Tensor, max_len: int) -> Tensor:
    """bool [B, max_len], True on real tokens."""
    lens = cu_lens[1:] - cu_lens[:-1]
    return torch.arange(max_len, device=cu_lens.device)[None, :] < lens[:, None]
