"""Variant-effect scoring by masked marginals (drop-in for esme/variant.py:10-165 of the
reference).

The reference builds one padded row per residue on the host, pads the output of
`predict_log_prob(..., pad_output=True)` back to [B, L, V], gathers the masked position
and then calls `.item()` once per (position, amino acid) -- 20 host syncs per residue.
Here the masked batch is assembled on the device from ONE copy of the wild-type row, the
packed forward runs up to `batch_size` masked copies at a time, only the masked rows go
through the LM head + log-softmax, and the whole [len(seq), 20] margin matrix comes back
in a single device->host copy.  Scores are identical to the reference's definition:
log p(aa | masked context) - log p(wild type | masked context), bf16 log-probabilities.
"""
from typing import List, Optional, Union

import pandas as pd
import torch

from . import _lib as L
from .alphabet import Alphabet3, tokenize
from .lora import has_lora


class MaskMarginDataset(torch.utils.data.Dataset):
    """One item per residue: the sequence with that residue replaced by <mask>, optionally cut to a
    window of `max_len` tokens centred on it (same fields and windowing rule as the reference,
    esme/variant.py:41-70)."""

    def __init__(self, seq: str, max_len: Optional[int] = None, alphabet=Alphabet3):
        super().__init__()
        self.seq = seq
        self.max_len = max_len
        self.token = tokenize([seq])[0]
        self.alphabet = alphabet

    def __len__(self):
        return len(self.seq)

    def window(self, idx: int):
        """(start, end, local_pos) of the token window for residue `idx` (0-based)."""
        tok_pos = idx + 1                       # <cls> occupies slot 0
        n = self.token.size(0)
        if self.max_len is not None and n > self.max_len:
            start = min(n - self.max_len, max(0, tok_pos - self.max_len // 2))
            end = min(n, start + self.max_len)
            return start, end, tok_pos - start
        return 0, n, tok_pos

    def __getitem__(self, idx):
        start, end, local = self.window(idx)
        token = self.token.clone()
        token[idx + 1] = self.alphabet.mask_idx
        wt = self.seq[idx]
        return {'token': token[start:end], 'local_pos': local, 'pos': idx + 1, 'wt': wt,
                'wt_token': self.alphabet.token_to_idx[wt]}


def _masked_position_log_probs(model, ds: MaskMarginDataset, batch_size: int) -> torch.Tensor:
    """log-probabilities [len(seq), V] (bf16, on the model's device) at each residue's masked position."""
    device = next(model.parameters()).device
    n = len(ds)
    base = ds.token.to(device)                                      # one H2D copy of the wild-type row
    wins = [ds.window(i) for i in range(n)]
    width = wins[0][1] - wins[0][0]
    starts = torch.tensor([w[0] for w in wins], device=device)
    local = torch.tensor([w[2] for w in wins], device=device)
    cols = torch.arange(width, device=device)
    out = []
    model.eval()
    with torch.no_grad():
        for b0 in range(0, n, batch_size):
            b1 = min(n, b0 + batch_size)
            rows = base[(starts[b0:b1, None] + cols[None, :])]      # [b, width] windows of the wild type
            rows[torch.arange(b1 - b0, device=device), local[b0:b1]] = ds.alphabet.mask_idx
            cu = torch.arange(0, (b1 - b0 + 1) * width, width, dtype=torch.int32, device=device)
            z = model.forward_representation(rows.reshape(-1), (cu, width))          # packed, no padding
            masked_rows = z[torch.arange(b1 - b0, device=device) * width + local[b0:b1]]
            masked_rows = masked_rows.contiguous()
            # (a model with LoRA adapters runs the operator path: its engine cannot be built over LoRA-wrapped linears)
            out.append(model._lm_head_ops(masked_rows, L.OUT_LOG_PROB) if has_lora(model)
                       else model.engine().lm_head(masked_rows, L.OUT_LOG_PROB))
    return torch.cat(out)


def predict_mask_margin(model, seq: Union[str, MaskMarginDataset], batch_size: int = 32, max_len=None,
                        alphabet=Alphabet3) -> pd.DataFrame:
    """DataFrame indexed by 'variant' (e.g. 'M1A') with one 'score' column: 20 rows per residue in
    `alphabet.amino_acids` order, residues in sequence order (reference: esme/variant.py:110-165)."""
    if isinstance(seq, str):
        ds = MaskMarginDataset(seq, max_len=max_len, alphabet=alphabet)
    elif isinstance(seq, MaskMarginDataset):
        ds = seq
    else:
        raise ValueError('seq must be str or MaskMarginDataset')
    logp = _masked_position_log_probs(model, ds, batch_size)        # [n, V]
    wt = torch.tensor([alphabet.token_to_idx[a] for a in ds.seq], device=logp.device)
    margin = logp - logp.gather(1, wt[:, None])                     # bf16 arithmetic, as in the reference
    aa_idx = torch.tensor(alphabet.amino_acids_idx, device=logp.device)
    scores = margin[:, aa_idx].float().cpu().numpy()                # the one device->host copy
    variants: List[str] = [f'{w}{i + 1}{aa}' for i, w in enumerate(ds.seq) for aa in alphabet.amino_acids]
    return pd.DataFrame({'variant': variants, 'score': scores.reshape(-1)}).set_index('variant')


def predict_pseudoperplexity(model, seq: Union[str, MaskMarginDataset], batch_size: int = 32, max_len=None,
                             alphabet=Alphabet3) -> float:
    """exp(mean over residues of -log p(wild type | residue masked)) (reference: esme/variant.py:168-215,
    which feeds the masked-position logits to torchmetrics' Perplexity)."""
    ds = seq if isinstance(seq, MaskMarginDataset) else MaskMarginDataset(seq, max_len=max_len, alphabet=alphabet)
    logp = _masked_position_log_probs(model, ds, batch_size).float()
    wt = torch.tensor([alphabet.token_to_idx[a] for a in ds.seq], device=logp.device)
    return float(torch.exp(-logp.gather(1, wt[:, None]).mean()))
