"""Masked-LM evaluation losses on the model's outputs (drop-in for esme/loss.py:5-54).  Pure selection + reduction
over the [T, V] outputs of `predict_log_prob` / `forward`; used by perplexity evaluation, not by the forward path."""
import torch.nn.functional as F

from .alphabet import Alphabet3


def _select(rows, tokens, mask):
    """Rows / targets at the masked positions, flattened over any leading dims."""
    rows = rows.reshape(-1, rows.size(-1))
    keep = mask.reshape(-1)
    return rows[keep], tokens.reshape(-1)[keep]


def nll_loss(log_probs, tokens, mask, nll_loss_kwargs=None, alphabet=Alphabet3):
    """Negative log-likelihood of the original `tokens` at the positions `mask` marks (from `mask_tokens`), given
    `model.predict_log_prob(masked_tokens, ...)`; padding targets are ignored."""
    picked, targets = _select(log_probs, tokens, mask)
    return F.nll_loss(picked, targets, ignore_index=alphabet.padding_idx, **(nll_loss_kwargs or {}))


def cross_entropy(logits, tokens, mask, cross_entropy_loss_kwargs=None, alphabet=Alphabet3):
    """Same from raw logits (`model(masked_tokens, ...)`)."""
    picked, targets = _select(logits, tokens, mask)
    return F.cross_entropy(picked, targets, ignore_index=alphabet.padding_idx, **(cross_entropy_loss_kwargs or {}))
