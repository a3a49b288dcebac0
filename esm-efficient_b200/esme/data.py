"""Token-budget batching of FASTA records into packed batches (drop-in for the inference part of
esme/data.py:12-162).  The training datamodules of the reference (Lightning,
masking, labels) are out of scope."""
from typing import List, Optional

import torch
from torch.utils.data import DataLoader, Dataset

from .alphabet import Alphabet3, pad_tokens, tokenize, tokenize_unpad
from .fasta import Fasta


class TokenSizeBatchSampler:
    """Greedy batches of sequence indices whose token count (residues + 2 per sequence) stays within
    `token_per_batch`: a batch is closed when the next sequence would overflow it -- the rule that defines
    the reference's "50k-token batch" (esme/data.py:32-54)."""

    def __init__(self, token_sizes, token_per_batch, drop_last=False, shuffle=True, random_state=None):
        self.token_sizes = token_sizes
        self.token_per_batch = token_per_batch
        self.drop_last = drop_last
        self.shuffle = shuffle
        self.random_state = random_state
        self._batches = list(self.batches())

    def batches(self):
        order = list(range(len(self.token_sizes)))
        if self.shuffle:
            import sklearn.utils                      # same permutation as the reference for a given seed
            order = sklearn.utils.shuffle(order, random_state=self.random_state)
        batch: List[int] = []
        used = 0
        for idx in order:
            need = self.token_sizes[idx] + 2
            if used + need > self.token_per_batch:
                yield batch
                batch, used = [idx], need
            else:
                batch.append(idx)
                used += need
        if batch and not self.drop_last:
            yield batch

    def __getitem__(self, idx):
        return self._batches[idx]

    def __len__(self):
        return len(self._batches)


class FastaDataset(Dataset):
    """One padded token row per record (reference: esme/data.py:81-112)."""

    def __init__(self, fasta, fai=None, k_sample=None, max_len=None, alphabet=Alphabet3):
        self.alphabet = alphabet
        self.fasta = Fasta(fasta, fai=fai, max_len=max_len, k_sample=k_sample)

    def read_seq(self, idx):
        return self.fasta[idx]

    def __len__(self):
        return len(self.fasta)

    def __getitem__(self, idx):
        return tokenize(self.read_seq(idx), alphabet=self.alphabet)

    @staticmethod
    def collate_fn(batch):
        return pad_tokens(batch)

    def to_dataloader(self, batch_size, shuffle=False, num_workers=0, **kwargs):
        return DataLoader(self, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers,
                          collate_fn=self.collate_fn, **kwargs)


class FastaTokenDataset(FastaDataset):
    """Item i = the i-th token-budget batch, already packed: `(tokens int64[T], (cu_lens int32[B+1], max_len))`,
    i.e. exactly the arguments of `model(tokens, pad_args)` (reference: esme/data.py:115-162)."""

    def __init__(self, fasta, fai=None, token_per_batch=50_000, k_sample=None, max_len=None, drop_last=False,
                 shuffle=True, random_state=None, alphabet=Alphabet3):
        super().__init__(fasta, fai=fai, k_sample=k_sample, max_len=max_len, alphabet=alphabet)
        self.token_per_batch = token_per_batch
        lengths = [row['length'] for row in self.fasta.fai]
        self.sampler = list(iter(TokenSizeBatchSampler(lengths, token_per_batch, drop_last=drop_last,
                                                       shuffle=shuffle, random_state=random_state)))

    def __len__(self):
        return len(self.sampler)

    def __getitem__(self, idx):
        token, _, cu_lens, max_len = tokenize_unpad([self.read_seq(i) for i in self.sampler[idx]],
                                                    alphabet=self.alphabet)
        return token, (cu_lens, max_len)

    def to_dataloader(self, num_workers=0, **kwargs):
        return DataLoader(self, num_workers=num_workers, batch_size=None, **kwargs)

    def device_batches(self, device, num_workers: int = 0):
        """Packed batches staged through pinned host memory and copied on a side stream, one batch ahead of
        the consumer, so the H2D copy of batch i+1 overlaps the forward of batch i."""
        dev = torch.device(device)
        copy_stream = torch.cuda.Stream(device=dev)

        def stage(item):
            tok, (cu, ml) = item
            with torch.cuda.stream(copy_stream):
                t = tok.pin_memory().to(dev, non_blocking=True)
                c = cu.pin_memory().to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy_stream)
            return t, c, ml, ev

        it = iter(self.to_dataloader(num_workers=num_workers))
        nxt = next(it, None)
        staged = stage(nxt) if nxt is not None else None
        while staged is not None:
            t, c, ml, ev = staged
            nxt = next(it, None)
            staged = stage(nxt) if nxt is not None else None
            torch.cuda.current_stream(dev).wait_event(ev)
            t.record_stream(torch.cuda.current_stream(dev))
            c.record_stream(torch.cuda.current_stream(dev))
            yield t, (c, ml)
