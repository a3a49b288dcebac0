"""Random-access FASTA reader over a samtools-style .fai index (drop-in for esme/fasta.py:5-100;
SURVEY.md §8f row 3, the step before the hot path).  No polars dependency: the index is a list of dicts,
and it is built in memory when no .fai file exists."""
import os
from typing import Dict, List, Optional

_COLS = ('id', 'length', 'offset', 'line_bases', 'line_width')


def read_fai(fai_path: str) -> List[Dict]:
    """[{'id', 'length', 'offset', 'line_bases', 'line_width'}, ...] in file order."""
    rows = []
    with open(fai_path) as f:
        for line in f:
            if not line.strip():
                continue
            p = line.rstrip('\n').split('\t')
            rows.append({'id': p[0], 'length': int(p[1]), 'offset': int(p[2]), 'line_bases': int(p[3]),
                         'line_width': int(p[4])})
    return rows


def build_fai(fasta_path: str) -> List[Dict]:
    """The same index computed from the FASTA file itself (what `samtools faidx` would write)."""
    rows, cur = [], None
    with open(fasta_path, 'rb') as f:
        pos = 0
        for raw in f:
            if raw.startswith(b'>'):
                cur = {'id': raw[1:].split()[0].decode(), 'length': 0, 'offset': pos + len(raw), 'line_bases': 0,
                       'line_width': 0}
                rows.append(cur)
            elif cur is not None and raw.strip():
                if cur['line_bases'] == 0:
                    cur['line_bases'], cur['line_width'] = len(raw.strip()), len(raw)
                cur['length'] += len(raw.strip())
            pos += len(raw)
    return rows


class Fasta:
    """`fasta[i]` / `fasta['id']` -> sequence string; `len(fasta)`; optional `max_len` filter and `k_sample`."""

    def __init__(self, fasta: str, fai: Optional[str] = None, max_len: Optional[int] = None,
                 k_sample: Optional[int] = None, random_state=None):
        if not os.path.exists(fasta):
            raise FileNotFoundError(f'File not found: {fasta}')
        self.fasta = fasta
        fai = fai or fasta + '.fai'
        self.fai = read_fai(fai) if os.path.exists(fai) else build_fai(fasta)
        if max_len is not None:
            self.fai = [r for r in self.fai if r['length'] <= max_len]
        if k_sample is not None:
            import numpy as np
            pick = np.random.RandomState(random_state).choice(len(self.fai), k_sample, replace=False)
            self.fai = [self.fai[i] for i in pick]
        self.proteins = {row['id']: i for i, row in enumerate(self.fai)}

    def __getitem__(self, idx):
        if isinstance(idx, int):
            return self.read_seq(idx)
        if isinstance(idx, str):
            return self.read_seq(self.proteins[idx])
        raise ValueError(f'Invalid index: {idx}')

    def read_seq(self, idx: int) -> str:
        row = self.fai[idx]
        n_lines = -(-row['length'] // max(row['line_bases'], 1))
        with open(self.fasta, 'rb') as f:
            f.seek(row['offset'])
            raw = f.read(row['length'] + n_lines * max(row['line_width'] - row['line_bases'], 1) + 2)
        seq = raw.split(b'>')[0].translate(None, b'\r\n \t').decode()[:row['length']]
        assert len(seq) == row['length']
        return seq

    def __len__(self):
        return len(self.fai)
