"""Caller side of oracle/ref_runner.py: run one job on the real reference (oracle/_ref) in a child interpreter.
TEST INFRASTRUCTURE ONLY (tests/ and bench.py's reference legs)."""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))


def ref_available() -> bool:
    return os.path.isfile(os.path.join(HERE, '_ref', 'esme_ref.zip'))


def decode(a: np.ndarray) -> torch.Tensor:
    """uint16 arrays are raw bf16 bit patterns -> float32 tensors holding exact bf16 values."""
    if a.dtype == np.uint16:
        return torch.from_numpy(a.view(np.int16).copy()).view(torch.bfloat16).float()
    return torch.from_numpy(a.copy())


def run_reference(job: dict, batch=None, tokens2d=None, timeout: int = 900):
    """-> (info dict printed by the runner, {name: tensor} outputs).  `batch` = (tokens, cu_lens, max_len) host
    tensors (written to a temporary .npz) unless job['batch'] is already given."""
    with tempfile.TemporaryDirectory(prefix='esmk_ref_') as tmp:
        job = dict(job)
        if batch is not None:
            tokens, cu, max_len = batch
            job['batch'] = os.path.join(tmp, 'batch.npz')
            np.savez(job['batch'], tokens=tokens.cpu().numpy(), cu_lens=cu.cpu().numpy().astype(np.int32),
                     max_len=np.int64(max_len))
        if tokens2d is not None:
            job['tokens2d'] = os.path.join(tmp, 'tokens2d.npz')
            np.savez(job['tokens2d'], tokens2d=tokens2d.cpu().numpy())
        job.setdefault('out', os.path.join(tmp, 'out.npz'))
        jpath = os.path.join(tmp, 'job.json')
        json.dump(job, open(jpath, 'w'))
        env = {k: v for k, v in os.environ.items() if k != 'PYTHONPATH'}
        r = subprocess.run([sys.executable, os.path.join(HERE, 'ref_runner.py'), jpath], capture_output=True,
                           text=True, timeout=timeout, env=env, cwd=tmp)
        lines = [l for l in r.stdout.splitlines() if l.startswith('@@REF@@')]
        if r.returncode != 0 or not lines:
            raise RuntimeError(f'reference runner failed (rc={r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-4000:]}')
        info = json.loads(lines[-1][len('@@REF@@'):])
        outs = {}
        if os.path.isfile(job['out']):
            with np.load(job['out']) as z:
                outs = {k: decode(z[k]) for k in z.files}
        return info, outs
