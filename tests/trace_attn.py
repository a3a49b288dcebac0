"""Latency trace of the attention kernel (debug aid): ESMK_ATTN_TRACE stamps -> per-phase cycle table."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))
os.environ['ESMK_ATTN_TRACE'] = os.path.join(ROOT, 'gpurun_out', 'attn_trace.txt')
import torch
from esme import ops
from oracle import esm_oracle as O
dev = 'cuda'
lens = O.synthetic_lengths(50000, seed=2)
T, H, hd = sum(lens), 20, 64
D = H * hd
cu = torch.zeros(len(lens) + 1, dtype=torch.int32); cu[1:] = torch.cumsum(torch.tensor(lens), 0); cu = cu.to(dev)
qkv = torch.randn(T, 3 * D, device=dev).bfloat16()
q, k, v = (qkv[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
_, info = ops.batch_meta(cu, T)
for _ in range(2):
    ops.attn_varlen(q, k, v, cu, max(lens), info)
torch.cuda.synchronize()
rows = [list(map(int, l.split())) for l in open(os.environ['ESMK_ATTN_TRACE'])]
print('records', len(rows))
import collections
for role, names in ((0, ['wait_s_full', 'ldtm', 'arrive+row max', 'rescale (+o_done wait)', 'wait_o_done', 'exp chunks 0-1', 'exp chunks 2-3 + wait_st']),
                    (1, ['wait_s_free', 'issue_S', 'wait_p_full', 'wait_v_full', 'issue_PV'])):
    acc = collections.defaultdict(list)
    per_block = []
    for cta in sorted(set(r[0] for r in rows)):
        rr = sorted([r for r in rows if r[0] == cta and r[1] == role], key=lambda r: r[2])
        for a, b in zip(rr[:-1], rr[1:]):
            per_block.append(b[3] - a[3])
        for r in rr:
            st = r[3:]
            for i, n in enumerate(names):
                if st[i + 1] and st[i]:
                    acc[n].append(st[i + 1] - st[i])
    print('role', 'softmax' if role == 0 else 'mma', 'cycles per block (median/mean):',
          sorted(per_block)[len(per_block) // 2] if per_block else None, sum(per_block) / max(1, len(per_block)))
    for n in names:
        v = sorted(acc[n])
        if v:
            print(f'   {n:16s} median {v[len(v)//2]:7d}  mean {sum(v)/len(v):9.1f}  p90 {v[int(len(v)*0.9)]:7d}  n={len(v)}')
