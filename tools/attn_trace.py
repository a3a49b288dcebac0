"""Per-block phase latencies of the round-2 attention kernel (debug aid; needs the tracing build:
make -C esm-efficient_b200/csrc OBJ_DIR=build_trace OUT_DIR=../esme/_lib_trace EXTRA=-DESMK_ATTN_TRACING).

    ESMK_LIB_PATH=esm-efficient_b200/esme/_lib_trace/libesmk.so python tools/attn_trace.py
"""
import collections
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))
path = os.path.join(ROOT, 'gpurun_out', 'attn_v3_trace.txt')
os.makedirs(os.path.dirname(path), exist_ok=True)
import torch
from esme import ops, synthetic

dev = 'cuda'
lens = synthetic.synthetic_lengths(50000, seed=2)
T, H, hd = sum(lens), 20, 64
D = H * hd
cu = torch.zeros(len(lens) + 1, dtype=torch.int32)
cu[1:] = torch.cumsum(torch.tensor(lens), 0)
cu = cu.to(dev)
qkv = torch.randn(T, 3 * D, device=dev).bfloat16()
q, k, v = (qkv[:, i * D:(i + 1) * D].unflatten(1, (H, hd)) for i in range(3))
_, info = ops.batch_meta(cu, T)
for _ in range(3):
    ops.attn_varlen(q, k, v, cu, max(lens), info)
torch.cuda.synchronize()
os.environ['ESMK_ATTN_TRACE'] = path
ops.attn_varlen(q, k, v, cu, max(lens), info)
torch.cuda.synchronize()
rows = [list(map(int, l.split())) for l in open(path)]
print('records', len(rows))
names = ['wait_s_full', 'softmax (max, rescale, exp, st)', 'named barrier', 'issue P.V + commits', 'wait kv_full (+q_full)', 'issue S MMAs', 'commit S (+q_empty)']
per = collections.defaultdict(list)
period = collections.defaultdict(list)
by = collections.defaultdict(list)
for r in rows:
    by[(r[0], r[1])].append(r)
for key, rr in by.items():
    rr.sort(key=lambda r: r[2])
    for i, r in enumerate(rr):
        st = r[3:]
        if min(st) == 0:
            continue
        for p in range(7):
            per[(p, 'issuer' if key[1] in (1, 5) else 'other')].append(st[p + 1] - st[p])
        if i + 1 < len(rr) and min(rr[i + 1][3:]) > 0:
            period['issuer' if key[1] in (1, 5) else 'other'].append(rr[i + 1][3] - st[0])
for role in ('issuer', 'other'):
    print(f'--- {role} warps: cycles per 64-key block, median / mean / p90   (period median {statistics.median(period[role]):.0f}, mean {statistics.mean(period[role]):.0f})')
    for p in range(7):
        v = sorted(per[(p, role)])
        print(f'   {names[p]:34s} {v[len(v) // 2]:7d} {statistics.mean(v):9.1f} {v[int(len(v) * .9)]:7d}   n={len(v)}')
