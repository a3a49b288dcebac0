"""CTA-level timeline of the attention kernel (debug aid)."""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'esm-efficient_b200'))
path = os.path.join(ROOT, 'gpurun_out', 'attn_cta_trace.txt')
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
for _ in range(3):
    ops.attn_varlen(q, k, v, cu, max(lens), info)
torch.cuda.synchronize()
os.environ['ESMK_ATTN_CTA_TRACE'] = path
ops.attn_varlen(q, k, v, cu, max(lens), info)
torch.cuda.synchronize()
rows = [list(map(int, l.split())) for l in open(path)]
t0 = min(r[1] for r in rows); t1 = max(r[3] for r in rows)
print('ctas', len(rows), 'kernel span us', (t1 - t0) / 1e3)
setup = sorted(r[2] - r[1] for r in rows)
print('setup ns: median', setup[len(setup)//2], 'p90', setup[int(len(setup)*.9)], 'max', setup[-1])
per_block = sorted((r[3] - r[2]) / r[5] for r in rows)
print('loop ns per block: median', per_block[len(per_block)//2], 'p10', per_block[len(per_block)//10], 'p90', per_block[int(len(per_block)*.9)])
by_nb = collections.defaultdict(list)
for r in rows: by_nb[r[5]].append((r[3] - r[2]) / r[5])
for nb in sorted(by_nb): print('  blocks', nb, 'n', len(by_nb[nb]), 'ns/block median', sorted(by_nb[nb])[len(by_nb[nb])//2])
sm = collections.defaultdict(list)
for r in rows: sm[r[4]].append(r)
busy = []; last_end = []
for s, rr in sm.items():
    tot = sum(r[3] - r[1] for r in rr)
    busy.append(tot / (2 * (t1 - t0)))
    last_end.append(max(r[3] for r in rr) - t0)
print('SMs', len(sm), 'slot occupancy mean %.3f min %.3f' % (sum(busy)/len(busy), min(busy)))
le = sorted(last_end)
print('SM finish time us: min %.1f median %.1f max %.1f' % (le[0]/1e3, le[len(le)//2]/1e3, le[-1]/1e3))
starts = sorted(r[1] - t0 for r in rows)
print('last CTA start us %.1f ; 50%% of CTAs started by us %.1f' % (starts[-1]/1e3, starts[len(starts)//2]/1e3))
long = sorted(rows, key=lambda r: -(r[3]-r[1]))[:5]
for r in long: print('  longest cta', r[0], 'blocks', r[5], 'start us %.1f dur us %.1f' % ((r[1]-t0)/1e3, (r[3]-r[1])/1e3))
