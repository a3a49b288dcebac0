import sys, json
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/esm-efficient_b200')
import torch
from esme import ops, synthetic
from flash_attn import flash_attn_varlen_func
dev='cuda'
lens=synthetic.synthetic_lengths(50000,seed=2)
H,hd=40,128; D=H*hd; T=sum(lens)
g=torch.Generator(device=dev).manual_seed(5)
qkv=torch.randn(T,3*D,generator=g,device=dev).to(torch.bfloat16)
q,k,v=(qkv[:,i*D:(i+1)*D].view(T,H,hd) for i in range(3))
cu=torch.zeros(len(lens)+1,dtype=torch.int32,device=dev); cu[1:]=torch.cumsum(torch.tensor(lens,dtype=torch.int32,device=dev),0)
_,ti=ops.batch_meta(cu,T)
fl=4.0*D*sum(l*l for l in lens)
def t(fn,n=10):
    for _ in range(2): y=fn()
    torch.cuda.synchronize(); e0,e1=torch.cuda.Event(True),torch.cuda.Event(True); e0.record()
    for _ in range(n): y=fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1)/n, y
ms_new,y=t(lambda: ops.attn_varlen(q,k,v,cu,max(lens),ti))
ms_fa,yf=t(lambda: flash_attn_varlen_func(q,k,v,cu,cu,max(lens),max(lens)))
ms_gen,yg=t(lambda: ops.attn_varlen(q,k,v,cu,max(lens),ti,impl=1),n=2)
print(json.dumps(dict(hd=128,H=H,T=T,tcgen05_ms=ms_new,tflops=fl/ms_new/1e9,flash_attn_ms=ms_fa,fa_tflops=fl/ms_fa/1e9,cuda_core_ms=ms_gen,max_abs_vs_fa=(y.float()-yf.reshape(T,D).float()).abs().max().item())))
