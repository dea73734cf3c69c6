import torch, numpy as np
def t(fn, it=10):
    for _ in range(3): fn()
    ts=[]
    for _ in range(it):
        a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True); a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b)*1e3)
    return float(np.median(ts))
n=671088640//4
x=torch.empty(n,device='cuda'); y=torch.empty(n,device='cuda')
us=t(lambda: x.zero_()); print(f'memset 671 MB: {us:.1f} us -> {671.09e6/us/1e3:.0f} GB/s')
us=t(lambda: y.copy_(x)); print(f'copy 671 MB (r+w): {us:.1f} us -> {2*671.09e6/us/1e3:.0f} GB/s')
src=torch.randn(40960,256,device='cuda'); idx=torch.randint(0,40960,(40960*16,),device='cuda')
us=t(lambda: torch.index_select(src,0,idx)); print(f'torch index_select gather 671 MB: {us:.1f} us -> {671.09e6/us/1e3:.0f} GB/s written')
